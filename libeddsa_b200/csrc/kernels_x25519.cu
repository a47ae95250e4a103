// X25519 variable-base ladder kernel (sm_100a): one Diffie-Hellman per thread, persistent grid-stride
// blocks, all state in registers, constant time (the scalar only feeds cswap masks).
// Replaces x25519() / DH(): /root/reference/lib/x25519.c:129-150, 215-222, 236-243.
#include "kernel_common.cuh"
using namespace edg;

#ifndef EDG_LB_X25519
#define EDG_LB_X25519 4     /* min resident blocks per SM the register allocator must allow (tuned, see profiles/) */
#endif
namespace {

__global__ void __launch_bounds__(kThreads, EDG_LB_X25519) k_x25519(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += T * EDG_BATCH) {
        fe X[EDG_BATCH], Z[EDG_BATCH];
        int cnt = 0;
#pragma unroll 1
        for (int k = 0; k < EDG_BATCH; k++) {                 // phase 1: ladders
            const size_t i = i0 + (size_t)k * T;
            if (i >= n) break;
            u32 s[8], p[8];
            load8(s, scalar, i);
            load8(p, point, i);
            x25519_front(X[k], Z[k], s, p);
            cnt++;
        }
        fe_batch_inv(Z, cnt);                                  // phase 2: one inversion for the whole batch
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {                        // phase 3: X / Z, canonical bytes
            u32 o[8];
            x25519_back(o, X[k], Z[k]);
            store8(out, i0 + (size_t)k * T, o);
        }
        scrub(X, EDG_BATCH);
        scrub(Z, EDG_BATCH);
    }
}

// Diagnostic kernel: one field operation per thread on raw 256-bit operands (weakly reduced inputs are
// legal), so the PTX carry-chain arithmetic of fe.cuh can be checked on the GPU against big integers.
__global__ void __launch_bounds__(kThreads) k_fe_selftest(size_t n, uint8_t *out, const uint8_t *a, const uint8_t *b, int op) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        fe x, y, r;
        load8(x.v, a, i);
        load8(y.v, b, i);
        switch (op) {
        case 0: fe_mul(r, x, y); break;
        case 1: fe_sq(r, x); break;
        case 2: fe_add(r, x, y); break;
        case 3: fe_sub(r, x, y); break;
        case 4: fe_mul121665(r, x); break;
        case 5: fe_canon(r, x); break;
        case 6: fe_inv(r, x); break;
        case 7: fe_pow2523(r, x); break;
        case 8: fe_neg(r, x); break;
        default: fe_copy(r, x); break;
        }
        store8(out, i, r.v);
    }
}

}  // namespace

extern "C" {

int edg_launch_fe_selftest(size_t n, uint8_t *out, const uint8_t *a, const uint8_t *b, int op, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_fe_selftest, n, 0, sm_count, nullptr);
    k_fe_selftest<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, out, a, b, op);
    return (int)cudaGetLastError();
}


int edg_launch_x25519(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_x25519, n, 0, sm_count, nullptr);
    k_x25519<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, out, scalar, point);
    return (int)cudaGetLastError();
}

}  // extern "C"
