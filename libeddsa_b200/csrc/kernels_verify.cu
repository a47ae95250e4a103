// Verification kernel (sm_100a): one signature per thread — SHA-512 challenge, decompression of A,
// Straus double-scalar multiplication S*B + t*(-A) with fixed signed 4-bit windows (uniform control
// flow), encoding and byte comparison.  Public data only, so table lookups are direct-indexed.
// Also hosts pk_ed25519_to_x25519 (it shares the decompression).
// Replaces ed25519_verify / pk_ed25519_to_x25519: /root/reference/lib/ed25519-sha512.c:148-237.
#include "kernel_common.cuh"
using namespace edg;

#ifndef EDG_LB_VERIFY
#define EDG_LB_VERIFY 4     /* min resident blocks per SM the register allocator must allow: 128 registers, 4 warps/SMSP
                               (measured +4 % over 3 blocks, profiles/r01_summary.md) */
#endif
namespace {

__global__ void __launch_bounds__(kThreads, EDG_LB_VERIFY) k_verify(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                                                     const unsigned long long *off, unsigned long long fixed_len, u32 *scratch,
                                                     const u32 *__restrict__ wtab) {
    u32 *qtab = scratch + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * EDG_QTAB_WORDS;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += T * EDG_BATCH) {
        fe X[EDG_BATCH], Y[EDG_BATCH], Z[EDG_BATCH];
        u32 on_curve = 0;                                      // bit k: public key of the k-th signature decoded to a curve point
        int cnt = 0;
#pragma unroll 1
        for (int k = 0; k < EDG_BATCH; k++) {                 // phase 1: C = S*B + t*(-A), projective
            const size_t i = i0 + (size_t)k * T;
            if (i >= n) break;
            const uint8_t *m; u64 len;
            msg_of(m, len, msgs, off, fixed_len, i);
            ge_p3 R;
            const u32 oc = ed25519_verify_front(R, reinterpret_cast<const u32 *>(sig + 64 * i), reinterpret_cast<const u32 *>(pub + 32 * i),
                                                m, len, qtab, wtab);
            on_curve |= (oc & 1u) << k;
            fe_copy(X[k], R.X); fe_copy(Y[k], R.Y); fe_copy(Z[k], R.Z);
            cnt++;
        }
        fe_batch_inv(Z, cnt);                                  // phase 2: one inversion per batch
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {                        // phase 3: encode and compare with the signature's R bytes
            const size_t i = i0 + (size_t)k * T;
            ok[i] = (uint8_t)ed25519_verify_back(X[k], Y[k], Z[k], (on_curve >> k) & 1u, reinterpret_cast<const u32 *>(sig + 64 * i));
        }
    }
}

// Window table of the base point (built once per device): entry e = e * B, e = 0 .. 2^(EDG_BWIN-1).
__global__ void k_wtab_base(u32 *base, int doublings) { wtab_base(base, doublings); }

__global__ void __launch_bounds__(kThreads) k_wtab_build(u32 *table, const u32 *base) {
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) {
        for (int i = 0; i < 24; i++) table[i] = (i == 0 || i == 8) ? 1u : 0u;      // neutral element (1, 1, 0)
    }
    if (8u * g + 8u < EDG_WTAB_ENTRIES) wtab_build8(table + 24u * (8u * g + 1u), base, 8u * g + 1u);
}

__global__ void __launch_bounds__(kThreads) k_pk_convert(size_t n, uint8_t *out, const uint8_t *in) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u32 p[8], o[8];
        load8(p, in, i);
        pk_ed25519_to_x25519_op(o, p);
        store8(out, i, o);
    }
}

}  // namespace

extern "C" {

size_t edg_verify_scratch_bytes(int sm_count) {
    int bps = 0;
    grid_for(k_verify, (size_t)1 << 40, 0, sm_count, &bps);
    return (size_t)sm_count * bps * kThreads * EDG_QTAB_WORDS * sizeof(u32);
}

size_t edg_verify_table_bytes(void) { return ((size_t)EDG_WTAB_WORDS + 24) * sizeof(u32); }

// table: edg_verify_table_bytes() of device memory; the last 24 words are scratch for the affine base point
int edg_verify_table_init(void *table, void *stream) {
    u32 *t = (u32 *)table, *base = t + EDG_WTAB_WORDS;
    k_wtab_base<<<1, 1, 0, (cudaStream_t)stream>>>(base, 0);
    const unsigned groups = (EDG_WTAB_ENTRIES - 1) / 8;
    k_wtab_build<<<(groups + kThreads - 1) / kThreads, kThreads, 0, (cudaStream_t)stream>>>(t, base);
    return (int)cudaGetLastError();
}

int edg_launch_verify(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                      const unsigned long long *off, unsigned long long fixed_len, void *scratch, const void *table,
                      int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_verify, n, 0, sm_count, nullptr);
    k_verify<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, ok, sig, pub, msgs, off, fixed_len, (u32 *)scratch, (const u32 *)table);
    return (int)cudaGetLastError();
}

int edg_launch_pk_convert(size_t n, uint8_t *out, const uint8_t *in, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_pk_convert, n, 0, sm_count, nullptr);
    k_pk_convert<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, out, in);
    return (int)cudaGetLastError();
}

}  // extern "C"
