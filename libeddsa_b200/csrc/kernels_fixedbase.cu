// Fixed-base kernels (sm_100a): ed25519_genpub, ed25519_sign, x25519_base and sk_ed25519_to_x25519.
// One operation per thread; the 49 152-byte signed radix-16 comb table of B is staged once per
// persistent block in shared memory and scanned with masks (constant time: no branch or address
// depends on a secret).  Replaces ed25519-sha512.c:53-137, 243-256 and x25519.c:158-208.
#define EDG_TABLE_QUAL __device__ const
#define EDG_WANT_BASE_COMB
#include "kernel_common.cuh"
using namespace edg;
#include "base_table.inc"

#ifndef EDG_LB_X25519_BASE
#define EDG_LB_X25519_BASE 1     /* min resident blocks per SM the register allocator must allow (tuned, see profiles/) */
#endif
#ifndef EDG_LB_GENPUB
#define EDG_LB_GENPUB 1     /* min resident blocks per SM the register allocator must allow (tuned, see profiles/) */
#endif
#ifndef EDG_LB_SIGN
#define EDG_LB_SIGN 1     /* min resident blocks per SM the register allocator must allow (tuned, see profiles/) */
#endif
namespace {

constexpr int kCombBytes = EDG_BASE_COMB_WORDS * 4;      // 49 152

__global__ void __launch_bounds__(kThreads, EDG_LB_X25519_BASE) k_x25519_base(size_t n, uint8_t *out, const uint8_t *scalar) {
    extern __shared__ __align__(16) u32 s_comb[];
    stage_table(s_comb, BASE_COMB, EDG_BASE_COMB_WORDS);
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += T * EDG_BATCH) {
        fe N[EDG_BATCH], D[EDG_BATCH];
        int cnt = 0;
#pragma unroll 1
        for (int k = 0; k < EDG_BATCH; k++) {
            const size_t i = i0 + (size_t)k * T;
            if (i >= n) break;
            u32 s[8];
            load8(s, scalar, i);
            x25519_base_front(N[k], D[k], s, s_comb);
            cnt++;
        }
        fe_batch_inv(D, cnt);
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            u32 o[8];
            x25519_back(o, N[k], D[k]);
            store8(out, i0 + (size_t)k * T, o);
        }
        scrub(N, EDG_BATCH);
        scrub(D, EDG_BATCH);
    }
}

__global__ void __launch_bounds__(kThreads, EDG_LB_GENPUB) k_genpub(size_t n, uint8_t *pub, const uint8_t *sec) {
    extern __shared__ __align__(16) u32 s_comb[];
    stage_table(s_comb, BASE_COMB, EDG_BASE_COMB_WORDS);
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += T * EDG_BATCH) {
        fe X[EDG_BATCH], Y[EDG_BATCH], Z[EDG_BATCH];
        int cnt = 0;
#pragma unroll 1
        for (int k = 0; k < EDG_BATCH; k++) {
            const size_t i = i0 + (size_t)k * T;
            if (i >= n) break;
            ge_p3 A;
            ed25519_genpub_front(A, sec + 32 * i, s_comb);
            fe_copy(X[k], A.X); fe_copy(Y[k], A.Y); fe_copy(Z[k], A.Z);
            cnt++;
        }
        fe_batch_inv(Z, cnt);
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            u32 o[8];
            ge_tobytes_zinv(o, X[k], Y[k], Z[k]);
            store8(pub, i0 + (size_t)k * T, o);
        }
        scrub(X, EDG_BATCH);
        scrub(Y, EDG_BATCH);
        scrub(Z, EDG_BATCH);
    }
}

// RAGGED = the batch has an offsets array: tiles of kThreads x EDG_BATCH consecutive signatures, visited in order of
// message length (kernel_common.cuh: block_sort_u32); otherwise the grid-stride mapping of the other kernels.
template <bool RAGGED>
__global__ void __launch_bounds__(kThreads, EDG_LB_SIGN) k_sign(size_t n, uint8_t *sig, const uint8_t *sec, const uint8_t *pub, const uint8_t *msgs,
                                                   const unsigned long long *off, unsigned long long fixed_len) {
    extern __shared__ __align__(16) u32 s_comb[];
    constexpr int TILE = kThreads * EDG_BATCH;
    constexpr int kTileBits = TILE == 512 ? 9 : TILE == 1024 ? 10 : TILE == 2048 ? 11 : 12;
    static_assert(TILE == (1 << kTileBits), "tile size must be a power of two");
    __shared__ u32 s_key[RAGGED ? TILE : 1];
    stage_table(s_comb, BASE_COMB, EDG_BASE_COMB_WORDS);
    const size_t T = (size_t)gridDim.x * blockDim.x;
    // RAGGED: full rounds of one TILE per block, then the remainder split evenly over the blocks (smaller tiles), so
    // that the last round costs every block the same; otherwise one grid-stride round per EDG_BATCH x T signatures
    const size_t full = RAGGED ? n / ((size_t)gridDim.x * TILE) : 0;
    const size_t rem0 = full * gridDim.x * TILE;
    const size_t share = RAGGED ? ((n - rem0 + gridDim.x - 1) / gridDim.x + 31) / 32 * 32 : 0;     // <= TILE
    const size_t rounds = RAGGED ? full + (n > rem0 ? 1 : 0) : 0;
    size_t round = 0, i0 = RAGGED ? 0 : (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; RAGGED ? round < rounds : i0 < n; round++, i0 += T * EDG_BATCH) {
        if (RAGGED) {
            size_t lim;                                        // tile [i0, lim)
            if (round < full) { i0 = (round * gridDim.x + blockIdx.x) * TILE; lim = i0 + TILE; }
            else { i0 = rem0 + blockIdx.x * share; lim = i0 + share < n ? i0 + share : n; }
            __syncthreads();                                   // the previous tile's keys are no longer needed
            for (int e = threadIdx.x; e < TILE; e += blockDim.x) s_key[e] = i0 < lim ? ragged_key<kTileBits>(off, i0, e, lim) : 0xffffffffu;
            block_sort_u32<TILE>(s_key);
        }
        auto op_index = [&](int k) -> size_t {
            if (!RAGGED) return i0 + (size_t)k * T;
            const u32 key = s_key[threadIdx.x + kThreads * k];
            return key == 0xffffffffu ? n : i0 + (key & (TILE - 1));
        };
        fe X[EDG_BATCH], Y[EDG_BATCH], Z[EDG_BATCH], AR[2 * EDG_BATCH];     // AR: secret scalar a and nonce r of each signature
        int cnt = 0;
#pragma unroll 1
        for (int k = 0; k < EDG_BATCH; k++) {
            const size_t i = op_index(k);
            if (i >= n) break;
            const uint8_t *m; u64 len;
            msg_of(m, len, msgs, off, fixed_len, i);
            ge_p3 R;
            ed25519_sign_front(AR[2 * k].v, AR[2 * k + 1].v, R, sec + 32 * i, m, len, s_comb);
            fe_copy(X[k], R.X); fe_copy(Y[k], R.Y); fe_copy(Z[k], R.Z);
            cnt++;
        }
        if (RAGGED && cnt == 0) continue;                      // (public: no work for this thread in the last round)
        fe_batch_inv(Z, cnt);
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            const size_t i = op_index(k);
            u32 p[8], o[16];
            const uint8_t *m; u64 len;
            msg_of(m, len, msgs, off, fixed_len, i);
            load8(p, pub, i);
            ed25519_sign_back(o, AR[2 * k].v, AR[2 * k + 1].v, X[k], Y[k], Z[k], p, m, len);
            store8(sig, 2 * i, o);
            store8(sig, 2 * i + 1, o + 8);
        }
        scrub(X, EDG_BATCH);
        scrub(Y, EDG_BATCH);
        scrub(Z, EDG_BATCH);
        scrub(AR, 2 * EDG_BATCH);
    }
}

__global__ void __launch_bounds__(kThreads) k_sk_convert(size_t n, uint8_t *out, const uint8_t *in) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u32 o[8];
        sk_ed25519_to_x25519_op(o, in + 32 * i);
        store8(out, i, o);
    }
}

}  // namespace

extern "C" {

int edg_fixedbase_init(void) {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_x25519_base, cudaFuncAttributeMaxDynamicSharedMemorySize, kCombBytes); if (e) return (int)e;
    e = cudaFuncSetAttribute(k_genpub, cudaFuncAttributeMaxDynamicSharedMemorySize, kCombBytes); if (e) return (int)e;
    e = cudaFuncSetAttribute(k_sign<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCombBytes); if (e) return (int)e;
    e = cudaFuncSetAttribute(k_sign<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCombBytes); if (e) return (int)e;
    return 0;
}

int edg_launch_x25519_base(size_t n, uint8_t *out, const uint8_t *scalar, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_x25519_base, n, kCombBytes, sm_count, nullptr);
    k_x25519_base<<<g, kThreads, kCombBytes, (cudaStream_t)stream>>>(n, out, scalar);
    return (int)cudaGetLastError();
}

int edg_launch_genpub(size_t n, uint8_t *pub, const uint8_t *sec, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_genpub, n, kCombBytes, sm_count, nullptr);
    k_genpub<<<g, kThreads, kCombBytes, (cudaStream_t)stream>>>(n, pub, sec);
    return (int)cudaGetLastError();
}

int edg_launch_sign(size_t n, uint8_t *sig, const uint8_t *sec, const uint8_t *pub, const uint8_t *msgs,
                    const unsigned long long *off, unsigned long long fixed_len, int sm_count, void *stream) {
    if (n == 0) return 0;
    if (off) {                                                 // ragged: tiles of kThreads x EDG_BATCH signatures per block
        int bps = 0;
        grid_for(k_sign<true>, n, kCombBytes, sm_count, &bps);
        const size_t tiles = (n + (size_t)kThreads * EDG_BATCH - 1) / ((size_t)kThreads * EDG_BATCH), cap = (size_t)sm_count * bps;
        k_sign<true><<<(unsigned)(tiles < cap ? tiles : cap), kThreads, kCombBytes, (cudaStream_t)stream>>>(n, sig, sec, pub, msgs, off, fixed_len);
    } else {
        int g = grid_for(k_sign<false>, n, kCombBytes, sm_count, nullptr);
        k_sign<false><<<g, kThreads, kCombBytes, (cudaStream_t)stream>>>(n, sig, sec, pub, msgs, off, fixed_len);
    }
    return (int)cudaGetLastError();
}

int edg_launch_sk_convert(size_t n, uint8_t *out, const uint8_t *in, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_sk_convert, n, 0, sm_count, nullptr);
    k_sk_convert<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, out, in);
    return (int)cudaGetLastError();
}

}  // extern "C"
