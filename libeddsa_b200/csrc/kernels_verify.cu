// Verification kernel (sm_100a): one signature per thread — SHA-512 challenge, decompression of A,
// Straus double-scalar multiplication S*B + t*(-A) with fixed signed 4-bit windows (uniform control
// flow), encoding and byte comparison.  Public data only, so table lookups are direct-indexed.
// Also hosts pk_ed25519_to_x25519 (it shares the decompression).
// Replaces ed25519_verify / pk_ed25519_to_x25519: /root/reference/lib/ed25519-sha512.c:148-237.
#define EDG_TABLE_QUAL __device__ const
#define EDG_WANT_BASE_SMALL
#include "kernel_common.cuh"
using namespace edg;
#include "base_table.inc"

#ifndef EDG_LB_VERIFY
#define EDG_LB_VERIFY 1     /* min resident blocks per SM the register allocator must allow (tuned, see profiles/) */
#endif
namespace {

__global__ void __launch_bounds__(kThreads, EDG_LB_VERIFY) k_verify(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                                                     const unsigned long long *off, unsigned long long fixed_len, u32 *scratch) {
    __shared__ __align__(16) u32 s_small[EDG_BASE_SMALL_WORDS + 2];
    stage_table(s_small, BASE_SMALL, EDG_BASE_SMALL_WORDS);
    u32 *qtab = scratch + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * EDG_QTAB_WORDS;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *m; u64 len;
        msg_of(m, len, msgs, off, fixed_len, i);
        ok[i] = (uint8_t)ed25519_verify_op(reinterpret_cast<const u32 *>(sig + 64 * i), reinterpret_cast<const u32 *>(pub + 32 * i),
                                           m, len, qtab, s_small);
    }
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

int edg_launch_verify(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                      const unsigned long long *off, unsigned long long fixed_len, void *scratch, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_verify, n, 0, sm_count, nullptr);
    k_verify<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, ok, sig, pub, msgs, off, fixed_len, (u32 *)scratch);
    return (int)cudaGetLastError();
}

int edg_launch_pk_convert(size_t n, uint8_t *out, const uint8_t *in, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_pk_convert, n, 0, sm_count, nullptr);
    k_pk_convert<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, out, in);
    return (int)cudaGetLastError();
}

}  // extern "C"
