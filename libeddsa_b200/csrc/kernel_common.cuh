// Shared helpers of the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ops.cuh"
#include "edg_internal.h"

namespace edg {

#ifndef EDG_THREADS
#define EDG_THREADS 128
#endif
constexpr int kThreads = EDG_THREADS;

__device__ __forceinline__ void load8(u32 w[8], const uint8_t *base, size_t i) {
    load_words8(w, reinterpret_cast<const u32 *>(base + 32 * i));
}

__device__ __forceinline__ void store8(uint8_t *base, size_t i, const u32 w[8]) {
    uint4 *p = reinterpret_cast<uint4 *>(base + 32 * i);
    p[0] = make_uint4(w[0], w[1], w[2], w[3]);
    p[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// cooperative copy of a table from global memory into shared memory (128-bit accesses)
__device__ __forceinline__ void stage_table(u32 *dst, const u32 *src, int words) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (int i = threadIdx.x; i < words / 4; i += blockDim.x) d[i] = s[i];
    for (int i = (words & ~3) + threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ void msg_of(const uint8_t *&m, u64 &len, const uint8_t *msgs, const unsigned long long *off,
                                        unsigned long long fixed_len, size_t i) {
    if (off) { m = msgs + off[i]; len = off[i + 1] - off[i]; }
    else { m = msgs + (size_t)i * fixed_len; len = fixed_len; }
}

// Ragged batches (an offsets array): lanes of a warp hash messages of different lengths and the warp pays for the
// longest.  The message kernels therefore work on tiles of consecutive operations and visit a tile in order of
// message length: keys (length << IDX_BITS | position in tile) are sorted in shared memory (bitonic, all threads of
// the block), so neighbouring lanes get neighbouring lengths.  Lengths are public; results go to their own slots.
template <int N>
__device__ __forceinline__ void block_sort_u32(u32 *key) {
#pragma unroll 1
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const u32 a = key[i], b = key[l];
                    if ((a > b) == ((i & k) == 0)) { key[i] = b; key[l] = a; }
                }
            }
        }
    }
    __syncthreads();
}

// key of operation `base + e` (e < N = 2^IDX_BITS) of a ragged batch; operations past the end sort last
template <int IDX_BITS>
__device__ __forceinline__ u32 ragged_key(const unsigned long long *off, size_t base, int e, size_t n) {
    if (base + e >= n) return 0xffffffffu;
    const unsigned long long len = off[base + e + 1] - off[base + e];
    const u32 cap = (1u << (32 - IDX_BITS)) - 2u;
    return ((len < cap ? (u32)len : cap) << IDX_BITS) | (u32)e;
}

// Overwrite per-thread scratch that held secret-derived values (the GPU analogue of the reference's
// burnstack(), lib/burnstack.c:12-19); volatile so the stores are not optimised away.
__device__ __forceinline__ void scrub(fe *p, int count) {
    volatile u32 *w = reinterpret_cast<volatile u32 *>(p);
    for (int i = 0; i < count * 8; i++) w[i] = 0;
}

// grid = min(blocks needed, SMs x resident blocks per SM): persistent blocks, grid-stride inside.
template <typename K>
int grid_for(K kernel, size_t n, int smem, int sm_count, int *blocks_per_sm_out) {
    int bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, kThreads, smem) != cudaSuccess || bps < 1) bps = 1;
    if (blocks_per_sm_out) *blocks_per_sm_out = bps;
    size_t need = (n + kThreads - 1) / kThreads;
    size_t cap = (size_t)sm_count * bps;
    return (int)(need < cap ? need : cap);
}

}  // namespace edg
