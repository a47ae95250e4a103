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
