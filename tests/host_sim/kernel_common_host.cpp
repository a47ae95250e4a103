// TEST INFRASTRUCTURE ONLY: the block-level helpers of libeddsa_b200/csrc/kernel_common.cuh (the length-ordering of ragged
// batches: ragged_key + block_sort_u32, exactly the code the message kernels run) executed on the host as a block of ONE
// thread under ptx_emul.h.  With one thread every `for (i = threadIdx.x; i < N; i += blockDim.x)` loop visits all i and
// __syncthreads() has nothing to wait for; a phase of the bitonic network touches every pair from exactly one i, so the
// sequential order gives the same result as the parallel one.
static inline void __syncthreads() {}
#include "../../libeddsa_b200/csrc/kernel_common.cuh"
using namespace edg;
extern "C" {
// keys of tile [base, base + 2^bits) of a ragged batch of n operations, sorted as the kernels sort them
void hs_tile_order(uint32_t *keys, const unsigned long long *off, uint64_t base, uint64_t n, int bits) {
    const int N = 1 << bits;
    for (int e = 0; e < N; e++)
        keys[e] = bits == 9 ? ragged_key<9>(off, base, e, n) : bits == 10 ? ragged_key<10>(off, base, e, n) : ragged_key<11>(off, base, e, n);
    if (bits == 9) block_sort_u32<512>(keys); else if (bits == 10) block_sort_u32<1024>(keys); else block_sort_u32<2048>(keys);
}
}
