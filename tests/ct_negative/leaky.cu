// NEGATIVE CONTROL for tools/ct_audit.py (test infrastructure only, never linked into the product):
// two kernels that are deliberately NOT constant time.  The audit must flag both.
#include <stdint.h>
extern "C" {
// secret-dependent table index (cache-timing leak)
__global__ void k_leaky_index(size_t n, uint32_t *out, const uint8_t *sec, const uint32_t *table) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = table[sec[i] & 15];
}
// secret-dependent branch (timing leak): square-and-multiply with a conditional multiply
__global__ void k_leaky_branch(size_t n, uint32_t *out, const uint8_t *sec, const uint32_t *table) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t acc = 1, e = sec[i];
    for (int b = 0; b < 8; b++) {
        acc = acc * acc;
        if ((e >> b) & 1) {
            for (int k = 0; k < 50; k++) acc = acc * table[k] + 12345u;   // heavy work only on 1-bits
        }
    }
    out[i] = acc;
}
// "masked select" written naively: nvcc turns the mask idiom into loads PREDICATED on the secret
// (@!P LDG) — which is why the audit runs on the SHIPPED SASS (not on the source) and treats a secret guard on a memory instruction as a violation.
__global__ void k_naive_select(size_t n, uint32_t *out, const uint8_t *sec, const uint32_t *table) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t idx = sec[i] & 15, r = 0;
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t m = 0u - ((((idx ^ k) - 1u) >> 31));
        r ^= (r ^ table[k]) & m;
    }
    out[i] = r;
}
// constant-time counterpart (must PASS): same loop with the mask hidden behind an optimisation barrier
__global__ void k_clean_select(size_t n, uint32_t *out, const uint8_t *sec, const uint32_t *table) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t idx = sec[i] & 15, r = 0;
    for (uint32_t k = 0; k < 16; k++) {
        uint32_t t = table[k];
        uint32_t m = 0u - ((((idx ^ k) - 1u) >> 31));
        asm volatile("" : "+r"(m));                      // the product's ct_mask(): the compiler no longer knows m is 0 / ~0
        r ^= (r ^ t) & m;
    }
    out[i] = r;
}
// a secret handed to ANOTHER lane through shared memory and used as a table index there: the audit lets secrets be
// stored to shared memory (the tensor-core lookup does) but must treat every later shared load as secret
__global__ void k_leaky_smem(size_t n, uint32_t *out, const uint8_t *sec, const uint32_t *table) {
    __shared__ uint32_t s[128];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    s[threadIdx.x & 127] = i < n ? sec[i] : 0;
    __syncthreads();
    if (i < n) out[i] = table[s[(threadIdx.x + 1) & 127] & 15];
}
// a secret-dependent source lane of a shuffle
__global__ void k_leaky_shfl(size_t n, uint32_t *out, const uint8_t *sec, const uint32_t *table) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    uint32_t v = table[threadIdx.x & 15], e = i < n ? sec[i] : 0;
    v = __shfl_sync(0xffffffffu, v, (int)(e & 31));
    if (i < n) out[i] = v;
}
}
