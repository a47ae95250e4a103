// Integer-pipe microbenchmark for B200 (sm_100a).
// Measures sustained per-SM-per-clock throughput of the instructions the GF(2^255-19)
// limb arithmetic is built from, so the IMAD roofline denominator is a measured number
// (SURVEY.md §7 step 2 / §8(d)).  Output: one JSON object on stdout.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int NACC = 8;      // independent dependency chains per thread
constexpr int UNROLL = 8;    // instructions per chain per loop trip

struct Res { unsigned long long cyc; };

#define KERNEL_HEAD(name) KERNEL_HEAD_(name, false)
#define KERNEL_HEADP(name) KERNEL_HEAD_(name, true)
#define KERNEL_HEAD_(name, PERTURB) \
__global__ void __launch_bounds__(512) name(uint32_t *out, unsigned long long *cyc, uint32_t seed) { \
    uint32_t a[NACC], b[NACC]; uint64_t w[NACC]; double d[NACC]; float f[NACC]; \
    _Pragma("unroll") for (int i = 0; i < NACC; i++) { a[i] = seed * (threadIdx.x + 1 + i); b[i] = a[i] ^ 0x9e3779b9u; \
        w[i] = ((uint64_t)a[i] << 32) | b[i]; d[i] = (double)a[i]; f[i] = (float)b[i]; } \
    uint32_t m = seed | 1u; double dm = 1.0000001; float fm = 1.0001f; (void)dm; (void)fm; (void)m; \
    __syncthreads(); unsigned long long t0 = clock64(); \
    for (int it = 0; it < ITERS; it++) { if (PERTURB) { _Pragma("unroll") for (int i = 0; i < NACC; i++) { a[i] += it; b[i] ^= a[i]; } } _Pragma("unroll") for (int u = 0; u < UNROLL; u++) { _Pragma("unroll") for (int i = 0; i < NACC; i++) {

#define KERNEL_TAIL \
    } } } \
    unsigned long long t1 = clock64(); \
    uint32_t r = 0; _Pragma("unroll") for (int i = 0; i < NACC; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)d[i] ^ (uint32_t)f[i]; \
    out[blockIdx.x * blockDim.x + threadIdx.x] = r; \
    if ((threadIdx.x & 31) == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0; }

KERNEL_HEAD(k_imad_lo)
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
KERNEL_TAIL

KERNEL_HEADP(k_imad_wide)
    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[(i + u) % NACC]), "r"(b[u]));
KERNEL_TAIL

KERNEL_HEADP(k_imad_wide_s)
    asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[(i + u) % NACC]), "r"(b[u]));
KERNEL_TAIL

KERNEL_HEAD(k_imad_hi)
    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
KERNEL_TAIL

KERNEL_HEADP(k_mul_wide)
    { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[(i + u) % NACC]), "r"(b[u])); w[i] ^= t; }
KERNEL_TAIL

// carry-chained lo/hi pair: what the saturated 8x32 representation needs
KERNEL_HEAD(k_madc_pair)
    asm volatile("mad.lo.cc.u32 %0, %0, %2, %1;\n\tmadc.hi.u32 %1, %0, %2, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(m));
KERNEL_TAIL

KERNEL_HEAD(k_iadd3)
    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) % NACC]));
KERNEL_TAIL

KERNEL_HEAD(k_lop3)
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m));
KERNEL_TAIL

KERNEL_HEAD(k_shf)
    asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
KERNEL_TAIL

KERNEL_HEAD(k_add64)
    asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % NACC]));
KERNEL_TAIL

KERNEL_HEAD(k_shr64)
    { asm volatile("shr.u64 %0, %1, 3;" : "=l"(w[i]) : "l"(w[i] | 0x8000000000000000ull)); }
KERNEL_TAIL

KERNEL_HEAD(k_dfma)
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(dm));
KERNEL_TAIL

KERNEL_HEAD(k_ffma)
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fm), "f"(fm));
KERNEL_TAIL

// co-issue tests (2 instrs per slot; reported rate counts BOTH)
KERNEL_HEADP(k_wide_plus_iadd)
    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[(i + u) % NACC]), "r"(b[u]));
    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) % NACC]));
KERNEL_TAIL

KERNEL_HEADP(k_wide_plus_2alu)
    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[(i + u) % NACC]), "r"(b[u]));
    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) % NACC]));
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(seed), "r"(m));
KERNEL_TAIL

KERNEL_HEADP(k_wide_plus_dfma)
    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[(i + u) % NACC]), "r"(b[u]));
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(dm));
KERNEL_TAIL

KERNEL_HEAD(k_lo_plus_iadd)
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(seed));
    asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(b[(i + 1) % NACC]));
KERNEL_TAIL

KERNEL_HEADP(k_wide_plus_ffma)
    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[(i + u) % NACC]), "r"(b[u]));
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fm), "f"(fm));
KERNEL_TAIL

typedef void (*kern_t)(uint32_t *, unsigned long long *, uint32_t);

static void run(const char *name, kern_t k, int per_slot, int nsm, int blocks_per_sm, int threads, bool last) {
    int blocks = nsm * blocks_per_sm;
    size_t nthreads = (size_t)blocks * threads;
    uint32_t *out; unsigned long long *cyc;
    CK(cudaMalloc(&out, nthreads * 4));
    CK(cudaMalloc(&cyc, nthreads / 32 * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<blocks, threads>>>(out, cyc, 12345u);   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k<<<blocks, threads>>>(out, cyc, 12345u + rep);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    size_t nw = nthreads / 32;
    unsigned long long *h = (unsigned long long *)malloc(nw * 8);
    CK(cudaMemcpy(h, cyc, nw * 8, cudaMemcpyDeviceToHost));
    double sum = 0; unsigned long long mx = 0;
    for (size_t i = 0; i < nw; i++) { sum += (double)h[i]; if (h[i] > mx) mx = h[i]; }
    double avg_cyc = sum / nw;
    double instr_per_thread = (double)ITERS * UNROLL * NACC * per_slot;
    // per SM: warps_per_sm * 32 lanes * instr / cycles
    double lanes_per_clk_sm = (double)blocks_per_sm * threads * instr_per_thread / avg_cyc;
    double total_per_s = (double)nthreads * instr_per_thread / (best * 1e-3);
    printf("  \"%s\": {\"thread_instr_per_clk_per_sm\": %.2f, \"total_thread_instr_per_s\": %.4e, \"ms\": %.4f, \"avg_cycles\": %.0f, \"max_cycles\": %llu, \"implied_mhz\": %.0f}%s\n",
           name, lanes_per_clk_sm, total_per_s, best, avg_cyc, mx, (double)mx / (best * 1e-3) / 1e6, last ? "" : ",");
    free(h); CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main(int argc, char **argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    int bps = argc > 1 ? atoi(argv[1]) : 2;
    int thr = argc > 2 ? atoi(argv[2]) : 512;
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"blocks_per_sm\": %d, \"threads\": %d,\n", p.name, nsm, bps, thr);
#define R(k, n) run(#k, k, n, nsm, bps, thr, false)
    R(k_imad_lo, 1); R(k_imad_wide, 1); R(k_imad_wide_s, 1); R(k_imad_hi, 1); R(k_mul_wide, 1);
    R(k_madc_pair, 2); R(k_iadd3, 1); R(k_lop3, 1); R(k_shf, 1); R(k_add64, 1); R(k_shr64, 1);
    R(k_dfma, 1); R(k_ffma, 1);
    R(k_wide_plus_iadd, 2); R(k_wide_plus_2alu, 3); R(k_wide_plus_dfma, 2); R(k_lo_plus_iadd, 2);
    run("k_wide_plus_ffma", k_wide_plus_ffma, 2, nsm, bps, thr, true);
    printf("}\n");
    return 0;
}
