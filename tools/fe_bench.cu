// Throughput of the field multiply / square as compiled for sm_100a: dependent chains per thread,
// all SMs busy.  Reports field mults/s and the implied wide-product (IMAD.WIDE) rate per SM per clock.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../libeddsa_b200/csrc -o fe_bench fe_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fe.cuh"
using namespace edg;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k_chain(uint32_t *out, unsigned long long *cyc, int iters) {
    fe x, y;
    for (int i = 0; i < 10; i++) { x.v[i] = (threadIdx.x * 2654435761u + i * 40503u + blockIdx.x) & 0x1ffffff; y.v[i] = (x.v[i] * 3u + 7u) & 0x1ffffff; }
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { fe_mul(x, x, y); fe_mul(y, y, x); }            // 2 M
        else if (MODE == 1) { fe_sq(x, x); fe_sq(y, y); }               // 2 S
        else {                                                          // ladder-like mix: 5 M + 4 S + adds
            fe a, b, c, d;
            fe_add(a, x, y); fe_sub(b, x, y);
            fe_sq(c, a); fe_sq(d, b);
            fe_mul(x, c, d);
            fe_sub(c, c, d);
            fe_mul121665(a, c); fe_add(a, a, d);
            fe_mul(y, c, a);
            fe_mul(a, x, b); fe_mul(b, y, d);
            fe_add(c, a, b); fe_sub(d, a, b);
            fe_sq(x, c); fe_sq(d, d); fe_mul(y, d, y);
        }
    }
    unsigned long long t1 = clock64();
    uint32_t r = 0;
    for (int i = 0; i < 10; i++) r ^= x.v[i] ^ y.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if ((threadIdx.x & 31) == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int nsm, int bps, int threads, int iters, double mults_per_iter, double prods_per_iter, bool last) {
    int blocks = nsm * bps; size_t nt = (size_t)blocks * threads;
    uint32_t *out; unsigned long long *cyc;
    CK(cudaMalloc(&out, nt * 4)); CK(cudaMalloc(&cyc, nt / 32 * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_chain<MODE><<<blocks, threads>>>(out, cyc, iters); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0)); k_chain<MODE><<<blocks, threads>>>(out, cyc, iters); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    size_t nw = nt / 32; unsigned long long *h = (unsigned long long *)malloc(nw * 8);
    CK(cudaMemcpy(h, cyc, nw * 8, cudaMemcpyDeviceToHost));
    double sum = 0; unsigned long long mx = 0; for (size_t i = 0; i < nw; i++) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
    double avg = sum / nw;
    double fm_per_s = (double)nt * iters * mults_per_iter / (best * 1e-3);
    double prod_clk_sm = (double)bps * threads * iters * prods_per_iter / avg;
    printf("  \"%s\": {\"blocks_per_sm\": %d, \"threads\": %d, \"field_mults_per_s\": %.4e, \"wide_products_per_clk_per_sm\": %.2f, \"ms\": %.3f, \"implied_mhz\": %.0f}%s\n",
           name, bps, threads, fm_per_s, prod_clk_sm, best, (double)mx / (best * 1e-3) / 1e6, last ? "" : ",");
    free(h); CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main(int argc, char **argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    int iters = 2000;
    printf("{\n  \"device\": \"%s\", \"sms\": %d,\n", p.name, nsm);
    int cfg[][2] = {{1, 128}, {1, 256}, {2, 256}, {3, 256}, {4, 256}};
    char nm[64];
    for (int c = 0; c < 5; c++) {
        snprintf(nm, 64, "mul_b%d_t%d", cfg[c][0], cfg[c][1]); run<0>(nm, nsm, cfg[c][0], cfg[c][1], iters, 2, 200, false);
        snprintf(nm, 64, "sq_b%d_t%d", cfg[c][0], cfg[c][1]);  run<1>(nm, nsm, cfg[c][0], cfg[c][1], iters, 2, 110, false);
        snprintf(nm, 64, "ladder_b%d_t%d", cfg[c][0], cfg[c][1]); run<2>(nm, nsm, cfg[c][0], cfg[c][1], iters, 9, 5 * 100 + 4 * 55 + 10, c == 4);
    }
    printf("}\n");
    return 0;
}
