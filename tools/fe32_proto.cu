// Prototype: saturated 8 x 32-bit limbs for GF(2^255-19) with IMAD.WIDE.U32(.X) carry chains.
// Checks results against host __int128 arithmetic and measures dependent-chain throughput.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fe32_proto fe32_proto.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// chain of CNT products a[0],a[2],.. (stride 2) times bi added into acc[0..2CNT) with carry; the carry out goes into acc[2CNT].
// ONE asm statement per chain: the compiler must not interleave anything that touches the carry flag.
template <int CNT> __device__ __forceinline__ void cmad(u32 *acc, const u32 *a, u32 bi);
template <> __device__ __forceinline__ void cmad<0>(u32 *, const u32 *, u32) {}
template <> __device__ __forceinline__ void cmad<1>(u32 *acc, const u32 *a, u32 bi) {
  asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(a[0]), "r"(bi));
}
template <> __device__ __forceinline__ void cmad<2>(u32 *acc, const u32 *a, u32 bi) {
  asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %7, %2; madc.hi.cc.u32 %3, %6, %7, %3; addc.u32 %4, %4, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]) : "r"(a[0]), "r"(a[2]), "r"(bi));
}
template <> __device__ __forceinline__ void cmad<3>(u32 *acc, const u32 *a, u32 bi) {
  asm("mad.lo.cc.u32 %0, %7, %10, %0; madc.hi.cc.u32 %1, %7, %10, %1; madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3; "
      "madc.lo.cc.u32 %4, %9, %10, %4; madc.hi.cc.u32 %5, %9, %10, %5; addc.u32 %6, %6, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]) : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(bi));
}
template <> __device__ __forceinline__ void cmad<4>(u32 *acc, const u32 *a, u32 bi) {
  asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1; madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3; "
      "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5; madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7; addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(bi));
}
// w[1..15] += od[0..14]  (one carry chain, one asm statement)
__device__ __forceinline__ void merge_odd(u32 *w, const u32 *od) {
  asm("add.cc.u32 %0, %0, %15; addc.cc.u32 %1, %1, %16; addc.cc.u32 %2, %2, %17; addc.cc.u32 %3, %3, %18; addc.cc.u32 %4, %4, %19; "
      "addc.cc.u32 %5, %5, %20; addc.cc.u32 %6, %6, %21; addc.cc.u32 %7, %7, %22; addc.cc.u32 %8, %8, %23; addc.cc.u32 %9, %9, %24; "
      "addc.cc.u32 %10, %10, %25; addc.cc.u32 %11, %11, %26; addc.cc.u32 %12, %12, %27; addc.cc.u32 %13, %13, %28; addc.u32 %14, %14, %29;"
      : "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]), "+r"(w[11]),
        "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15])
      : "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]),
        "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
}

// r = lo + 38 * hi  (mod 2^256 - 38), weakly reduced (< 2^256)
__device__ __forceinline__ void fold512(u32 r[8], const u32 w[16]) {
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) { u64 t = (u64)w[8 + j] * 38u + w[j] + c; r[j] = (u32)t; c = t >> 32; }
  u32 cc = (u32)c * 38u, last;
  asm("add.cc.u32 %0, %0, %9; addc.cc.u32 %1, %1, 0; addc.cc.u32 %2, %2, 0; addc.cc.u32 %3, %3, 0; addc.cc.u32 %4, %4, 0; "
      "addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0; addc.u32 %8, 0, 0;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=r"(last) : "r"(cc));
  r[0] += last * 38u;
}

__device__ __forceinline__ void fe_mul32(u32 r[8], const u32 a[8], const u32 b[8]) {
  u32 ev[17], od[17];
#pragma unroll
  for (int i = 0; i < 17; i++) ev[i] = od[i] = 0;
  // product a_j b_i sits at word i+j: even positions accumulate in ev[i+j], odd positions in od[i+j-1]
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if ((i & 1) == 0) { cmad<4>(ev + i, a, b[i]); cmad<4>(od + i, a + 1, b[i]); }
    else              { cmad<4>(ev + i + 1, a + 1, b[i]); cmad<4>(od + i - 1, a, b[i]); }
  }
  merge_odd(ev, od);
  fold512(r, ev);
}

__device__ __forceinline__ void fe_sq32(u32 r[8], const u32 a[8]) {
  u32 ev[17], od[17];
#pragma unroll
  for (int i = 0; i < 17; i++) ev[i] = od[i] = 0;
  // off-diagonal products a_j a_i, j > i
  // row i, j = i+1, i+3, ..: word i+j odd  -> od[i+j-1] = od[2i + 0, 2, ...]   ; count = ceil((7-i)/2)
  // row i, j = i+2, i+4, ..: word i+j even -> ev[2i+2, ...]                    ; count = floor((7-i)/2)
#define ROW(i) cmad<(8 - (i)) / 2>(od + 2 * (i), a + (i) + 1, a[i]); cmad<(7 - (i)) / 2>(ev + 2 * (i) + 2, a + (i) + 2, a[i]);
  ROW(0) ROW(1) ROW(2) ROW(3) ROW(4) ROW(5) ROW(6)
#undef ROW
  merge_odd(ev, od);
  // double
#pragma unroll
  for (int k = 15; k > 0; k--) ev[k] = (ev[k] << 1) | (ev[k - 1] >> 31);
  ev[0] = 0;   // word 0 has no off-diagonal contribution
  // add diagonal squares a_i^2 at word 2i
  asm("mad.lo.cc.u32 %0, %16, %16, %0; madc.hi.cc.u32 %1, %16, %16, %1; madc.lo.cc.u32 %2, %17, %17, %2; madc.hi.cc.u32 %3, %17, %17, %3; "
      "madc.lo.cc.u32 %4, %18, %18, %4; madc.hi.cc.u32 %5, %18, %18, %5; madc.lo.cc.u32 %6, %19, %19, %6; madc.hi.cc.u32 %7, %19, %19, %7; "
      "madc.lo.cc.u32 %8, %20, %20, %8; madc.hi.cc.u32 %9, %20, %20, %9; madc.lo.cc.u32 %10, %21, %21, %10; madc.hi.cc.u32 %11, %21, %21, %11; "
      "madc.lo.cc.u32 %12, %22, %22, %12; madc.hi.cc.u32 %13, %22, %22, %13; madc.lo.cc.u32 %14, %23, %23, %14; madc.hi.u32 %15, %23, %23, %15;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]), "+r"(ev[8]), "+r"(ev[9]),
        "+r"(ev[10]), "+r"(ev[11]), "+r"(ev[12]), "+r"(ev[13]), "+r"(ev[14]), "+r"(ev[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
  fold512(r, ev);
}

template <int MODE> __global__ void __launch_bounds__(256) k_chain(u32 *out, int iters) {
  u32 a[8], b[8];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x; b[i] = a[i] * 3u + 7u; }
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) { fe_mul32(a, a, b); fe_mul32(b, b, a); }
    else { fe_sq32(a, a); fe_sq32(b, b); }
  }
  u32 x = 0; for (int i = 0; i < 8; i++) x ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_check(u32 *r_mul, u32 *r_sq, const u32 *a, const u32 *b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  u32 x[8], y[8], z[8];
  for (int k = 0; k < 8; k++) { x[k] = a[8 * i + k]; y[k] = b[8 * i + k]; }
  fe_mul32(z, x, y); for (int k = 0; k < 8; k++) r_mul[8 * i + k] = z[k];
  fe_sq32(z, x); for (int k = 0; k < 8; k++) r_sq[8 * i + k] = z[k];
}
// host reference: (a*b) mod p compared modulo p via simple bignum
typedef unsigned __int128 u128;
static void mulmod_host(u32 r[8], const u32 a[8], const u32 b[8]) {
  u64 w[17] = {0};
  for (int i = 0; i < 8; i++) { u64 c = 0; for (int j = 0; j < 8; j++) { u64 t = (u64)a[i] * b[j] + w[i + j] + c; w[i + j] = (u32)t; c = t >> 32; } w[i + 8] = c; }
  // fold to < 2^256 then canonical mod p
  u64 c = 0; u32 t8[8];
  for (int j = 0; j < 8; j++) { u64 t = w[8 + j] * 38 + w[j] + c; t8[j] = (u32)t; c = t >> 32; }
  while (c) { u64 cc = c * 38; c = 0; for (int j = 0; j < 8; j++) { u64 t = (u64)t8[j] + (j == 0 ? cc : 0) + c; t8[j] = (u32)t; c = t >> 32; if (j == 0) cc = 0; } }
  memcpy(r, t8, 32);
}
static void canon_host(u32 v[8]) {   // v < 2^256 -> [0,p)
  for (int round = 0; round < 2; round++) {
    u32 top = v[7] >> 31; v[7] &= 0x7fffffff; u64 c = 19ull * top;
    for (int j = 0; j < 8; j++) { u64 t = (u64)v[j] + c; v[j] = (u32)t; c = t >> 32; }
  }
  u32 t[8]; u64 c = 19; for (int j = 0; j < 8; j++) { u64 s = (u64)v[j] + c; t[j] = (u32)s; c = s >> 32; }
  if (t[7] >> 31) { t[7] &= 0x7fffffff; memcpy(v, t, 32); }
}
int main() {
  int n = 1 << 16;
  u32 *ha = (u32 *)malloc(n * 32), *hb = (u32 *)malloc(n * 32), *hm = (u32 *)malloc(n * 32), *hs = (u32 *)malloc(n * 32);
  srand(1);
  for (int i = 0; i < n * 8; i++) { ha[i] = ((u32)rand() << 16) ^ rand(); hb[i] = ((u32)rand() << 16) ^ rand(); }
  for (int k = 0; k < 8; k++) { ha[k] = 0xffffffff; hb[k] = 0xffffffff; ha[8 + k] = 0; hb[8 + k] = 0xffffffff; ha[16 + k] = k == 0 ? 0xffffffed : (k == 7 ? 0x7fffffff : 0xffffffff); hb[16 + k] = ha[16 + k]; }
  u32 *da, *db, *dm, *ds; CK(cudaMalloc(&da, n * 32)); CK(cudaMalloc(&db, n * 32)); CK(cudaMalloc(&dm, n * 32)); CK(cudaMalloc(&ds, n * 32));
  CK(cudaMemcpy(da, ha, n * 32, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb, n * 32, cudaMemcpyHostToDevice));
  k_check<<<n / 256, 256>>>(dm, ds, da, db, n); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hm, dm, n * 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hs, ds, n * 32, cudaMemcpyDeviceToHost));
  int bad_m = 0, bad_s = 0;
  for (int i = 0; i < n; i++) {
    u32 e[8], g[8];
    mulmod_host(e, ha + 8 * i, hb + 8 * i); canon_host(e); memcpy(g, hm + 8 * i, 32); canon_host(g); bad_m += memcmp(e, g, 32) != 0;
    mulmod_host(e, ha + 8 * i, ha + 8 * i); canon_host(e); memcpy(g, hs + 8 * i, 32); canon_host(g); bad_s += memcmp(e, g, 32) != 0;
  }
  printf("{\"check_n\": %d, \"mul_mismatch\": %d, \"sq_mismatch\": %d,\n", n, bad_m, bad_s);
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int nsm = p.multiProcessorCount, iters = 2000;
  u32 *out; CK(cudaMalloc(&out, (size_t)nsm * 8 * 256 * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int bps = 1; bps <= 8; bps *= 2) for (int mode = 0; mode < 2; mode++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      CK(cudaEventRecord(e0));
      if (mode == 0) k_chain<0><<<nsm * bps, 256>>>(out, iters); else k_chain<1><<<nsm * bps, 256>>>(out, iters);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
    }
    printf(" \"%s_b%d\": %.4e,\n", mode ? "sq" : "mul", bps, (double)nsm * bps * 256 * iters * 2 / (best * 1e-3));
  }
  printf(" \"unit\": \"field ops/s\"}\n");
  return 0;
}
