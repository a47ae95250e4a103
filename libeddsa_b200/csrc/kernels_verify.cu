// Verification kernels (sm_100a), one signature per thread, in two stages (ops.cuh: ed25519_verify_front / _loop):
//   k_verify_front  SHA-512 challenge, half-gcd (hgcd.cuh), decompression of A and R, the two 8-entry tables
//   k_verify        Straus multi-scalar multiplication over ~33 signed 4-bit windows (uniform control flow) and the
//                   projective comparison with the neutral element
// Public data only, so table lookups are direct-indexed.
// Also hosts pk_ed25519_to_x25519 (it shares the decompression).
// Replaces ed25519_verify / pk_ed25519_to_x25519: /root/reference/lib/ed25519-sha512.c:148-237.
#define EDG_SHA_MSGWORD_NOINLINE
#include "kernel_common.cuh"
using namespace edg;

#ifndef EDG_VERIFY_WAVES
#define EDG_VERIFY_WAVES 4  /* waves of resident threads per pass: sizes the per-signature records in scratch */
#endif
#ifndef EDG_LB_VERIFY
#define EDG_LB_VERIFY 4     /* front kernel: min resident blocks per SM the register allocator must allow: 128 registers, 4 warps
                               per scheduler */
#endif
#ifndef EDG_VTHREADS
#define EDG_VTHREADS EDG_THREADS   /* block size of the window-loop kernel (64 .. 256 threads and 2 .. 8 resident blocks all land
                                      within 2 % of each other, profiles/r01_pipe_microbench.md) */
#endif
#ifndef EDG_LB_VLOOP
#define EDG_LB_VLOOP EDG_LB_VERIFY
#endif
namespace {
constexpr int kVThreads = EDG_VTHREADS;

// stage 1: one signature per thread -> one EDG_VSTATE_WORDS-word record.  Records are handed out sorted by window
// count: signatures needing more than EDG_NWIN_SPLIT windows (5 %) fill the record array from the front, the others
// from the back, so a warp of the loop kernel (which runs the maximum over its lanes) almost never waits for a
// single long lane: 33.9 -> 33.05 windows on average.  Slots come from two counters (one warp-aggregated atomic each).
#define EDG_NWIN_SPLIT 33
__global__ void __launch_bounds__(kThreads, EDG_LB_VERIFY) k_verify_front(size_t n, size_t first, const uint8_t *sig, const uint8_t *pub,
                                                     const uint8_t *msgs, const unsigned long long *off, unsigned long long fixed_len,
                                                     u32 *state, unsigned int *counters, int full_scalars) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((k & ~(size_t)31) >= n) return;                    // whole warps stay (full-mask votes below)
    const bool live = k < n;
    const size_t i = first + (live ? k : n - 1);
    const uint8_t *m; u64 len;
    msg_of(m, len, msgs, off, fixed_len, i);
    const u32 *sg = reinterpret_cast<const u32 *>(sig + 64 * i), *pk = reinterpret_cast<const u32 *>(pub + 32 * i);
    verify_scalars v;
    const int nwin = ed25519_verify_front_scalars(v, sg, pk, m, len, full_scalars != 0);
    __syncwarp();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lo = __ballot_sync(0xffffffffu, live && nwin <= EDG_NWIN_SPLIT);
    const unsigned hi = __ballot_sync(0xffffffffu, live && nwin > EDG_NWIN_SPLIT);
    unsigned base_lo = 0, base_hi = 0;
    if (lane == 0) {
        if (lo) base_lo = atomicAdd(&counters[0], (unsigned)__popc(lo));
        if (hi) base_hi = atomicAdd(&counters[1], (unsigned)__popc(hi));
    }
    base_lo = __shfl_sync(0xffffffffu, base_lo, 0);
    base_hi = __shfl_sync(0xffffffffu, base_hi, 0);
    if (!live) return;
    const unsigned below = (1u << lane) - 1u;
    // the few long ones go to the FRONT of the array (their warps start first, the kernel's tail is made of short ones)
    const size_t slot = nwin <= EDG_NWIN_SPLIT ? n - 1 - ((size_t)base_lo + __popc(lo & below)) : (size_t)base_hi + __popc(hi & below);
    ed25519_verify_front_points(state + slot * EDG_VSTATE_WORDS, v, nwin, (u32)k, sg, pk);
}

// stage 2: the window loop.  Whole warps stay together (the trip count is agreed per warp with a full-mask
// reduction): lanes past the end redo the last record and drop the result.
__global__ void __launch_bounds__(kVThreads, EDG_LB_VLOOP) k_verify(size_t n, uint8_t *ok, const u32 *state, const u32 *__restrict__ wtab) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((k & ~(size_t)31) >= n) return;
    const u32 *rec = state + (k < n ? k : n - 1) * EDG_VSTATE_WORDS;
    __shared__ uint4 s_stage[16 * kVThreads];                  // 2 table entries x 8 x 16 bytes per thread, chunk-major
    const u32 r = ed25519_verify_loop(rec, wtab, reinterpret_cast<u32 *>(s_stage));
    if (k < n) ok[rec[602]] = (uint8_t)r;
}

// Window tables of B and 2^128 B (built once per device): entry e = e * P, e = 0 .. 2^15.
__global__ void k_wtab_base(u32 *base) { wtab_base(base + 24 * threadIdx.x, 128 * (int)threadIdx.x); }

__global__ void __launch_bounds__(kThreads) k_wtab_build(u32 *tables, const u32 *bases) {
    u32 *table = tables + (size_t)blockIdx.y * EDG_WTAB_WORDS;
    const u32 *base = bases + 24 * blockIdx.y;
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

// signatures per pass: a whole number of waves of the loop kernel (resident threads of the device)
static size_t verify_chunk(int sm_count) {
    int bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_verify, kVThreads, 0) != cudaSuccess || bps < 1) bps = 1;
    return (size_t)sm_count * bps * kVThreads * EDG_VERIFY_WAVES;
}

size_t edg_verify_record_bytes(void) { return EDG_VSTATE_WORDS * sizeof(u32); }
unsigned edg_verify_waves(void) { return EDG_VERIFY_WAVES; }

// the record slab of one pass; its last 256 bytes hold the slot counters of the passes (two per pass, zeroed by the launcher)
size_t edg_verify_scratch_bytes(int sm_count) { return verify_chunk(sm_count) * EDG_VSTATE_WORDS * sizeof(u32) + 256; }

size_t edg_verify_table_bytes(void) { return (2 * (size_t)EDG_WTAB_WORDS + 48) * sizeof(u32); }

// table: edg_verify_table_bytes() of device memory = the tables of B and 2^128 B; the last 48 words are the two
// base points in affine form
int edg_verify_table_init(void *table, void *stream) {
    u32 *t = (u32 *)table, *bases = t + 2 * (size_t)EDG_WTAB_WORDS;
    k_wtab_base<<<1, 2, 0, (cudaStream_t)stream>>>(bases);
    const unsigned groups = (EDG_WTAB_ENTRIES - 1) / 8;
    k_wtab_build<<<dim3((groups + kThreads - 1) / kThreads, 2), kThreads, 0, (cudaStream_t)stream>>>(t, bases);
    return (int)cudaGetLastError();
}

// test hook (EDDSA_B200_DEBUG_FULL_SCALARS=1): every signature takes the full-length fallback (rho, tau) = (1, t)
static int g_full_scalars = 0;
void edg_verify_debug_full_scalars(int on) { g_full_scalars = on; }

// kernels edg_launch_verify(n, ..) launches: two per pass
unsigned edg_verify_launches(size_t n, int sm_count) {
    const size_t chunk = verify_chunk(sm_count);
    return (unsigned)(2 * ((n + chunk - 1) / chunk));
}

int edg_launch_verify(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                      const unsigned long long *off, unsigned long long fixed_len, void *scratch, const void *table,
                      int sm_count, void *stream) {
    const size_t chunk = verify_chunk(sm_count);
    unsigned int *counters = (unsigned int *)((u32 *)scratch + chunk * EDG_VSTATE_WORDS);
    for (size_t first = 0; first < n; first += chunk) {
        const size_t m = n - first < chunk ? n - first : chunk;
        cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), (cudaStream_t)stream);
        k_verify_front<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(m, first, sig, pub, msgs, off, fixed_len, (u32 *)scratch, counters, g_full_scalars);
        k_verify<<<(unsigned)((m + kVThreads - 1) / kVThreads), kVThreads, 0, (cudaStream_t)stream>>>(m, ok + first, (const u32 *)scratch, (const u32 *)table);
    }
    return (int)cudaGetLastError();
}

int edg_launch_pk_convert(size_t n, uint8_t *out, const uint8_t *in, int sm_count, void *stream) {
    if (n == 0) return 0;
    int g = grid_for(k_pk_convert, n, 0, sm_count, nullptr);
    k_pk_convert<<<g, kThreads, 0, (cudaStream_t)stream>>>(n, out, in);
    return (int)cudaGetLastError();
}

}  // extern "C"
