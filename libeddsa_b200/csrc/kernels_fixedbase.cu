// Fixed-base kernels (sm_100a): ed25519_genpub, ed25519_sign, x25519_base and sk_ed25519_to_x25519.
// Replaces ed25519-sha512.c:53-137, 243-256 and x25519.c:158-208.
//
// The work of these operations is of two kinds that want different kernels:
//   * the comb  x -> x * B  (+ shared inversion + encoding): bound by the integer multiplier.  ONE kernel, k_comb, serves
//     all three operations.  The signed radix-2^W table of B (sc.cuh: EDG_COMB_W; built once per device by k_comb_base /
//     k_comb_rows) is staged in shared memory once per persistent block and scanned with masks — constant time: no
//     branch or address depends on a secret.  Small code, <= 128 registers: 4 warps per scheduler.
//   * SHA-512 / arithmetic mod L (key expansion, nonce, challenge, S = r + t a): ALU-bound, no field arithmetic.  Small
//     kernels of their own (k_expand_key, k_sign_nonce, k_sign_finish), one operation per thread; ragged batches are
//     visited in tiles sorted by message length so that the lanes of a warp hash messages of similar length.
// Secret scalars travel between the stages through a scratch buffer the last reader overwrites with zeros.
// (Round 1 ran everything in one 240 KB kernel per operation at 188 / 161 registers: the multiplier pipe was 58-63 % busy.)
#include "kernel_common.cuh"
using namespace edg;

#ifndef EDG_COMB_THREADS
#define EDG_COMB_THREADS 512
#endif
#ifndef EDG_COMB_BLOCKS
#define EDG_COMB_BLOCKS 1        /* resident blocks per SM: 512 threads x 128 registers = 4 warps per scheduler, one copy of the table
                                    (132 KB) + 16 exchange areas (56 KB) per SM (2 x 256 threads: -2 %) */
#endif
#ifndef EDG_COMB_LB_THREADS
#define EDG_COMB_LB_THREADS EDG_COMB_THREADS     /* register budget of k_comb = 65536 / this */
#endif
#ifndef EDG_FIXEDBASE_PASS_LOG2
#define EDG_FIXEDBASE_PASS_LOG2 21   /* operations per pass of the staged kernels: bounds the scratch (64 B per signature) */
#endif
#ifndef EDG_MSG_TILE
#define EDG_MSG_TILE 512         /* ragged batches: consecutive operations sorted by message length together (4 per thread: small
                                    tiles keep enough blocks in flight — a verify pass of 303 104 signatures is 592 tiles) */
#endif
#ifndef EDG_LB_HASH
#define EDG_LB_HASH 4            /* hash kernels: min resident blocks per SM the register allocator must allow */
#endif
#ifndef EDG_COMB_MMA
#define EDG_COMB_MMA (EDG_COMB_W >= 5)   /* table lookups as one-hot x table products on the tensor cores (ge.cuh) */
#endif
namespace {

constexpr int kCombThreads = EDG_COMB_THREADS;
constexpr int kCombBytes = EDG_COMB_WORDS * 4 + (EDG_COMB_MMA ? (EDG_COMB_THREADS / 32) * 32 * EDG_XCHG_STRIDE * 4 : 0);
constexpr size_t kPass = (size_t)1 << EDG_FIXEDBASE_PASS_LOG2;
constexpr int kMsgTile = EDG_MSG_TILE;
constexpr int kMsgTileBits = kMsgTile == 512 ? 9 : kMsgTile == 1024 ? 10 : kMsgTile == 2048 ? 11 : 12;
static_assert(kMsgTile == (1 << kMsgTileBits), "tile size must be a power of two between 512 and 4096");

__device__ __forceinline__ void store_words8(u32 *dst, const u32 w[8]) {
    uint4 *p = reinterpret_cast<uint4 *>(dst);
    p[0] = make_uint4(w[0], w[1], w[2], w[3]);
    p[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Scratch that a kernel reads AND overwrites (the wipe) must not go through the read-only data path: ld.global.nc tells
// the compiler the memory never changes during the kernel, so it may sink such a load below the wipe and read zeros.
__device__ __forceinline__ void load_scratch8(u32 w[8], const u32 *src) {
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
    const uint4 a = p[0], b = p[1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

__device__ __forceinline__ void wipe_words8(u32 *dst) {
    volatile uint4 *p = reinterpret_cast<volatile uint4 *>(dst);
    p[0].x = 0; p[0].y = 0; p[0].z = 0; p[0].w = 0; p[1].x = 0; p[1].y = 0; p[1].z = 0; p[1].w = 0;
}

// MODE 0: out[i] = Edwards encoding of scalars[i] * B, scalars already reduced mod L (from k_expand_key / k_sign_nonce)
// MODE 1: out[i] = Montgomery u of (clamp(scalars[i]) mod L) * B for raw 32-byte scalars (x25519_base)
// Every thread runs up to EDG_BATCH operations with one shared inversion (ops.cuh: fe_batch_inv).
template <int MODE>
__global__ void __launch_bounds__(EDG_COMB_LB_THREADS, EDG_COMB_BLOCKS) k_comb(size_t n, uint8_t *out, unsigned out_stride, u32 *scalars, int wipe,
                                                                      const u32 *__restrict__ comb_g) {
    extern __shared__ __align__(16) u32 s_comb[];                // the table | one exchange area per warp (EDG_COMB_MMA)
    stage_table(s_comb, comb_g, EDG_COMB_WORDS);
#if EDG_COMB_MMA
    u32 *xchg = s_comb + EDG_COMB_WORDS + (threadIdx.x >> 5) * (32 * EDG_XCHG_STRIDE);
#endif
    const size_t T = (size_t)gridDim.x * blockDim.x;
    // whole warps stay together (the tensor-core lookup is warp-synchronous): lanes past the end of the batch redo the
    // last operation and drop the result
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; (i0 & ~(size_t)31) < n; i0 += T * EDG_BATCH) {
        fe U[EDG_BATCH], V[MODE == 0 ? EDG_BATCH : 1], Z[EDG_BATCH];     // MODE 0: X, Y, Z;  MODE 1: Z + Y, -, Z - Y
        int cnt = 0;
#pragma unroll 1
        for (int k = 0; k < EDG_BATCH; k++) {
            const size_t i = i0 + (size_t)k * T;
            if ((i & ~(size_t)31) >= n) break;                 // (warp-uniform)
            const bool live = i < n;
            const size_t src = live ? i : n - 1;
            u32 x[8];
            load_scratch8(x, scalars + 8 * src);
            if (MODE == 1) {
                u32 e[8];
#pragma unroll
                for (int w = 0; w < 8; w++) e[w] = x[w];
                clamp_words(e);                                // x25519.c:170-172
                sc_reduce256(x, e);                            // Q9
            }
            ge_p3 R;
#if EDG_COMB_MMA
            ge_scalarmult_base_ct_mma(R, x, s_comb, xchg);
#else
            ge_scalarmult_base_ct(R, x, s_comb);
#endif
            if (MODE == 0) { fe_copy(U[k], R.X); fe_copy(V[k], R.Y); fe_copy(Z[k], R.Z); }
            else { fe_add(U[k], R.Z, R.Y); fe_sub(Z[k], R.Z, R.Y); }                       // x25519.c:190-194
            cnt += live ? 1 : 0;
        }
        fe_batch_inv(Z, cnt);
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            u32 o[8];
            if (MODE == 0) ge_tobytes_zinv(o, U[k], V[k], Z[k]);
            else x25519_back(o, U[k], Z[k]);
            uint4 *p = reinterpret_cast<uint4 *>(out + (size_t)out_stride * (i0 + (size_t)k * T));
            p[0] = make_uint4(o[0], o[1], o[2], o[3]);
            p[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        scrub(U, EDG_BATCH);
        scrub(V, MODE == 0 ? EDG_BATCH : 1);
        scrub(Z, EDG_BATCH);
    }
    // the scalars are wiped once every thread is past its last read: a lane past the end re-reads ANOTHER lane's scalar
    if (wipe) {
        __syncthreads();
        // (grid-wide ordering is not needed: only lanes of the last warp of the batch re-read a scalar, and it belongs to
        //  their own warp)
        for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += T) wipe_words8(scalars + 8 * i0);
    }
}

// comb table in fragment order for ge_pre_select_mma (ops.cuh: comb_mma_word)
__global__ void k_comb_layout(u32 *mma, const u32 *table) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < EDG_COMB_WORDS) mma[idx] = comb_mma_word(table, idx);
}

// a[i] = clamp(SHA512(sec[i])[0..31]) mod L                                     [ed25519_key_setup, ed25519-sha512.c:31-47, :62]
__global__ void __launch_bounds__(kThreads) k_expand_key(size_t n, u32 *a_out, const uint8_t *sec) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u32 a[8];
        u64 prefix[4];
        ed25519_expand_key(a, prefix, sec + 32 * i);
        store_words8(a_out + 8 * i, a);
    }
}

// Visits operations [0, n) once each: one per thread per grid stride, or (RAGGED: the batch has an offsets array) in
// tiles of kMsgTile consecutive operations sorted by message length in shared memory (kernel_common.cuh), so that
// neighbouring lanes get messages of neighbouring lengths.  Lengths are public.
template <bool RAGGED, typename F>
__device__ __forceinline__ void for_each_message(size_t n, const unsigned long long *off, F body) {
    if constexpr (!RAGGED) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) body(i);
    } else {
        __shared__ u32 s_key[kMsgTile];
        for (size_t t0 = (size_t)blockIdx.x * kMsgTile; t0 < n; t0 += (size_t)gridDim.x * kMsgTile) {
            __syncthreads();                                   // the previous tile's keys are no longer needed
            for (int e = threadIdx.x; e < kMsgTile; e += blockDim.x) s_key[e] = ragged_key<kMsgTileBits>(off, t0, e, n);
            block_sort_u32<kMsgTile>(s_key);
            for (int e = threadIdx.x; e < kMsgTile; e += blockDim.x) {
                const u32 key = s_key[e];
                if (key == 0xffffffffu) break;                 // past the end of the batch (sorted last)
                body(t0 + (key & (kMsgTile - 1)));
            }
        }
    }
}

// sign, stage 1: a[i] (as k_expand_key) and the nonce r[i] = H(prefix || M) mod L                  [ed25519-sha512.c:96-105]
template <bool RAGGED>
__global__ void __launch_bounds__(kThreads, EDG_LB_HASH) k_sign_nonce(size_t n, u32 *a_out, u32 *r_out, const uint8_t *sec, const uint8_t *msgs,
                                                         const unsigned long long *off, unsigned long long fixed_len) {
    for_each_message<RAGGED>(n, off, [&](size_t i) {
        const uint8_t *m; u64 len;
        msg_of(m, len, msgs, off, fixed_len, i);
        u32 a[8], r[8];
        ed25519_sign_nonce(a, r, sec + 32 * i, m, len);
        store_words8(a_out + 8 * i, a);
        store_words8(r_out + 8 * i, r);
    });
}

// sign, stage 3: S = r + H(R || pub || M) a mod L next to the R bytes k_comb wrote into sig[i]; wipes a[i], r[i].  [:112-122]
template <bool RAGGED>
__global__ void __launch_bounds__(kThreads, EDG_LB_HASH) k_sign_finish(size_t n, uint8_t *sig, u32 *a_in, u32 *r_in, const uint8_t *pub, const uint8_t *msgs,
                                                          const unsigned long long *off, unsigned long long fixed_len) {
    for_each_message<RAGGED>(n, off, [&](size_t i) {
        const uint8_t *m; u64 len;
        msg_of(m, len, msgs, off, fixed_len, i);
        u32 R[8], p[8], t[8];
        load_scratch8(R, reinterpret_cast<const u32 *>(sig + 64 * i));     // (sig is written below: not read-only either)
        load8(p, pub, i);
        ed25519_sign_challenge(t, R, p, m, len);               // public data only
        u32 a[8], r[8], S[8];
        load_scratch8(a, a_in + 8 * i);                        // the secrets enter here
        load_scratch8(r, r_in + 8 * i);
        sc_muladd(S, t, a, r);
        store8(sig, 2 * i + 1, S);
        wipe_words8(a_in + 8 * i);
        wipe_words8(r_in + 8 * i);
    });
}

__global__ void __launch_bounds__(kThreads) k_sk_convert(size_t n, uint8_t *out, const uint8_t *in) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u32 o[8];
        sk_ed25519_to_x25519_op(o, in + 32 * i);
        store8(out, i, o);
    }
}

// comb table: bases[j] = 2^(W j) B (one thread per row), then row j = 1 .. 2^(W-1) multiples of bases[j]
__global__ void k_comb_base(u32 *bases) { wtab_base(bases + 24 * threadIdx.x, EDG_COMB_W * (int)threadIdx.x); }

__global__ void k_comb_rows(u32 *table, const u32 *bases) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x, row = g / (EDG_COMB_ENTRIES / 8), grp = g % (EDG_COMB_ENTRIES / 8);
    if (row < EDG_COMB_ROWS) wtab_build8(table + (size_t)row * (EDG_COMB_ENTRIES * 24) + 24u * 8u * grp, bases + 24 * row, 8u * grp + 1u);
}

template <typename K>
int msg_grid(K kernel, size_t n, bool ragged, int sm_count) {
    int bps = 0;
    grid_for(kernel, n, 0, sm_count, &bps);
    const size_t per = ragged ? (size_t)kMsgTile : (size_t)kThreads, need = (n + per - 1) / per, cap = (size_t)sm_count * bps;
    return (int)(need < cap ? need : cap);
}

int comb_grid(size_t n, int sm_count) {
    const size_t need = (n + kCombThreads - 1) / kCombThreads, cap = (size_t)sm_count * EDG_COMB_BLOCKS;
    return (int)(need < cap ? need : cap);
}

}  // namespace

extern "C" {

int edg_kernels_init(void) {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_comb<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCombBytes); if (e) return (int)e;
    e = cudaFuncSetAttribute(k_comb<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCombBytes); if (e) return (int)e;
    return 0;
}

size_t edg_comb_table_payload_bytes(void) { return (size_t)EDG_COMB_WORDS * sizeof(u32); }
// device allocation: the table | the row base points | the table in fragment order (what the kernels stage when EDG_COMB_MMA)
size_t edg_comb_table_bytes(void) { return (2 * (size_t)EDG_COMB_WORDS + 24u * EDG_COMB_ROWS) * sizeof(u32); }
static const u32 *comb_for_kernels(const void *table) { return (const u32 *)table + (EDG_COMB_MMA ? EDG_COMB_WORDS + 24u * EDG_COMB_ROWS : 0); }
void edg_comb_geometry(int *rows, int *entries) { *rows = EDG_COMB_ROWS; *entries = EDG_COMB_ENTRIES; }

int edg_comb_table_init(void *table, void *stream) {
    u32 *t = (u32 *)table, *bases = t + EDG_COMB_WORDS;
    k_comb_base<<<1, EDG_COMB_ROWS, 0, (cudaStream_t)stream>>>(bases);
    const unsigned groups = EDG_COMB_ROWS * (EDG_COMB_ENTRIES / 8);
    k_comb_rows<<<(groups + 63) / 64, 64, 0, (cudaStream_t)stream>>>(t, bases);
    k_comb_layout<<<(EDG_COMB_WORDS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bases + 24 * EDG_COMB_ROWS, t);
    return (int)cudaGetLastError();
}

// scratch of one pass: the secret scalar a (genpub, sign) and the nonce r (sign), 32 bytes each per operation
size_t edg_fixedbase_scratch_bytes(int is_sign, size_t n) { return (n < kPass ? n : kPass) * (is_sign ? 64 : 32); }

int edg_launch_x25519_base(size_t n, uint8_t *out, const uint8_t *scalar, const void *comb, int sm_count, void *stream) {
    if (n == 0) return 0;
    k_comb<1><<<comb_grid(n, sm_count), kCombThreads, kCombBytes, (cudaStream_t)stream>>>(n, out, 32u, (u32 *)scalar, 0, comb_for_kernels(comb));
    return (int)cudaGetLastError();
}

int edg_launch_genpub(size_t n, uint8_t *pub, const uint8_t *sec, void *scratch, const void *comb, int sm_count, void *stream, unsigned *launches) {
    cudaStream_t st = (cudaStream_t)stream;
    u32 *a = (u32 *)scratch;
    for (size_t first = 0; first < n; first += kPass) {
        const size_t m = n - first < kPass ? n - first : kPass;
        k_expand_key<<<msg_grid(k_expand_key, m, false, sm_count), kThreads, 0, st>>>(m, a, sec + 32 * first);
        k_comb<0><<<comb_grid(m, sm_count), kCombThreads, kCombBytes, st>>>(m, pub + 32 * first, 32u, a, 1, comb_for_kernels(comb));
        *launches += 2;
    }
    return (int)cudaGetLastError();
}

int edg_launch_sign(size_t n, uint8_t *sig, const uint8_t *sec, const uint8_t *pub, const uint8_t *msgs, const unsigned long long *off,
                    unsigned long long fixed_len, void *scratch, const void *comb, int sm_count, void *stream, unsigned *launches) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cap = n < kPass ? n : kPass;
    u32 *a = (u32 *)scratch, *r = a + 8 * cap;
    for (size_t first = 0; first < n; first += kPass) {
        const size_t m = n - first < kPass ? n - first : kPass;
        const uint8_t *mp = off ? msgs : msgs + first * fixed_len;
        const unsigned long long *op = off ? off + first : nullptr;
        if (off) k_sign_nonce<true><<<msg_grid(k_sign_nonce<true>, m, true, sm_count), kThreads, 0, st>>>(m, a, r, sec + 32 * first, mp, op, fixed_len);
        else k_sign_nonce<false><<<msg_grid(k_sign_nonce<false>, m, false, sm_count), kThreads, 0, st>>>(m, a, r, sec + 32 * first, mp, op, fixed_len);
        k_comb<0><<<comb_grid(m, sm_count), kCombThreads, kCombBytes, st>>>(m, sig + 64 * first, 64u, r, 0, comb_for_kernels(comb));
        if (off) k_sign_finish<true><<<msg_grid(k_sign_finish<true>, m, true, sm_count), kThreads, 0, st>>>(m, sig + 64 * first, a, r, pub + 32 * first, mp, op, fixed_len);
        else k_sign_finish<false><<<msg_grid(k_sign_finish<false>, m, false, sm_count), kThreads, 0, st>>>(m, sig + 64 * first, a, r, pub + 32 * first, mp, op, fixed_len);
        *launches += 3;
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
