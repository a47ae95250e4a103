// SHA-512 (FIPS 180-4), one message stream per thread, for the three hashes of the Ed25519 path:
//   SHA512(sk)                      -> key setup          (ed25519-sha512.c:31-47)
//   SHA512(prefix32 || M)           -> nonce r            (ed25519-sha512.c:101-105)
//   SHA512(R32 || A32 || M)         -> challenge t        (ed25519-sha512.c:112-117, 166-171)
// Role of the reference's lib/sha512.c (compress sha512.c:83-124, init/add/final :127-210).
// Design for the GPU: no streaming context and no byte buffer — the hashed string is described as
// "PRE_WORDS big-endian 64-bit words held in registers, followed by `len` message bytes in global
// memory"; blocks are assembled directly in the 16-word rolling schedule.  64-bit rotates become
// SHF funnel-shift pairs, 64-bit adds IADD3/IADD3.X on the ALU pipe.  Full message blocks that are
// 16-byte aligned are fetched with 128-bit loads (the fixed 64 B / 1 KB layouts of the BASELINE
// configs always are); ragged tails and unaligned rows fall back to byte loads.
#pragma once
#include "fe.cuh"

namespace edg {

#if defined(__CUDACC__)
#define EDG_K512 c_K512
__constant__ u64 c_K512[80] = {
#else
#define EDG_K512 h_K512
static const u64 h_K512[80] = {
#endif
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL,
    0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL,
    0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL,
    0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL,
    0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL,
    0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, 0x06ca6351e003826fULL, 0x142929670a0e6e70ULL,
    0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL,
    0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL,
    0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL,
    0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL,
    0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL,
    0xca273eceea26619cULL, 0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL,
    0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL,
    0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};

// rotate right by a compile-time constant: two funnel shifts on the device (the generic expression compiles to five
// shift / multiply instructions per rotate)
EDG_HD u64 ror64(u64 x, int n) {
#if defined(__CUDA_ARCH__)
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    const u32 a = n < 32 ? lo : hi, b = n < 32 ? hi : lo;             // (b:a) >> (n mod 32) gives the low word, (a:b) the high
    const u32 rl = __funnelshift_r(a, b, (u32)n & 31u), rh = __funnelshift_r(b, a, (u32)n & 31u);
    return ((u64)rh << 32) | rl;
#else
    return (x >> n) | (x << (64 - n));
#endif
}

EDG_HD u32 bswap32(u32 x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}

// low 32 bits of (hi:lo) >> sh, 0 <= sh < 32
EDG_HD u32 funnel_r(u32 lo, u32 hi, u32 sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// big-endian 64-bit word from two little-endian 32-bit words holding bytes [0..3], [4..7]
EDG_HD u64 be64_from_le_words(u32 w0, u32 w1) { return ((u64)bswap32(w0) << 32) | bswap32(w1); }

EDG_HD void sha512_iv(u64 s[8]) {
    s[0] = 0x6a09e667f3bcc908ULL; s[1] = 0xbb67ae8584caa73bULL; s[2] = 0x3c6ef372fe94f82bULL; s[3] = 0xa54ff53a5f1d36f1ULL;
    s[4] = 0x510e527fade682d1ULL; s[5] = 0x9b05688c2b3e6c1fULL; s[6] = 0x1f83d9abfb41bd6bULL; s[7] = 0x5be0cd19137e2179ULL;
}

// one compression; w[16] is consumed (used as the rolling schedule).   [reference: compress, sha512.c:83-124]
EDG_NOINLINE void sha512_compress(u64 h[8], u64 w[16]) {
    u64 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
    for (int i = 0; i < 80; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (i > 0) {
                const u64 w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                w[j] += (ror64(w15, 1) ^ ror64(w15, 8) ^ (w15 >> 7)) + w[(j + 9) & 15] + (ror64(w2, 19) ^ ror64(w2, 61) ^ (w2 >> 6));
            }
            const u64 t1 = hh + (ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41)) + ((e & f) ^ (~e & g)) + EDG_K512[i + j] + w[j];
            const u64 t2 = (ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// The 8-byte big-endian word at message offset m of the PADDED message (0x80 after the last byte,
// zeros elsewhere; the length field is patched in by the caller).
// Inlined 32 times this helper is most of the hash's code; the verify kernels (public data only) call it
// out of line to keep their instruction footprint small, the secret-key kernels keep it inline so that no
// secret-holding register is ever saved to the stack around a call.
#if defined(EDG_SHA_MSGWORD_NOINLINE) && defined(__CUDACC__)
static __host__ __device__ __noinline__ u64 sha512_msg_word(const uint8_t *msg, u64 len, u64 m) {
#else
EDG_HD u64 sha512_msg_word(const uint8_t *msg, u64 len, u64 m) {
#endif
    if (m + 8 <= len) {
        const uint8_t *p = msg + m;
        if ((((uintptr_t)p) & 7) == 0) {
            const u32 *q = (const u32 *)p;
            return be64_from_le_words(q[0], q[1]);
        }
        if (m + 12 <= len) {
            // unaligned, at least 4 more message bytes behind the word: three aligned 32-bit loads and two funnel
            // shifts.  The loads stay inside [message start rounded down to 4, message end) — allocations start
            // 4-byte aligned, so nothing outside the caller's buffer is touched.
            const uintptr_t ad = (uintptr_t)p;
            const u32 *q = (const u32 *)(ad & ~(uintptr_t)3);
            const u32 sh = (u32)(ad & 3) * 8;
            const u32 w0 = q[0], w1 = q[1], w2 = q[2];
            return be64_from_le_words(funnel_r(w0, w1, sh), funnel_r(w1, w2, sh));
        }
        u64 v = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) v = (v << 8) | p[j];
        return v;
    }
    if (m > len) return 0;
    u64 v = 0;
    for (int j = 0; j < 8; j++) {
        const u64 o = m + j;
        const u64 byte = o < len ? msg[o] : (o == len ? 0x80u : 0u);
        v = (v << 8) | byte;
    }
    return v;
}

// state = SHA512( pre[0..PRE_WORDS) as big-endian words  ||  msg[0..len) ).
// PRE_WORDS is 0, 4 (32-byte prefix) or 8 (64-byte prefix).          [reference: sha512_init/add/final, sha512.c:127-210]
template <int PRE_WORDS>
EDG_HD void sha512_prefixed(u64 state[8], const u64 *pre, const uint8_t *msg, u64 len) {
    const u64 pre_bytes = 8 * PRE_WORDS;
    const u64 total = pre_bytes + len;
    const u64 nblocks = (total + 144) >> 7;       // data + 0x80 + 16-byte length, rounded up to 128
    sha512_iv(state);
    u64 w[16];
    for (u64 blk = 0; blk < nblocks; blk++) {
        if (blk == 0) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k < PRE_WORDS) w[k] = pre[k];
                else w[k] = sha512_msg_word(msg, len, (u64)(8 * (k - PRE_WORDS)));
            }
        } else {
            const u64 m0 = (blk << 7) - pre_bytes;
            const uint8_t *p = msg + m0;
            if (m0 + 128 <= len && (((uintptr_t)p) & 15) == 0) {
#if defined(__CUDA_ARCH__)
                const uint4 *q = (const uint4 *)p;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 v = __ldg(q + k);
                    w[2 * k] = be64_from_le_words(v.x, v.y);
                    w[2 * k + 1] = be64_from_le_words(v.z, v.w);
                }
#else
                const u32 *q = (const u32 *)p;
                for (int k = 0; k < 16; k++) w[k] = be64_from_le_words(q[2 * k], q[2 * k + 1]);
#endif
            } else if (m0 + 144 <= len) {
                // unaligned full block with 16 more message bytes behind it: nine aligned 128-bit loads (the block's
                // bytes start 0..15 bytes into them) and funnel shifts; word offset and byte shift are per message
                const uintptr_t ad = (uintptr_t)p;
                const u32 ws = (u32)(ad >> 2) & 3u, sh = (u32)(ad & 3) * 8;
                u32 x[36];
#if defined(__CUDA_ARCH__)
                const uint4 *q = (const uint4 *)(ad & ~(uintptr_t)15);
#pragma unroll
                for (int k = 0; k < 9; k++) { const uint4 v = __ldg(q + k); x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w; }
#else
                const u32 *q = (const u32 *)(ad & ~(uintptr_t)15);
                for (int k = 0; k < 36; k++) x[k] = q[k];
#endif
                // y[k] = x[k + ws] for k = 0..32 (two conditional moves per word: shift by 2, then by 1)
#pragma unroll
                for (int k = 0; k < 34; k++) x[k] = (ws & 2u) ? x[k + 2] : x[k];
#pragma unroll
                for (int k = 0; k < 33; k++) x[k] = (ws & 1u) ? x[k + 1] : x[k];
#pragma unroll
                for (int k = 0; k < 16; k++)
                    w[k] = be64_from_le_words(funnel_r(x[2 * k], x[2 * k + 1], sh), funnel_r(x[2 * k + 1], x[2 * k + 2], sh));
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) w[k] = sha512_msg_word(msg, len, m0 + 8 * k);
            }
        }
        if (blk == nblocks - 1) {                 // 128-bit big-endian bit length   [sha512.c:196-203]
            w[14] = total >> 61;
            w[15] = total << 3;
        }
        sha512_compress(state, w);
    }
}

// The 64 digest bytes read as a little-endian 512-bit integer, as 16 words (input of sc_reduce512).
EDG_HD void sha512_state_to_le_words(u32 x[16], const u64 state[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[2 * i] = bswap32((u32)(state[i] >> 32));
        x[2 * i + 1] = bswap32((u32)state[i]);
    }
}

}  // namespace edg
