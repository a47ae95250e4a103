// GF(2^255-19) arithmetic for sm_100a — one field element per thread, held in registers.
//
// Role of the reference's lib/fld.c + lib/fld.h (32-bit path: fld.c:282-531, common code
// fld.c:540-709), re-designed for the B200 integer pipes:
//   * 10 UNSIGNED limbs, alternating 26/25 bits (radix 2^25.5), value = sum v[i] * 2^ceil(25.5 i).
//   * every limb product is one 32x32->64 IMAD.WIDE.U32; the x19 wrap-around and the x2 of
//     odd*odd pairs are folded into 32-bit pre-multiplied operands, so a multiplication is exactly
//     100 wide products (55 for a squaring) and the accumulators never leave registers.
//   * additions / subtractions are lazy (no carry); subtraction adds a multiple of p limb-wise so
//     limbs stay non-negative (the reference uses signed limbs instead; only canonical bytes have
//     to agree, SURVEY.md §7 hard part 1).
//
// Limb bounds ("tight" = output of fe_mul/fe_sq/fe_carry):
//   even limbs <= 2^26 + 2^13, odd limbs <= 2^25 + 2^17.
// fe_mul(a, b):  a even <= 2^28.3, odd <= 2^27.3 ; b even <= 2^27.7, odd <= 2^26.7 (b gets the x19).
// fe_sq(a):      a even <= 2^27.7, odd <= 2^26.7.
// These are checked by tests/test_fe_host.py (interval test through the host build of this header).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EDG_HD __host__ __device__ __forceinline__
#define EDG_NOINLINE __host__ __device__ __forceinline__   /* compact loop bodies: cheap to inline, keeps secrets off the stack */
#else
#define EDG_HD static inline
#define EDG_NOINLINE static
#endif

namespace edg {

typedef uint32_t u32;
typedef uint64_t u64;

struct fe { u32 v[10]; };

#define EDG_M26 0x3ffffffu
#define EDG_M25 0x1ffffffu

EDG_HD u64 mulw(u32 a, u32 b) { return (u64)a * (u64)b; }

// Optimisation barrier for constant-time masks: hides from the compiler that m is 0 / ~0, so mask
// arithmetic (x ^ ((x ^ y) & m)) can never be turned into a select, a branch or a load predicated on
// a secret (nvcc does exactly that to naive mask idioms — see tests/ct_negative/leaky.cu).
EDG_HD u32 ct_mask(u32 m) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(m));
#endif
    return m;
}

// host-only instrumentation used by tests/host_sim to pin the field-operation counts quoted in DESIGN.md
#if defined(EDG_COUNT_OPS) && !defined(__CUDA_ARCH__)
static unsigned long edg_cnt_mul = 0, edg_cnt_sq = 0;
#define EDG_COUNT_MUL() (++edg_cnt_mul)
#define EDG_COUNT_SQ() (++edg_cnt_sq)
#else
#define EDG_COUNT_MUL() ((void)0)
#define EDG_COUNT_SQ() ((void)0)
#endif

EDG_HD void fe_set_u32(fe &r, u32 x) {
    r.v[0] = x;
#pragma unroll
    for (int i = 1; i < 10; i++) r.v[i] = 0;
}

EDG_HD void fe_copy(fe &r, const fe &a) {
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = a.v[i];
}

// r = a + b (lazy)                                   [reference: fld_add, fld.h:94]
EDG_HD void fe_add(fe &r, const fe &a, const fe &b) {
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = a.v[i] + b.v[i];
}

// limbs of 2p: every limb of b that is <= these can be subtracted without going negative.
#define EDG_2P0 0x7ffffdau   /* 2*(2^26-19) */
#define EDG_2PE 0x7fffffeu   /* 2*(2^26-1)  */
#define EDG_2PO 0x3fffffeu   /* 2*(2^25-1)  */

// r = a - b + 2p (lazy); requires b even <= 2^27-38, b odd <= 2^26-2   [reference: fld_sub, fld.h:102]
EDG_HD void fe_sub(fe &r, const fe &a, const fe &b) {
    r.v[0] = a.v[0] + EDG_2P0 - b.v[0];
#pragma unroll
    for (int i = 1; i < 10; i++) r.v[i] = a.v[i] + ((i & 1) ? EDG_2PO : EDG_2PE) - b.v[i];
}

// r = a - b + 4p (lazy); requires b even <= 2^28-76, b odd <= 2^27-4
EDG_HD void fe_sub4(fe &r, const fe &a, const fe &b) {
    r.v[0] = a.v[0] + 2u * EDG_2P0 - b.v[0];
#pragma unroll
    for (int i = 1; i < 10; i++) r.v[i] = a.v[i] + 2u * ((i & 1) ? EDG_2PO : EDG_2PE) - b.v[i];
}

// r = 2p - a (lazy negate); requires a tight            [reference: fld_neg, fld.h:138]
EDG_HD void fe_neg(fe &r, const fe &a) {
    r.v[0] = EDG_2P0 - a.v[0];
#pragma unroll
    for (int i = 1; i < 10; i++) r.v[i] = ((i & 1) ? EDG_2PO : EDG_2PE) - a.v[i];
}

// r = 2a (lazy)                                        [reference: fld_scale2, fld.h:126]
EDG_HD void fe_dbl(fe &r, const fe &a) {
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = a.v[i] << 1;
}

// Carry a 10-column 64-bit accumulator set down to tight limbs.  Two interleaved chains
// (0->1->..->5 and 5->6->..->9->0->1) so that consecutive steps are independent.
// Role of the CARRY macro, fld.c:318-330.
EDG_HD void fe_carry64(fe &r, u64 h[10]) {
    u64 c0, c5;
    c0 = h[0] >> 26; h[1] += c0; h[0] &= EDG_M26;
    c5 = h[5] >> 25; h[6] += c5; h[5] &= EDG_M25;
    c0 = h[1] >> 25; h[2] += c0; h[1] &= EDG_M25;
    c5 = h[6] >> 26; h[7] += c5; h[6] &= EDG_M26;
    c0 = h[2] >> 26; h[3] += c0; h[2] &= EDG_M26;
    c5 = h[7] >> 25; h[8] += c5; h[7] &= EDG_M25;
    c0 = h[3] >> 25; h[4] += c0; h[3] &= EDG_M25;
    c5 = h[8] >> 26; h[9] += c5; h[8] &= EDG_M26;
    c0 = h[4] >> 26; h[5] += c0; h[4] &= EDG_M26;
    c5 = h[9] >> 25; h[0] += c5 * 19u; h[9] &= EDG_M25;
    c0 = h[5] >> 25; h[6] += c0; h[5] &= EDG_M25;
    c5 = h[0] >> 26; h[1] += c5; h[0] &= EDG_M26;
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = (u32)h[i];
}

// Cheap 32-bit carry pass for lazily added values (limbs < 2^32): result tight.
EDG_HD void fe_carry(fe &r, const fe &a) {
    u32 t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = a.v[i];
    u32 c;
    c = t[9] >> 25; t[9] &= EDG_M25; t[0] += 19u * c;      // c < 2^7
#pragma unroll
    for (int i = 0; i < 9; i++) {
        if (i & 1) { c = t[i] >> 25; t[i] &= EDG_M25; }
        else       { c = t[i] >> 26; t[i] &= EDG_M26; }
        t[i + 1] += c;
    }
    // t[9] < 2^25 + 2^7 ; leave (well inside tight bound)
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = t[i];
}

// r = a * b mod p, tight output.  100 IMAD.WIDE.U32.            [reference: fld_mul, fld.c:448]
EDG_HD void fe_mul(fe &r, const fe &a, const fe &b) {
    EDG_COUNT_MUL();
    u32 b19[10], a2[10];
#pragma unroll
    for (int j = 1; j < 10; j++) b19[j] = 19u * b.v[j];
#pragma unroll
    for (int i = 1; i < 10; i += 2) a2[i] = a.v[i] << 1;
    u64 h[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        u64 s = 0;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const int j = (k - i + 10) % 10;
            const bool wrap = i > k;
            const bool both_odd = (i & 1) && (j & 1);
            const u32 x = both_odd ? a2[i] : a.v[i];
            const u32 y = wrap ? b19[j] : b.v[j];
            s += mulw(x, y);
        }
        h[k] = s;
    }
    fe_carry64(r, h);
}

// r = a^2 mod p, tight output.  55 IMAD.WIDE.U32.                [reference: fld_sq, fld.c:503]
EDG_HD void fe_sq(fe &r, const fe &a) {
    EDG_COUNT_SQ();
    // d[i] = 2 a_i ; w[j] = 19 a_j (j even) or 38 a_j (j odd) for the wrapped half
    u32 d[10], w[10];
#pragma unroll
    for (int i = 0; i < 10; i++) d[i] = a.v[i] << 1;
#pragma unroll
    for (int j = 5; j < 10; j++) w[j] = ((j & 1) ? 38u : 19u) * a.v[j];
    u64 h[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        u64 s = 0;
#pragma unroll
        for (int i = 0; i < 10; i++) {
#pragma unroll
            for (int j = i; j < 10; j++) {
                if ((i + j) % 10 != k) continue;
                const bool wrap = (i + j) >= 10;
                const bool both_odd = (i & 1) && (j & 1);
                // coefficient of a_i a_j: (i<j ? 2 : 1) * (both_odd ? 2 : 1) * (wrap ? 19 : 1)
                u32 x, y;
                if (!wrap) {
                    if (i == j) { x = both_odd ? d[i] : a.v[i]; y = a.v[j]; }
                    else        { x = d[i]; y = both_odd ? d[j] : a.v[j]; }
                } else {
                    // wrap implies j >= 5; w[j] already carries 19 (j even) or 38 (j odd)
                    const int twos = (i < j ? 1 : 0) + (both_odd ? 1 : 0) - ((j & 1) ? 1 : 0);  // 0 or 1
                    x = twos ? d[i] : a.v[i];
                    y = w[j];
                }
                s += mulw(x, y);
            }
        }
        h[k] = s;
    }
    fe_carry64(r, h);
}

// r = a * 121665 mod p (tight a), tight output.         [reference: fld_scale, fld.c:430; only s=121665 is used, x25519.c:77]
EDG_HD void fe_mul121665(fe &r, const fe &a) {
    u64 h[10];
#pragma unroll
    for (int i = 0; i < 10; i++) h[i] = mulw(a.v[i], 121665u);
    fe_carry64(r, h);
}

// n successive squarings
EDG_HD void fe_sqn(fe &r, const fe &a, int n) {
    fe_sq(r, a);
#pragma unroll 1
    for (int i = 1; i < n; i++) fe_sq(r, r);
}

// Unique representative in [0, p), limbs exactly 26/25 bits.          [reference: fld_reduce, fld.c:342]
EDG_HD void fe_canon(fe &r, const fe &a) {
    u32 t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = a.v[i];
    u32 c;
    // two plain carry rounds bring the value into [0, 2^255 + small)
#pragma unroll
    for (int round = 0; round < 2; round++) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            if (i & 1) { c = t[i] >> 25; t[i] &= EDG_M25; }
            else       { c = t[i] >> 26; t[i] &= EDG_M26; }
            t[i + 1] += c;
        }
        c = t[9] >> 25; t[9] &= EDG_M25; t[0] += 19u * c;
    }
    // now value < 2^255 + 19*2^7 and all limbs but t[0] are in range; t[0] < 2^26 + 19*2^7.
    // q = 1 iff value >= p  <=>  value + 19 >= 2^255 : propagate the carry of (value + 19).
    c = (t[0] + 19u) >> 26;
#pragma unroll
    for (int i = 1; i < 10; i++) c = (t[i] + c) >> ((i & 1) ? 25 : 26);
    // c is 0 or 1 (or 2 only if value >= 2^256-ish, impossible here)
    t[0] += 19u * c;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        u32 cc;
        if (i & 1) { cc = t[i] >> 25; t[i] &= EDG_M25; }
        else       { cc = t[i] >> 26; t[i] &= EDG_M26; }
        t[i + 1] += cc;
    }
    t[9] &= EDG_M25;   // drops the 2^255 that "value - p = value + 19 - 2^255" subtracts
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = t[i];
}

// 32 little-endian bytes (as 8 LE words) -> fe.  ALL 256 bits are taken; bit 255 folds in as +19
// (SURVEY Q6).                                                        [reference: fld_import, fld.c:383]
EDG_HD void fe_from_words(fe &r, const u32 w[8]) {
    // limb i starts at bit offset o_i = ceil(25.5 i): 0,26,51,77,102,128,153,179,204,230
    r.v[0] = w[0] & EDG_M26;
    r.v[1] = ((w[0] >> 26) | (w[1] << 6)) & EDG_M25;
    r.v[2] = ((w[1] >> 19) | (w[2] << 13)) & EDG_M26;
    r.v[3] = ((w[2] >> 13) | (w[3] << 19)) & EDG_M25;
    r.v[4] = (w[3] >> 6) & EDG_M26;
    r.v[5] = w[4] & EDG_M25;
    r.v[6] = ((w[4] >> 25) | (w[5] << 7)) & EDG_M26;
    r.v[7] = ((w[5] >> 19) | (w[6] << 13)) & EDG_M25;
    r.v[8] = ((w[6] >> 12) | (w[7] << 20)) & EDG_M26;
    r.v[9] = (w[7] >> 6) & EDG_M25;
    r.v[0] += 19u * (w[7] >> 31);
}

// fe -> canonical 32 bytes (8 LE words).                              [reference: fld_export, fld.c:406]
EDG_HD void fe_to_words(u32 w[8], const fe &a) {
    fe t;
    fe_canon(t, a);
    w[0] = t.v[0] | (t.v[1] << 26);
    w[1] = (t.v[1] >> 6) | (t.v[2] << 19);
    w[2] = (t.v[2] >> 13) | (t.v[3] << 13);
    w[3] = (t.v[3] >> 19) | (t.v[4] << 6);
    w[4] = t.v[5] | (t.v[6] << 25);
    w[5] = (t.v[6] >> 7) | (t.v[7] << 19);
    w[6] = (t.v[7] >> 13) | (t.v[8] << 12);
    w[7] = (t.v[8] >> 20) | (t.v[9] << 6);
}

// 1 if a == 0 mod p else 0; branch-free.                              [reference: fld_eq, fld.c:547]
EDG_HD u32 fe_is_zero(const fe &a) {
    fe t;
    fe_canon(t, a);
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) x |= t.v[i];
    return (u32)(((u64)x - 1) >> 63);
}

// a, b tight -> 1 if equal mod p
EDG_HD u32 fe_eq(const fe &a, const fe &b) {
    fe t;
    fe_sub(t, a, b);
    return fe_is_zero(t);
}

// r = mask ? b : a  (mask all-ones or zero), branch-free   [reference: memselect, ed.c:80]
EDG_HD void fe_select(fe &r, const fe &a, const fe &b, u32 mask) {
#pragma unroll
    for (int i = 0; i < 10; i++) r.v[i] = a.v[i] ^ ((a.v[i] ^ b.v[i]) & mask);
}

// conditional swap under mask (all-ones or zero), branch-free   [reference: ctmemswap, x25519.c:36]
EDG_HD void fe_cswap(fe &a, fe &b, u32 mask) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        u32 d = (a.v[i] ^ b.v[i]) & mask;
        a.v[i] ^= d;
        b.v[i] ^= d;
    }
}

// Fixed exponentiations z^(p-2) (which = 1: inverse, 0 -> 0) and z^((p-5)/8) (which = 0), as ONE
// compact, table-driven routine so that a kernel carries a single copy of the chain instead of ~20
// inlined multiplies (instruction-cache footprint; the chain is 254 S + 11 M / 251 S + 11 M exactly as
// in the reference).  Program step: acc = acc^(2^n) * S[m] (m = 4: no multiply), then S[st] = acc
// (st = 4: no store).  Control flow depends only on these public constants.
//                                                           [reference: fld_inv fld.c:579-645, fld_pow2523 fld.c:658-709]
EDG_NOINLINE void fe_pow_chain(fe *out, const fe *zin, int which) {
    // program packed into immediates (no table in local memory):     step: 0   1   2    3   4    5    6    7    8     9    10   11
    //   squarings n                                                         1   2   0    1   5   10   20   10   50   100   50   (5|2)
    //   multiplier slot m (4 = none)                                         4   0   1    2   2    2    3    2    2     3    2   (1|0)
    //   store slot st (4 = none)                                             1   2   1    2   2    3    4    2    3     4    4    4
    const u64 PROG_N_LO = 0x0a140a0501000201ULL, PROG_N_HI = 0x00326432ULL;
    const u64 PROG_M = 0x023223222104ULL;          // 4 bits per step, step 0 in the low nibble
    const u64 PROG_S = 0x444324322121ULL;
    fe s0, s1, s2, s3, acc;
    fe_copy(s0, *zin);
    fe_copy(acc, *zin);
    fe_copy(s1, *zin); fe_copy(s2, *zin); fe_copy(s3, *zin);
#pragma unroll 1
    for (int step = 0; step < 12; step++) {
        int n = (int)(((step < 8 ? PROG_N_LO >> (8 * step) : PROG_N_HI >> (8 * (step - 8)))) & 0xff);
        int m = (int)((PROG_M >> (4 * step)) & 0xf);
        const int st = (int)((PROG_S >> (4 * step)) & 0xf);
        if (step == 11) { n = which ? 5 : 2; m = which ? 1 : 0; }   // * z11 (inverse) or * z (pow2523)
#pragma unroll 1
        for (int i = 0; i < n; i++) fe_sq(acc, acc);
        if (m != 4) {
            fe t;
            if (m == 0) fe_copy(t, s0); else if (m == 1) fe_copy(t, s1); else if (m == 2) fe_copy(t, s2); else fe_copy(t, s3);
            fe_mul(acc, acc, t);
        }
        if (st == 1) fe_copy(s1, acc); else if (st == 2) fe_copy(s2, acc); else if (st == 3) fe_copy(s3, acc);
    }
    fe_copy(*out, acc);
}

// r = z^(p-2) = z^-1 (0 -> 0).  254 S + 11 M.                      [reference: fld_inv, fld.c:579]
EDG_HD void fe_inv(fe &r, const fe &z) { fe_pow_chain(&r, &z, 1); }

// r = z^((p-5)/8) = z^(2^252 - 3).  251 S + 11 M.                  [reference: fld_pow2523, fld.c:658]
EDG_HD void fe_pow2523(fe &r, const fe &z) { fe_pow_chain(&r, &z, 0); }

}  // namespace edg
