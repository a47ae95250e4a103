// Scalars modulo the group order L = 2^252 + 27742317777372353535851937790883648493, one per thread.
//
// Role of the reference's lib/sc.c: sc_barrett (sc.c:79-158), sc_import (sc.c:191), sc_export /
// sc_reduce (sc.c:221,164), sc_mul + sc_add (sc.c:241, sc.h:53) and the signed radix-16 recoding
// that ed_scale_base derives from con_off (sc.c:40, ed.c:406-422).  Re-designed for 32-bit
// registers: eight saturated 32-bit words, Barrett (HAC 14.42) with b = 2^32, k = 8,
// mu = floor(2^512 / L); every loop has constant trip count, every select is a mask — no branch
// or address depends on the scalar (needed by the sign / genpub / x25519_base kernels).
// The reference's vartime JSF recoding (sc.c:272-324) is deliberately not reproduced: verify uses
// fixed signed 4-bit windows (uniform control flow across a warp; SURVEY.md §7 hard part 2).
#pragma once
#include "fe.cuh"

namespace edg {

#define EDG_SC_L_INIT  {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0x00000000u, 0x00000000u, 0x00000000u, 0x10000000u, 0x00000000u}
#define EDG_SC_MU_INIT {0x0a2c131bu, 0xed9ce5a3u, 0x086329a7u, 0x2106215du, 0xffffffebu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x0000000fu}

// r[0..na+nb) = a * b (schoolbook, operand scanning; all indices compile-time after unrolling)
template <int NA, int NB>
EDG_HD void sc_mul_words(u32 *r, const u32 *a, const u32 *b) {
#pragma unroll
    for (int i = 0; i < NA + NB; i++) r[i] = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) {
        u32 carry = 0;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            u64 t = mulw(a[i], b[j]) + r[i + j] + carry;
            r[i + j] = (u32)t;
            carry = (u32)(t >> 32);
        }
        r[i + NB] = carry;
    }
}

// x (16 words, < 2^512) mod L -> 8 words in [0, L).   [reference: sc_barrett, sc.c:79-158]
EDG_HD void sc_reduce512(u32 r[8], const u32 x[16]) {
    const u32 Lw[9] = EDG_SC_L_INIT;
    const u32 MU[9] = EDG_SC_MU_INIT;
    // q1 = floor(x / b^(k-1)) : 9 words ; q2 = q1 * mu ; q3 = floor(q2 / b^(k+1))
    u32 q2[18];
    sc_mul_words<9, 9>(q2, x + 7, MU);
    // r2 = (q3 * L) mod b^(k+1)  (low 9 words only)
    u32 r2[9];
#pragma unroll
    for (int i = 0; i < 9; i++) r2[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        u32 carry = 0;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            if (i + j < 9) {
                u64 t = mulw(q2[9 + i], Lw[j]) + r2[i + j] + carry;
                r2[i + j] = (u32)t;
                carry = (u32)(t >> 32);
            }
        }
    }
    // v = (x mod b^(k+1)) - r2  mod b^(k+1)
    u32 v[9];
    u32 borrow = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        u64 d = (u64)x[i] - r2[i] - borrow;
        v[i] = (u32)d;
        borrow = (u32)(d >> 63);
    }
    // HAC 14.42 step 4: at most two subtractions of L; done unconditionally with a mask
#pragma unroll
    for (int round = 0; round < 2; round++) {
        u32 t[9];
        borrow = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            u64 d = (u64)v[i] - Lw[i] - borrow;
            t[i] = (u32)d;
            borrow = (u32)(d >> 63);
        }
        const u32 keep = ct_mask(0u - borrow);   // all-ones if v < L (keep v)
#pragma unroll
        for (int i = 0; i < 9; i++) v[i] = t[i] ^ ((t[i] ^ v[i]) & keep);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = v[i];
}

// 32-byte value (8 words, any value < 2^256) mod L.  No range check.  [sc_import, sc.c:191]
EDG_HD void sc_reduce256(u32 r[8], const u32 x[8]) {
    u32 w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = x[i]; w[8 + i] = 0; }
    sc_reduce512(r, w);
}

// r = (a * b + c) mod L  for a, b, c < 2^256 with a*b + c < 2^512.     [sc_mul sc.c:241 + sc_add sc.h:53 + sc_export sc.c:221]
EDG_HD void sc_muladd(u32 r[8], const u32 a[8], const u32 b[8], const u32 c[8]) {
    u32 w[16];
    sc_mul_words<8, 8>(w, a, b);
    u32 carry = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        u64 t = (u64)w[i] + (i < 8 ? c[i] : 0u) + carry;
        w[i] = (u32)t;
        carry = (u32)(t >> 32);
    }
    sc_reduce512(r, w);
}

// Signed radix-16 digits of x in [0, L): e = x + 0x888...8 (no overflow since x < 2^253), nibble_j(e) - 8
// in [-8, 7] and sum (nibble_j - 8) 16^j = x.  The 64 nibbles are returned packed (8 per word); callers
// take (e[j>>3] >> 4*(j&7)) & 15 and subtract 8.        [reference: con_off trick, sc.c:40 + ed.c:406-422]
EDG_HD void sc_recode_radix16(u32 e[8], const u32 x[8]) {
    u32 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 t = (u64)x[i] + 0x88888888u + carry;
        e[i] = (u32)t;
        carry = (u32)(t >> 32);
    }
}

}  // namespace edg
