// Half-size scalars for verification (public data, variable time).
//
// The reference checks  encode(S*B - t*A) == R-bytes  (ed25519-sha512.c:174-180) by a full-length
// double-scalar multiplication (ed_dual_scale, ed.c:455-507: ~252 doublings).  The same predicate is
//      R-bytes is the canonical encoding of a curve point R'   and   S*B - t*A - R' = O,
// and for any ODD integer rho the point D = S*B - t*A - R' is O  iff  rho*D = O: the curve group has
// order 8 L, so D = D_L + D_T with D_L of order 1 or L and D_T of order 1, 2, 4 or 8; an odd
// 0 < |rho| < L kills neither part.  With tau = rho * t (mod 8 L) — modulo the FULL group order, so the
// identity also holds when A carries a torsion component —
//      rho*D = (rho S mod L)*B - tau*A - rho*R'.
// The vectors (rho, tau) with tau = rho t (mod 8L) form a lattice of determinant 8L ~ 2^255.6; the
// extended Euclidean algorithm on (8L, t), stopped when the remainder drops below 2^128, yields
// |rho| < 2^128 and 0 <= tau < 2^128.  The double-scalar multiplication then needs ~130 doublings
// instead of 252 (the base-point scalar rho S stays full length, but B is fixed: its upper half uses a
// second window table for 2^128 B).  This is the "TODO: batch verify"-free way to halve the work: every
// signature is still decided on its own, exactly.
//
// If the Euclidean vector has an even rho the next vector of the sequence (always odd then) is used;
// in the (never observed, ~2^-30) case that it would not fit, (rho, tau) = (1, t) — the plain
// full-length computation — is the fallback, so the result never depends on luck.
#pragma once
#include "fe.cuh"

namespace edg {

#define EDG_SC_8L_INIT {0xe7ae9f68u, 0xc09318d2u, 0x17bce6b2u, 0xa6f7cef5u, 0x00000000u, 0x00000000u, 0x00000000u, 0x80000000u}

EDG_HD int hg_clz(u32 x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// number of significant bits of an 8-word value
EDG_HD int hg_bitlen8(const u32 x[8]) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) n = x[i] ? 32 * i + 32 - hg_clz(x[i]) : n;
    return n;
}

// (x[h], x[h-1]) as one 64-bit value, 3 <= h <= 7 (no dynamically indexed registers)
EDG_HD u64 hg_top2(const u32 x[8], int h) {
    u32 a, b;
    if (h == 7) { a = x[7]; b = x[6]; }
    else if (h == 6) { a = x[6]; b = x[5]; }
    else if (h == 5) { a = x[5]; b = x[4]; }
    else if (h == 4) { a = x[4]; b = x[3]; }
    else { a = x[3]; b = x[2]; }
    return ((u64)a << 32) | b;
}

// u -= q x  on the remainders,  mu += q xm  on the cofactor magnitudes (q x <= ru is the caller's business)
EDG_HD void hg_step(u32 ru[8], u32 mu[5], const u32 x[8], const u32 xm[5], u32 q) {
    u64 carry = 0;
    u32 borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u64 p = mulw(q, x[i]) + carry;
        carry = p >> 32;
        const u64 d = (u64)ru[i] - (u32)p - borrow;
        ru[i] = (u32)d;
        borrow = (u32)(d >> 63);
    }
    carry = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const u64 p = mulw(q, xm[i]) + mu[i] + carry;
        mu[i] = (u32)p;
        carry = p >> 32;
    }
}

// (rho, tau): tau = rho * t (mod 8L), rho odd, rho = (rho_neg ? -1 : 1) * rho_abs.  t < L.
// Normal case: rho_abs < 2^160 (typically < 2^129), tau < 2^128.  Fallback: rho = 1, tau = t.
EDG_HD void half_gcd(u32 rho_abs[8], u32 &rho_neg, u32 tau[8], const u32 t[8]) {
    const u32 N8L[8] = EDG_SC_8L_INIT;
    // lattice vectors u = (-+mu, ru), v = (+-mv, rv) with ru >= rv >= 0 and mu rv + mv ru = 8L throughout
    u32 ru[8], rv[8], mu[5], mv[5];
#pragma unroll
    for (int i = 0; i < 8; i++) { ru[i] = N8L[i]; rv[i] = t[i]; }
#pragma unroll
    for (int i = 0; i < 5; i++) { mu[i] = 0; mv[i] = 0; }
    mv[0] = 1;
    u32 sv = 0;                                            // sign of v's first coordinate (1 = negative)
    bool fallback = false;
#pragma unroll 1
    for (;;) {
        const bool big = (rv[4] | rv[5] | rv[6] | rv[7]) != 0;      // rv >= 2^128: keep reducing
        if (!big) {
            if (mv[0] & 1u) break;                                   // odd rho: done
            if (rv[3] == 0) { fallback = true; break; }              // next vector could exceed 160 bits (rv < 2^96)
        }
        // quotient estimate from the leading 32 bits of ru and the aligned bits of rv: 1 <= q <= floor(ru / rv).
        // Here ru >= rv >= 2^96, so the leading word of ru is word 3 or higher.
        const int h = ru[7] ? 7 : (ru[6] ? 6 : (ru[5] ? 5 : (ru[4] ? 4 : 3)));
        const u64 u2 = hg_top2(ru, h);
        const int c = hg_clz((u32)(u2 >> 32));
        const u32 U = (u32)((u2 << c) >> 32);
        const u32 V = (u32)((hg_top2(rv, h) << c) >> 32);
        if (V != 0) {
            u32 q = (V == 0xffffffffu) ? 1u : U / (V + 1u);
            q = q ? q : 1u;                                          // ru >= rv
            hg_step(ru, mu, rv, mv, q);
        } else {
            // rare (a partial quotient of 32 bits or more): work with v shifted left by whole words, as long as
            // the shifted remainder keeps at least 33 bits of distance to ru (then 2^32 x <= ru still holds)
            u32 x[8], xm[5];
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = rv[i];
#pragma unroll
            for (int i = 0; i < 5; i++) xm[i] = mv[i];
            const int lu = hg_bitlen8(ru);
#pragma unroll 1
            while (hg_bitlen8(x) + 33 <= lu) {
#pragma unroll
                for (int i = 7; i > 0; i--) x[i] = x[i - 1];
                x[0] = 0;
#pragma unroll
                for (int i = 4; i > 0; i--) xm[i] = xm[i - 1];
                xm[0] = 0;
            }
            const u32 V2 = (u32)((hg_top2(x, h) << c) >> 32);
            u32 q = (V2 == 0xffffffffu) ? 1u : U / (V2 + 1u);
            q = q ? q : 1u;
            hg_step(ru, mu, x, xm, q);
        }
        // keep ru >= rv
        u32 borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u64 d = (u64)ru[i] - rv[i] - borrow;
            borrow = (u32)(d >> 63);
        }
        if (borrow) {
#pragma unroll
            for (int i = 0; i < 8; i++) { const u32 y = ru[i]; ru[i] = rv[i]; rv[i] = y; }
#pragma unroll
            for (int i = 0; i < 5; i++) { const u32 y = mu[i]; mu[i] = mv[i]; mv[i] = y; }
            sv ^= 1u;
        }
    }
    if (fallback) {
#pragma unroll
        for (int i = 0; i < 8; i++) { rho_abs[i] = (i == 0) ? 1u : 0u; tau[i] = t[i]; }
        rho_neg = 0;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) { rho_abs[i] = (i < 5) ? mv[i] : 0u; tau[i] = rv[i]; }
        rho_neg = sv;
    }
}

// Signed radix-16 digits over exactly nwin windows, in place: x += 0x88..8 (nwin nibbles); needs x < 2^(4 nwin - 2).
// digit_j = nibble_j - 8 in [-8, 7], sum digit_j 16^j = the original x.
EDG_HD void sc_recode_radix16_n(u32 x[8], int nwin) {
    u32 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int nib = nwin - 8 * i;                        // nibbles of this word that belong to the recoding
        const u32 off = nib >= 8 ? 0x88888888u : (nib <= 0 ? 0u : (0x88888888u & ((1u << (4 * nib)) - 1u)));
        const u64 s = (u64)x[i] + off + carry;
        x[i] = (u32)s;
        carry = (u32)(s >> 32);
    }
}

}  // namespace edg
