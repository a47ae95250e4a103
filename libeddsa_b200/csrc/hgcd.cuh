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
// If the Euclidean vector has an even rho the next vector of the sequence (always odd then) is used.  That vector can be
// long: the fallback (rho, tau) = (1, t) — the plain full-length computation, ~1.7x the work — is taken exactly when the
// short vector has an even rho AND its remainder is below 2^96 (the next cofactor could then exceed 160 bits).  For
// uniformly random t (what a SHA-512 challenge is) that needs a partial quotient above 2^32 at the crossing of 2^128: probability
// ~2^-33 per signature, never seen in 3 x 10^5 random challenges (tests/test_host_sim.py::test_half_gcd).  For STRUCTURED t
// (small multiples or neighbours of L / 2^k, inverses of small numbers mod 8L ...) it is common — 28 % of a structured
// fuzz set — but t is a hash output, so nobody can steer a signature there; either way the result never depends on luck:
// the fallback decides the same equation.  The records of such signatures sort to the front of a pass (64 windows).
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

// bits [s, s + 32) of an 8-word value, 0 <= s <= 224 (no dynamically indexed registers: selection chain)
EDG_HD u32 hg_extract32(const u32 x[8], int s) {
    const int w = s >> 5, sh = s & 31;
    u32 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        lo = (w == i) ? x[i] : lo;
        if (i + 1 < 8) hi = (w == i) ? x[i + 1] : hi;
    }
    return (u32)(((((u64)hi) << 32) | lo) >> sh);
}

// out = a x - b y  modulo 2^(32 W): two's complement W-word values, small non-negative multipliers.
// Device form: a x = E + 2^32 O with E the products of the even words and O those of the odd words — the 64-bit
// products inside E (and inside O) do not overlap, so all of them are independent wide multiplies without an
// addend; then three carry chains (E + O for both terms, and the difference).  The portable form below it (host
// build of the tests) computes the same value word by word.
#if defined(__CUDA_ARCH__)
template <int W> __device__ __forceinline__ void hg_lin(u32 *out, const u32 *x, u32 a, const u32 *y, u32 b);

// t[0..W) = low W words of m * v   (W = 8 or 6)
template <int W> __device__ __forceinline__ void hg_mul1(u32 *t, const u32 *v, u32 m) {
    u32 e[W], o[W];                                        // o[k] = word k of 2^32 O (o[0] = 0)
#pragma unroll
    for (int i = 0; i < W; i += 2) { const u64 p = mulw(m, v[i]); e[i] = (u32)p; e[i + 1] = (u32)(p >> 32); }
    o[0] = 0;
#pragma unroll
    for (int i = 1; i < W; i += 2) {
        if (i + 1 < W) { const u64 p = mulw(m, v[i]); o[i] = (u32)p; o[i + 1] = (u32)(p >> 32); }
        else o[i] = m * v[i];
    }
    t[0] = e[0];
    if (W == 8)
        asm("add.cc.u32 %0, %7, %14; addc.cc.u32 %1, %8, %15; addc.cc.u32 %2, %9, %16; addc.cc.u32 %3, %10, %17; "
            "addc.cc.u32 %4, %11, %18; addc.cc.u32 %5, %12, %19; addc.u32 %6, %13, %20;"
            : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]),
              "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
    else
        asm("add.cc.u32 %0, %5, %10; addc.cc.u32 %1, %6, %11; addc.cc.u32 %2, %7, %12; addc.cc.u32 %3, %8, %13; addc.u32 %4, %9, %14;"
            : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5])
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]));
}

template <> __device__ __forceinline__ void hg_lin<8>(u32 *out, const u32 *x, u32 a, const u32 *y, u32 b) {
    u32 p[8], q[8];
    hg_mul1<8>(p, x, a);
    hg_mul1<8>(q, y, b);
    asm("sub.cc.u32 %0, %8, %16; subc.cc.u32 %1, %9, %17; subc.cc.u32 %2, %10, %18; subc.cc.u32 %3, %11, %19; "
        "subc.cc.u32 %4, %12, %20; subc.cc.u32 %5, %13, %21; subc.cc.u32 %6, %14, %22; subc.u32 %7, %15, %23;"
        : "=r"(out[0]), "=r"(out[1]), "=r"(out[2]), "=r"(out[3]), "=r"(out[4]), "=r"(out[5]), "=r"(out[6]), "=r"(out[7])
        : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]),
          "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]));
}
template <> __device__ __forceinline__ void hg_lin<6>(u32 *out, const u32 *x, u32 a, const u32 *y, u32 b) {
    u32 p[6], q[6];
    hg_mul1<6>(p, x, a);
    hg_mul1<6>(q, y, b);
    asm("sub.cc.u32 %0, %6, %12; subc.cc.u32 %1, %7, %13; subc.cc.u32 %2, %8, %14; subc.cc.u32 %3, %9, %15; "
        "subc.cc.u32 %4, %10, %16; subc.u32 %5, %11, %17;"
        : "=r"(out[0]), "=r"(out[1]), "=r"(out[2]), "=r"(out[3]), "=r"(out[4]), "=r"(out[5])
        : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]));
}

// x = -x over W words (W = 8 or 6): one borrow chain from zero
template <int W> __device__ __forceinline__ void hg_neg(u32 *x) {
    if (W == 8)
        asm("sub.cc.u32 %0, 0, %0; subc.cc.u32 %1, 0, %1; subc.cc.u32 %2, 0, %2; subc.cc.u32 %3, 0, %3; "
            "subc.cc.u32 %4, 0, %4; subc.cc.u32 %5, 0, %5; subc.cc.u32 %6, 0, %6; subc.u32 %7, 0, %7;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]));
    else
        asm("sub.cc.u32 %0, 0, %0; subc.cc.u32 %1, 0, %1; subc.cc.u32 %2, 0, %2; subc.cc.u32 %3, 0, %3; subc.cc.u32 %4, 0, %4; subc.u32 %5, 0, %5;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]));
}

// 1 if x < y (8-word unsigned): the borrow out of x - y
__device__ __forceinline__ u32 hg_less8(const u32 x[8], const u32 y[8]) {
    u32 d, bw;
    asm("sub.cc.u32 %0, %2, %10; subc.cc.u32 %0, %3, %11; subc.cc.u32 %0, %4, %12; subc.cc.u32 %0, %5, %13; "
        "subc.cc.u32 %0, %6, %14; subc.cc.u32 %0, %7, %15; subc.cc.u32 %0, %8, %16; subc.cc.u32 %0, %9, %17; subc.u32 %1, 0, 0;"
        : "=&r"(d), "=r"(bw)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
    return bw & 1u;
}
#else
template <int W>
EDG_HD void hg_lin(u32 *out, const u32 *x, u32 a, const u32 *y, u32 b) {
    u64 ca = 0, cb = 0;
    u32 borrow = 0;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const u64 pa = mulw(a, x[i]) + ca;
        const u64 pb = mulw(b, y[i]) + cb;
        ca = pa >> 32;
        cb = pb >> 32;
        const u64 d = (u64)(u32)pa - (u32)pb - borrow;
        out[i] = (u32)d;
        borrow = (u32)(d >> 63);
    }
}

template <int W>
EDG_HD void hg_neg(u32 *x) {
    u32 carry = 1;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const u64 t = (u64)(~x[i]) + carry;
        x[i] = (u32)t;
        carry = (u32)(t >> 32);
    }
}

// 1 if x < y (8-word unsigned)
EDG_HD u32 hg_less8(const u32 x[8], const u32 y[8]) {
    u32 borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u64 d = (u64)x[i] - y[i] - borrow;
        borrow = (u32)(d >> 63);
    }
    return borrow;
}
#endif

// (rho, tau): tau = rho * t (mod 8L), rho odd, rho = (rho_neg ? -1 : 1) * rho_abs.  t < L.
// Normal case: rho_abs < 2^160 (typically < 2^129), tau < 2^128.  Fallback: rho = 1, tau = t.
//
// Lattice basis u = (bu, ru), v = (bv, rv): remainders ru >= rv >= 0 as 8-word unsigned values, cofactors bu, bv
// as 6-word two's complement values; every step is unimodular, so u and v always generate
// {(b, r) : r = b t mod 8L} — the result's validity never depends on the quotients being exact.
//   * Lehmer step: run Euclid on the leading 32 bits (x, y) of (ru, rv) while the accumulated cofactors stay
//     below 2^16 and the remainder stays above the 2^128 target, then apply the 2 x 2 matrix to the long numbers
//     at once (signs fixed up afterwards if the approximation went one step too far);
//   * single step with an under-estimated quotient when the leading words allow no Lehmer step (large partial
//     quotient, or the last step across 2^128).
EDG_HD void half_gcd(u32 rho_abs[8], u32 &rho_neg, u32 tau[8], const u32 t[8], bool force_fallback = false) {
    const u32 N8L[8] = EDG_SC_8L_INIT;
    u32 ru[8], rv[8], bu[6], bv[6];
#pragma unroll
    for (int i = 0; i < 8; i++) { ru[i] = N8L[i]; rv[i] = t[i]; }
#pragma unroll
    for (int i = 0; i < 6; i++) { bu[i] = 0; bv[i] = 0; }
    bv[0] = 1;
    bool fallback = force_fallback;                        // (test hook: exercise the full-length path on any input)
    // Every pass strictly decreases max(ru, rv) (a Lehmer pass by at least one bit), so ~130 passes suffice; the
    // cap is a belt-and-braces bound so that no input can ever keep a GPU thread spinning.
#pragma unroll 1
    for (int pass = 0; !fallback; pass++) {
        if (pass >= 1024) { fallback = true; break; }
        const bool big = (rv[4] | rv[5] | rv[6] | rv[7]) != 0;      // rv >= 2^128: keep reducing
        if (!big) {
            if (bv[0] & 1u) break;                                   // odd rho: done
            if (rv[3] == 0) { fallback = true; break; }              // next vector could exceed 160 bits (rv < 2^96)
        }
        // leading 32 bits of ru and the aligned bits of rv (ru >= rv >= 2^96 here)
        const int lu = hg_bitlen8(ru);
        const int s = lu - 32;
        const u32 X = hg_extract32(ru, s), Y = hg_extract32(rv, s);
        u32 a = 1, b = 0, c = 0, d = 1, x = X, y = Y;
        int k = 0;
        if (big) {
            // Euclid on (x, y):  x_k = (-1)^k (a X - b Y),  y_k = (-1)^(k+1) (c X - d Y)
            const u32 ylim = s < 112 ? (1u << (128 - s)) : 0x10000u; // stay above 2^128 / keep 16 significant bits
#pragma unroll 1
            while (y >= ylim) {
                const u32 q = x / y, r = x - q * y;
                if (r < ylim || q >= 0x10000u) break;
                const u32 c2 = a + q * c, d2 = b + q * d;            // a, b, c, d, q < 2^16: no overflow
                if (c2 >= 0x10000u || d2 >= 0x10000u) break;
                a = c; b = d; c = c2; d = d2;
                x = y; y = r;
                k++;
            }
        }
        if (k > 0) {
            // u' = +-(a u - b v), v' = +-(d v - c u): the sign that makes the remainder non-negative (for an odd
            // number of steps both come out negative; an overshooting approximation can flip one more)
            u32 nu[8], nv[8], mu[6], mv[6];
            hg_lin<8>(nu, ru, a, rv, b);
            hg_lin<6>(mu, bu, a, bv, b);
            hg_lin<8>(nv, rv, d, ru, c);
            hg_lin<6>(mv, bv, d, bu, c);
            if (nu[7] >> 31) { hg_neg<8>(nu); hg_neg<6>(mu); }
            if (nv[7] >> 31) { hg_neg<8>(nv); hg_neg<6>(mv); }
#pragma unroll
            for (int i = 0; i < 8; i++) { ru[i] = nu[i]; rv[i] = nv[i]; }
#pragma unroll
            for (int i = 0; i < 6; i++) { bu[i] = mu[i]; bv[i] = mv[i]; }
        } else {
            // one step u -= q 2^(32 j) v with 1 <= q 2^(32 j) <= floor(ru / rv): v is shifted left by whole words while
            // it keeps at least 33 bits of distance to ru (then 2^32 x <= ru still holds)
            u32 xr[8], xb[6];
#pragma unroll
            for (int i = 0; i < 8; i++) xr[i] = rv[i];
#pragma unroll
            for (int i = 0; i < 6; i++) xb[i] = bv[i];
            u32 V = Y;
            if (V == 0) {
#pragma unroll 1
                while (hg_bitlen8(xr) + 33 <= lu) {
#pragma unroll
                    for (int i = 7; i > 0; i--) xr[i] = xr[i - 1];
                    xr[0] = 0;
#pragma unroll
                    for (int i = 5; i > 0; i--) xb[i] = xb[i - 1];
                    xb[0] = 0;
                }
                V = hg_extract32(xr, s);
            }
            u32 q = (V == 0xffffffffu) ? 1u : X / (V + 1u);
            q = q ? q : 1u;                                          // ru >= rv
            hg_lin<8>(ru, ru, 1u, xr, q);
            hg_lin<6>(bu, bu, 1u, xb, q);
        }
        if (hg_less8(ru, rv)) {
#pragma unroll
            for (int i = 0; i < 8; i++) { const u32 z = ru[i]; ru[i] = rv[i]; rv[i] = z; }
#pragma unroll
            for (int i = 0; i < 6; i++) { const u32 z = bu[i]; bu[i] = bv[i]; bv[i] = z; }
        }
    }
    if (fallback) {
#pragma unroll
        for (int i = 0; i < 8; i++) { rho_abs[i] = (i == 0) ? 1u : 0u; tau[i] = t[i]; }
        rho_neg = 0;
    } else {
        rho_neg = bv[5] >> 31;
        if (rho_neg) hg_neg<6>(bv);
#pragma unroll
        for (int i = 0; i < 8; i++) { rho_abs[i] = (i < 6) ? bv[i] : 0u; tau[i] = rv[i]; }
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
