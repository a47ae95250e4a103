// Twisted Edwards group  -x^2 + y^2 = 1 + d x^2 y^2  over GF(2^255-19), one point per thread.
//
// Role of the reference's lib/ed.c: struct ed / struct pced (ed.h:13-18, ed.c:30-34), ed_import
// (ed.c:100-149), ed_export (:155-169), ed_add/ed_double/ed_sub (:175-273), ed_add_pc/ed_sub_pc
// (:282-335), scale16 (:346-391).  Same complete a = -1 extended-coordinate law (Hisil et al.), but
//   * unsigned lazy limbs: every formula below is annotated with the limb magnitude alpha (limb <=
//     alpha * 2^26 on even limbs, half that on odd limbs) of each operand; fe_mul(a, b) needs
//     alpha_a * alpha_b <= 32 and alpha_b <= 3.3 (b gets the x19), fe_sq needs alpha <= 3.2;
//   * the doubling skips T when the next operation is another doubling (4S+3M instead of 4S+5M);
//   * precomputed points keep 2dT (not T) so an addition is 8M / 7M (mixed) — the reference spends 9M;
//   * constant-time table selection works on register words with LOP3 masks (sign / genpub / x25519_base).
#pragma once
#include "fe.cuh"

namespace edg {

struct ge_p3 { fe X, Y, Z, T; };                 // extended (X:Y:Z:T), XY = ZT          [struct ed, ed.h:13]
struct ge_pre { fe ypx, ymx, xy2d; };            // affine precomputed: y+x, y-x, 2dxy   [struct pced, ed.c:30]
struct ge_cached { fe ypx, ymx, z2, t2d; };      // projective cached: Y+X, Y-X, 2Z, 2dT

#define EDG_FE_D    {{0x35978a3u, 0x0d37284u, 0x3156ebdu, 0x06a0a0eu, 0x001c029u, 0x179e898u, 0x3a03cbbu, 0x1ce7198u, 0x2e2b6ffu, 0x1480db3u}}
#define EDG_FE_2D   {{0x2b2f159u, 0x1a6e509u, 0x22add7au, 0x0d4141du, 0x0038052u, 0x0f3d130u, 0x3407977u, 0x19ce331u, 0x1c56dffu, 0x0901b67u}}
#define EDG_FE_SQRTM1 {{0x20ea0b0u, 0x186c9d2u, 0x08f189du, 0x035697fu, 0x0bd0c60u, 0x1fbd7a7u, 0x2804c9eu, 0x1e16569u, 0x004fc1du, 0x0ae0c92u}}

EDG_HD void ge_identity(ge_p3 &p) {
    fe_set_u32(p.X, 0); fe_set_u32(p.Y, 1); fe_set_u32(p.Z, 1); fe_set_u32(p.T, 0);
}

// r = 2p.  4S + 3M (+1M when need_t).  Inputs tight (alpha 1).                  [ed_double, ed.c:211]
EDG_HD void ge_dbl(ge_p3 &r, const ge_p3 &p, bool need_t) {
    fe a, b, c, s, e, f, g, h;
    fe_sq(a, p.X);                      // A = X^2            (1)
    fe_sq(b, p.Y);                      // B = Y^2            (1)
    fe_sq(c, p.Z);
    fe_dbl(c, c);                       // C = 2 Z^2          (2)
    fe_add(s, p.X, p.Y);                //                    (2)
    fe_sq(s, s);                        // (X+Y)^2            (1)
    fe_add(h, a, b);                    // H = A + B          (2)
    fe_sub(e, h, s);                    // E = H - (X+Y)^2    (4)
    fe_carry(e, e);                     //                    (1)  keeps E usable as the x19 operand
    fe_sub(g, a, b);                    // G = A - B          (3)
    fe_add(f, c, g);                    // F = C + G          (5)
    fe_mul(r.X, f, e);                  // X3 = E F
    fe_mul(r.Y, g, h);                  // Y3 = G H
    fe_mul(r.Z, f, g);                  // Z3 = F G
    if (need_t) fe_mul(r.T, h, e);      // T3 = E H
}

// r = p + q, q affine precomputed.  7M (6M when !need_t).                       [ed_add_pc, ed.c:282]
EDG_HD void ge_madd(ge_p3 &r, const ge_p3 &p, const ge_pre &q, bool need_t) {
    fe a, b, c, d, e, f, g, h;
    fe_sub(a, p.Y, p.X);                // (3)
    fe_mul(a, a, q.ymx);                // A = (Y1-X1)(y2-x2)
    fe_add(b, p.Y, p.X);                // (2)
    fe_mul(b, b, q.ypx);                // B = (Y1+X1)(y2+x2)
    fe_mul(c, p.T, q.xy2d);             // C = T1 * 2d x2 y2     (xy2d alpha <= 2 after a lazy negate)
    fe_dbl(d, p.Z);                     // D = 2 Z1            (2)
    fe_sub(e, b, a);                    // E = B - A           (3)
    fe_sub(f, d, c);                    // F = D - C           (4)
    fe_add(g, d, c);                    // G = D + C           (3)
    fe_add(h, b, a);                    // H = B + A           (2)
    fe_mul(r.X, f, e);
    fe_mul(r.Y, g, h);
    fe_mul(r.Z, f, g);
    if (need_t) fe_mul(r.T, e, h);
}

// r = p + q, q projective cached.  8M (7M when !need_t).                        [ed_add, ed.c:175]
EDG_HD void ge_add_cached(ge_p3 &r, const ge_p3 &p, const ge_cached &q, bool need_t) {
    fe a, b, c, d, e, f, g, h;
    fe_sub(a, p.Y, p.X);                // (3)
    fe_mul(a, a, q.ymx);                // q.ymx alpha <= 3
    fe_add(b, p.Y, p.X);                // (2)
    fe_mul(b, b, q.ypx);                // q.ypx alpha <= 3 (2 normally, 3 when it is a swapped ymx)
    fe_mul(c, p.T, q.t2d);              // q.t2d alpha <= 2
    fe_mul(d, p.Z, q.z2);               // q.z2 alpha 2
    fe_sub(e, b, a);                    // (3)
    fe_sub(f, d, c);                    // (3)
    fe_add(g, d, c);                    // (2)
    fe_add(h, b, a);                    // (2)
    fe_mul(r.X, e, f);
    fe_mul(r.Y, g, h);
    fe_mul(r.Z, f, g);
    if (need_t) fe_mul(r.T, e, h);
}

// cached form of p (tight coordinates): Y+X (2), Y-X (3), 2Z (2), 2dT (1).      [ed_precompute, ed.c:436]
EDG_HD void ge_to_cached(ge_cached &c, const ge_p3 &p) {
    const fe d2 = EDG_FE_2D;
    fe_add(c.ypx, p.Y, p.X);
    fe_sub(c.ymx, p.Y, p.X);
    fe_dbl(c.z2, p.Z);
    fe_mul(c.t2d, p.T, d2);
}

// cached point <-> 40 contiguous words of per-thread scratch (16-byte aligned): 10 x 128-bit accesses
EDG_HD void ge_cached_store(u32 *dst, const ge_cached &c) {
    u32 w[40];
#pragma unroll
    for (int i = 0; i < 10; i++) { w[i] = c.ypx.v[i]; w[10 + i] = c.ymx.v[i]; w[20 + i] = c.z2.v[i]; w[30 + i] = c.t2d.v[i]; }
#if defined(__CUDA_ARCH__)
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < 10; i++) d4[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
#else
    for (int i = 0; i < 40; i++) dst[i] = w[i];
#endif
}

EDG_HD void ge_cached_load(ge_cached &c, const u32 *src) {
    u32 w[40];
#if defined(__CUDA_ARCH__)
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int i = 0; i < 10; i++) { const uint4 v = s4[i]; w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
#else
    for (int i = 0; i < 40; i++) w[i] = src[i];
#endif
#pragma unroll
    for (int i = 0; i < 10; i++) { c.ypx.v[i] = w[i]; c.ymx.v[i] = w[10 + i]; c.z2.v[i] = w[20 + i]; c.t2d.v[i] = w[30 + i]; }
}

// q = neg ? -q : q for a cached point, neg an all-ones/zero mask (public data in verify, but branch-free anyway)
EDG_HD void ge_cached_cneg(ge_cached &q, u32 neg) {
    fe_cswap(q.ypx, q.ymx, neg);
    fe nt;
    fe_neg(nt, q.t2d);                  // (2)
    fe_select(q.t2d, q.t2d, nt, neg);
}

// q = neg ? -q : q for an affine precomputed point (canonical limbs).           [scale16, ed.c:384-390]
EDG_HD void ge_pre_cneg(ge_pre &q, u32 neg) {
    fe_cswap(q.ypx, q.ymx, neg);
    fe nt;
    fe_neg(nt, q.xy2d);                 // (2)
    fe_select(q.xy2d, q.xy2d, nt, neg);
}

// Constant-time lookup of digit * (row point), digit in [-8, 7]; row = 8 entries x 30 words holding
// 1P..8P.  Every entry is read and folded in with a mask; no branch or address depends on digit.
//                                                                                 [scale16, ed.c:346-391]
EDG_HD void ge_pre_select_ct(ge_pre &t, const u32 *row, int digit) {
    const u32 neg = ct_mask((u32)(digit >> 31));         // all-ones if digit < 0
    const u32 absd = ((u32)digit ^ neg) - neg;           // 0..8
    u32 w[30];
#pragma unroll
    for (int i = 0; i < 30; i++) w[i] = (i == 0 || i == 10) ? 1u : 0u;    // neutral element (1, 1, 0)
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u32 m = ct_mask(0u - ((((absd ^ (u32)(k + 1)) - 1u) >> 31)));   // all-ones iff absd == k+1
#if defined(__CUDA_ARCH__)
        // entry k starts at word 30k: 8-byte aligned -> 15 x 64-bit broadcast loads (same address in every lane)
        const uint2 *e2 = reinterpret_cast<const uint2 *>(row + 30 * k);
#pragma unroll
        for (int i = 0; i < 15; i++) {
            const uint2 v = e2[i];
            w[2 * i] ^= (w[2 * i] ^ v.x) & m;
            w[2 * i + 1] ^= (w[2 * i + 1] ^ v.y) & m;
        }
#else
#pragma unroll
        for (int i = 0; i < 30; i++) w[i] ^= (w[i] ^ row[30 * k + i]) & m;
#endif
    }
#pragma unroll
    for (int i = 0; i < 10; i++) { t.ypx.v[i] = w[i]; t.ymx.v[i] = w[10 + i]; t.xy2d.v[i] = w[20 + i]; }
    ge_pre_cneg(t, neg);
}

// Variable-time (public data) lookup for the verify kernel: entry |digit| of a 9-entry table
// (0 = neutral element), negated when digit < 0.
EDG_HD void ge_pre_load(ge_pre &t, const u32 *tbl, int digit) {
    const u32 neg = (u32)(digit >> 31);
    const u32 absd = ((u32)digit ^ neg) - neg;
    const u32 *e = tbl + 30 * absd;
#if defined(__CUDA_ARCH__)
    // 64-bit shared-memory loads: entries are 120 B apart, so the <= 9 distinct entries a warp touches
    // fall in distinct bank pairs (no conflicts); 128-bit loads would 2-way conflict.
    const uint2 *e2 = reinterpret_cast<const uint2 *>(e);
    u32 w[30];
#pragma unroll
    for (int i = 0; i < 15; i++) { const uint2 v = e2[i]; w[2 * i] = v.x; w[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 10; i++) { t.ypx.v[i] = w[i]; t.ymx.v[i] = w[10 + i]; t.xy2d.v[i] = w[20 + i]; }
#else
#pragma unroll
    for (int i = 0; i < 10; i++) { t.ypx.v[i] = e[i]; t.ymx.v[i] = e[10 + i]; t.xy2d.v[i] = e[20 + i]; }
#endif
    ge_pre_cneg(t, neg);
}

// Decompress 32 bytes (8 LE words) into an affine point (Z = 1).  Never fails, exactly like the
// reference: bit 255 is the sign, the low 255 bits are y and are NOT range-checked (y >= p is
// reduced), x = 0 with sign 1 stays 0.  Returns the all-ones mask if (x, y) is on the curve; for an
// off-curve y the reference goes on with x = sqrt(-1)*beta (garbage) — callers apply the SURVEY Q5
// policy.  If negate is set the result is -P (x and t negated), which is what verify needs.
//                                                                                 [ed_import, ed.c:100-149]
EDG_HD u32 ge_frombytes(ge_p3 &p, const u32 in[8], bool negate) {
    const fe d = EDG_FE_D, sqrtm1 = EDG_FE_SQRTM1;
    u32 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = in[i];
    const u32 sign = w[7] >> 31;
    w[7] &= 0x7fffffffu;                                 // ed.c:107-108
    fe one, yy, u, v, v3, v7, beta, chk, x, t;
    fe_set_u32(one, 1);
    fe_from_words(p.Y, w);                               // tight (limb0 < 2^26)
    fe_sq(yy, p.Y);
    fe_mul(v, yy, d);
    fe_sub(u, yy, one);                                  // u = y^2 - 1        (3)
    fe_add(v, v, one);                                   // v = d y^2 + 1      (1)
    fe_sq(v3, v);
    fe_mul(v3, v3, v);                                   // v^3
    fe_sq(v7, v3);
    fe_mul(v7, v7, v);                                   // v^7
    fe_mul(t, u, v7);
    fe_pow2523(t, t);                                    // (u v^7)^((p-5)/8)
    fe_mul(t, t, v3);
    fe_mul(beta, u, t);                                  // beta = u v^3 (u v^7)^((p-5)/8)    ed.c:121-131
    fe_sq(chk, beta);
    fe_mul(chk, chk, v);                                 // v beta^2
    fe_sub4(t, chk, u);
    const u32 is_root = 0u - fe_is_zero(t);              // v beta^2 == u                      ed.c:134-137
    fe_add(t, chk, u);
    const u32 is_jroot = 0u - fe_is_zero(t);             // v beta^2 == -u  <=>  v (j beta)^2 == u
    fe_mul(t, beta, sqrtm1);
    fe_select(x, t, beta, is_root);                      // x = is_root ? beta : j beta        ed.c:140-142
    fe_canon(x, x);
    u32 flip = 0u - ((sign ^ (x.v[0] & 1u)) & 1u);       // ed.c:143-144
    if (negate) flip = ~flip;
    fe_neg(t, x);                                        // (2), 0 stays == 0 mod p
    fe_select(p.X, x, t, flip);
    fe_carry(p.X, p.X);                                  // back to tight (the lazy negate has alpha 2)
    fe_mul(p.T, p.X, p.Y);
    fe_set_u32(p.Z, 1);
    return is_root | is_jroot;
}

// Affine (x, y) -> 32 bytes: canonical y, bit 255 = lsb of canonical x.         [ed_export tail, ed.c:160-168]
EDG_HD void ge_affine_tobytes(u32 out[8], const fe &x, const fe &y) {
    fe cx;
    fe_to_words(out, y);
    fe_canon(cx, x);
    out[7] |= (cx.v[0] & 1u) << 31;
}

// Extended point -> 32 bytes (one inversion, 254S + 13M).                       [ed_export, ed.c:155-169]
EDG_HD void ge_tobytes(u32 out[8], const ge_p3 &p) {
    fe zi, x, y;
    fe_inv(zi, p.Z);
    fe_mul(x, p.X, zi);
    fe_mul(y, p.Y, zi);
    ge_affine_tobytes(out, x, y);
}

}  // namespace edg
