/*
 * oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the Ed25519 / X25519 hot path of
 * phlay/libeddsa v0.8, written from the published algorithms; it defines the expected bytes
 * and accept/reject decisions the CUDA path is compared with.
 *
 *   * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 *     may load this.  The product library (libeddsa_b200.so) never links, loads or calls it.
 *   * Pinned (tests/test_oracle.py): all 1024 rows of the reference's own KAT table
 *     test/x25519-table.h (tests/golden/x25519_kat.bin), RFC 8032 §7.1 vectors, and fixtures
 *     produced by the compiled reference itself (oracle/_ref/libeddsa_ref.so, built by
 *     oracle/Makefile from /root/reference/lib) for genpub / sign / verify / x25519_base,
 *     including the adversarial classes of SURVEY.md §0 (Q1–Q9).
 *   * Representation is deliberately different from both reference builds (5x51-bit unsigned
 *     limbs with unsigned __int128, binary double-and-add with the complete addition law, bitwise
 *     long division for mod L): the reference's comb / JSF / Barrett are *performance* choices
 *     whose outputs are mathematically determined for every on-curve input, so the oracle states
 *     the function they compute, not their schedule.  One documented policy: an off-curve public
 *     key makes verify return 0 (SURVEY.md Q5 — the reference returns schedule-dependent garbage
 *     that differs between its own 32/64-bit builds and can never compare equal to R).
 *
 * Each function cites the reference file:line whose behaviour it restates (paths relative to
 * /root/reference/lib).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

/* ------------------------------------------------------------------------------------------
 * SHA-512 (FIPS 180-4)                         restates sha512.c:83-210 (compress/init/add/final)
 * ------------------------------------------------------------------------------------------ */
static const u64 K512[80] = {
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

typedef struct {
    u64 h[8];
    u8 buf[128];
    size_t fill;
    u64 total; /* bytes */
} osha512;

static u64 ror64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }

static void osha_block(u64 h[8], const u8 *p)
{
    u64 w[16], s[8], t1, t2;
    int i;
    for (i = 0; i < 16; i++) {
        u64 v = 0;
        int j;
        for (j = 0; j < 8; j++) v = (v << 8) | p[8 * i + j];
        w[i] = v;
    }
    memcpy(s, h, sizeof s);
    for (i = 0; i < 80; i++) {
        if (i >= 16) {
            u64 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            w[i & 15] += (ror64(w15, 1) ^ ror64(w15, 8) ^ (w15 >> 7)) + w[(i + 9) & 15] +
                         (ror64(w2, 19) ^ ror64(w2, 61) ^ (w2 >> 6));
        }
        t1 = s[7] + (ror64(s[4], 14) ^ ror64(s[4], 18) ^ ror64(s[4], 41)) + ((s[4] & s[5]) ^ (~s[4] & s[6])) + K512[i] + w[i & 15];
        t2 = (ror64(s[0], 28) ^ ror64(s[0], 34) ^ ror64(s[0], 39)) + ((s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]));
        s[7] = s[6]; s[6] = s[5]; s[5] = s[4]; s[4] = s[3] + t1;
        s[3] = s[2]; s[2] = s[1]; s[1] = s[0]; s[0] = t1 + t2;
    }
    for (i = 0; i < 8; i++) h[i] += s[i];
}

static void osha_init(osha512 *c) /* sha512.c:127 */
{
    static const u64 iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                              0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    memcpy(c->h, iv, sizeof iv);
    c->fill = 0;
    c->total = 0;
}

static void osha_add(osha512 *c, const u8 *d, size_t n) /* sha512.c:143 */
{
    c->total += n;
    while (n > 0) {
        size_t take = 128 - c->fill;
        if (take > n) take = n;
        memcpy(c->buf + c->fill, d, take);
        c->fill += take; d += take; n -= take;
        if (c->fill == 128) { osha_block(c->h, c->buf); c->fill = 0; }
    }
}

static void osha_final(osha512 *c, u8 out[64]) /* sha512.c:175 (length = 128-bit big-endian bit count) */
{
    u64 bits_lo = c->total << 3, bits_hi = c->total >> 61;
    int i;
    c->buf[c->fill++] = 0x80;
    if (c->fill > 112) {
        memset(c->buf + c->fill, 0, 128 - c->fill);
        osha_block(c->h, c->buf);
        c->fill = 0;
    }
    memset(c->buf + c->fill, 0, 112 - c->fill);
    for (i = 0; i < 8; i++) {
        c->buf[112 + i] = (u8)(bits_hi >> (56 - 8 * i));
        c->buf[120 + i] = (u8)(bits_lo >> (56 - 8 * i));
    }
    osha_block(c->h, c->buf);
    for (i = 0; i < 64; i++) out[i] = (u8)(c->h[i >> 3] >> (56 - 8 * (i & 7)));
}

void oracle_sha512(u8 out[64], const u8 *data, size_t len)
{
    osha512 c;
    osha_init(&c);
    osha_add(&c, data, len);
    osha_final(&c, out);
}

/* ------------------------------------------------------------------------------------------
 * GF(2^255-19): 5 unsigned limbs of 51 bits.
 * ------------------------------------------------------------------------------------------ */
typedef struct { u64 v[5]; } ofe;
#define M51 ((1ULL << 51) - 1)

static const ofe OFE_ZERO = {{0, 0, 0, 0, 0}};
static const ofe OFE_ONE = {{1, 0, 0, 0, 0}};

static u64 load64(const u8 *p)
{
    u64 v = 0;
    int i;
    for (i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}

/* fld_import (fld.c:137 / :383): all 256 bits are used; bit 255 contributes 2^255 = +19 (Q6). */
static void ofe_frombytes(ofe *r, const u8 s[32])
{
    u64 a = load64(s), b = load64(s + 8), c = load64(s + 16), d = load64(s + 24);
    r->v[0] = a & M51;
    r->v[1] = ((a >> 51) | (b << 13)) & M51;
    r->v[2] = ((b >> 38) | (c << 26)) & M51;
    r->v[3] = ((c >> 25) | (d << 39)) & M51;
    r->v[4] = (d >> 12) & M51;
    r->v[0] += 19 * (d >> 63);
}

static void ofe_carry(ofe *r)
{
    int i, k;
    for (k = 0; k < 2; k++) {
        for (i = 0; i < 4; i++) { r->v[i + 1] += r->v[i] >> 51; r->v[i] &= M51; }
        r->v[0] += 19 * (r->v[4] >> 51); r->v[4] &= M51;
    }
}

/* fld_reduce + fld_export (fld.c:54,163 / :342,406): canonical value in [0,p), little endian. */
static void ofe_tobytes(u8 s[32], const ofe *a)
{
    ofe t = *a;
    u64 q, w[4];
    int i;
    ofe_carry(&t);
    ofe_carry(&t);
    /* q = 1 iff t >= p */
    q = (t.v[0] + 19) >> 51;
    for (i = 1; i < 5; i++) q = (t.v[i] + q) >> 51;
    t.v[0] += 19 * q;
    for (i = 0; i < 4; i++) { t.v[i + 1] += t.v[i] >> 51; t.v[i] &= M51; }
    t.v[4] &= M51;
    w[0] = t.v[0] | (t.v[1] << 51);
    w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
    w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
    w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
    for (i = 0; i < 32; i++) s[i] = (u8)(w[i >> 3] >> (8 * (i & 7)));
}

static void ofe_add(ofe *r, const ofe *a, const ofe *b) /* fld.h:94 */
{
    int i;
    for (i = 0; i < 5; i++) r->v[i] = a->v[i] + b->v[i];
}

/* r = a - b; adds 4p so limbs stay non-negative (inputs carried, < 2^52). fld.h:102 */
static void ofe_sub(ofe *r, const ofe *a, const ofe *b)
{
    r->v[0] = a->v[0] + 0x1fffffffffffb4ULL - b->v[0];
    r->v[1] = a->v[1] + 0x1ffffffffffffcULL - b->v[1];
    r->v[2] = a->v[2] + 0x1ffffffffffffcULL - b->v[2];
    r->v[3] = a->v[3] + 0x1ffffffffffffcULL - b->v[3];
    r->v[4] = a->v[4] + 0x1ffffffffffffcULL - b->v[4];
}

/* fld_mul (fld.c:210 / :448).  Inputs < 2^54 per limb, output carried (< 2^51 + small). */
static void ofe_mul(ofe *r, const ofe *a, const ofe *b)
{
    u128 t[5];
    u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3], a4 = a->v[4];
    u64 b0 = b->v[0], b1 = b->v[1], b2 = b->v[2], b3 = b->v[3], b4 = b->v[4];
    u64 b1_19 = 19 * b1, b2_19 = 19 * b2, b3_19 = 19 * b3, b4_19 = 19 * b4, c;
    int i;
    t[0] = (u128)a0 * b0 + (u128)a1 * b4_19 + (u128)a2 * b3_19 + (u128)a3 * b2_19 + (u128)a4 * b1_19;
    t[1] = (u128)a0 * b1 + (u128)a1 * b0 + (u128)a2 * b4_19 + (u128)a3 * b3_19 + (u128)a4 * b2_19;
    t[2] = (u128)a0 * b2 + (u128)a1 * b1 + (u128)a2 * b0 + (u128)a3 * b4_19 + (u128)a4 * b3_19;
    t[3] = (u128)a0 * b3 + (u128)a1 * b2 + (u128)a2 * b1 + (u128)a3 * b0 + (u128)a4 * b4_19;
    t[4] = (u128)a0 * b4 + (u128)a1 * b3 + (u128)a2 * b2 + (u128)a3 * b1 + (u128)a4 * b0;
    for (i = 0; i < 4; i++) { t[i + 1] += (u64)(t[i] >> 51); r->v[i] = (u64)t[i] & M51; }
    c = (u64)(t[4] >> 51);
    r->v[4] = (u64)t[4] & M51;
    r->v[0] += 19 * c;
    r->v[1] += r->v[0] >> 51;
    r->v[0] &= M51;
}

static void ofe_sq(ofe *r, const ofe *a) { ofe_mul(r, a, a); } /* fld_sq, fld.c:250 / :503 */

static void ofe_sqn(ofe *r, const ofe *a, int n)
{
    *r = *a;
    while (n-- > 0) ofe_sq(r, r);
}

/* z^(2^250-1) and z^11: common part of fld_inv (fld.c:579) and fld_pow2523 (fld.c:658). */
static void ofe_pow250(ofe *r, ofe *z11, const ofe *z)
{
    ofe z2, z9, a, b, c;
    ofe_sq(&z2, z);
    ofe_sqn(&a, &z2, 2);
    ofe_mul(&z9, &a, z);
    ofe_mul(z11, &z9, &z2);
    ofe_sq(&a, z11);
    ofe_mul(&a, &a, &z9);        /* 2^5-1 */
    ofe_sqn(&b, &a, 5);  ofe_mul(&a, &b, &a);   /* 2^10-1 */
    ofe_sqn(&b, &a, 10); ofe_mul(&b, &b, &a);   /* 2^20-1 */
    ofe_sqn(&c, &b, 20); ofe_mul(&b, &c, &b);   /* 2^40-1 */
    ofe_sqn(&b, &b, 10); ofe_mul(&a, &b, &a);   /* 2^50-1 */
    ofe_sqn(&b, &a, 50); ofe_mul(&b, &b, &a);   /* 2^100-1 */
    ofe_sqn(&c, &b, 100); ofe_mul(&b, &c, &b);  /* 2^200-1 */
    ofe_sqn(&b, &b, 50); ofe_mul(r, &b, &a);    /* 2^250-1 */
}

static void ofe_inv(ofe *r, const ofe *z) /* fld_inv, fld.c:579: z^(p-2); inv(0) = 0 (Q7) */
{
    ofe t, z11;
    ofe_pow250(&t, &z11, z);
    ofe_sqn(&t, &t, 5);
    ofe_mul(r, &t, &z11);
}

static void ofe_pow2523(ofe *r, const ofe *z) /* fld_pow2523, fld.c:658: z^((p-5)/8) */
{
    ofe t, z11;
    ofe_pow250(&t, &z11, z);
    ofe_sqn(&t, &t, 2);
    ofe_mul(r, &t, z);
}

static int ofe_iszero(const ofe *a) /* role of fld_eq, fld.c:547 */
{
    u8 s[32];
    int i, acc = 0;
    ofe_tobytes(s, a);
    for (i = 0; i < 32; i++) acc |= s[i];
    return acc == 0;
}

static int ofe_eq(const ofe *a, const ofe *b)
{
    ofe t;
    ofe_sub(&t, a, b);
    return ofe_iszero(&t);
}

static int ofe_lsb(const ofe *a)
{
    u8 s[32];
    ofe_tobytes(s, a);
    return s[0] & 1;
}

static void ofe_neg(ofe *r, const ofe *a) { ofe_sub(r, &OFE_ZERO, a); }

/* curve constants, little-endian bytes: d = -121665/121666, sqrt(-1)   (fld.c:24-41 / :295-306) */
static const u8 D_BYTES[32] = {0xa3, 0x78, 0x59, 0x13, 0xca, 0x4d, 0xeb, 0x75, 0xab, 0xd8, 0x41, 0x41, 0x4d, 0x0a, 0x70, 0x00,
                               0x98, 0xe8, 0x79, 0x77, 0x79, 0x40, 0xc7, 0x8c, 0x73, 0xfe, 0x6f, 0x2b, 0xee, 0x6c, 0x03, 0x52};
static const u8 SQRTM1_BYTES[32] = {0xb0, 0xa0, 0x0e, 0x4a, 0x27, 0x1b, 0xee, 0xc4, 0x78, 0xe4, 0x2f, 0xad, 0x06, 0x18, 0x43, 0x2f,
                                    0xa7, 0xd7, 0xfb, 0x3d, 0x99, 0x00, 0x4d, 0x2b, 0x0b, 0xdf, 0xc1, 0x4f, 0x80, 0x24, 0x83, 0x2b};
/* base point B: y = 4/5, x positive (even)                                  (ed.c:46-68 pced_B) */
static const u8 B_BYTES[32] = {0x58, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66,
                               0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66, 0x66};

/* ------------------------------------------------------------------------------------------
 * Scalars mod L = 2^252 + 27742317777372353535851937790883648493       (sc.c)
 * ------------------------------------------------------------------------------------------ */
static const u32 L_WORDS[9] = {0x5cf5d3ed, 0x5812631a, 0xa2f79cd6, 0x14def9de, 0, 0, 0, 0x10000000, 0};

/* x (nw little-endian 32-bit words) mod L -> 8 words, fully reduced.
 * Restates sc_barrett's contract (sc.c:79-158): result in [0,L) for any x < 2^512. */
static void osc_mod(u32 r[8], const u32 *x, int nw)
{
    u32 acc[9] = {0};
    int bit, i;
    for (bit = nw * 32 - 1; bit >= 0; bit--) {
        u32 carry = (x[bit >> 5] >> (bit & 31)) & 1, t[9];
        u64 bw = 0;
        for (i = 0; i < 9; i++) { u32 n = (acc[i] << 1) | carry; carry = acc[i] >> 31; acc[i] = n; }
        for (i = 0; i < 9; i++) {
            u64 d = (u64)acc[i] - L_WORDS[i] - bw;
            t[i] = (u32)d;
            bw = (d >> 63) & 1;
        }
        if (!bw) memcpy(acc, t, sizeof t);
    }
    memcpy(r, acc, 32);
}

/* sc_import (sc.c:191): up to 64 little-endian bytes -> reduced scalar. No range check (Q1). */
static void osc_frombytes(u32 r[8], const u8 *s, size_t len)
{
    u32 w[16] = {0};
    size_t i;
    for (i = 0; i < len && i < 64; i++) w[i >> 2] |= (u32)s[i] << (8 * (i & 3));
    osc_mod(r, w, 16);
}

static void osc_tobytes(u8 out[32], const u32 a[8]) /* sc_export, sc.c:221 */
{
    int i;
    for (i = 0; i < 32; i++) out[i] = (u8)(a[i >> 2] >> (8 * (i & 3)));
}

/* r = (a*b + c) mod L : sc_mul (sc.c:241) followed by sc_add (sc.h:53) and the reduce in sc_export */
static void osc_muladd(u32 r[8], const u32 a[8], const u32 b[8], const u32 c[8])
{
    u32 w[17] = {0};
    int i, j;
    for (i = 0; i < 8; i++) {
        u64 carry = 0;
        for (j = 0; j < 8; j++) {
            u64 t = (u64)a[i] * b[j] + w[i + j] + carry;
            w[i + j] = (u32)t;
            carry = t >> 32;
        }
        w[i + 8] = (u32)carry;
    }
    {
        u64 carry = 0;
        for (i = 0; i < 16; i++) {
            u64 t = (u64)w[i] + (i < 8 ? c[i] : 0) + carry;
            w[i] = (u32)t;
            carry = t >> 32;
        }
    }
    osc_mod(r, w, 16);
}

void oracle_sc_reduce(u8 out[32], const u8 *in, size_t len)
{
    u32 r[8];
    osc_frombytes(r, in, len);
    osc_tobytes(out, r);
}

void oracle_sc_muladd(u8 out[32], const u8 a[32], const u8 b[32], const u8 c[32])
{
    u32 x[8], y[8], z[8], r[8];
    osc_frombytes(x, a, 32);
    osc_frombytes(y, b, 32);
    osc_frombytes(z, c, 32);
    osc_muladd(r, x, y, z);
    osc_tobytes(out, r);
}

/* ------------------------------------------------------------------------------------------
 * Twisted Edwards group, extended coordinates (X:Y:Z:T), a = -1          (ed.c)
 * ------------------------------------------------------------------------------------------ */
typedef struct { ofe x, y, z, t; } oed;

static void oed_identity(oed *p)
{
    p->x = OFE_ZERO; p->y = OFE_ONE; p->z = OFE_ONE; p->t = OFE_ZERO;
}

/* unified complete addition (ed_add ed.c:175; also covers ed_double :211, ed_sub :245,
 * ed_add_pc :282, ed_sub_pc :310 — all specialisations of this one law) */
static void oed_add(oed *r, const oed *p, const oed *q)
{
    ofe a, b, c, d, e, f, g, h, t, d2;
    ofe_frombytes(&d2, D_BYTES);
    ofe_add(&d2, &d2, &d2);
    ofe_sub(&a, &p->y, &p->x); ofe_sub(&t, &q->y, &q->x); ofe_carry(&a); ofe_carry(&t); ofe_mul(&a, &a, &t);
    ofe_add(&b, &p->y, &p->x); ofe_add(&t, &q->y, &q->x); ofe_mul(&b, &b, &t);
    ofe_mul(&c, &p->t, &q->t); ofe_carry(&d2); ofe_mul(&c, &c, &d2);
    ofe_mul(&d, &p->z, &q->z); ofe_add(&d, &d, &d);
    ofe_sub(&e, &b, &a); ofe_sub(&f, &d, &c); ofe_add(&g, &d, &c); ofe_add(&h, &b, &a);
    ofe_carry(&e); ofe_carry(&f); ofe_carry(&g); ofe_carry(&h);
    ofe_mul(&r->x, &e, &f);
    ofe_mul(&r->y, &g, &h);
    ofe_mul(&r->t, &e, &h);
    ofe_mul(&r->z, &f, &g);
}

static void oed_neg(oed *r, const oed *p)
{
    *r = *p;
    ofe_neg(&r->x, &p->x); ofe_carry(&r->x);
    ofe_neg(&r->t, &p->t); ofe_carry(&r->t);
}

/* ed_import (ed.c:100-149).  Never fails in the reference (Q3); *on_curve reports whether the
 * recovered x satisfies the curve equation (only used for the documented Q5 policy). */
static void oed_frombytes(oed *p, const u8 in[32], int *on_curve)
{
    u8 tmp[32];
    ofe d, j, u, v, v3, v7, beta, chk, x;
    int flag, sign = in[31] >> 7;
    memcpy(tmp, in, 32);
    tmp[31] &= 0x7f;                            /* ed.c:107-108: sign bit removed, y NOT range-checked */
    ofe_frombytes(&p->y, tmp);
    ofe_frombytes(&d, D_BYTES);
    ofe_frombytes(&j, SQRTM1_BYTES);
    ofe_sq(&u, &p->y);
    ofe_mul(&v, &u, &d);
    ofe_sub(&u, &u, &OFE_ONE); ofe_carry(&u);   /* u = y^2 - 1 */
    ofe_add(&v, &v, &OFE_ONE);                  /* v = d y^2 + 1 */
    ofe_sq(&v3, &v); ofe_mul(&v3, &v3, &v);     /* v^3 */
    ofe_sq(&v7, &v3); ofe_mul(&v7, &v7, &v);    /* v^7 */
    ofe_mul(&beta, &u, &v7);
    ofe_pow2523(&beta, &beta);
    ofe_mul(&beta, &beta, &v3);
    ofe_mul(&beta, &beta, &u);                  /* beta = u v^3 (u v^7)^((p-5)/8)   ed.c:121-131 */
    ofe_sq(&chk, &beta); ofe_mul(&chk, &chk, &v);
    flag = ofe_eq(&chk, &u);                    /* ed.c:134-137 */
    if (flag) x = beta; else ofe_mul(&x, &beta, &j);   /* ed.c:140-142 */
    ofe_sq(&chk, &x); ofe_mul(&chk, &chk, &v);
    *on_curve = ofe_eq(&chk, &u);
    if (sign ^ ofe_lsb(&x)) { ofe_neg(&x, &x); ofe_carry(&x); }   /* ed.c:143-144; x = 0 stays 0 */
    p->x = x;
    ofe_mul(&p->t, &p->x, &p->y);
    p->z = OFE_ONE;
}

/* ed_export (ed.c:155-169) */
static void oed_tobytes(u8 out[32], const oed *p)
{
    ofe zi, x, y;
    ofe_inv(&zi, &p->z);
    ofe_mul(&x, &p->x, &zi);
    ofe_mul(&y, &p->y, &zi);
    ofe_tobytes(out, &y);
    out[31] |= (u8)(ofe_lsb(&x) << 7);
}

/* r = k * P, k given as 8 little-endian words (binary double-and-add, complete law). */
static void oed_scale(oed *r, const u32 k[8], const oed *p)
{
    int bit;
    oed_identity(r);
    for (bit = 255; bit >= 0; bit--) {
        oed_add(r, r, r);
        if ((k[bit >> 5] >> (bit & 31)) & 1) oed_add(r, r, p);
    }
}

/* ed_scale_base (ed.c:397-430): x * B.  The reference's signed radix-16 comb over ed_lookup
 * computes exactly this group element. */
static void oed_scale_base(oed *r, const u32 k[8])
{
    oed b;
    int oc;
    oed_frombytes(&b, B_BYTES, &oc);
    oed_scale(r, k, &b);
}

/* ------------------------------------------------------------------------------------------
 * Ed25519                                                               (ed25519-sha512.c)
 * ------------------------------------------------------------------------------------------ */
static void okey_setup(u8 h[64], const u8 sk[32]) /* ed25519_key_setup, ed25519-sha512.c:31-47 */
{
    oracle_sha512(h, sk, 32);
    h[31] &= 0x7f;
    h[31] |= 0x40;
    h[0] &= 0xf8;
}

void oracle_ed25519_genpub(u8 pub[32], const u8 sec[32]) /* genpub, ed25519-sha512.c:53-67 */
{
    u8 h[64];
    u32 a[8];
    oed A;
    okey_setup(h, sec);
    osc_frombytes(a, h, 32);
    oed_scale_base(&A, a);
    oed_tobytes(pub, &A);
}

/* sign, ed25519-sha512.c:84-123.  pub is hashed as given, never re-derived (Q8). */
void oracle_ed25519_sign(u8 sig[64], const u8 sec[32], const u8 pub[32], const u8 *data, size_t len)
{
    u8 h[64];
    u32 a[8], r[8], t[8], s[8];
    osha512 c;
    oed R;
    okey_setup(h, sec);
    osc_frombytes(a, h, 32);
    osha_init(&c); osha_add(&c, h + 32, 32); osha_add(&c, data, len); osha_final(&c, h);
    osc_frombytes(r, h, 64);
    oed_scale_base(&R, r);
    oed_tobytes(sig, &R);
    osha_init(&c); osha_add(&c, sig, 32); osha_add(&c, pub, 32); osha_add(&c, data, len); osha_final(&c, h);
    osc_frombytes(t, h, 64);
    osc_muladd(s, t, a, r);
    osc_tobytes(sig + 32, s);
}

/* ed25519_verify, ed25519-sha512.c:148-181: cofactorless; S reduced mod L without range check
 * (Q1); C = S*B + t*(-A) encoded canonically and compared with sig[0..31] as bytes (Q2);
 * A decoded without any failure path (Q3); t = SHA512(R||A||M) mod L fully reduced (Q4). */
int oracle_ed25519_verify(const u8 sig[64], const u8 pub[32], const u8 *data, size_t len)
{
    u8 h[64], check[32];
    u32 S[8], t[8];
    osha512 c;
    oed A, nA, P, Q, C;
    int on_curve;
    oed_frombytes(&A, pub, &on_curve);
    if (!on_curve) return 0;                     /* documented policy, see header / SURVEY Q5 */
    osc_frombytes(S, sig + 32, 32);
    osha_init(&c); osha_add(&c, sig, 32); osha_add(&c, pub, 32); osha_add(&c, data, len); osha_final(&c, h);
    osc_frombytes(t, h, 64);
    oed_neg(&nA, &A);
    oed_scale_base(&P, S);
    oed_scale(&Q, t, &nA);
    oed_add(&C, &P, &Q);
    oed_tobytes(check, &C);
    return memcmp(check, sig, 32) == 0;
}

/* pk_ed25519_to_x25519, ed25519-sha512.c:187-237: u = (z+y)/(z-y) */
void oracle_pk_ed25519_to_x25519(u8 out[32], const u8 in[32])
{
    oed P;
    ofe u, t;
    int oc;
    oed_frombytes(&P, in, &oc);
    ofe_add(&u, &P.z, &P.y);
    ofe_sub(&t, &P.z, &P.y); ofe_carry(&t);
    ofe_inv(&t, &t);
    ofe_mul(&u, &u, &t);
    ofe_tobytes(out, &u);
}

/* sk_ed25519_to_x25519, ed25519-sha512.c:243-256 */
void oracle_sk_ed25519_to_x25519(u8 out[32], const u8 in[32])
{
    u8 h[64];
    okey_setup(h, in);
    memcpy(out, h, 32);
}

/* ------------------------------------------------------------------------------------------
 * X25519                                                                        (x25519.c)
 * ------------------------------------------------------------------------------------------ */
static void ofe_cswap(ofe *a, ofe *b, u64 mask) /* ctmemswap, x25519.c:36 */
{
    int i;
    for (i = 0; i < 5; i++) {
        u64 d = (a->v[i] ^ b->v[i]) & mask;
        a->v[i] ^= d;
        b->v[i] ^= d;
    }
}

/* do_x25519 (x25519.c:129-150) with mg_scale (:104-123) and montgomery (:60-94):
 * scalar clamped here; u taken as the full 256-bit integer mod p (Q6); 256 ladder steps from
 * bit 255 down with swap-before / swap-after (Q7); result X * Z^(p-2), so Z = 0 gives 0. */
void oracle_x25519(u8 out[32], const u8 scalar[32], const u8 point[32])
{
    u8 e[32];
    ofe x1, x2 = OFE_ONE, z2 = OFE_ZERO, x3, z3 = OFE_ONE;
    ofe a24 = {{121665, 0, 0, 0, 0}};
    int pos;
    memcpy(e, scalar, 32);
    e[0] &= 0xf8;
    e[31] &= 0x7f;
    e[31] |= 0x40;
    ofe_frombytes(&x1, point);
    ofe_carry(&x1);
    x3 = x1;
    for (pos = 255; pos >= 0; pos--) {
        u64 mask = (u64)0 - (u64)((e[pos >> 3] >> (pos & 7)) & 1);
        ofe sa, da, aa, bb, ee, sb, db, da_sb, sa_db, t;
        ofe_cswap(&x2, &x3, mask);
        ofe_cswap(&z2, &z3, mask);
        ofe_add(&sa, &x2, &z2);
        ofe_sub(&da, &x2, &z2); ofe_carry(&da);
        ofe_sq(&aa, &sa);
        ofe_sq(&bb, &da);
        ofe_add(&sb, &x3, &z3);
        ofe_sub(&db, &x3, &z3); ofe_carry(&db);
        ofe_mul(&x2, &aa, &bb);
        ofe_sub(&ee, &aa, &bb); ofe_carry(&ee);
        ofe_mul(&t, &ee, &a24);
        ofe_add(&t, &t, &aa);
        ofe_mul(&z2, &ee, &t);
        ofe_mul(&da_sb, &da, &sb);
        ofe_mul(&sa_db, &sa, &db);
        ofe_add(&t, &da_sb, &sa_db);
        ofe_sq(&x3, &t);
        ofe_sub(&t, &da_sb, &sa_db); ofe_carry(&t);
        ofe_sq(&t, &t);
        ofe_mul(&z3, &t, &x1);
        ofe_cswap(&x2, &x3, mask);
        ofe_cswap(&z2, &z3, mask);
    }
    ofe_inv(&z2, &z2);
    ofe_mul(&x2, &x2, &z2);
    ofe_tobytes(out, &x2);
}

/* do_x25519_base (x25519.c:158-197): clamp, reduce mod L (Q9), x*B on the Edwards curve,
 * u = (Z+Y)/(Z-Y). */
void oracle_x25519_base(u8 out[32], const u8 scalar[32])
{
    u8 e[32];
    u32 k[8];
    oed R;
    ofe u, t;
    memcpy(e, scalar, 32);
    e[0] &= 0xf8;
    e[31] &= 0x7f;
    e[31] |= 0x40;
    osc_frombytes(k, e, 32);
    oed_scale_base(&R, k);
    ofe_sub(&t, &R.z, &R.y); ofe_carry(&t);
    ofe_inv(&t, &t);
    ofe_add(&u, &R.z, &R.y);
    ofe_mul(&u, &u, &t);
    ofe_tobytes(out, &u);
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers (serial loops; the threaded timing harness lives in oracle/harness.c)
 * ------------------------------------------------------------------------------------------ */
void oracle_ed25519_verify_batch(size_t n, u8 *ok, const u8 *sig, const u8 *pub, const u8 *msgs, const size_t *off, size_t fixed_len)
{
    size_t i;
    for (i = 0; i < n; i++) {
        const u8 *m = off ? msgs + off[i] : msgs + i * fixed_len;
        size_t len = off ? off[i + 1] - off[i] : fixed_len;
        ok[i] = (u8)oracle_ed25519_verify(sig + 64 * i, pub + 32 * i, m, len);
    }
}
