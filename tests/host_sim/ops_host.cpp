// TEST INFRASTRUCTURE ONLY: host build of the per-thread operation bodies (ops.cuh), so the exact
// code the CUDA kernels execute can be checked against the oracle / golden fixtures without a GPU.
// Nothing in the product links this file.
#include <string.h>
#include <stdint.h>
#define EDG_COUNT_OPS
#include "../../libeddsa_b200/csrc/ops.cuh"
#if defined(__CUDA_ARCH__)   // PTX-emulation build (ptx_rewrite.py + ptx_emul.h): the device paths carry no instrumentation
static unsigned long edg_cnt_mul = 0, edg_cnt_sq = 0;
static int edg_last_nwin = 0;
#endif
using namespace edg;
// the fixed-base comb table, built exactly as the device does (k_comb_base / k_comb_rows); not counted as field work
static const u32 *host_comb() {
    static u32 *tab = 0;
    if (!tab) {
        const unsigned long m = edg_cnt_mul, q = edg_cnt_sq;
        tab = new u32[EDG_COMB_WORDS];
        for (int j = 0; j < EDG_COMB_ROWS; j++) {
            u32 base[24];
            wtab_base(base, EDG_COMB_W * j);
            comb_table_row(tab + (size_t)j * EDG_COMB_ENTRIES * 24, base);
        }
        edg_cnt_mul = m; edg_cnt_sq = q;
    }
    return tab;
}
#define BASE_COMB host_comb()
extern "C" {
const uint32_t *hs_comb(int *rows, int *entries) { *rows = EDG_COMB_ROWS; *entries = EDG_COMB_ENTRIES; return host_comb(); }
// Emulation of the tensor-core table lookup (ge.cuh: ge_pre_select_mma) for one warp, with the fragment layouts of
// mma.sync.m16n8k16 / m16n8k32 .u8 as the PTX ISA defines them (A: thread (g, q) holds rows g, g + 8 at k = 4q + i (+16);
// B: k = 4q + i (+16), column g; D: rows g, g + 8, columns 2q, 2q + 1).  It reads the fragment-order table built by the
// SAME comb_mma_word() the device uses and reproduces the device's one-hot construction, byte packing and exchange-area
// indexing: out[lane][24] must be entry |digit_lane| of the row (all zero for digit 0).
void hs_mma_select(uint32_t *out, int row, const uint32_t *absd) {
    const u32 *std_tab = host_comb();
    const int H = EDG_COMB_ENTRIES / 16;
    static u32 *mma = 0;
    if (!mma) { mma = new u32[EDG_COMB_WORDS]; for (unsigned i = 0; i < EDG_COMB_WORDS; i++) mma[i] = comb_mma_word(std_tab, i); }
    const u32 *row_mma = mma + (size_t)row * EDG_COMB_ENTRIES * 24;
    u32 xchg[32 * 28], A[32][4][2];
    for (unsigned lane = 0; lane < 32; lane++) {           // every thread's A fragments, by the device's formula
        const unsigned g = lane >> 2, q = lane & 3;
        for (int r = 0; r < 4; r++)
            for (int h = 0; h < H; h++) {
                const u32 x = absd[g + 8 * r] - 1u - 16u * h - 4u * q;
                const u32 hit = 0u - ((((x >> 2) - 1u) >> 31));
                A[lane][r][h] = (1u << ((x & 3u) * 8u)) & hit;
            }
    }
    for (unsigned lane = 0; lane < 32; lane++) {
        const unsigned g = lane >> 2, q = lane & 3;
        for (int m = 0; m < 6; m++)
            for (int mt = 0; mt < 2; mt++) {
                u32 d[2][4];                                   // the two byte-tiles 2m, 2m + 1
                for (int s = 0; s < 2; s++)
                    for (int c = 0; c < 4; c++) {              // c0, c1: row g; c2, c3: row g + 8; columns 2q + (c & 1)
                        const int rr = 2 * mt + (c >> 1), col = 2 * q + (c & 1);
                        // D[row][col] = sum_k A[row][k] B[k][col] over k = 0 .. 16H - 1.  A[row g + 8 rr][k] is byte k % 4 of word
                        // k / 16 of thread (g, (k % 16) / 4); B[k][col] is byte k % 4 of word k / 16 of thread (col, (k % 16) / 4)
                        u32 sum = 0;
                        for (int k = 0; k < 16 * H; k++) {
                            const u32 aw = A[4 * g + (k % 16) / 4][rr][k / 16];
                            const u32 bw = row_mma[((2 * m + s) * H + k / 16) * 32 + 4 * col + (k % 16) / 4];
                            sum += ((aw >> (8 * (k % 4))) & 0xffu) * ((bw >> (8 * (k % 4))) & 0xffu);
                        }
                        d[s][c] = sum;
                    }
                const u32 lo = (d[0][0] & 0xff) | (d[0][1] & 0xff) << 8 | (d[1][0] & 0xff) << 16 | (d[1][1] & 0xff) << 24;
                const u32 hi = (d[0][2] & 0xff) | (d[0][3] & 0xff) << 8 | (d[1][2] & 0xff) << 16 | (d[1][3] & 0xff) << 24;
                xchg[(g + 16 * mt) * 28 + 6 * q + m] = lo;
                xchg[(g + 8 + 16 * mt) * 28 + 6 * q + m] = hi;
            }
    }
    for (int lane = 0; lane < 32; lane++)
        for (int i = 0; i < 24; i++) out[24 * lane + i] = xchg[lane * 28 + i];
}
void hs_x25519(uint8_t *out, const uint8_t *scalar, const uint8_t *point) {
    u32 o[8], s[8], p[8]; memcpy(s, scalar, 32); memcpy(p, point, 32); x25519_op(o, s, p); memcpy(out, o, 32);
}
void hs_genpub(uint8_t *pub, const uint8_t *sk) { u32 o[8]; ed25519_genpub_op(o, sk, BASE_COMB); memcpy(pub, o, 32); }
void hs_sign(uint8_t *sig, const uint8_t *sk, const uint8_t *pub, const uint8_t *msg, uint64_t len) {
    u32 o[16], p[8]; memcpy(p, pub, 32); ed25519_sign_op(o, sk, p, msg, len, BASE_COMB); memcpy(sig, o, 64);
}
// the base-point window table, built exactly as the device does (k_wtab_base / k_wtab_build); not counted as field work
static const u32 *host_wtab() {
    static u32 *tab = 0;
    if (!tab) {
        const unsigned long m = edg_cnt_mul, q = edg_cnt_sq;
        tab = new u32[2 * (size_t)EDG_WTAB_WORDS];
        for (int m = 0; m < 2; m++) {
            u32 base[24], *tb = tab + m * (size_t)EDG_WTAB_WORDS;
            wtab_base(base, 128 * m);
            for (int i = 0; i < 24; i++) tb[i] = (i == 0 || i == 8) ? 1u : 0u;
            for (u32 g = 0; 8u * g + 8u < EDG_WTAB_ENTRIES; g++) wtab_build8(tb + 24u * (8u * g + 1u), base, 8u * g + 1u);
        }
        edg_cnt_mul = m; edg_cnt_sq = q;
    }
    return tab;
}
const uint32_t *hs_wtab(uint32_t *entries) { *entries = EDG_WTAB_ENTRIES; return host_wtab(); }
int hs_verify(const uint8_t *sig, const uint8_t *pub, const uint8_t *msg, uint64_t len) {
    u32 s[16], p[8], qtab[EDG_VSTATE_WORDS]; memcpy(s, sig, 64); memcpy(p, pub, 32);
    const u32 *wt = host_wtab();
    return (int)ed25519_verify_op(s, p, msg, len, qtab, wt);
}
void hs_x25519_base(uint8_t *out, const uint8_t *scalar) { u32 o[8], s[8]; memcpy(s, scalar, 32); x25519_base_op(o, s, BASE_COMB); memcpy(out, o, 32); }
void hs_pk_conv(uint8_t *out, const uint8_t *in) { u32 o[8], p[8]; memcpy(p, in, 32); pk_ed25519_to_x25519_op(o, p); memcpy(out, o, 32); }
void hs_sk_conv(uint8_t *out, const uint8_t *in) { u32 o[8]; sk_ed25519_to_x25519_op(o, in); memcpy(out, o, 32); }
void hs_batch_inv(uint8_t *z, int cnt) {   // z: cnt x 32 bytes, in place
    fe t[EDG_BATCH]; for (int k = 0; k < cnt; k++) memcpy(t[k].v, z + 32 * k, 32);
    fe_batch_inv(t, cnt);
    for (int k = 0; k < cnt; k++) { u32 w[8]; fe_to_words(w, t[k]); memcpy(z + 32 * k, w, 32); }
}
int hs_verify_full(const uint8_t *sig, const uint8_t *pub, const uint8_t *msg, uint64_t len) {   // the (1, t) fallback path
    u32 s[16], p[8], qtab[EDG_VSTATE_WORDS]; memcpy(s, sig, 64); memcpy(p, pub, 32);
    return (int)ed25519_verify_op(s, p, msg, len, qtab, host_wtab(), true);
}
int hs_last_nwin(void) { return edg_last_nwin; }
#if defined(__CUDA_ARCH__)
// PTX-emulation build only: the verify loop with its cp.async staging of the window's two scratch-table entries through
// "shared memory" (what k_verify does with s_stage), as thread 0 of a block of one
int hs_verify_staged(const uint8_t *sig, const uint8_t *pub, const uint8_t *msg, uint64_t len) {
    u32 s[16], p[8], qtab[EDG_VSTATE_WORDS];
    static uint4 stage[16];
    memcpy(s, sig, 64); memcpy(p, pub, 32);
    ed25519_verify_front(qtab, s, p, msg, len, false);
    return (int)ed25519_verify_loop(qtab, host_wtab(), reinterpret_cast<u32 *>(stage));
}
#endif
void hs_half_gcd(uint32_t *rho_abs, uint32_t *rho_neg, uint32_t *tau, const uint32_t *t) { u32 n; half_gcd(rho_abs, n, tau, t); *rho_neg = n; }
void hs_counts(unsigned long *mul, unsigned long *sq, int reset) { *mul = edg_cnt_mul; *sq = edg_cnt_sq; if (reset) edg_cnt_mul = edg_cnt_sq = 0; }
}
