// TEST INFRASTRUCTURE ONLY: host build of the per-thread operation bodies (ops.cuh), so the exact
// code the CUDA kernels execute can be checked against the oracle / golden fixtures without a GPU.
// Nothing in the product links this file.
#include <string.h>
#include <stdint.h>
#define EDG_COUNT_OPS
#include "../../libeddsa_b200/csrc/ops.cuh"
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
void hs_half_gcd(uint32_t *rho_abs, uint32_t *rho_neg, uint32_t *tau, const uint32_t *t) { u32 n; half_gcd(rho_abs, n, tau, t); *rho_neg = n; }
void hs_counts(unsigned long *mul, unsigned long *sq, int reset) { *mul = edg_cnt_mul; *sq = edg_cnt_sq; if (reset) edg_cnt_mul = edg_cnt_sq = 0; }
}
