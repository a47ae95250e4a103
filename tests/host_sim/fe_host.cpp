// TEST INFRASTRUCTURE ONLY: builds the device headers for the host so that the limb
// arithmetic can be unit-tested against Python big integers without a GPU.  Nothing in the
// product (libeddsa_b200.so) links this file.
#include "../../libeddsa_b200/csrc/fe.cuh"
#include <string.h>
using namespace edg;
extern "C" {
void h_fe_mul(uint32_t *r, const uint32_t *a, const uint32_t *b) { fe x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); fe_mul(z, x, y); memcpy(r, z.v, 32); }
void h_fe_sq(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 32); fe_sq(z, x); memcpy(r, z.v, 32); }
void h_fe_mul121665(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 32); fe_mul121665(z, x); memcpy(r, z.v, 32); }
void h_fe_canon(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 32); fe_canon(z, x); memcpy(r, z.v, 32); }
void h_fe_add(uint32_t *r, const uint32_t *a, const uint32_t *b) { fe x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); fe_add(z, x, y); memcpy(r, z.v, 32); }
void h_fe_sub(uint32_t *r, const uint32_t *a, const uint32_t *b) { fe x, y, z; memcpy(x.v, a, 32); memcpy(y.v, b, 32); fe_sub(z, x, y); memcpy(r, z.v, 32); }
void h_fe_neg(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 32); fe_neg(z, x); memcpy(r, z.v, 32); }
void h_fe_from_bytes(uint32_t *r, const uint8_t *in) { uint32_t w[8]; memcpy(w, in, 32); fe z; fe_from_words(z, w); memcpy(r, z.v, 32); }
void h_fe_to_bytes(uint8_t *out, const uint32_t *a) { fe x; memcpy(x.v, a, 32); uint32_t w[8]; fe_to_words(w, x); memcpy(out, w, 32); }
void h_fe_inv(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 32); fe_inv(z, x); memcpy(r, z.v, 32); }
void h_fe_pow2523(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 32); fe_pow2523(z, x); memcpy(r, z.v, 32); }
uint32_t h_fe_is_zero(const uint32_t *a) { fe x; memcpy(x.v, a, 32); return fe_is_zero(x); }
}
#include "../../libeddsa_b200/csrc/sc.cuh"
extern "C" {
void h_sc_reduce512(uint32_t *r, const uint32_t *x) { sc_reduce512(r, x); }
void h_sc_reduce256(uint32_t *r, const uint32_t *x) { sc_reduce256(r, x); }
void h_sc_muladd(uint32_t *r, const uint32_t *a, const uint32_t *b, const uint32_t *c) { sc_muladd(r, a, b, c); }
void h_sc_recode(uint32_t *e, const uint32_t *x) { uint32_t t[EDG_COMB_EW]; sc_recode_comb(t, x); for (int i = 0; i < 9; i++) e[i] = i < EDG_COMB_EW ? t[i] : 0; }
int h_comb_w(void) { return EDG_COMB_W; }
}
#include "../../libeddsa_b200/csrc/sha512.cuh"
extern "C" {
// digest = SHA512(pre (npre bytes: 0, 32 or 64) || msg)
void h_sha512(uint8_t *out, const uint8_t *pre, int npre, const uint8_t *msg, uint64_t len) {
    u64 st[8], pw[8];
    for (int k = 0; k < npre / 8; k++) { u32 a, b; memcpy(&a, pre + 8 * k, 4); memcpy(&b, pre + 8 * k + 4, 4); pw[k] = be64_from_le_words(a, b); }
    if (npre == 0) sha512_prefixed<0>(st, pw, msg, len);
    else if (npre == 32) sha512_prefixed<4>(st, pw, msg, len);
    else sha512_prefixed<8>(st, pw, msg, len);
    u32 x[16]; sha512_state_to_le_words(x, st); memcpy(out, x, 64);
}
}
