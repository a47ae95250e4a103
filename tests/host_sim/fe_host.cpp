// TEST INFRASTRUCTURE ONLY: builds the device headers for the host so that the limb
// arithmetic can be unit-tested against Python big integers without a GPU.  Nothing in the
// product (libeddsa_b200.so) links this file.
#include "../../libeddsa_b200/csrc/fe.cuh"
#include <string.h>
using namespace edg;
extern "C" {
void h_fe_mul(uint32_t *r, const uint32_t *a, const uint32_t *b) { fe x, y, z; memcpy(x.v, a, 40); memcpy(y.v, b, 40); fe_mul(z, x, y); memcpy(r, z.v, 40); }
void h_fe_sq(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 40); fe_sq(z, x); memcpy(r, z.v, 40); }
void h_fe_mul121665(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 40); fe_mul121665(z, x); memcpy(r, z.v, 40); }
void h_fe_carry(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 40); fe_carry(z, x); memcpy(r, z.v, 40); }
void h_fe_canon(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 40); fe_canon(z, x); memcpy(r, z.v, 40); }
void h_fe_sub(uint32_t *r, const uint32_t *a, const uint32_t *b) { fe x, y, z; memcpy(x.v, a, 40); memcpy(y.v, b, 40); fe_sub(z, x, y); memcpy(r, z.v, 40); }
void h_fe_sub4(uint32_t *r, const uint32_t *a, const uint32_t *b) { fe x, y, z; memcpy(x.v, a, 40); memcpy(y.v, b, 40); fe_sub4(z, x, y); memcpy(r, z.v, 40); }
void h_fe_neg(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 40); fe_neg(z, x); memcpy(r, z.v, 40); }
void h_fe_from_bytes(uint32_t *r, const uint8_t *in) { uint32_t w[8]; memcpy(w, in, 32); fe z; fe_from_words(z, w); memcpy(r, z.v, 40); }
void h_fe_to_bytes(uint8_t *out, const uint32_t *a) { fe x; memcpy(x.v, a, 40); uint32_t w[8]; fe_to_words(w, x); memcpy(out, w, 32); }
void h_fe_inv(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 40); fe_inv(z, x); memcpy(r, z.v, 40); }
void h_fe_pow2523(uint32_t *r, const uint32_t *a) { fe x, z; memcpy(x.v, a, 40); fe_pow2523(z, x); memcpy(r, z.v, 40); }
uint32_t h_fe_is_zero(const uint32_t *a) { fe x; memcpy(x.v, a, 40); return fe_is_zero(x); }
}
