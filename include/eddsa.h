/*
 * eddsa.h — public API of the B200-native Ed25519 / X25519 engine.
 *
 * Drop-in for the header of phlay/libeddsa v0.8: every prototype below has the name, argument
 * order, sizes and semantics of the corresponding declaration in /root/reference/lib/eddsa.h
 * (cited per function), so existing callers and the reference's own test/selftest-*.c compile and
 * link unchanged.  Each call runs as a batch of one on the GPU; throughput-oriented callers use
 * eddsa_batch.h.  There is no CPU fallback: if no usable CUDA device is present these functions
 * print a diagnostic and abort() rather than return unverified data.
 */
#ifndef EDDSA_H
#define EDDSA_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#if !defined(EDDSA_DECL)
#  if defined(EDDSA_BUILD) && defined(__GNUC__)
#    define EDDSA_DECL __attribute__((visibility("default")))
#  else
#    define EDDSA_DECL
#  endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ED25519_KEY_LEN 32 /* reference eddsa.h:41 */
#define ED25519_SIG_LEN 64 /* reference eddsa.h:42 */
#define X25519_KEY_LEN 32  /* reference eddsa.h:62 */

/* reference eddsa.h:44 — public key of secret key `sec` */
EDDSA_DECL void ed25519_genpub(uint8_t pub[ED25519_KEY_LEN], const uint8_t sec[ED25519_KEY_LEN]);

/* reference eddsa.h:47 — deterministic signature; `pub` is hashed as given, never re-derived */
EDDSA_DECL void ed25519_sign(uint8_t sig[ED25519_SIG_LEN], const uint8_t sec[ED25519_KEY_LEN],
                             const uint8_t pub[ED25519_KEY_LEN], const uint8_t *data, size_t len);

/* reference eddsa.h:52 — true iff the signature is accepted by the reference's rules */
EDDSA_DECL bool ed25519_verify(const uint8_t sig[ED25519_SIG_LEN], const uint8_t pub[ED25519_KEY_LEN],
                               const uint8_t *data, size_t len);

/* reference eddsa.h:64 — X25519 public value of `scalar` (fixed base 9) */
EDDSA_DECL void x25519_base(uint8_t out[X25519_KEY_LEN], const uint8_t scalar[X25519_KEY_LEN]);

/* reference eddsa.h:67 — X25519 shared secret */
EDDSA_DECL void x25519(uint8_t out[X25519_KEY_LEN], const uint8_t scalar[X25519_KEY_LEN],
                       const uint8_t point[X25519_KEY_LEN]);

/* reference eddsa.h:77,80 — key conversion Ed25519 -> X25519 */
EDDSA_DECL void pk_ed25519_to_x25519(uint8_t out[X25519_KEY_LEN], const uint8_t in[ED25519_KEY_LEN]);
EDDSA_DECL void sk_ed25519_to_x25519(uint8_t out[X25519_KEY_LEN], const uint8_t in[ED25519_KEY_LEN]);

/* reference eddsa.h:92-114 — obsolete names, forwarded */
EDDSA_DECL void eddsa_genpub(uint8_t pub[32], const uint8_t sec[32]);
EDDSA_DECL void eddsa_sign(uint8_t sig[64], const uint8_t sec[32], const uint8_t pub[32], const uint8_t *data, size_t len);
EDDSA_DECL bool eddsa_verify(const uint8_t sig[64], const uint8_t pub[32], const uint8_t *data, size_t len);
EDDSA_DECL void DH(uint8_t out[32], const uint8_t sec[32], const uint8_t point[32]);
EDDSA_DECL void eddsa_pk_eddsa_to_dh(uint8_t out[32], const uint8_t in[32]);
EDDSA_DECL void eddsa_sk_eddsa_to_dh(uint8_t out[32], const uint8_t in[32]);

#ifdef __cplusplus
}
#endif
#endif
