/* Internal interface between the C host layer (host.c) and the CUDA launchers (kernels.cu). */
#ifndef EDG_INTERNAL_H
#define EDG_INTERNAL_H
#include <stddef.h>
#include <stdint.h>


#ifdef __cplusplus
extern "C" {
#endif

/* All launchers: device pointers valid on the CURRENT device, 16-byte aligned; `stream` is a
 * cudaStream_t; returns 0 or a cudaError_t.  Asynchronous. */
int edg_fixedbase_init(void);
size_t edg_verify_scratch_bytes(int sm_count);   /* = signatures per pass x edg_verify_record_bytes() */
size_t edg_verify_record_bytes(void);
unsigned edg_verify_waves(void);                  /* a pass = this many waves of resident threads of the loop kernel */
int edg_launch_x25519(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point, int sm_count, void *stream);
int edg_launch_x25519_base(size_t n, uint8_t *out, const uint8_t *scalar, int sm_count, void *stream);
int edg_launch_genpub(size_t n, uint8_t *pub, const uint8_t *sec, int sm_count, void *stream);
int edg_launch_sign(size_t n, uint8_t *sig, const uint8_t *sec, const uint8_t *pub, const uint8_t *msgs,
                    const unsigned long long *off, unsigned long long fixed_len, int sm_count, void *stream);
/* window table of the base point used by verify: built once per device, read-only afterwards */
size_t edg_verify_table_bytes(void);
int edg_verify_table_init(void *table, void *stream);
void edg_verify_debug_full_scalars(int on);      /* test hook: force the full-length (rho, tau) = (1, t) path */
unsigned edg_verify_launches(size_t n, int sm_count); /* kernels one edg_launch_verify(n) launches */
int edg_launch_verify(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                      const unsigned long long *off, unsigned long long fixed_len, void *scratch, const void *table,
                      int sm_count, void *stream);
int edg_launch_pk_convert(size_t n, uint8_t *out, const uint8_t *in, int sm_count, void *stream);
int edg_launch_fe_selftest(size_t n, uint8_t *out, const uint8_t *a, const uint8_t *b, int op, int sm_count, void *stream);
int edg_launch_sk_convert(size_t n, uint8_t *out, const uint8_t *in, int sm_count, void *stream);

#ifdef __cplusplus
}
#endif
#endif
