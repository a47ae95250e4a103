/*
 * host.c — C host layer of the B200-native Ed25519 / X25519 engine.
 *
 * Implements the public API of include/eddsa.h (the reference's /root/reference/lib/eddsa.h:44-114
 * surface, single operations executed as batches of one) and the batch C-ABI of
 * include/eddsa_batch.h on top of the CUDA launchers in kernels_*.cu.
 *
 *   - one context per device: NSLOT streams, page-locked staging buffers, device buffers and the
 *     verify kernel's per-thread scratch; created lazily, guarded by a mutex;
 *   - a host-buffer batch is sharded by contiguous index ranges over the devices (one host thread
 *     per device, no inter-device traffic) and each shard is streamed in chunks: stage -> H2D ->
 *     kernel -> D2H -> unstage, the slots rotating so copies of one chunk overlap the kernel of
 *     another; caller memory that is already page-locked is used directly;
 *   - no CPU fallback anywhere: without a working device the batch calls return an error and the
 *     void single-operation calls abort().
 */
#define _GNU_SOURCE
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "eddsa.h"
#include "eddsa_batch.h"
#include "edg_internal.h"

#define EDG_MAX_DEV 16
#define EDG_NSLOT 3
#define EDG_ALIGN 256
#define EDG_MAX_USER_STREAMS 16

typedef enum { OP_GENPUB, OP_SIGN, OP_VERIFY, OP_X25519, OP_X25519_BASE, OP_PK_CONV, OP_SK_CONV, OP_FE_TEST } edg_op_t;

typedef struct {
    edg_op_t op;
    size_t n;
    int nin;
    const uint8_t *in[3];
    size_t in_item[3];
    int has_msgs;
    const uint8_t *msgs;
    const size_t *off;
    size_t fixed_len;
    uint8_t *out;
    size_t out_item;
} edg_job_t;

typedef struct {
    int dev, sm_count, ready;
    pthread_mutex_t lock; /* the host-buffer pipeline of this device is exclusive */
    cudaStream_t stream[EDG_NSLOT];
    cudaEvent_t done[EDG_NSLOT];
    cudaStream_t kstream;                       /* verify: all kernels of the host-buffer pipeline run here, one after the other */
    cudaEvent_t in_ready[EDG_NSLOT], k_done[EDG_NSLOT];
    size_t verify_pass;                         /* signatures per pass of the two verify kernels (whole waves) */
    uint8_t *h_in[EDG_NSLOT], *h_out[EDG_NSLOT];
    uint8_t *d_in[EDG_NSLOT], *d_out[EDG_NSLOT];
    size_t in_cap, out_cap;
    void *scratch[EDG_NSLOT];
    size_t scratch_bytes;
    void *wtab; /* verify: window table of the base point, built at context creation */
    pthread_mutex_t us_lock;
    struct { void *stream; void *buf; int used; } user_scratch[EDG_MAX_USER_STREAMS];
} edg_dev_t;

static edg_dev_t g_dev[EDG_MAX_DEV]; /* indexed by CUDA device ordinal */
static int g_list[EDG_MAX_DEV];      /* devices the host-buffer API shards over */
static int g_nlist = 0, g_nactive = 0, g_init_rc = 0;
static size_t g_chunk_bytes = (size_t)128 << 20;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static pthread_mutex_t g_ctx_lock = PTHREAD_MUTEX_INITIALIZER;
static unsigned long long g_launches = 0;
static __thread char t_err[256];

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return code ? code : EDDSA_B200_EINVAL;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    rc = fail((int)e_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); goto out; } } while (0)

static size_t align_up(size_t x) { return (x + EDG_ALIGN - 1) & ~(size_t)(EDG_ALIGN - 1); }

static void global_init(void)
{
    int count = 0, i;
    const char *env = getenv("EDDSA_B200_DEVICES");
    const char *chunk = getenv("EDDSA_B200_CHUNK_MB");
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1) {
        g_init_rc = e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice;
        return;
    }
    if (count > EDG_MAX_DEV) count = EDG_MAX_DEV;
    for (i = 0; i < EDG_MAX_DEV; i++) {
        g_dev[i].dev = i;
        pthread_mutex_init(&g_dev[i].lock, NULL);
        pthread_mutex_init(&g_dev[i].us_lock, NULL);
    }
    if (env && strchr(env, ',')) {
        char buf[128], *tok, *save = NULL;
        snprintf(buf, sizeof buf, "%s", env);
        for (tok = strtok_r(buf, ",", &save); tok && g_nlist < EDG_MAX_DEV; tok = strtok_r(NULL, ",", &save)) {
            int d = atoi(tok);
            if (d >= 0 && d < count) g_list[g_nlist++] = d;
        }
    } else {
        int want = env && *env ? atoi(env) : count;
        if (want < 1 || want > count) want = count;
        for (i = 0; i < want; i++) g_list[g_nlist++] = i;
    }
    if (g_nlist == 0) g_list[g_nlist++] = 0;
    g_nactive = g_nlist;
    if (chunk && atoi(chunk) > 0) g_chunk_bytes = (size_t)atoi(chunk) << 20;
    if (getenv("EDDSA_B200_DEBUG_FULL_SCALARS") && atoi(getenv("EDDSA_B200_DEBUG_FULL_SCALARS")) > 0) edg_verify_debug_full_scalars(1);
}

static int engine_ready(void)
{
    pthread_once(&g_once, global_init);
    if (g_init_rc) return fail(g_init_rc, "no usable CUDA device: %s", cudaGetErrorString((cudaError_t)g_init_rc));
    return 0;
}

/* per-device context without staging buffers (enough for the _dev API) */
static int dev_basic(int dev, edg_dev_t **out_ctx)
{
    edg_dev_t *c;
    int rc = engine_ready();
    if (rc) return rc;
    if (dev < 0 || dev >= EDG_MAX_DEV) return fail(EDDSA_B200_EINVAL, "device ordinal %d out of range", dev);
    c = &g_dev[dev];
    if (!c->ready) {
        pthread_mutex_lock(&g_ctx_lock);
        if (!c->ready) {
            int prev = -1, sms = 0, i;
            cudaGetDevice(&prev);
            CU(cudaSetDevice(dev));
            CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            c->sm_count = sms;
            rc = edg_fixedbase_init();
            if (rc) { rc = fail(rc, "kernel attribute setup failed: %s", cudaGetErrorString((cudaError_t)rc)); goto out; }
            c->scratch_bytes = edg_verify_scratch_bytes(sms);
            for (i = 0; i < EDG_NSLOT; i++) {
                CU(cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&c->in_ready[i], cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&c->k_done[i], cudaEventDisableTiming));
            }
            CU(cudaStreamCreateWithFlags(&c->kstream, cudaStreamNonBlocking));
            c->verify_pass = (c->scratch_bytes - 256) / edg_verify_record_bytes();
            CU(cudaMalloc(&c->wtab, edg_verify_table_bytes()));
            rc = edg_verify_table_init(c->wtab, c->stream[0]);
            if (rc) { rc = fail(rc, "window table build failed: %s", cudaGetErrorString((cudaError_t)rc)); goto out; }
            CU(cudaStreamSynchronize(c->stream[0]));
            __sync_synchronize();
            c->ready = 1;
        out:
            if (prev >= 0 && prev != dev) cudaSetDevice(prev);
        }
        pthread_mutex_unlock(&g_ctx_lock);
        if (rc) return rc;
    }
    *out_ctx = c;
    return 0;
}

/* make sure the staging / device buffers of c hold at least the given capacities (device current) */
static int dev_reserve(edg_dev_t *c, size_t in_need, size_t out_need, int need_scratch)
{
    int rc = 0, i;
    if (in_need < ((size_t)1 << 20)) in_need = (size_t)1 << 20;   /* avoid re-allocating for small growing batches */
    if (out_need < ((size_t)1 << 16)) out_need = (size_t)1 << 16;
    if (in_need > c->in_cap) {
        for (i = 0; i < EDG_NSLOT; i++) {
            if (c->h_in[i]) cudaFreeHost(c->h_in[i]);
            if (c->d_in[i]) cudaFree(c->d_in[i]);
            c->h_in[i] = c->d_in[i] = NULL;
        }
        c->in_cap = 0;
        for (i = 0; i < EDG_NSLOT; i++) {
            CU(cudaMallocHost((void **)&c->h_in[i], in_need));
            CU(cudaMalloc((void **)&c->d_in[i], in_need));
        }
        c->in_cap = in_need;
    }
    if (out_need > c->out_cap) {
        for (i = 0; i < EDG_NSLOT; i++) {
            if (c->h_out[i]) cudaFreeHost(c->h_out[i]);
            if (c->d_out[i]) cudaFree(c->d_out[i]);
            c->h_out[i] = c->d_out[i] = NULL;
        }
        c->out_cap = 0;
        for (i = 0; i < EDG_NSLOT; i++) {
            CU(cudaMallocHost((void **)&c->h_out[i], out_need));
            CU(cudaMalloc((void **)&c->d_out[i], out_need));
        }
        c->out_cap = out_need;
    }
    if (need_scratch && !c->scratch[0])
        CU(cudaMalloc(&c->scratch[0], c->scratch_bytes));      /* one slab: the pipeline's verify kernels run one after the other */
out:
    return rc;
}

static int is_pinned(const void *p)
{
    struct cudaPointerAttributes a;
    if (!p) return 0;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost;
}

static int launch(edg_dev_t *c, edg_op_t op, size_t n, uint8_t *d_out, uint8_t *const d_in[3], const uint8_t *d_msgs,
                  const unsigned long long *d_off, size_t fixed_len, void *scratch, void *stream)
{
    int rc;
    switch (op) {
    case OP_GENPUB: rc = edg_launch_genpub(n, d_out, d_in[0], c->sm_count, stream); break;
    case OP_SIGN: rc = edg_launch_sign(n, d_out, d_in[0], d_in[1], d_msgs, d_off, fixed_len, c->sm_count, stream); break;
    case OP_VERIFY: rc = edg_launch_verify(n, d_out, d_in[0], d_in[1], d_msgs, d_off, fixed_len, scratch, c->wtab, c->sm_count, stream); break;
    case OP_X25519: rc = edg_launch_x25519(n, d_out, d_in[0], d_in[1], c->sm_count, stream); break;
    case OP_X25519_BASE: rc = edg_launch_x25519_base(n, d_out, d_in[0], c->sm_count, stream); break;
    case OP_PK_CONV: rc = edg_launch_pk_convert(n, d_out, d_in[0], c->sm_count, stream); break;
    case OP_SK_CONV: rc = edg_launch_sk_convert(n, d_out, d_in[0], c->sm_count, stream); break;
    case OP_FE_TEST: rc = edg_launch_fe_selftest(n, d_out, d_in[0], d_in[1], (int)fixed_len, c->sm_count, stream); break;
    default: return fail(EDDSA_B200_EINVAL, "unknown operation");
    }
    if (rc) return fail(rc, "kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    if (n) __sync_fetch_and_add(&g_launches, op == OP_VERIFY ? (unsigned long long)edg_verify_launches(n, c->sm_count) : 1ULL);
    return 0;
}

static size_t msg_bytes(const edg_job_t *j, size_t lo, size_t hi)
{
    if (!j->has_msgs) return 0;
    return j->off ? j->off[hi] - j->off[lo] : (hi - lo) * j->fixed_len;
}

/* bytes of device input for items [lo, hi): fixed-size arrays + rebased offsets + message bytes */
static size_t chunk_in_bytes(const edg_job_t *j, size_t lo, size_t hi)
{
    size_t b = 0;
    int k;
    for (k = 0; k < j->nin; k++) b += align_up((hi - lo) * j->in_item[k]);
    if (j->has_msgs) {
        if (j->off) b += align_up((hi - lo + 1) * sizeof(unsigned long long));
        b += align_up(msg_bytes(j, lo, hi) + 16);
    }
    return b;
}

/* Run items [lo, hi) of job j on device context c (chunked, pipelined over EDG_NSLOT slots). */
static int run_shard(edg_dev_t *c, const edg_job_t *j, size_t lo, size_t hi)
{
    int rc = 0, k, s;
    size_t pos, nshard = hi - lo, target, c_lo[EDG_NSLOT], c_hi[EDG_NSLOT];
    int inflight[EDG_NSLOT] = {0};
    int pin_in[3] = {0, 0, 0}, pin_msgs = 0, pin_out = 0;
    size_t chunk_no = 0;
    if (nshard == 0) return 0;
    pthread_mutex_lock(&c->lock);
    CU(cudaSetDevice(c->dev));
    for (k = 0; k < j->nin; k++) pin_in[k] = is_pinned(j->in[k]);
    pin_msgs = j->has_msgs && is_pinned(j->msgs);
    pin_out = is_pinned(j->out);
    /* aim for >= 4 chunks per shard (copy/compute overlap) but never tiny ones */
    target = (nshard + 3) / 4;
    if (target < 65536) target = 65536;
    /* verify: chunks of whole passes of its two kernels (whole waves of resident threads), kernels serialised on one
     * stream so that the two stages never share an SM (instruction cache) while copies overlap on the slot streams */
    if (j->op == OP_VERIFY && nshard > c->verify_pass) target = c->verify_pass;
    if (target > nshard) target = nshard;
    for (pos = lo; pos < hi;) {
        size_t m = target, in_bytes, out_bytes, ofs;
        uint8_t *d_in[3] = {NULL, NULL, NULL};
        const uint8_t *d_msgs = NULL;
        const unsigned long long *d_off = NULL;
        /* verify: a short first chunk (one wave) so that the kernels start early; its copy is the only exposed one */
        if (j->op == OP_VERIFY && chunk_no == 0 && target == c->verify_pass && c->verify_pass >= 8) m = c->verify_pass / edg_verify_waves();
        if (pos + m > hi) m = hi - pos;
        /* shrink the chunk until it fits the staging budget (ragged messages); a single oversized item grows the buffers */
        while (m > 1 && chunk_in_bytes(j, pos, pos + m) > g_chunk_bytes) m = (m + 1) / 2;
        in_bytes = chunk_in_bytes(j, pos, pos + m);
        out_bytes = align_up(m * j->out_item);
        s = (int)(chunk_no % EDG_NSLOT);
        if (inflight[s]) { /* retire the chunk that used this slot */
            CU(cudaEventSynchronize(c->done[s]));
            if (!pin_out) memcpy(j->out + c_lo[s] * j->out_item, c->h_out[s], (c_hi[s] - c_lo[s]) * j->out_item);
            inflight[s] = 0;
        }
        if (in_bytes > c->in_cap || out_bytes > c->out_cap || (j->op == OP_VERIFY && !c->scratch[0])) {
            /* (re)allocation frees buffers other slots may still use: drain first */
            for (k = 0; k < EDG_NSLOT; k++)
                if (inflight[k]) {
                    CU(cudaEventSynchronize(c->done[k]));
                    if (!pin_out) memcpy(j->out + c_lo[k] * j->out_item, c->h_out[k], (c_hi[k] - c_lo[k]) * j->out_item);
                    inflight[k] = 0;
                }
            rc = dev_reserve(c, in_bytes > c->in_cap ? in_bytes : c->in_cap, out_bytes > c->out_cap ? out_bytes : c->out_cap,
                             j->op == OP_VERIFY);
            if (rc) goto out;
        }
        ofs = 0;
        for (k = 0; k < j->nin; k++) {
            size_t bytes = m * j->in_item[k];
            const uint8_t *src = j->in[k] + pos * j->in_item[k];
            if (!pin_in[k]) { memcpy(c->h_in[s] + ofs, src, bytes); src = c->h_in[s] + ofs; }
            CU(cudaMemcpyAsync(c->d_in[s] + ofs, src, bytes, cudaMemcpyHostToDevice, c->stream[s]));
            d_in[k] = c->d_in[s] + ofs;
            ofs += align_up(bytes);
        }
        if (j->has_msgs) {
            size_t mb = msg_bytes(j, pos, pos + m);
            const uint8_t *src = j->msgs + (j->off ? j->off[pos] : pos * j->fixed_len);
            if (j->off) {
                unsigned long long *ho = (unsigned long long *)(c->h_in[s] + ofs);
                size_t i;
                for (i = 0; i <= m; i++) ho[i] = (unsigned long long)(j->off[pos + i] - j->off[pos]);
                CU(cudaMemcpyAsync(c->d_in[s] + ofs, ho, (m + 1) * sizeof *ho, cudaMemcpyHostToDevice, c->stream[s]));
                d_off = (const unsigned long long *)(c->d_in[s] + ofs);
                ofs += align_up((m + 1) * sizeof *ho);
            }
            if (mb) {
                if (!pin_msgs) { memcpy(c->h_in[s] + ofs, src, mb); src = c->h_in[s] + ofs; }
                CU(cudaMemcpyAsync(c->d_in[s] + ofs, src, mb, cudaMemcpyHostToDevice, c->stream[s]));
            }
            d_msgs = c->d_in[s] + ofs;
        }
        if (j->op == OP_VERIFY) {
            CU(cudaEventRecord(c->in_ready[s], c->stream[s]));
            CU(cudaStreamWaitEvent(c->kstream, c->in_ready[s], 0));
            rc = launch(c, j->op, m, c->d_out[s], d_in, d_msgs, d_off, j->fixed_len, c->scratch[0], c->kstream);
            if (rc) goto out;
            CU(cudaEventRecord(c->k_done[s], c->kstream));
            CU(cudaStreamWaitEvent(c->stream[s], c->k_done[s], 0));
        } else {
            rc = launch(c, j->op, m, c->d_out[s], d_in, d_msgs, d_off, j->fixed_len, NULL, c->stream[s]);
            if (rc) goto out;
        }
        CU(cudaMemcpyAsync(pin_out ? j->out + pos * j->out_item : c->h_out[s], c->d_out[s], m * j->out_item,
                           cudaMemcpyDeviceToHost, c->stream[s]));
        CU(cudaEventRecord(c->done[s], c->stream[s]));
        inflight[s] = 1; c_lo[s] = pos; c_hi[s] = pos + m;
        pos += m;
        chunk_no++;
    }
out:
    for (k = 0; k < EDG_NSLOT; k++) {
        s = (int)((chunk_no + k) % EDG_NSLOT); /* oldest first */
        if (inflight[s]) {
            cudaError_t e = cudaEventSynchronize(c->done[s]);
            if (e != cudaSuccess && !rc) rc = fail((int)e, "cudaEventSynchronize failed: %s", cudaGetErrorString(e));
            if (!rc && !pin_out) memcpy(j->out + c_lo[s] * j->out_item, c->h_out[s], (c_hi[s] - c_lo[s]) * j->out_item);
        }
    }
    pthread_mutex_unlock(&c->lock);
    return rc;
}

typedef struct { edg_dev_t *c; const edg_job_t *j; size_t lo, hi; int rc; char err[256]; } shard_arg_t;

static void *shard_thread(void *p)
{
    shard_arg_t *a = (shard_arg_t *)p;
    t_err[0] = 0;
    a->rc = run_shard(a->c, a->j, a->lo, a->hi);
    memcpy(a->err, t_err, sizeof a->err);
    return NULL;
}

static int run_job(const edg_job_t *j)
{
    int rc = engine_ready(), g, ndev, k;
    shard_arg_t args[EDG_MAX_DEV];
    pthread_t th[EDG_MAX_DEV];
    if (rc) return rc;
    t_err[0] = 0;
    if (j->n == 0) return 0;
    for (k = 0; k < j->nin; k++)
        if (!j->in[k]) return fail(EDDSA_B200_EINVAL, "NULL input array");
    if (!j->out) return fail(EDDSA_B200_EINVAL, "NULL output array");
    if (j->has_msgs && !j->msgs && msg_bytes(j, 0, j->n) != 0) return fail(EDDSA_B200_EINVAL, "NULL message blob");
    if (j->has_msgs && j->off)
        for (size_t i = 0; i < j->n; i++)
            if (j->off[i + 1] < j->off[i]) return fail(EDDSA_B200_EINVAL, "message offsets must be non-decreasing");
    ndev = g_nactive;
    if ((size_t)ndev > (j->n + 16383) / 16384) ndev = (int)((j->n + 16383) / 16384); /* small batches: fewer devices */
    if (ndev < 1) ndev = 1;
    for (g = 0; g < ndev; g++) {
        rc = dev_basic(g_list[g], &args[g].c);
        if (rc) return rc;
        args[g].j = j;
        args[g].lo = j->n * (size_t)g / ndev;
        args[g].hi = j->n * (size_t)(g + 1) / ndev;
        args[g].rc = 0;
    }
    if (ndev == 1) return run_shard(args[0].c, j, 0, j->n);
    for (g = 0; g < ndev; g++) pthread_create(&th[g], NULL, shard_thread, &args[g]);
    for (g = 0; g < ndev; g++) {
        pthread_join(th[g], NULL);
        if (args[g].rc && !rc) { rc = args[g].rc; memcpy(t_err, args[g].err, sizeof t_err); }
    }
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * host-buffer batch API
 * --------------------------------------------------------------------------------------------- */
int ed25519_genpub_batch(size_t n, uint8_t *pub, const uint8_t *sec)
{
    edg_job_t j = {OP_GENPUB, n, 1, {sec, NULL, NULL}, {32, 0, 0}, 0, NULL, NULL, 0, pub, 32};
    return run_job(&j);
}

int ed25519_sign_batch(size_t n, uint8_t *sig, const uint8_t *sec, const uint8_t *pub, const uint8_t *msgs,
                       const size_t *off, size_t fixed_len)
{
    edg_job_t j = {OP_SIGN, n, 2, {sec, pub, NULL}, {32, 32, 0}, 1, msgs, off, fixed_len, sig, 64};
    return run_job(&j);
}

int ed25519_verify_batch(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                         const size_t *off, size_t fixed_len)
{
    edg_job_t j = {OP_VERIFY, n, 2, {sig, pub, NULL}, {64, 32, 0}, 1, msgs, off, fixed_len, ok, 1};
    return run_job(&j);
}

int x25519_batch(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point)
{
    edg_job_t j = {OP_X25519, n, 2, {scalar, point, NULL}, {32, 32, 0}, 0, NULL, NULL, 0, out, 32};
    return run_job(&j);
}

int x25519_base_batch(size_t n, uint8_t *out, const uint8_t *scalar)
{
    edg_job_t j = {OP_X25519_BASE, n, 1, {scalar, NULL, NULL}, {32, 0, 0}, 0, NULL, NULL, 0, out, 32};
    return run_job(&j);
}

int pk_ed25519_to_x25519_batch(size_t n, uint8_t *out, const uint8_t *in)
{
    edg_job_t j = {OP_PK_CONV, n, 1, {in, NULL, NULL}, {32, 0, 0}, 0, NULL, NULL, 0, out, 32};
    return run_job(&j);
}

int sk_ed25519_to_x25519_batch(size_t n, uint8_t *out, const uint8_t *in)
{
    edg_job_t j = {OP_SK_CONV, n, 1, {in, NULL, NULL}, {32, 0, 0}, 0, NULL, NULL, 0, out, 32};
    return run_job(&j);
}

/* diagnostic: one GF(2^255-19) operation per item on the device (op: 0 mul, 1 sq, 2 add, 3 sub, 4 x121665,
 * 5 canonical form, 6 inverse, 7 ^((p-5)/8), 8 negate); a, b, out: n x 32 bytes, any 256-bit values */
int eddsa_b200_fe_selftest(size_t n, uint8_t *out, const uint8_t *a, const uint8_t *b, int op)
{
    edg_job_t j = {OP_FE_TEST, n, 2, {a, b, NULL}, {32, 32, 0}, 0, NULL, NULL, (size_t)op, out, 32};
    return run_job(&j);
}

/* ---------------------------------------------------------------------------------------------
 * device-buffer batch API (current device, asynchronous)
 * --------------------------------------------------------------------------------------------- */
static int cur_ctx(edg_dev_t **c);

/* diagnostic: read back the base-point window tables the verify kernels use on the current device */
size_t eddsa_b200_verify_tables(uint8_t *out, size_t cap)
{
    edg_dev_t *c;
    size_t bytes = edg_verify_table_bytes() - 48 * sizeof(uint32_t);   /* without the two base points at the end */
    if (!out || cap < bytes || cur_ctx(&c)) return 0;
    if (cudaMemcpy(out, c->wtab, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { fail(1, "cudaMemcpy of the window tables failed"); return 0; }
    return bytes;
}

static int cur_ctx(edg_dev_t **c)
{
    int dev = 0, rc = engine_ready();
    cudaError_t e;
    if (rc) return rc;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail((int)e, "cudaGetDevice failed: %s", cudaGetErrorString(e));
    return dev_basic(dev, c);
}

static int misaligned(const void *p) { return ((uintptr_t)p & 15) != 0; }

static int user_scratch(edg_dev_t *c, void *stream, void **buf)
{
    int i, rc = 0;
    pthread_mutex_lock(&c->us_lock);
    for (i = 0; i < EDG_MAX_USER_STREAMS; i++)
        if (c->user_scratch[i].used && c->user_scratch[i].stream == stream) { *buf = c->user_scratch[i].buf; goto out; }
    for (i = 0; i < EDG_MAX_USER_STREAMS; i++)
        if (!c->user_scratch[i].used) {
            CU(cudaMalloc(&c->user_scratch[i].buf, c->scratch_bytes));
            c->user_scratch[i].used = 1;
            c->user_scratch[i].stream = stream;
            *buf = c->user_scratch[i].buf;
            goto out;
        }
    rc = fail(EDDSA_B200_EINVAL, "too many distinct streams used with ed25519_verify_batch_dev (max %d per device)", EDG_MAX_USER_STREAMS);
out:
    pthread_mutex_unlock(&c->us_lock);
    return rc;
}

static int dev_call(edg_op_t op, size_t n, uint8_t *out, const uint8_t *a, const uint8_t *b, const uint8_t *msgs,
                    const uint64_t *off, size_t fixed_len, void *stream)
{
    edg_dev_t *c;
    uint8_t *d_in[3] = {(uint8_t *)a, (uint8_t *)b, NULL};
    void *scratch = NULL;
    int rc = cur_ctx(&c);
    if (rc) return rc;
    t_err[0] = 0;
    if (n == 0) return 0;
    if (!out || !a || misaligned(out) || misaligned(a) || (b && misaligned(b)))
        return fail(EDDSA_B200_EINVAL, "device arrays must be non-NULL and 16-byte aligned");
    if (op == OP_VERIFY) {
        rc = user_scratch(c, stream, &scratch);
        if (rc) return rc;
    }
    return launch(c, op, n, out, d_in, msgs, (const unsigned long long *)off, fixed_len, scratch, stream);
}

int ed25519_genpub_batch_dev(size_t n, uint8_t *pub, const uint8_t *sec, void *stream)
{
    return dev_call(OP_GENPUB, n, pub, sec, NULL, NULL, NULL, 0, stream);
}

int ed25519_sign_batch_dev(size_t n, uint8_t *sig, const uint8_t *sec, const uint8_t *pub, const uint8_t *msgs,
                           const uint64_t *off, size_t fixed_len, void *stream)
{
    if (n && !pub) return fail(EDDSA_B200_EINVAL, "NULL pub array");
    return dev_call(OP_SIGN, n, sig, sec, pub, msgs, off, fixed_len, stream);
}

int ed25519_verify_batch_dev(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs,
                             const uint64_t *off, size_t fixed_len, void *stream)
{
    edg_dev_t *c;
    uint8_t *d_in[3] = {(uint8_t *)sig, (uint8_t *)pub, NULL};
    void *scratch = NULL;
    int rc = cur_ctx(&c);
    if (rc) return rc;
    t_err[0] = 0;
    if (n == 0) return 0;
    if (!ok || !sig || !pub || misaligned(sig) || misaligned(pub))
        return fail(EDDSA_B200_EINVAL, "device arrays must be non-NULL and sig/pub 16-byte aligned");
    rc = user_scratch(c, stream, &scratch);
    if (rc) return rc;
    return launch(c, OP_VERIFY, n, ok, d_in, msgs, (const unsigned long long *)off, fixed_len, scratch, stream);
}

int x25519_batch_dev(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point, void *stream)
{
    if (n && !point) return fail(EDDSA_B200_EINVAL, "NULL point array");
    return dev_call(OP_X25519, n, out, scalar, point, NULL, NULL, 0, stream);
}

int x25519_base_batch_dev(size_t n, uint8_t *out, const uint8_t *scalar, void *stream)
{
    return dev_call(OP_X25519_BASE, n, out, scalar, NULL, NULL, NULL, 0, stream);
}

int pk_ed25519_to_x25519_batch_dev(size_t n, uint8_t *out, const uint8_t *in, void *stream)
{
    return dev_call(OP_PK_CONV, n, out, in, NULL, NULL, NULL, 0, stream);
}

int sk_ed25519_to_x25519_batch_dev(size_t n, uint8_t *out, const uint8_t *in, void *stream)
{
    return dev_call(OP_SK_CONV, n, out, in, NULL, NULL, NULL, 0, stream);
}

/* ---------------------------------------------------------------------------------------------
 * engine control
 * --------------------------------------------------------------------------------------------- */
int eddsa_b200_init(void)
{
    edg_dev_t *c;
    int rc = engine_ready(), g;
    if (rc) return rc;
    for (g = 0; g < g_nlist; g++) {
        rc = dev_basic(g_list[g], &c);
        if (rc) return rc;
    }
    return 0;
}

void eddsa_b200_shutdown(void)
{
    int d, i;
    if (g_init_rc || !g_nlist) return;
    for (d = 0; d < EDG_MAX_DEV; d++) {
        edg_dev_t *c = &g_dev[d];
        if (!c->ready) continue;
        pthread_mutex_lock(&c->lock);
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        for (i = 0; i < EDG_NSLOT; i++) {
            if (c->h_in[i]) cudaFreeHost(c->h_in[i]);
            if (c->h_out[i]) cudaFreeHost(c->h_out[i]);
            if (c->d_in[i]) cudaFree(c->d_in[i]);
            if (c->d_out[i]) cudaFree(c->d_out[i]);
            if (c->scratch[i]) cudaFree(c->scratch[i]);
            c->h_in[i] = c->h_out[i] = c->d_in[i] = c->d_out[i] = NULL;
            c->scratch[i] = NULL;
            cudaStreamDestroy(c->stream[i]);
            cudaEventDestroy(c->done[i]);
            cudaEventDestroy(c->in_ready[i]);
            cudaEventDestroy(c->k_done[i]);
        }
        for (i = 0; i < EDG_MAX_USER_STREAMS; i++)
            if (c->user_scratch[i].used) { cudaFree(c->user_scratch[i].buf); c->user_scratch[i].used = 0; }
        cudaStreamDestroy(c->kstream);
        if (c->wtab) cudaFree(c->wtab);
        c->wtab = NULL;
        c->in_cap = c->out_cap = 0;
        c->ready = 0;
        pthread_mutex_unlock(&c->lock);
    }
}

int eddsa_b200_device_count(void)
{
    if (engine_ready()) return 0;
    return g_nactive;
}

int eddsa_b200_set_device_count(int count)
{
    int rc = engine_ready();
    if (rc) return rc;
    if (count < 0 || count > g_nlist) return fail(EDDSA_B200_EINVAL, "device count %d outside 0..%d", count, g_nlist);
    g_nactive = count ? count : g_nlist;
    return 0;
}

unsigned long long eddsa_b200_launch_count(void) { return g_launches; }

const char *eddsa_b200_last_error(void) { return t_err; }

/* ---------------------------------------------------------------------------------------------
 * eddsa.h single-operation API: a batch of one.  void functions cannot report failure, so a device
 * error is fatal (never silently wrong, never a CPU fallback).
 * --------------------------------------------------------------------------------------------- */
static void must(int rc, const char *what)
{
    if (rc) {
        fprintf(stderr, "libeddsa_b200: %s failed (%d): %s\n", what, rc, t_err);
        abort();
    }
}

void ed25519_genpub(uint8_t pub[32], const uint8_t sec[32]) { must(ed25519_genpub_batch(1, pub, sec), "ed25519_genpub"); }

void ed25519_sign(uint8_t sig[64], const uint8_t sec[32], const uint8_t pub[32], const uint8_t *data, size_t len)
{
    static const uint8_t empty[1] = {0};
    must(ed25519_sign_batch(1, sig, sec, pub, data ? data : empty, NULL, len), "ed25519_sign");
}

bool ed25519_verify(const uint8_t sig[64], const uint8_t pub[32], const uint8_t *data, size_t len)
{
    static const uint8_t empty[1] = {0};
    uint8_t ok = 0;
    must(ed25519_verify_batch(1, &ok, sig, pub, data ? data : empty, NULL, len), "ed25519_verify");
    return ok != 0;
}

void x25519_base(uint8_t out[32], const uint8_t scalar[32]) { must(x25519_base_batch(1, out, scalar), "x25519_base"); }

void x25519(uint8_t out[32], const uint8_t scalar[32], const uint8_t point[32])
{
    must(x25519_batch(1, out, scalar, point), "x25519");
}

void pk_ed25519_to_x25519(uint8_t out[32], const uint8_t in[32]) { must(pk_ed25519_to_x25519_batch(1, out, in), "pk_ed25519_to_x25519"); }
void sk_ed25519_to_x25519(uint8_t out[32], const uint8_t in[32]) { must(sk_ed25519_to_x25519_batch(1, out, in), "sk_ed25519_to_x25519"); }

/* obsolete names (reference ed25519-sha512.c:262-324, x25519.c:236-243) */
void eddsa_genpub(uint8_t pub[32], const uint8_t sec[32]) { ed25519_genpub(pub, sec); }
void eddsa_sign(uint8_t sig[64], const uint8_t sec[32], const uint8_t pub[32], const uint8_t *data, size_t len) { ed25519_sign(sig, sec, pub, data, len); }
bool eddsa_verify(const uint8_t sig[64], const uint8_t pub[32], const uint8_t *data, size_t len) { return ed25519_verify(sig, pub, data, len); }
void DH(uint8_t out[32], const uint8_t sec[32], const uint8_t point[32]) { x25519(out, sec, point); }
void eddsa_pk_eddsa_to_dh(uint8_t out[32], const uint8_t in[32]) { pk_ed25519_to_x25519(out, in); }
void eddsa_sk_eddsa_to_dh(uint8_t out[32], const uint8_t in[32]) { sk_ed25519_to_x25519(out, in); }
