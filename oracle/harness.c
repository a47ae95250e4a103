/*
 * harness.c — TEST / BENCH INFRASTRUCTURE ONLY.
 * Threaded batch driver for a CPU implementation of the eddsa.h API (either the compiled
 * reference oracle/_ref/libeddsa_ref.so or the restatement in oracle.c), used
 *   (1) by tests to produce expected outputs for a batch quickly on all host cores, and
 *   (2) by bench.py for the cpu_baseline / --impl reference numbers: one instance per core,
 *       each thread looping the single-op API over its own contiguous slice.
 * The function pointers are passed in by the caller (ctypes), so this file binds to no library.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void (*genpub_fn)(uint8_t *, const uint8_t *);
typedef void (*sign_fn)(uint8_t *, const uint8_t *, const uint8_t *, const uint8_t *, size_t);
typedef bool (*verify_fn)(const uint8_t *, const uint8_t *, const uint8_t *, size_t);
typedef void (*x25519_fn)(uint8_t *, const uint8_t *, const uint8_t *);
typedef void (*base_fn)(uint8_t *, const uint8_t *);

enum { OP_GENPUB = 0, OP_SIGN = 1, OP_VERIFY = 2, OP_X25519 = 3, OP_X25519_BASE = 4 };

typedef struct {
    int op;
    void *fn;
    size_t lo, hi;
    uint8_t *out;          /* pub[32] / sig[64] / ok[1] / out[32] per op */
    const uint8_t *a;      /* sec / sec / sig / scalar / scalar */
    const uint8_t *b;      /* - / pub / pub / point / - */
    const uint8_t *msgs;
    const size_t *off;     /* n+1 offsets or NULL */
    size_t fixed_len;
    double seconds;        /* >0: loop over the slice repeatedly for this long (timing mode) */
    size_t done;           /* ops completed (timing mode) */
    double elapsed;
} job_t;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void run_one(job_t *j, size_t i)
{
    const uint8_t *m = NULL;
    size_t len = 0;
    if (j->op == OP_SIGN || j->op == OP_VERIFY) {
        m = j->off ? j->msgs + j->off[i] : j->msgs + i * j->fixed_len;
        len = j->off ? j->off[i + 1] - j->off[i] : j->fixed_len;
    }
    switch (j->op) {
    case OP_GENPUB: ((genpub_fn)j->fn)(j->out + 32 * i, j->a + 32 * i); break;
    case OP_SIGN: ((sign_fn)j->fn)(j->out + 64 * i, j->a + 32 * i, j->b + 32 * i, m, len); break;
    case OP_VERIFY: j->out[i] = ((verify_fn)j->fn)(j->a + 64 * i, j->b + 32 * i, m, len) ? 1 : 0; break;
    case OP_X25519: ((x25519_fn)j->fn)(j->out + 32 * i, j->a + 32 * i, j->b + 32 * i); break;
    case OP_X25519_BASE: ((base_fn)j->fn)(j->out + 32 * i, j->a + 32 * i); break;
    }
}

static void *worker(void *p)
{
    job_t *j = (job_t *)p;
    size_t i;
    double t0 = now_s();
    if (j->seconds <= 0) {
        for (i = j->lo; i < j->hi; i++) run_one(j, i);
        j->done = j->hi - j->lo;
    } else if (j->hi > j->lo) {
        i = j->lo;
        for (;;) {
            run_one(j, i);
            j->done++;
            if (++i == j->hi) i = j->lo;
            if ((j->done & 15) == 0 && now_s() - t0 >= j->seconds) break;
        }
    }
    j->elapsed = now_s() - t0;
    return NULL;
}

/*
 * Run op over n items on nthreads threads.  seconds <= 0: each item exactly once (outputs valid
 * for all n), returns wall seconds.  seconds > 0: timing mode — every thread loops over its
 * slice for that long; *ops_done gets the total and the return value is the wall time.
 */
double harness_run(int op, void *fn, int nthreads, double seconds, size_t n, uint8_t *out, const uint8_t *a,
                   const uint8_t *b, const uint8_t *msgs, const size_t *off, size_t fixed_len, size_t *ops_done)
{
    pthread_t *th;
    job_t *jobs;
    int t;
    size_t total = 0;
    double t0, wall;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n && n > 0) nthreads = (int)n;
    th = (pthread_t *)calloc(nthreads, sizeof *th);
    jobs = (job_t *)calloc(nthreads, sizeof *jobs);
    t0 = now_s();
    for (t = 0; t < nthreads; t++) {
        job_t *j = &jobs[t];
        j->op = op; j->fn = fn;
        j->lo = n * (size_t)t / nthreads;
        j->hi = n * (size_t)(t + 1) / nthreads;
        j->out = out; j->a = a; j->b = b; j->msgs = msgs; j->off = off; j->fixed_len = fixed_len;
        j->seconds = seconds;
        pthread_create(&th[t], NULL, worker, j);
    }
    for (t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        total += jobs[t].done;
    }
    wall = now_s() - t0;
    if (ops_done) *ops_done = total;
    free(th);
    free(jobs);
    return wall;
}
