// TEST INFRASTRUCTURE ONLY — a CUDA runtime SIMULATOR for the host layer (libeddsa_b200/csrc/host.c).
//
// host.c is plain C on top of the CUDA runtime API and the launchers of edg_internal.h.  This file implements both on
// the CPU so that the whole C-ABI (include/eddsa.h, include/eddsa_batch.h) — device sharding, worker threads, chunk
// schedule, slot rotation, pinned / pageable staging, event dependencies, scrubbing of secrets, error paths — can be
// exercised by `pytest -m "not gpu"` in a container without a GPU.  tests/host_sim/Makefile links host.c (unchanged,
// compiled from where it lies) with this file into tests/host_sim/libeddsa_sim.so; nothing in the product links,
// loads or mentions it, and the product library has no CPU path of any kind.
//
//   * streams are queues of operations that run LAZILY: an operation executes only when the host waits for it
//     (cudaEventSynchronize / cudaStreamSynchronize / cudaDeviceSynchronize / cudaFree) or when an operation that
//     depends on it through cudaStreamWaitEvent executes.  This is the most adversarial legal schedule: a missing event
//     dependency, a staging buffer reused before its copy ran, or an output read before its D2H copy all give wrong
//     results (CUDASIM_SCHEDULE=eager runs everything at once instead, =random something in between, =others-first
//     runs every other stream as far as it can before the awaited one advances: work without a dependency runs EARLY);
//   * every copy, memset and kernel argument is bounds-checked against the live allocations of the right device
//     (violations are counted and described: cudasim_error_count / cudasim_first_error); fresh and freed memory is
//     poisoned; asynchronous copies from or to pageable memory (which the real runtime serialises) are counted;
//   * the "kernels" run the per-thread operation bodies of ops.cuh compiled for the host — the same code the CUDA
//     kernels execute per thread — so results are bit-exact and can be checked against the oracle;
//   * any runtime call can be made to fail (cudasim_fail), including a kernel that faults asynchronously and poisons
//     the device until cudasim_clear_faults().
#include <cuda_runtime_api.h>

#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../libeddsa_b200/csrc/ops.cuh"
#include "../../libeddsa_b200/csrc/edg_internal.h"

using namespace edg;

namespace {

enum { K_DEVICE = 0, K_PINNED = 1 };
enum { API_MALLOC, API_MALLOC_HOST, API_MEMCPY_ASYNC, API_MEMSET_ASYNC, API_EVENT_RECORD, API_EVENT_SYNC, API_STREAM_WAIT,
       API_POOL_ALLOC, API_FREE_ASYNC, API_STREAM_CREATE, API_EVENT_CREATE, API_LAUNCH, API_SET_DEVICE, API_POOL_CREATE,
       API_KERNEL_FAULT, API_STREAM_SYNC, API_COUNT };
enum { SCHED_LAZY, SCHED_EAGER, SCHED_RANDOM, SCHED_OTHERS_FIRST };
constexpr int MAX_DEV = 16;

struct Alloc { size_t size; int kind, dev; bool lib; };
struct Stream;
struct Event { Stream *rec_stream = nullptr; uint64_t rec_ticket = 0, done_ticket = 0; };
struct Op {
    enum Type { COPY, MEMSET, KERNEL, RECORD, WAIT, FREE } type;
    void *dst = nullptr; const void *src = nullptr; size_t bytes = 0; int val = 0; int kind = 0;
    std::vector<uint8_t> snapshot;                  // pageable source of an async H2D copy: taken at call time, as the runtime does
    std::function<void()> fn;
    Event *ev = nullptr; uint64_t ticket = 0; Stream *other = nullptr;
};
struct Stream { int dev; std::deque<Op> q; };
struct Stats { uint64_t h2d_bytes, d2h_bytes, h2d_copies, d2h_copies, kernels, pageable_async, memsets, dev_allocs, host_allocs, pool_allocs, max_chunk_items, event_syncs; };
struct Launch { int dev, op; uint64_t n; };

std::recursive_mutex G;
std::map<uintptr_t, Alloc> g_allocs;
std::vector<Stream *> g_streams;
Stream *g_default[MAX_DEV];
int g_ndev = 4, g_sms = 2, g_sched = SCHED_LAZY, g_threads = 8;
bool g_configured = false;
thread_local int t_dev = 0;
thread_local cudaError_t t_last = cudaSuccess;
uint64_t g_ticket = 0, g_errors = 0, g_rng = 0x9e3779b97f4a7c15ull;
std::string g_first_error;
Stats g_stats;
std::vector<Launch> g_launch_log;
long g_fail_countdown[API_COUNT];
int g_fault[MAX_DEV];
int g_waves = 16, g_full_scalars = 0;
bool g_keep_freed = false;
int g_resident = 4 * 128;          // resident threads per SM of the verify loop kernel; CUDASIM_RESIDENT: smaller waves make the
                                   // verify chunk schedule cheap to exercise

void sim_error(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (g_errors++ == 0) g_first_error = buf;
    if (getenv("CUDASIM_VERBOSE")) fprintf(stderr, "cudasim: %s\n", buf);
}

void configure() {
    if (g_configured) return;
    g_configured = true;
    if (getenv("CUDASIM_DEVICES")) g_ndev = atoi(getenv("CUDASIM_DEVICES"));
    if (getenv("CUDASIM_SMS")) g_sms = atoi(getenv("CUDASIM_SMS"));
    if (getenv("CUDASIM_THREADS")) g_threads = atoi(getenv("CUDASIM_THREADS"));
    if (getenv("CUDASIM_RESIDENT")) g_resident = atoi(getenv("CUDASIM_RESIDENT"));
    if (getenv("CUDASIM_KEEP_FREED")) g_keep_freed = atoi(getenv("CUDASIM_KEEP_FREED")) != 0;
    if (g_ndev > MAX_DEV) g_ndev = MAX_DEV;
    if (g_sms < 1) g_sms = 1;
    if (g_threads < 1) g_threads = 1;
    const char *s = getenv("CUDASIM_SCHEDULE");
    if (s && !strcmp(s, "eager")) g_sched = SCHED_EAGER;
    if (s && !strcmp(s, "random")) g_sched = SCHED_RANDOM;
    if (s && !strcmp(s, "others-first")) g_sched = SCHED_OTHERS_FIRST;
    for (int d = 0; d < MAX_DEV; d++) { g_default[d] = new Stream{d, {}}; g_streams.push_back(g_default[d]); }
}

// should this call fail?  (cudasim_fail(api, nth): the nth call from now does, once)
bool inject(int api) {
    if (g_fail_countdown[api] > 0 && --g_fail_countdown[api] == 0) return true;
    return false;
}

cudaError_t ret(cudaError_t e) { if (e != cudaSuccess) t_last = e; return e; }

const Alloc *find_alloc(const void *p, size_t bytes, uintptr_t *base = nullptr) {
    const uintptr_t a = (uintptr_t)p;
    auto it = g_allocs.upper_bound(a);
    if (it == g_allocs.begin()) return nullptr;
    --it;
    if (a < it->first || a + bytes > it->first + it->second.size) return nullptr;
    if (base) *base = it->first;
    return &it->second;
}

bool need_dev(const void *p, size_t bytes, int dev, const char *what, size_t align = 1) {
    if (bytes == 0) return true;
    const Alloc *al = find_alloc(p, bytes);
    if (!al || al->kind != K_DEVICE) { sim_error("%s: %zu bytes at %p are not inside a live device allocation", what, bytes, p); return false; }
    if (al->dev != dev) { sim_error("%s: memory of device %d used on device %d", what, al->dev, dev); return false; }
    if ((uintptr_t)p % align) { sim_error("%s: pointer %p is not %zu-byte aligned", what, p, align); return false; }
    return true;
}

void *sim_alloc(size_t size, int kind, int dev, bool lib, uint8_t poison) {
    void *p = nullptr;
    if (posix_memalign(&p, 512, size ? size : 1) != 0) return nullptr;
    memset(p, poison, size);
    g_allocs[(uintptr_t)p] = Alloc{size, kind, dev, lib};
    return p;
}

bool sim_free(void *p, int kind) {
    auto it = g_allocs.find((uintptr_t)p);
    if (it == g_allocs.end() || it->second.kind != kind) { sim_error("free of %p: not the start of a live %s allocation", p, kind == K_DEVICE ? "device" : "pinned"); return false; }
    memset(p, 0xDE, it->second.size);
    g_allocs.erase(it);
    free(p);
    return true;
}

void exec_front(Stream *s);

// SCHED_OTHERS_FIRST: before the stream the host waits for advances by one operation, every OTHER stream runs as far as
// it can without help (a wait on an event whose record has not executed blocks it).  Work that lacks a dependency on
// the driven stream therefore runs as EARLY as possible — the complement of the lazy schedule, where it runs late.
bool g_pumping = false;
std::vector<Stream *> g_driving;       // streams with an operation in progress (a wait that is driving its dependency): not to be advanced
void pump_others() {
    if (g_pumping) return;
    g_pumping = true;
    for (bool progress = true; progress;) {
        progress = false;
        for (size_t i = 0; i < g_streams.size(); i++) {
            Stream *t = g_streams[i];
            bool busy = false;
            for (Stream *d : g_driving) busy |= d == t;
            if (busy) continue;
            while (!t->q.empty()) {
                const Op &f = t->q.front();
                if (f.type == Op::WAIT && f.ev->done_ticket < f.ticket) break;
                exec_front(t);
                progress = true;
            }
        }
    }
    g_pumping = false;
}

template <typename Pred>
void drive_until(Stream *s, Pred done) {
    g_driving.push_back(s);
    while (!done() && !s->q.empty()) {
        if (g_sched == SCHED_OTHERS_FIRST) pump_others();
        if (!s->q.empty()) exec_front(s);
    }
    g_driving.pop_back();
}

void drive_all(Stream *s) { drive_until(s, [] { return false; }); }

void exec_front(Stream *s) {
    Op op = std::move(s->q.front());
    s->q.pop_front();
    switch (op.type) {
    case Op::COPY: {
        const void *src = op.snapshot.empty() ? op.src : op.snapshot.data();
        if (op.kind == cudaMemcpyHostToDevice) {
            if (!need_dev(op.dst, op.bytes, s->dev, "H2D copy destination")) break;
            g_stats.h2d_bytes += op.bytes; g_stats.h2d_copies++;
        } else {
            if (!need_dev(op.src, op.bytes, s->dev, "D2H copy source")) break;
            g_stats.d2h_bytes += op.bytes; g_stats.d2h_copies++;
        }
        memcpy(op.dst, src, op.bytes);
        break;
    }
    case Op::MEMSET:
        if (need_dev(op.dst, op.bytes, s->dev, "memset")) memset(op.dst, op.val, op.bytes);
        g_stats.memsets++;
        break;
    case Op::KERNEL:
        if (g_fault[s->dev]) break;                  // a faulted device executes nothing more
        if (inject(API_KERNEL_FAULT)) { g_fault[s->dev] = (int)cudaErrorIllegalAddress; break; }
        g_stats.kernels++;
        op.fn();
        break;
    case Op::RECORD:
        if (op.ev->done_ticket < op.ticket) op.ev->done_ticket = op.ticket;
        break;
    case Op::WAIT:
        if (op.other) drive_until(op.other, [&] { return op.ev->done_ticket >= op.ticket; });
        if (op.ev->done_ticket < op.ticket) sim_error("stream wait on an event whose record can never complete");
        break;
    case Op::FREE:
        // CUDASIM_KEEP_FREED=1: a block returned to the pool keeps its contents and stays visible to cudasim_find, as pool memory
        // does on the device — what a kernel left in its scratch can then be searched for secrets
        if (!g_keep_freed) sim_free(op.dst, K_DEVICE);
        break;
    }
}

uint64_t rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return g_rng; }

// called after an operation has been queued: the schedule decides how much runs now
void after_enqueue(Stream *s) {
    if (g_sched == SCHED_EAGER) drive_all(s);
    else if (g_sched == SCHED_RANDOM) {
        Stream *v = g_streams[rnd() % g_streams.size()];
        size_t k = rnd() % 4;
        while (k-- && !v->q.empty()) exec_front(v);
    }
}

Stream *stream_of(cudaStream_t st) { return st ? (Stream *)st : g_default[t_dev]; }

void enqueue(Stream *s, Op &&op) { s->q.push_back(std::move(op)); after_enqueue(s); }

void parallel_for(size_t n, const std::function<void(size_t)> &body) {
    const size_t nt = n < 8 ? 1 : n < (size_t)g_threads ? n : (size_t)g_threads;
    if (nt == 1) { for (size_t i = 0; i < n; i++) body(i); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&, t] { for (size_t i = n * t / nt; i < n * (t + 1) / nt; i++) body(i); });
    for (auto &x : th) x.join();
}

#if !defined(CUDASIM_REAL_KERNELS)   // (the flavour with the REAL kernels links the rewritten kernels_*.cu instead: simt_emul.h)
// ---- the tables the kernels read, built exactly as the device builds them (k_wtab_base / k_wtab_build, k_comb_*) ----
const u32 *host_comb() {
    static u32 *tab = nullptr;
    if (!tab) {
        tab = new u32[EDG_COMB_WORDS];
        for (int j = 0; j < EDG_COMB_ROWS; j++) {
            u32 base[24];
            wtab_base(base, EDG_COMB_W * j);
            comb_table_row(tab + (size_t)j * EDG_COMB_ENTRIES * 24, base);
        }
    }
    return tab;
}

const u32 *host_wtab() {
    static u32 *tab = nullptr;
    if (!tab) {
        tab = new u32[2 * (size_t)EDG_WTAB_WORDS + 48];
        for (int m = 0; m < 2; m++) {
            u32 *base = tab + 2 * (size_t)EDG_WTAB_WORDS + 24 * m, *tb = tab + m * (size_t)EDG_WTAB_WORDS;
            wtab_base(base, 128 * m);
            for (int i = 0; i < 24; i++) tb[i] = (i == 0 || i == 8) ? 1u : 0u;
            const size_t groups = (EDG_WTAB_ENTRIES - 1) / 8;
            parallel_for(groups, [&](size_t g) { wtab_build8(tb + 24u * (8u * (u32)g + 1u), base, 8u * (u32)g + 1u); });
        }
    }
    return tab;
}

// mirrors of constants private to the kernel translation units (kernels_fixedbase.cu: kPass; kernels_verify.cu: resident
// blocks x threads of the loop kernel on sm_100a — the default of g_resident)
constexpr size_t kPass = (size_t)1 << 21;

enum { L_X25519, L_X25519_BASE, L_GENPUB, L_SIGN, L_VERIFY, L_PK_CONV, L_SK_CONV, L_FE_TEST, L_SC_TEST, L_TABLES };

int launch(void *stream, int what, size_t n, std::function<void(int)> body) {
    std::lock_guard<std::recursive_mutex> lk(G);
    configure();
    Stream *s = stream_of((cudaStream_t)stream);
    if (g_fault[s->dev]) return g_fault[s->dev];
    if (inject(API_LAUNCH)) return (int)cudaErrorLaunchFailure;
    g_launch_log.push_back(Launch{s->dev, what, n});
    if (what != L_TABLES && n > g_stats.max_chunk_items) g_stats.max_chunk_items = n;
    Op op;
    op.type = Op::KERNEL;
    const int dev = s->dev;
    op.fn = [body, dev] { body(dev); };
    enqueue(s, std::move(op));
    return 0;
}

void msg_of(const uint8_t *&m, u64 &len, const uint8_t *msgs, const unsigned long long *off, unsigned long long fixed_len, size_t i) {
    if (off) { m = msgs + off[i]; len = off[i + 1] - off[i]; }
    else { m = msgs + i * fixed_len; len = fixed_len; }
}

bool need_msgs(const uint8_t *msgs, const unsigned long long *off, unsigned long long fixed_len, size_t n, int dev, const char *what) {
    if (off) {
        if (!need_dev(off, (n + 1) * sizeof *off, dev, what, 8)) return false;
        for (size_t i = 0; i < n; i++)
            if (off[i + 1] < off[i]) { sim_error("%s: offsets decrease at %zu", what, i); return false; }
        return need_dev(msgs + off[0], off[n] - off[0], dev, what);
    }
    return need_dev(msgs, n * fixed_len, dev, what);
}

#endif
}  // namespace

#if !defined(CUDASIM_REAL_KERNELS)
// =====================================================================================================================
// launchers (edg_internal.h)
// =====================================================================================================================
extern "C" {

int edg_kernels_init(void) { return 0; }
size_t edg_comb_table_payload_bytes(void) { return (size_t)EDG_COMB_WORDS * sizeof(u32); }
size_t edg_comb_table_bytes(void) { return (2 * (size_t)EDG_COMB_WORDS + 24u * EDG_COMB_ROWS) * sizeof(u32); }
void edg_comb_geometry(int *rows, int *entries) { *rows = EDG_COMB_ROWS; *entries = EDG_COMB_ENTRIES; }
size_t edg_fixedbase_scratch_bytes(int is_sign, size_t n) { return (n < kPass ? n : kPass) * (is_sign ? 64 : 32); }
size_t edg_verify_pass(int sm_count) { return (size_t)sm_count * g_resident * g_waves; }
void edg_verify_set_waves(int waves) { g_waves = waves < 1 ? 1 : waves > 16 ? 16 : waves; }
unsigned edg_verify_waves(void) { return (unsigned)g_waves; }
size_t edg_verify_record_bytes(void) { return EDG_VSTATE_WORDS * sizeof(u32); }
static size_t perm_offset(size_t records) { return records * EDG_VSTATE_WORDS * sizeof(u32); }
static size_t counters_offset(size_t records) { return perm_offset(records) + ((records * sizeof(unsigned) + 255) & ~(size_t)255); }
size_t edg_verify_scratch_bytes(size_t records) { return counters_offset(records) + 256; }
size_t edg_verify_table_bytes(void) { return (2 * (size_t)EDG_WTAB_WORDS + 48) * sizeof(u32); }
void edg_verify_debug_full_scalars(int on) { g_full_scalars = on; }

int edg_comb_table_init(void *table, void *stream) {
    return launch(stream, L_TABLES, 0, [=](int dev) {
        if (!need_dev(table, edg_comb_table_bytes(), dev, "comb table", 16)) return;
        u32 *t = (u32 *)table, *bases = t + EDG_COMB_WORDS, *mma = bases + 24 * EDG_COMB_ROWS;
        memcpy(t, host_comb(), (size_t)EDG_COMB_WORDS * sizeof(u32));
        for (int j = 0; j < EDG_COMB_ROWS; j++) wtab_base(bases + 24 * j, EDG_COMB_W * j);
        for (unsigned i = 0; i < EDG_COMB_WORDS; i++) mma[i] = comb_mma_word(t, i);
    });
}

int edg_verify_table_init(void *table, void *stream) {
    return launch(stream, L_TABLES, 0, [=](int dev) {
        if (!need_dev(table, edg_verify_table_bytes(), dev, "window tables", 16)) return;
        memcpy(table, host_wtab(), edg_verify_table_bytes());
    });
}

int edg_launch_x25519(size_t n, uint8_t *out, const uint8_t *scalar, const uint8_t *point, int, void *stream) {
    if (n == 0) return 0;
    return launch(stream, L_X25519, n, [=](int dev) {
        if (!need_dev(out, 32 * n, dev, "x25519 out", 16) || !need_dev(scalar, 32 * n, dev, "x25519 scalar", 16) || !need_dev(point, 32 * n, dev, "x25519 point", 16)) return;
        parallel_for(n, [&](size_t i) {
            u32 o[8], s[8], p[8];
            memcpy(s, scalar + 32 * i, 32); memcpy(p, point + 32 * i, 32);
            x25519_op(o, s, p);
            memcpy(out + 32 * i, o, 32);
        });
    });
}

int edg_launch_x25519_base(size_t n, uint8_t *out, const uint8_t *scalar, const void *comb, int, void *stream) {
    if (n == 0) return 0;
    return launch(stream, L_X25519_BASE, n, [=](int dev) {
        if (!need_dev(out, 32 * n, dev, "x25519_base out", 16) || !need_dev(scalar, 32 * n, dev, "x25519_base scalar", 16) ||
            !need_dev(comb, edg_comb_table_bytes(), dev, "comb table", 16)) return;
        parallel_for(n, [&](size_t i) {
            u32 o[8], s[8];
            memcpy(s, scalar + 32 * i, 32);
            x25519_base_op(o, s, (const u32 *)comb);
            memcpy(out + 32 * i, o, 32);
        });
    });
}

int edg_launch_genpub(size_t n, uint8_t *pub, const uint8_t *sec, void *scratch, const void *comb, int, void *stream, unsigned *launches) {
    for (size_t first = 0; first < n; first += kPass) *launches += 2;
    if (n == 0) return 0;
    return launch(stream, L_GENPUB, n, [=](int dev) {
        if (!need_dev(pub, 32 * n, dev, "genpub pub", 16) || !need_dev(sec, 32 * n, dev, "genpub sec", 16) ||
            !need_dev(scratch, edg_fixedbase_scratch_bytes(0, n), dev, "genpub scratch", 16) || !need_dev(comb, edg_comb_table_bytes(), dev, "comb table", 16)) return;
        parallel_for(n, [&](size_t i) {
            u32 o[8];
            ed25519_genpub_op(o, sec + 32 * i, (const u32 *)comb);
            memcpy(pub + 32 * i, o, 32);
        });
        memset(scratch, 0, edg_fixedbase_scratch_bytes(0, n));      // the kernels leave their scalar scratch wiped
    });
}

int edg_launch_sign(size_t n, uint8_t *sig, const uint8_t *sec, const uint8_t *pub, const uint8_t *msgs, const unsigned long long *off,
                    unsigned long long fixed_len, void *scratch, const void *comb, int, void *stream, unsigned *launches) {
    for (size_t first = 0; first < n; first += kPass) *launches += 3;
    if (n == 0) return 0;
    return launch(stream, L_SIGN, n, [=](int dev) {
        if (!need_dev(sig, 64 * n, dev, "sign sig", 16) || !need_dev(sec, 32 * n, dev, "sign sec", 16) || !need_dev(pub, 32 * n, dev, "sign pub", 16) ||
            !need_msgs(msgs, off, fixed_len, n, dev, "sign messages") ||
            !need_dev(scratch, edg_fixedbase_scratch_bytes(1, n), dev, "sign scratch", 16) || !need_dev(comb, edg_comb_table_bytes(), dev, "comb table", 16)) return;
        parallel_for(n, [&](size_t i) {
            u32 o[16], p[8];
            const uint8_t *m; u64 len;
            msg_of(m, len, msgs, off, fixed_len, i);
            memcpy(p, pub + 32 * i, 32);
            ed25519_sign_op(o, sec + 32 * i, p, m, len, (const u32 *)comb);
            memcpy(sig + 64 * i, o, 64);
        });
        memset(scratch, 0, edg_fixedbase_scratch_bytes(1, n));
    });
}

int edg_launch_verify(size_t n, uint8_t *ok, const uint8_t *sig, const uint8_t *pub, const uint8_t *msgs, const unsigned long long *off,
                      unsigned long long fixed_len, void *scratch, size_t records, const void *table, int, void *stream, unsigned *launches) {
    if (records == 0) return (int)cudaErrorInvalidValue;
    for (size_t first = 0; first < n; first += records) *launches += 3;
    if (n == 0) return 0;
    const int full = g_full_scalars;
    return launch(stream, L_VERIFY, n, [=](int dev) {
        if (!need_dev(ok, n, dev, "verify ok") || !need_dev(sig, 64 * n, dev, "verify sig", 16) || !need_dev(pub, 32 * n, dev, "verify pub", 16) ||
            !need_msgs(msgs, off, fixed_len, n, dev, "verify messages") ||
            !need_dev(scratch, edg_verify_scratch_bytes(records), dev, "verify scratch", 16) || !need_dev(table, edg_verify_table_bytes(), dev, "window tables", 16)) return;
        memset(scratch, 0xEE, edg_verify_scratch_bytes(records));   // the records, the permutation and the counters are written by every pass
        parallel_for(n, [&](size_t i) {
            u32 s[16], p[8], qtab[EDG_VSTATE_WORDS];
            const uint8_t *m; u64 len;
            msg_of(m, len, msgs, off, fixed_len, i);
            memcpy(s, sig + 64 * i, 64); memcpy(p, pub + 32 * i, 32);
            ok[i] = (uint8_t)ed25519_verify_op(s, p, m, len, qtab, (const u32 *)table, full != 0);
        });
    });
}

int edg_launch_pk_convert(size_t n, uint8_t *out, const uint8_t *in, int, void *stream) {
    if (n == 0) return 0;
    return launch(stream, L_PK_CONV, n, [=](int dev) {
        if (!need_dev(out, 32 * n, dev, "pk_convert out", 16) || !need_dev(in, 32 * n, dev, "pk_convert in", 16)) return;
        parallel_for(n, [&](size_t i) { u32 o[8], p[8]; memcpy(p, in + 32 * i, 32); pk_ed25519_to_x25519_op(o, p); memcpy(out + 32 * i, o, 32); });
    });
}

int edg_launch_sk_convert(size_t n, uint8_t *out, const uint8_t *in, int, void *stream) {
    if (n == 0) return 0;
    return launch(stream, L_SK_CONV, n, [=](int dev) {
        if (!need_dev(out, 32 * n, dev, "sk_convert out", 16) || !need_dev(in, 32 * n, dev, "sk_convert in", 16)) return;
        parallel_for(n, [&](size_t i) { u32 o[8]; sk_ed25519_to_x25519_op(o, in + 32 * i); memcpy(out + 32 * i, o, 32); });
    });
}

int edg_launch_fe_selftest(size_t n, uint8_t *out, const uint8_t *a, const uint8_t *b, int op, int, void *stream) {
    if (n == 0) return 0;
    return launch(stream, L_FE_TEST, n, [=](int dev) {
        if (!need_dev(out, 32 * n, dev, "fe_selftest out", 16) || !need_dev(a, 32 * n, dev, "fe_selftest a", 16) || !need_dev(b, 32 * n, dev, "fe_selftest b", 16)) return;
        if (op == 9) {
            const size_t groups = (n + EDG_BATCH - 1) / EDG_BATCH;
            parallel_for(groups, [&](size_t g) {
                fe z[EDG_BATCH];
                const int cnt = (int)(n - g * EDG_BATCH < EDG_BATCH ? n - g * EDG_BATCH : EDG_BATCH);
                for (int k = 0; k < cnt; k++) memcpy(z[k].v, a + 32 * (g * EDG_BATCH + k), 32);
                fe_batch_inv(z, cnt);
                for (int k = 0; k < cnt; k++) { u32 w[8]; fe_to_words(w, z[k]); memcpy(out + 32 * (g * EDG_BATCH + k), w, 32); }
            });
            return;
        }
        parallel_for(n, [&](size_t i) {
            fe x, y, r;
            memcpy(x.v, a + 32 * i, 32); memcpy(y.v, b + 32 * i, 32);
            switch (op) {
            case 0: fe_mul(r, x, y); break;
            case 1: fe_sq(r, x); break;
            case 2: fe_add(r, x, y); break;
            case 3: fe_sub(r, x, y); break;
            case 4: fe_mul121665(r, x); break;
            case 5: fe_canon(r, x); break;
            case 6: fe_inv(r, x); break;
            case 7: fe_pow2523(r, x); break;
            case 8: fe_neg(r, x); break;
            default: fe_copy(r, x); break;
            }
            memcpy(out + 32 * i, r.v, 32);
        });
    });
}

int edg_launch_sc_selftest(size_t n, uint8_t *out, const uint8_t *a, const uint8_t *b, const uint8_t *c, int op, int, void *stream) {
    if (n == 0) return 0;
    return launch(stream, L_SC_TEST, n, [=](int dev) {
        if (!need_dev(out, 32 * n, dev, "sc_selftest out", 16) || !need_dev(a, 32 * n, dev, "sc_selftest a", 16) || !need_dev(b, 32 * n, dev, "sc_selftest b", 16) ||
            !need_dev(c, 32 * n, dev, "sc_selftest c", 16)) return;
        parallel_for(n, [&](size_t i) {
            u32 x[16], z[8], r[8];
            memcpy(x, a + 32 * i, 32); memcpy(x + 8, b + 32 * i, 32); memcpy(z, c + 32 * i, 32);
            if (op == 0) sc_reduce512(r, x);
            else if (op == 1) sc_reduce256(r, x);
            else sc_muladd(r, x, x + 8, z);
            memcpy(out + 32 * i, r, 32);
        });
    });
}

#else
extern "C" {
// the REAL kernels (tests/host_sim/simt_emul.h): a launch arrives as a closure that runs the whole grid
int cudasim_enqueue_kernel(void *stream, void (*fn)(void *), void *arg) {
    std::lock_guard<std::recursive_mutex> lk(G);
    configure();
    Stream *s = stream_of((cudaStream_t)stream);
    if (g_fault[s->dev] || inject(API_LAUNCH)) { t_last = g_fault[s->dev] ? (cudaError_t)g_fault[s->dev] : cudaErrorLaunchFailure; return (int)t_last; }
    Op op;
    op.type = Op::KERNEL;
    op.fn = [fn, arg] { fn(arg); };
    enqueue(s, std::move(op));
    return 0;
}
#endif

// =====================================================================================================================
// CUDA runtime API (the subset host.c uses)
// =====================================================================================================================
#define LOCK std::lock_guard<std::recursive_mutex> lk(G); configure()

const char *cudaGetErrorString(cudaError_t e) {
    static thread_local char buf[64];
    if (e == cudaSuccess) return "no error";
    if (e == cudaErrorMemoryAllocation) return "out of memory (simulated)";
    if (e == cudaErrorNoDevice) return "no CUDA-capable device is detected (simulated)";
    if (e == cudaErrorIllegalAddress) return "an illegal memory access was encountered (simulated)";
    if (e == cudaErrorLaunchFailure) return "unspecified launch failure (simulated)";
    snprintf(buf, sizeof buf, "simulated CUDA error %d", (int)e);
    return buf;
}

cudaError_t cudaGetLastError(void) { cudaError_t e = t_last; t_last = cudaSuccess; return e; }

cudaError_t cudaGetDeviceCount(int *count) {
    LOCK;
    *count = g_ndev;
    return g_ndev > 0 ? cudaSuccess : ret(cudaErrorNoDevice);
}

cudaError_t cudaGetDevice(int *dev) { *dev = t_dev; return cudaSuccess; }

cudaError_t cudaSetDevice(int dev) {
    LOCK;
    if (dev < 0 || dev >= g_ndev) return ret(cudaErrorInvalidDevice);
    if (inject(API_SET_DEVICE)) return ret(cudaErrorDevicesUnavailable);
    t_dev = dev;
    return cudaSuccess;
}

cudaError_t cudaDeviceGetAttribute(int *value, enum cudaDeviceAttr attr, int device) {
    LOCK;
    if (device < 0 || device >= g_ndev) return ret(cudaErrorInvalidDevice);
    if (attr != cudaDevAttrMultiProcessorCount) return ret(cudaErrorInvalidValue);
    *value = g_sms;
    return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned int) {
    LOCK;
    if (inject(API_STREAM_CREATE)) return ret(cudaErrorMemoryAllocation);
    Stream *s = new Stream{t_dev, {}};
    g_streams.push_back(s);
    *st = (cudaStream_t)s;
    return cudaSuccess;
}

cudaError_t cudaStreamDestroy(cudaStream_t st) {
    LOCK;
    Stream *s = (Stream *)st;
    drive_all(s);                                   // the runtime lets queued work finish
    for (size_t i = 0; i < g_streams.size(); i++)
        if (g_streams[i] == s) { g_streams.erase(g_streams.begin() + i); break; }
    // the object stays allocated: events recorded on it may still name it
    return cudaSuccess;
}

cudaError_t cudaStreamSynchronize(cudaStream_t st) {
    LOCK;
    Stream *s = stream_of(st);
    if (inject(API_STREAM_SYNC)) return ret(cudaErrorLaunchFailure);
    drive_all(s);
    return g_fault[s->dev] ? ret((cudaError_t)g_fault[s->dev]) : cudaSuccess;
}

cudaError_t cudaDeviceSynchronize(void) {
    LOCK;
    for (size_t i = 0; i < g_streams.size(); i++)
        if (g_streams[i]->dev == t_dev) drive_all(g_streams[i]);
    return g_fault[t_dev] ? ret((cudaError_t)g_fault[t_dev]) : cudaSuccess;
}

cudaError_t cudaEventCreateWithFlags(cudaEvent_t *ev, unsigned int) {
    LOCK;
    if (inject(API_EVENT_CREATE)) return ret(cudaErrorMemoryAllocation);
    *ev = (cudaEvent_t) new Event();
    return cudaSuccess;
}

cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }   // kept alive: queued operations may still name it

cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st) {
    LOCK;
    Event *e = (Event *)ev;
    Stream *s = stream_of(st);
    if (inject(API_EVENT_RECORD)) return ret(cudaErrorLaunchFailure);
    e->rec_stream = s;
    e->rec_ticket = ++g_ticket;
    Op op;
    op.type = Op::RECORD; op.ev = e; op.ticket = e->rec_ticket;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned int) {
    LOCK;
    Event *e = (Event *)ev;
    Stream *s = stream_of(st);
    if (inject(API_STREAM_WAIT)) return ret(cudaErrorLaunchFailure);
    if (!e->rec_stream) return cudaSuccess;          // never recorded: no dependency
    Op op;
    op.type = Op::WAIT; op.ev = e; op.ticket = e->rec_ticket; op.other = e->rec_stream;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaEventSynchronize(cudaEvent_t ev) {
    LOCK;
    Event *e = (Event *)ev;
    g_stats.event_syncs++;
    if (inject(API_EVENT_SYNC)) return ret(cudaErrorLaunchFailure);
    if (!e->rec_stream) return cudaSuccess;
    const uint64_t want = e->rec_ticket;
    drive_until(e->rec_stream, [&] { return e->done_ticket >= want; });
    return g_fault[e->rec_stream->dev] ? ret((cudaError_t)g_fault[e->rec_stream->dev]) : cudaSuccess;
}

cudaError_t cudaMalloc(void **p, size_t size) {
    LOCK;
    if (inject(API_MALLOC)) return ret(cudaErrorMemoryAllocation);
    *p = sim_alloc(size, K_DEVICE, t_dev, true, 0xDB);
    g_stats.dev_allocs++;
    return *p ? cudaSuccess : ret(cudaErrorMemoryAllocation);
}

cudaError_t cudaMallocHost(void **p, size_t size) {
    LOCK;
    if (inject(API_MALLOC_HOST)) return ret(cudaErrorMemoryAllocation);
    *p = sim_alloc(size, K_PINNED, t_dev, true, 0xDC);
    g_stats.host_allocs++;
    return *p ? cudaSuccess : ret(cudaErrorMemoryAllocation);
}

static void sync_device_of(int dev) {
    for (size_t i = 0; i < g_streams.size(); i++)
        if (g_streams[i]->dev == dev) drive_all(g_streams[i]);
}

cudaError_t cudaFree(void *p) {
    LOCK;
    if (!p) return cudaSuccess;
    const Alloc *al = find_alloc(p, 0);
    if (al) sync_device_of(al->dev);               // cudaFree synchronises the device
    return sim_free(p, K_DEVICE) ? cudaSuccess : ret(cudaErrorInvalidValue);
}

cudaError_t cudaFreeHost(void *p) {
    LOCK;
    if (!p) return cudaSuccess;
    const Alloc *al = find_alloc(p, 0);
    if (al) sync_device_of(al->dev);
    return sim_free(p, K_PINNED) ? cudaSuccess : ret(cudaErrorInvalidValue);
}

cudaError_t cudaMemPoolCreate(cudaMemPool_t *pool, const struct cudaMemPoolProps *props) {
    LOCK;
    if (inject(API_POOL_CREATE)) return ret(cudaErrorMemoryAllocation);
    if (props->location.type != cudaMemLocationTypeDevice || props->location.id != t_dev) sim_error("memory pool created for device %d while device %d is current", props->location.id, t_dev);
    *pool = (cudaMemPool_t)(new int(props->location.id));
    return cudaSuccess;
}

cudaError_t cudaMemPoolDestroy(cudaMemPool_t pool) { delete (int *)pool; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, enum cudaMemPoolAttr, void *) { return cudaSuccess; }

cudaError_t cudaMallocFromPoolAsync(void **p, size_t size, cudaMemPool_t pool, cudaStream_t st) {
    LOCK;
    Stream *s = stream_of(st);
    if (inject(API_POOL_ALLOC)) return ret(cudaErrorMemoryAllocation);
    if (*(int *)pool != s->dev) sim_error("pool of device %d used on a stream of device %d", *(int *)pool, s->dev);
    *p = sim_alloc(size, K_DEVICE, s->dev, true, 0xDD);
    g_stats.pool_allocs++;
    return *p ? cudaSuccess : ret(cudaErrorMemoryAllocation);
}

cudaError_t cudaFreeAsync(void *p, cudaStream_t st) {
    LOCK;
    Stream *s = stream_of(st);
    if (inject(API_FREE_ASYNC)) return ret(cudaErrorLaunchFailure);
    Op op;
    op.type = Op::FREE; op.dst = p;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes *a, const void *p) {
    LOCK;
    const Alloc *al = find_alloc(p, 1);
    memset(a, 0, sizeof *a);
    a->type = !al ? cudaMemoryTypeUnregistered : al->kind == K_PINNED ? cudaMemoryTypeHost : cudaMemoryTypeDevice;
    a->device = al ? al->dev : -1;
    return cudaSuccess;
}

cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind) {
    LOCK;
    // legacy default stream: does not wait for non-blocking streams, so nothing is driven here
    if (kind == cudaMemcpyDeviceToHost && !need_dev(src, bytes, t_dev, "cudaMemcpy source")) return ret(cudaErrorInvalidValue);
    if (kind == cudaMemcpyHostToDevice && !need_dev(dst, bytes, t_dev, "cudaMemcpy destination")) return ret(cudaErrorInvalidValue);
    memcpy(dst, src, bytes);
    return cudaSuccess;
}

cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t st) {
    LOCK;
    Stream *s = stream_of(st);
    if (g_fault[s->dev]) return ret((cudaError_t)g_fault[s->dev]);
    if (inject(API_MEMCPY_ASYNC)) return ret(cudaErrorLaunchFailure);
    Op op;
    op.type = Op::COPY; op.dst = dst; op.src = src; op.bytes = bytes; op.kind = (int)kind;
    const void *host_side = kind == cudaMemcpyHostToDevice ? src : dst;
    const Alloc *al = find_alloc(host_side, bytes);
    const bool pinned = al && al->kind == K_PINNED;
    if (kind != cudaMemcpyHostToDevice && kind != cudaMemcpyDeviceToHost) { sim_error("cudaMemcpyAsync kind %d is not modelled", (int)kind); return ret(cudaErrorInvalidValue); }
    if (!pinned) {
        // pageable memory: the runtime stages the source before returning (H2D) / completes the copy before returning (D2H)
        g_stats.pageable_async++;
        if (kind == cudaMemcpyHostToDevice) op.snapshot.assign((const uint8_t *)src, (const uint8_t *)src + bytes);
        s->q.push_back(std::move(op));
        if (kind == cudaMemcpyDeviceToHost) drive_all(s); else after_enqueue(s);
        return cudaSuccess;
    }
    enqueue(s, std::move(op));
    return cudaSuccess;
}

cudaError_t cudaMemset(void *p, int val, size_t bytes) {
    LOCK;
    if (!need_dev(p, bytes, t_dev, "cudaMemset")) return ret(cudaErrorInvalidValue);
    memset(p, val, bytes);
    return cudaSuccess;
}

cudaError_t cudaMemsetAsync(void *p, int val, size_t bytes, cudaStream_t st) {
    LOCK;
    Stream *s = stream_of(st);
    if (inject(API_MEMSET_ASYNC)) return ret(cudaErrorLaunchFailure);
    Op op;
    op.type = Op::MEMSET; op.dst = p; op.val = val; op.bytes = bytes;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

// =====================================================================================================================
// control surface for the tests
// =====================================================================================================================
void cudasim_config(int devices, int sms, int schedule) {
    LOCK;
    if (devices >= 0) g_ndev = devices > MAX_DEV ? MAX_DEV : devices;
    if (sms > 0) g_sms = sms;
    if (schedule >= 0) g_sched = schedule;
}

uint64_t cudasim_error_count(void) { LOCK; return g_errors; }
const char *cudasim_first_error(void) { LOCK; return g_first_error.c_str(); }
void cudasim_clear_errors(void) { LOCK; g_errors = 0; g_first_error.clear(); }
void cudasim_fail(int api, long nth) { LOCK; if (api >= 0 && api < API_COUNT) g_fail_countdown[api] = nth; }
void cudasim_clear_faults(void) { LOCK; memset(g_fault, 0, sizeof g_fault); memset(g_fail_countdown, 0, sizeof g_fail_countdown); }
void cudasim_stats(uint64_t out[12]) { LOCK; memcpy(out, &g_stats, sizeof g_stats); }
void cudasim_reset_stats(void) { LOCK; memset(&g_stats, 0, sizeof g_stats); g_launch_log.clear(); }
int cudasim_current_device(void) { return t_dev; }

// launches since the last reset: triples (device, operation, items); returns how many there were
size_t cudasim_launch_log(uint64_t *out, size_t cap) {
    LOCK;
    for (size_t i = 0; i < g_launch_log.size() && i < cap; i++) {
        out[3 * i] = (uint64_t)g_launch_log[i].dev; out[3 * i + 1] = (uint64_t)g_launch_log[i].op; out[3 * i + 2] = g_launch_log[i].n;
    }
    return g_launch_log.size();
}

// operations queued on any stream and not yet executed
size_t cudasim_pending_ops(void) {
    LOCK;
    size_t k = 0;
    for (Stream *s : g_streams) k += s->q.size();
    return k;
}

void cudasim_sync_all(void) {
    LOCK;
    for (size_t i = 0; i < g_streams.size(); i++) drive_all(g_streams[i]);
}

// live allocations made by the LIBRARY (kind 0 device, 1 pinned, -1 both): count, and bytes through *bytes
size_t cudasim_live_allocs(int kind, uint64_t *bytes) {
    LOCK;
    size_t k = 0;
    uint64_t b = 0;
    for (auto &kv : g_allocs)
        if (kv.second.lib && (kind < 0 || kv.second.kind == kind)) { k++; b += kv.second.size; }
    if (bytes) *bytes = b;
    return k;
}

// occurrences of any of `count` byte patterns (`len` >= 8 bytes each, back to back in `patterns`) in the library's live
// allocations — the secret-scrubbing tests.  One pass over the memory: the first 8 bytes of every pattern sit in a small
// hash table that is probed at every byte offset.
size_t cudasim_find(const uint8_t *patterns, size_t count, size_t len) {
    LOCK;
    if (len < 8 || count == 0) return 0;
    size_t cap = 64;
    while (cap < 4 * count) cap <<= 1;
    std::vector<int> slot(cap, -1);
    auto hash = [&](uint64_t v) { return (size_t)((v * 0x9e3779b97f4a7c15ull) >> 40) & (cap - 1); };
    for (size_t k = 0; k < count; k++) {
        uint64_t v;
        memcpy(&v, patterns + k * len, 8);
        size_t h = hash(v);
        while (slot[h] >= 0) h = (h + 1) & (cap - 1);
        slot[h] = (int)k;
    }
    size_t hits = 0;
    for (auto &kv : g_allocs) {
        if (!kv.second.lib || kv.second.size < len) continue;
        const uint8_t *p = (const uint8_t *)kv.first;
        const size_t last = kv.second.size - len;
        for (size_t i = 0; i <= last; i++) {
            uint64_t v;
            memcpy(&v, p + i, 8);
            for (size_t h = hash(v); slot[h] >= 0; h = (h + 1) & (cap - 1))
                if (!memcmp(p + i, patterns + (size_t)slot[h] * len, len)) { hits++; break; }
        }
    }
    return hits;
}

// memory owned by the TEST (a caller's page-locked buffers, a caller's device arrays for the _dev API)
void *cudasim_user_alloc(size_t size, int kind, int dev) { LOCK; return sim_alloc(size, kind, dev, false, 0xCC); }
void cudasim_user_free(void *p) { LOCK; auto it = g_allocs.find((uintptr_t)p); if (it != g_allocs.end()) sim_free(p, it->second.kind); }
void *cudasim_stream_create(int dev) { LOCK; Stream *s = new Stream{dev, {}}; g_streams.push_back(s); return s; }
void cudasim_set_device(int dev) { t_dev = dev; }

}  // extern "C"
