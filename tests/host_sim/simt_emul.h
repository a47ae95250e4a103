// TEST INFRASTRUCTURE ONLY — a small SIMT emulator: the REAL kernels of libeddsa_b200/csrc/kernels_*.cu on the CPU.
//
// tests/host_sim/ptx_rewrite.py turns every `kernel<<<grid, block, smem, stream>>>(args)` of the .cu files into
// edg_simt::launch(grid, block, smem, stream, [=] { kernel(args); }) and every inline-PTX statement into a call of the PTX
// interpreter (ptx_emul.h).  A launch is queued on the simulated stream (cudasim.cpp) as a closure that runs the grid block
// after block; the threads of a block are OS threads, so __syncthreads(), the warp collectives (__shfl_sync,
// __reduce_max_sync, __ballot_sync, __syncwarp) and the warp-wide tensor-core product mma.sync are real rendezvous between
// lanes, with the fragment layouts the PTX ISA defines.  __shared__ arrays are static storage (blocks run one at a time),
// dynamic shared memory is a per-block buffer.  Everything the launchers do on the host — passes, grids, scratch layout,
// tile sorting, the permutation of the verify loop — runs as shipped.
//
// __activemask() reports the calling lane alone (a legal answer wherever the code may be diverged), and collectives on a
// one-lane mask are local; full-mask collectives need every live lane of the warp, as on the device.
#pragma once
#define EDG_SIMT 1
#include "ptx_emul.h"

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace edg_simt {

struct Rendezvous {                                   // a reusable barrier whose party can shrink (lanes that have returned)
    std::mutex m;
    std::condition_variable cv;
    int live = 0, arrived = 0;
    unsigned gen = 0;
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++arrived >= live) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
    void leave() {
        std::unique_lock<std::mutex> lk(m);
        live--;
        if (live > 0 && arrived >= live) { arrived = 0; gen++; cv.notify_all(); }
    }
};
struct Warp : Rendezvous { uint32_t slot[32][8]; int lanes = 0; };     // lanes: the warp's width (live shrinks as lanes return)
struct Block : Rendezvous { std::vector<Warp> warps; char *dyn = nullptr; };
struct Lane { dim3 tid, bid, bdim, gdim; Block *blk; Warp *warp; unsigned lane; };
inline thread_local Lane *cur = nullptr;

// every lane contributes `words` words and gets everybody's
inline void warp_all(const uint32_t *mine, int words, uint32_t all[32][8]) {
    Warp &w = *cur->warp;
    memcpy(w.slot[cur->lane], mine, 4 * (size_t)words);
    w.arrive_and_wait();
    memcpy(all, w.slot, sizeof w.slot);
    w.arrive_and_wait();
}

inline void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    if (block.y != 1 || block.z != 1) { fprintf(stderr, "simt_emul: only one-dimensional blocks are modelled\n"); abort(); }
    const unsigned nthreads = block.x, nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Block blk;
                blk.live = (int)nthreads;
                blk.warps = std::vector<Warp>(nwarps);
                for (unsigned w = 0; w < nwarps; w++) blk.warps[w].live = blk.warps[w].lanes = (int)(w + 1 < nwarps ? 32 : nthreads - 32 * w);
                std::vector<char> dyn(smem + 64, (char)0xA5);
                blk.dyn = dyn.data() + (64 - ((uintptr_t)dyn.data() & 63)) % 64;
                std::vector<std::thread> th;
                th.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; t++)
                    th.emplace_back([&, t] {
                        Lane lane = {dim3(t), dim3(bx, by, bz), block, grid, &blk, &blk.warps[t / 32], t % 32};
                        cur = &lane;
                        edg_ptx::shared_base = nullptr;
                        body();
                        lane.warp->leave();
                        blk.leave();
                        cur = nullptr;
                    });
                for (auto &x : th) x.join();
            }
}

extern "C" int cudasim_enqueue_kernel(void *stream, void (*fn)(void *), void *arg);

inline void launch(dim3 grid, dim3 block, size_t smem, void *stream, std::function<void()> body) {
    auto *job = new std::function<void()>([=] { run_grid(grid, block, smem, body); });
    if (cudasim_enqueue_kernel(stream, [](void *p) { auto *f = (std::function<void()> *)p; (*f)(); delete f; }, job) != 0) delete job;
}

}  // namespace edg_simt

// field operations executed by the kernels since the last reset (builds with -DEDG_EMUL_COUNT)
extern "C" __attribute__((weak, visibility("default"))) void edg_emul_counts(unsigned long long *mul, unsigned long long *sq, int reset) {
    *mul = edg_ptx::cnt_mul.load();
    *sq = edg_ptx::cnt_sq.load();
    if (reset) { edg_ptx::cnt_mul = 0; edg_ptx::cnt_sq = 0; }
}

#define threadIdx (edg_simt::cur->tid)
#define blockIdx (edg_simt::cur->bid)
#define blockDim (edg_simt::cur->bdim)
#define gridDim (edg_simt::cur->gdim)
#define __shared__ static
#define __align__(n) alignas(n)
#define __launch_bounds__(...)

static inline void __syncthreads() { edg_simt::cur->blk->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { edg_simt::cur->warp->arrive_and_wait(); }
static inline unsigned __activemask() { return 1u << edg_simt::cur->lane; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline bool edg_one_lane(unsigned mask) { return mask == (1u << edg_simt::cur->lane); }
static inline unsigned __shfl_sync(unsigned mask, unsigned v, int src, int = 32) {
    if (edg_one_lane(mask)) return v;
    uint32_t all[32][8];
    edg_simt::warp_all(&v, 1, all);
    return all[src & 31][0];
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    if (edg_one_lane(mask)) return pred ? mask : 0u;
    uint32_t all[32][8], p = pred ? 1u : 0u, r = 0;
    edg_simt::warp_all(&p, 1, all);
    for (int l = 0; l < edg_simt::cur->warp->lanes; l++) r |= all[l][0] << l;
    return r & mask;
}
static inline int __reduce_max_sync(unsigned mask, int v) {
    if (edg_one_lane(mask)) return v;
    uint32_t all[32][8], x = (uint32_t)v;
    edg_simt::warp_all(&x, 1, all);
    int m = v;
    for (int l = 0; l < edg_simt::cur->warp->lanes; l++) m = (int)all[l][0] > m ? (int)all[l][0] : m;
    return m;
}

// mma.sync.aligned.m16n8k{16,32}.row.col.s32.u8.u8.s32 (PTX ISA fragment layouts; g = lane / 4, q = lane % 4):
//   A (16 x k, row-major): register 2 h + r of lane (g, q) holds row g + 8 r, columns 16 h + 4 q .. + 3 (one byte each)
//   B (k x 8, column-major): register h of lane (g, q) holds rows 16 h + 4 q .. + 3 of column g
//   C / D (16 x 8): register c of lane (g, q) is row g + 8 (c / 2), column 2 q + c % 2
inline void edg_ptx::mma_u8(int k, uint64_t *reg, const uint64_t *idx) {
    const int na = k / 8, nb = k / 16;
    uint32_t mine[8], all[32][8];
    for (int i = 0; i < na; i++) mine[i] = (uint32_t)reg[idx[4 + i]];
    for (int i = 0; i < nb; i++) mine[na + i] = (uint32_t)reg[idx[4 + na + i]];
    edg_simt::warp_all(mine, na + nb, all);
    const unsigned lane = edg_simt::cur->lane, g = lane >> 2, q = lane & 3;
    for (int c = 0; c < 4; c++) {
        const unsigned row = g + 8 * (c >> 1), col = 2 * q + (c & 1);
        int32_t sum = (int32_t)reg[idx[4 + na + nb + c]];
        for (int kk = 0; kk < k; kk++) {
            const uint32_t aw = all[4 * (row & 7) + (kk & 15) / 4][2 * (kk / 16) + (row >> 3)];
            const uint32_t bw = all[4 * col + (kk & 15) / 4][na + kk / 16];
            sum += (int32_t)(((aw >> (8 * (kk & 3))) & 0xffu) * ((bw >> (8 * (kk & 3))) & 0xffu));
        }
        reg[idx[c]] = (uint32_t)sum;
    }
}
