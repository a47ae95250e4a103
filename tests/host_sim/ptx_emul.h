// TEST INFRASTRUCTURE ONLY — host shims for compiling the DEVICE code paths of libeddsa_b200/csrc/*.cuh on the CPU.
//
// tests/host_sim/ptx_rewrite.py turns every inline-PTX statement of the headers into a call of the small PTX interpreter
// below; this header supplies that interpreter plus the CUDA keywords and intrinsics the device paths use, with the
// semantics the PTX ISA / CUDA math API documents.  A translation unit compiled with
//     g++ -D__CUDA_ARCH__=1000 -include ptx_emul.h -I<rewritten headers>
// then runs the exact device-side C++ of the kernels' per-thread operation bodies — carry chains, funnel shifts, byte
// permutes, 128-bit loads, the cp.async staging of the verify loop — as ONE lane (thread 0 of a block of 1), so the unit
// tests of tests/test_host_sim.py cover them without a GPU.  Nothing in the product includes this file.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __restrict__

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v = {x, y, z, w}; return v; }
template <typename T> static inline T __ldg(const T *p) { return *p; }

#if !defined(EDG_SIMT)
// one lane of a one-thread block
struct edg_dim3 { unsigned x, y, z; };
static const edg_dim3 blockDim = {1, 1, 1}, gridDim = {1, 1, 1}, threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static inline unsigned __shfl_sync(unsigned, unsigned v, int, int = 32) { return v; }
static inline int __reduce_max_sync(unsigned, int v) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
#else
// many lanes: tests/host_sim/simt_emul.h supplies the thread geometry and the warp collectives
namespace edg_ptx { inline void mma_u8(int k, uint64_t *reg, const uint64_t *idx); }
#endif

// CUDA integer intrinsics (CUDA math API semantics)
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (shift & 31u));
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }

namespace edg_ptx {

// field multiplications / squarings executed (EDG_EMUL_COUNT builds: the real kernels on the SIMT emulator)
inline std::atomic<unsigned long long> cnt_mul{0}, cnt_sq{0};

// "shared memory": the window handed out by __cvta_generic_to_shared is addressed by 32-bit offsets, as on the device
static thread_local char *shared_base = nullptr;
static const uint32_t kSharedBias = 0x400;

struct Arg { void *p; uint64_t val; int size; };
template <typename T> static inline Arg out(T &x, const char *c) {
    (void)c;
    Arg a = {(void *)&x, 0, (int)sizeof(T)};
    return a;
}
template <typename T> static inline Arg in(const T &x, const char *c) {
    (void)c;
    Arg a = {nullptr, 0, (int)sizeof(T)};
    if (sizeof(T) == 8) { uint64_t v; memcpy(&v, &x, 8); a.val = v; }
    else { uint32_t v = 0; memcpy(&v, &x, sizeof(T) < 4 ? sizeof(T) : 4); a.val = v; }
    return a;
}
template <typename T> static inline Arg in(T *const &x, const char *) { Arg a = {nullptr, (uint64_t)(uintptr_t)x, 8}; return a; }

enum Op { ADD, ADDC, SUB, SUBC, MAD_LO, MADC_LO, MADC_HI, MAD_HI, CP_ASYNC, NOP, MMA_K16, MMA_K32 };
struct Operand { bool imm; uint64_t v; };            // register index or immediate
struct Ins { Op op; bool cc; int nops; Operand o[16]; };
struct Prog { std::vector<Ins> ins; };

static inline Operand parse_operand(std::string t) {
    while (!t.empty() && (t[0] == ' ' || t[0] == '[')) t.erase(0, 1);
    while (!t.empty() && (t.back() == ' ' || t.back() == ']')) t.pop_back();
    Operand o;
    if (!t.empty() && t[0] == '%') { o.imm = false; o.v = strtoull(t.c_str() + 1, nullptr, 10); }
    else { o.imm = true; o.v = strtoull(t.c_str(), nullptr, 0); }
    return o;
}

static inline Prog compile(const char *tmpl) {
    Prog p;
    std::string s(tmpl);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t semi = s.find(';', pos);
        if (semi == std::string::npos) semi = s.size();
        std::string st = s.substr(pos, semi - pos);
        pos = semi + 1;
        size_t b = st.find_first_not_of(" \t\n");
        if (b == std::string::npos) continue;
        st = st.substr(b);
        size_t sp = st.find_first_of(" \t");
        std::string opc = st.substr(0, sp), rest = sp == std::string::npos ? "" : st.substr(sp + 1);
        Ins in;
        in.cc = false;
        in.nops = 0;
        if (opc == "add.cc.u32") { in.op = ADD; in.cc = true; }
        else if (opc == "addc.cc.u32") { in.op = ADDC; in.cc = true; }
        else if (opc == "addc.u32") { in.op = ADDC; }
        else if (opc == "sub.cc.u32") { in.op = SUB; in.cc = true; }
        else if (opc == "subc.cc.u32") { in.op = SUBC; in.cc = true; }
        else if (opc == "subc.u32") { in.op = SUBC; }
        else if (opc == "mad.lo.cc.u32") { in.op = MAD_LO; in.cc = true; }
        else if (opc == "madc.lo.cc.u32") { in.op = MADC_LO; in.cc = true; }
        else if (opc == "madc.hi.cc.u32") { in.op = MADC_HI; in.cc = true; }
        else if (opc == "madc.hi.u32") { in.op = MADC_HI; }
        else if (opc == "cp.async.cg.shared.global") { in.op = CP_ASYNC; }
        else if (opc == "cp.async.commit_group" || opc == "cp.async.wait_group") { in.op = NOP; rest = ""; }
        else if (opc == "mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32") { in.op = MMA_K16; }
        else if (opc == "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32") { in.op = MMA_K32; }
        else { fprintf(stderr, "ptx_emul: instruction '%s' is not modelled\n", opc.c_str()); abort(); }
        for (char &ch : rest)
            if (ch == '{' || ch == '}') ch = ' ';           // operand groups of mma: D(4), A(2 | 4), B(1 | 2), C(4), in order
        size_t q = 0;
        while (q < rest.size() && in.nops < 16) {
            size_t comma = rest.find(',', q);
            if (comma == std::string::npos) comma = rest.size();
            std::string t = rest.substr(q, comma - q);
            if (t.find_first_not_of(" \t") != std::string::npos) in.o[in.nops++] = parse_operand(t);
            q = comma + 1;
        }
        p.ins.push_back(in);
    }
    return p;
}

// Executes one asm statement.  CC.CF lives for the duration of the statement (every carry chain of the headers is one
// statement).  add / addc: carry-out; sub / subc: CF = borrow-out, subc subtracts it (PTX ISA, "Extended-precision integer
// arithmetic"); mad{c}.lo/hi: d = lo / hi 32 bits of a * b, + c (+ CF), carry-out of that 32-bit addition.
static inline void run(const Prog &p, Arg *a, int n) {
    uint64_t reg[40];
    for (int i = 0; i < n; i++) {
        if (a[i].p) { if (a[i].size == 8) memcpy(&reg[i], a[i].p, 8); else { uint32_t v; memcpy(&v, a[i].p, 4); reg[i] = v; } }
        else reg[i] = a[i].val;
    }
    auto val = [&](const Operand &o) -> uint64_t { return o.imm ? o.v : reg[o.v]; };
    uint32_t cf = 0;
    for (const Ins &in : p.ins) {
        switch (in.op) {
        case ADD: case ADDC: {
            const uint64_t t = (uint64_t)(uint32_t)val(in.o[1]) + (uint32_t)val(in.o[2]) + (in.op == ADDC ? cf : 0u);
            reg[in.o[0].v] = (uint32_t)t;
            if (in.cc) cf = (uint32_t)(t >> 32);
            break;
        }
        case SUB: case SUBC: {
            const uint64_t t = (uint64_t)(uint32_t)val(in.o[1]) - (uint32_t)val(in.o[2]) - (in.op == SUBC ? cf : 0u);
            reg[in.o[0].v] = (uint32_t)t;
            if (in.cc) cf = (uint32_t)(t >> 63);
            break;
        }
        case MAD_LO: case MADC_LO: case MADC_HI: case MAD_HI: {
            const uint64_t prod = (uint64_t)(uint32_t)val(in.o[1]) * (uint32_t)val(in.o[2]);
            const uint32_t part = (in.op == MADC_HI || in.op == MAD_HI) ? (uint32_t)(prod >> 32) : (uint32_t)prod;
            const uint64_t t = (uint64_t)part + (uint32_t)val(in.o[3]) + ((in.op == MADC_LO || in.op == MADC_HI) ? cf : 0u);
            reg[in.o[0].v] = (uint32_t)t;
            if (in.cc) cf = (uint32_t)(t >> 32);
            break;
        }
        case CP_ASYNC:                                   // [shared offset], [global address], 16
            if (!shared_base) { fprintf(stderr, "ptx_emul: cp.async without a shared window\n"); abort(); }
            memcpy(shared_base + ((uint32_t)val(in.o[0]) - kSharedBias), (const void *)(uintptr_t)val(in.o[1]), (size_t)val(in.o[2]));
            break;
        case NOP: break;
        case MMA_K16: case MMA_K32: {
#if defined(EDG_SIMT)
            uint64_t idx[16];
            for (int k = 0; k < in.nops; k++) idx[k] = in.o[k].imm ? ~0ull : in.o[k].v;
            mma_u8(in.op == MMA_K16 ? 16 : 32, reg, idx);
#else
            fprintf(stderr, "ptx_emul: mma.sync needs the multi-lane build (simt_emul.h)\n");
            abort();
#endif
            break;
        }
        }
    }
    for (int i = 0; i < n; i++)
        if (a[i].p) { if (a[i].size == 8) memcpy(a[i].p, &reg[i], 8); else { const uint32_t v = (uint32_t)reg[i]; memcpy(a[i].p, &v, 4); } }
}

}  // namespace edg_ptx

// the generic address of a shared-memory object as a shared-window offset; the first object seen defines the window
static inline size_t __cvta_generic_to_shared(const void *p) {
    if (!edg_ptx::shared_base) edg_ptx::shared_base = (char *)p;
    return (size_t)((const char *)p - edg_ptx::shared_base) + edg_ptx::kSharedBias;
}
