#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 through the library's OWN multi-device path: one process, one call per operation,
host buffers, the batch sharded by host.c (run_job) over EDDSA_B200_DEVICES — what a C / cgo caller gets.

  config 4: ed25519_genpub_batch / x25519_base_batch on 2^24 secret keys (512 MiB in, 512 MiB out per operation)
  config 5: ed25519_verify_batch on 2^24 signatures over 1024-byte messages (17.9 GB of host input), a deterministic
            10 % mutated (eight classes; S + L is an accept-quirk), valid signatures made by ed25519_sign_batch over the
            same devices; EVERY mutated row and 2^16 random unmutated rows compared with the CPU reference

Prints ONE JSON line.  bench.py runs this as a subprocess of rank 0 (the other ranks wait) and files the result under
"also.inproc"; every number is end to end (wall clock around the synchronous call; copies, staging and kernels inside).
The limiter is named with measured numbers: host-to-device copy bandwidth of plain cudaMemcpyAsync from pinned memory
per device and with all devices copying at once, and the host's memcpy bandwidth (the pageable path stages through it).

usage: tools/inproc_bench.py --devices N [--log2n 24] [--reps 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    os.environ["EDDSA_B200_DEVICES"] = ",".join(str(i) for i in range(args.devices)) + ","
    import psutil
    import torch
    import libeddsa_b200 as ed
    from cpu_ref import best_cpu_impl
    from edmodel import L

    G = args.devices
    assert torch.cuda.device_count() >= G
    avail = psutil.virtual_memory().available
    log2n = args.log2n
    while log2n > 18 and (1 << log2n) * 1200 * 1.5 > 0.6 * avail:    # config 5 needs ~1.2 KB of host memory per signature
        log2n -= 1
    n = 1 << log2n
    res = {"devices": G, "log2n": log2n, "host_cores": os.cpu_count(), "host_ram_available_gb": round(avail / 1e9, 1)}
    if log2n != args.log2n:
        res["note"] = f"batch reduced from 2^{args.log2n} to 2^{log2n}: not enough free host memory"
    ed.init()
    assert ed.device_count() == G
    cpu = best_cpu_impl()
    dev0 = torch.device("cuda:0")
    L_ = ed.lib()
    import ctypes
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def timed(fn, reps):
        fn()                                                     # warm-up: staging buffers, scratch pool, contexts
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best

    def call(name, *a):
        rc = getattr(L_, name)(*a)
        if rc:
            raise RuntimeError(f"{name} failed: {rc} {L_.eddsa_b200_last_error()}")

    # ---- copy-bandwidth context (names the limiter) -----------------------------------------------------------
    chunk = 1 << 30
    hbuf = torch.empty(chunk, dtype=torch.uint8, pin_memory=True)
    dbufs = [torch.empty(chunk, dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
    def h2d(devs):
        for g in devs:
            dbufs[g].copy_(hbuf, non_blocking=True)
        for g in devs:
            torch.cuda.synchronize(g)
    one = timed(lambda: h2d([0]), 3)
    allg = timed(lambda: h2d(range(G)), 3)
    src = np.empty(chunk, np.uint8); dst = np.empty(chunk, np.uint8); src[:] = 1; dst[:] = 0
    mc = timed(lambda: np.copyto(dst, src), 3)
    res["copy_bandwidth"] = {"h2d_pinned_one_device_gbs": round(chunk / one / 1e9, 1), "h2d_pinned_all_devices_aggregate_gbs": round(G * chunk / allg / 1e9, 1),
                             "host_memcpy_one_thread_gbs": round(chunk / mc / 1e9, 1)}
    del hbuf, dbufs, src, dst

    rng = np.random.default_rng(0x5EED0004)
    pick = rng.integers(0, n, 1 << 16)

    # ---- config 4: keygen --------------------------------------------------------------------------------------
    sec_t = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True)
    g = torch.Generator(device=dev0); g.manual_seed(0x5EED0004)
    sec_t.copy_(torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev0, generator=g))
    out_t = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True)
    sec_np, out_np = sec_t.numpy(), out_t.numpy()
    c4 = {"shards": [n * (k + 1) // G - n * k // G for k in range(G)], "io_bytes_per_op": 64}
    for name, fn, ref in (("genpub", "ed25519_genpub_batch", cpu.genpub), ("x25519_base", "x25519_base_batch", cpu.x25519_base)):
        dt = timed(lambda: call(fn, n, vp(out_t), vp(sec_t)), args.reps)
        ok = bool((out_np[pick] == ref(sec_np[pick])).all())
        c4[name] = {"ops_per_s": n / dt, "seconds": dt, "h2d_gbs": n * 32 / dt / 1e9, "parity_sample": "ok" if ok else "MISMATCH", "parity_rows": len(pick)}
    # the same from ordinary (pageable) memory: the library stages through its pinned slots
    sec_pg, out_pg = np.array(sec_np), np.empty_like(out_np)
    dt = timed(lambda: call("ed25519_genpub_batch", n, ctypes.c_void_p(out_pg.ctypes.data), ctypes.c_void_p(sec_pg.ctypes.data)), 3)
    c4["genpub_pageable"] = {"ops_per_s": n / dt, "seconds": dt}
    res["config4_keygen"] = c4
    pub_t = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True)
    call("ed25519_genpub_batch", n, vp(pub_t), vp(sec_t))
    del out_t, out_np, sec_pg, out_pg

    # ---- config 5: verify, 1 KB messages, 10 % mutated -----------------------------------------------------------
    mlen = 1024
    msg_t = torch.empty((n, mlen), dtype=torch.uint8, pin_memory=True)
    step = 1 << 20
    for lo in range(0, n, step):                                 # generated on device 0, copied down in slices
        msg_t[lo:lo + step].copy_(torch.randint(0, 256, (min(step, n - lo), mlen), dtype=torch.uint8, device=dev0, generator=g))
    torch.cuda.synchronize()
    sig_t = torch.empty((n, 64), dtype=torch.uint8, pin_memory=True)
    sign_s = timed(lambda: call("ed25519_sign_batch", n, vp(sig_t), vp(sec_t), vp(pub_t), vp(msg_t), None, mlen), 1)
    sig, pub, msg = sig_t.numpy(), pub_t.numpy(), msg_t.numpy()
    h = (np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B1) + np.uint64(0x5EED)) & np.uint64(0xFFFFFFFF)
    h = ((h ^ (h >> np.uint64(15))) * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    idx = np.nonzero(h % np.uint64(10) == 0)[0]
    cls = ((h[idx] // np.uint64(10)) % np.uint64(8)).astype(np.int64)
    Lb = np.frombuffer(L.to_bytes(32, "little"), np.uint8).astype(np.int64)
    for c in range(8):
        r = idx[cls == c]
        if c == 0: sig[r, 3] ^= 1
        elif c == 1: sig[r, 40] ^= 0x20
        elif c == 2: msg[r, mlen - 1] ^= 0x10
        elif c == 3: pub[r, 5] ^= 4
        elif c == 4: sig[r, :32] = 0xFF                          # non-canonical R (y >= p)
        elif c == 5:                                             # S + L: accepted, S is never range-checked (Q1)
            s = sig[r, 32:].astype(np.int64) + Lb
            for b in range(31):
                s[:, b + 1] += s[:, b] >> 8
                s[:, b] &= 0xFF
            sig[r, 32:] = s.astype(np.uint8)
        elif c == 6: sig[r, :] = 0
        elif c == 7: pub[r, 31] ^= 0x80
    ok_t = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
    dt = timed(lambda: call("ed25519_verify_batch", n, vp(ok_t), vp(sig_t), vp(pub_t), vp(msg_t), None, mlen), args.reps)
    ok = ok_t.numpy()
    keep = np.ones(n, bool); keep[idx] = False
    rest = rng.choice(np.nonzero(keep)[0], 1 << 16, replace=False)
    rows = np.concatenate([idx, rest])
    t0 = time.perf_counter()
    want = np.empty(len(rows), np.uint8)
    for lo in range(0, len(rows), 1 << 18):
        r = rows[lo:lo + (1 << 18)]
        want[lo:lo + len(r)] = cpu.verify(sig[r], pub[r], msg[r], fixed_len=mlen)
    cpu_s = time.perf_counter() - t0
    same = bool((ok[rows] == want).all())
    acc = ok[idx]
    quirk = bool(acc[cls == 5].all() and not acc[cls != 5].any())
    in_bytes = n * (64 + 32 + mlen)
    res["config5_verify_1kb_10pct_mutated"] = {
        "ops_per_s": n / dt, "seconds": dt, "h2d_bytes": in_bytes, "h2d_gbs_aggregate": in_bytes / dt / 1e9, "h2d_gbs_per_device": in_bytes / dt / 1e9 / G,
        "shards": [n * (k + 1) // G - n * k // G for k in range(G)], "accepted": int(ok.sum()), "mutated_rows": int(len(idx)),
        "parity_sample": "ok" if same and quirk else "MISMATCH", "parity_rows": int(len(rows)),
        "parity_protocol": "all mutated rows + 2^16 random unmutated rows vs the CPU reference (%s, %d threads, %.1f s)" % (cpu.kind, cpu.threads, cpu_s),
        "sign_1kb_ops_per_s_e2e": n / sign_s,
    }
    c5 = res["config5_verify_1kb_10pct_mutated"]
    # the same call from ordinary (pageable) memory: every byte is staged through the pinned slots (memcpy split over helper threads)
    msg_pg, sig_pg, pub_pg, ok_pg = np.array(msg), np.array(sig), np.array(pub), np.empty(n, np.uint8)
    cp = lambda a: ctypes.c_void_p(a.ctypes.data)
    dtp = timed(lambda: call("ed25519_verify_batch", n, cp(ok_pg), cp(sig_pg), cp(pub_pg), cp(msg_pg), None, mlen), 2)
    c5["pageable"] = {"ops_per_s": n / dtp, "seconds": dtp, "staged_gbs": in_bytes / dtp / 1e9, "same_decisions": bool((ok_pg == ok).all())}
    del msg_pg, sig_pg, pub_pg
    bw = res["copy_bandwidth"]
    c5["limiter"] = ("host-to-device copies: %.1f GB/s per device moved by the call vs %.1f GB/s per device that plain pinned cudaMemcpyAsync reaches "
                     "with all %d devices copying at once (%.1f GB/s alone); the kernels alone would need %.1f GB/s per device at 47 M verify/s" % (
                         c5["h2d_gbs_per_device"], bw["h2d_pinned_all_devices_aggregate_gbs"] / G, G, bw["h2d_pinned_one_device_gbs"], 47e6 * (96 + mlen) / 1e9))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
