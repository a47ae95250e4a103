"""Probe: run-to-run variation of host-buffer calls from ordinary (pageable) memory, with and without copy helpers."""
import os, sys, time, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libeddsa_b200 as ed
n = 1 << 24
rng = np.random.default_rng(1)
sec = rng.integers(0, 256, (n, 32), dtype=np.uint8); out = np.empty_like(sec)
L = ed.lib(); import ctypes
cp = lambda a: ctypes.c_void_p(a.ctypes.data)
ts = []
for i in range(12):
    t0 = time.perf_counter(); rc = L.ed25519_genpub_batch(n, cp(out), cp(sec)); ts.append(round((time.perf_counter() - t0) * 1e3, 1)); assert rc == 0
print(json.dumps({"copy_threads": os.environ.get("EDDSA_B200_COPY_THREADS", "default"), "genpub_2^24_pageable_ms": ts}))
