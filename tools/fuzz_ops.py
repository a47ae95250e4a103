"""One-off large differential of the secret-key operations against the compiled reference on the GPU box:
usage: tools/fuzz_ops.py <first seed> <last seed+1>   (2^20 operations of each kind per seed)"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import libeddsa_b200 as ed
from cpu_ref import best_cpu_impl
cpu = best_cpu_impl(); tot = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(seed); n = 1 << 20
    mlen = int(rng.integers(0, 300))
    sec = rng.integers(0, 256, (n, 32), dtype=np.uint8); pts = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, max(mlen, 1)), dtype=np.uint8)
    # structured rows: all-zero / all-one scalars and points, low-order and non-canonical u, sparse values
    k = np.arange(0, n, 64)
    sec[k[0::4]] = 0; sec[k[1::4]] = 0xff; pts[k[2::4]] = 0; pts[k[3::4]] = 0xff
    sparse = np.arange(7, n, 97); pts[sparse, 1:31] = 0; sec[sparse, 2:30] = 0
    pub = ed.ed25519_genpub_batch(sec)
    wrong = pub.copy(); wrong[::5] = rng.integers(0, 256, (len(wrong[::5]), 32), dtype=np.uint8)      # sign hashes the caller's pub as given (Q8)
    res = {"genpub": int((pub != cpu.genpub(sec)).any(axis=1).sum()),
           "sign": int((ed.ed25519_sign_batch(sec, wrong, msgs, fixed_len=msgs.shape[1]) != cpu.sign(sec, wrong, msgs, fixed_len=msgs.shape[1])).any(axis=1).sum()),
           "x25519": int((ed.x25519_batch(sec, pts) != cpu.x25519(sec, pts)).any(axis=1).sum()),
           "x25519_base": int((ed.x25519_base_batch(sec) != cpu.x25519_base(sec)).any(axis=1).sum())}
    print(seed, "msg len", msgs.shape[1], "mismatches", res, flush=True)
    tot += sum(res.values())
print("TOTAL MISMATCHES", tot)
