import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import libeddsa_b200 as ed, edmodel as em
from cpu_ref import best_cpu_impl
cpu = best_cpu_impl(); L = em.L
small = [np.frombuffer(em.enc(p), np.uint8) for p in em.small_order_points()]
small_nc = [np.frombuffer(em.enc(p, noncanonical=True), np.uint8) for p in em.small_order_points() if p[1] < 19]
tot = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(seed); n = 1 << 20
    mlen = int(rng.integers(0, 200))
    sec, msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8), rng.integers(0, 256, (n, max(mlen, 1)), dtype=np.uint8)
    pub = ed.ed25519_genpub_batch(sec); sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=msgs.shape[1])
    cls = rng.integers(0, 16, n)
    pos = rng.integers(0, 32, n); bit = (1 << rng.integers(0, 8, n)).astype(np.uint8)
    m = cls == 0; sig[m, pos[m]] ^= bit[m]
    m = cls == 1; sig[m, 32 + pos[m]] ^= bit[m]
    m = cls == 2; pub[m, pos[m]] ^= bit[m]
    m = cls == 3; msgs[m, pos[m] % msgs.shape[1]] ^= bit[m]
    for i in np.nonzero(cls == 4)[0][:20000]:
        v = int.from_bytes(sig[i, 32:].tobytes(), "little") + int(rng.integers(1, 16)) * L
        if v < 2**256: sig[i, 32:] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    m = cls == 7; sig[m, :32] = np.stack([small[k] for k in rng.integers(0, 8, m.sum())])
    m = cls == 8; sig[m, :32] = sig[np.roll(np.nonzero(m)[0], 1), :32]
    m = cls == 9; pub[m] = np.stack([small[k] for k in rng.integers(0, 8, m.sum())])
    m = cls == 10; pub[m, 31] ^= 0x80
    m = cls == 11; pub[m] = np.stack([small_nc[k] for k in rng.integers(0, len(small_nc), m.sum())])
    m = cls == 6; sig[m, :32] = 0xff
    got = ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=msgs.shape[1]); want = cpu.verify(sig, pub, msgs, fixed_len=msgs.shape[1])
    bad = np.nonzero(got != want)[0]
    print(seed, "len", msgs.shape[1], "accepted", int(want.sum()), "mismatches", len(bad), [(int(i), int(cls[i])) for i in bad[:5]], flush=True)
    tot += len(bad)
print("TOTAL MISMATCHES", tot)
