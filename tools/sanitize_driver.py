"""Small mixed workload for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libeddsa_b200 as ed
rng = np.random.default_rng(3)
for n in (1, 33, 700):
    sec = rng.integers(0, 256, (n, 32), dtype=np.uint8); pts = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    lens = rng.integers(0, 300, n); msgs = [rng.integers(0, 256, int(l), dtype=np.uint8).tobytes() for l in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64); blob = np.frombuffer(b"".join(msgs) + b"\0", np.uint8)
    pub = ed.ed25519_genpub_batch(sec)
    sig = ed.ed25519_sign_batch(sec, pub, blob, off=off)
    ok = ed.ed25519_verify_batch(sig, pub, blob, off=off); assert ok.all()
    fixed = np.frombuffer(rng.bytes(n * 64), np.uint8)
    sig2 = ed.ed25519_sign_batch(sec, pub, fixed, fixed_len=64); assert ed.ed25519_verify_batch(sig2, pub, fixed, fixed_len=64).all()
    ed.x25519_batch(sec, pts); ed.x25519_base_batch(sec); ed.pk_ed25519_to_x25519_batch(pub); ed.sk_ed25519_to_x25519_batch(sec)
print("sanitize driver ok")
