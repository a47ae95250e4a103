"""Quick GPU sanity + first timings (developer tool; the real checks live in tests/)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libeddsa_b200 as ed
from cpu_ref import best_cpu_impl
import torch

cpu = best_cpu_impl()
print("cpu impl:", cpu.kind, flush=True)
G = os.path.join(ROOT, "tests", "golden")
kat = np.fromfile(os.path.join(G, "x25519_kat.bin"), np.uint8).reshape(-1, 3, 32)
out = ed.x25519_batch(kat[:, 1], kat[:, 0])
print("x25519 KAT mismatches:", int((out != kat[:, 2]).any(axis=1).sum()), flush=True)
rng = np.random.default_rng(7)
n = 4096
sec = rng.integers(0, 256, (n, 32), dtype=np.uint8)
msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
pub = ed.ed25519_genpub_batch(sec); rp = cpu.genpub(sec)
print("genpub mismatches:", int((pub != rp).any(axis=1).sum()), flush=True)
sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64); rs = cpu.sign(sec, rp, msgs, fixed_len=64)
print("sign mismatches:", int((sig != rs).any(axis=1).sum()), flush=True)
bad = sig.copy(); bad[::3, 5] ^= 1
ok = ed.ed25519_verify_batch(bad, pub, msgs, fixed_len=64); rok = cpu.verify(bad, rp, msgs, fixed_len=64)
print("verify mismatches:", int((ok != rok).sum()), "accepted", int(ok.sum()), flush=True)
xb = ed.x25519_base_batch(sec); print("x25519_base mismatches:", int((xb != cpu.x25519_base(sec)).any(axis=1).sum()), flush=True)
adv = np.fromfile(os.path.join(G, "verify_adv.bin"), np.uint8).reshape(-1, 228)
lens = adv[:, 96].astype(np.uint64) | (adv[:, 97].astype(np.uint64) << 8)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
blob = np.concatenate([adv[i, 100:100 + int(lens[i])] for i in range(len(adv))])
ok = ed.ed25519_verify_batch(adv[:, :64], adv[:, 64:96], blob, off=off)
print("adversarial mismatches:", int((ok != adv[:, 99]).sum()), "of", len(adv), flush=True)

# timings on device-resident data
dev = torch.device("cuda:0")
N = 1 << 20
sec_t = torch.randint(0, 256, (N, 32), dtype=torch.uint8, device=dev)
msg_t = torch.randint(0, 256, (N, 64), dtype=torch.uint8, device=dev)
pub_t = torch.empty((N, 32), dtype=torch.uint8, device=dev)
sig_t = torch.empty((N, 64), dtype=torch.uint8, device=dev)
ok_t = torch.empty((N,), dtype=torch.uint8, device=dev)
out_t = torch.empty((N, 32), dtype=torch.uint8, device=dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
res = {}
t = timeit(lambda: ed.ed25519_genpub_batch_dev(pub_t, sec_t)); res["genpub"] = N / t * 1e3
t = timeit(lambda: ed.ed25519_sign_batch_dev(sig_t, sec_t, pub_t, msg_t, fixed_len=64)); res["sign"] = N / t * 1e3
t = timeit(lambda: ed.ed25519_verify_batch_dev(ok_t, sig_t, pub_t, msg_t, fixed_len=64)); res["verify"] = N / t * 1e3
res["verify_all_ok"] = bool(ok_t.all().item())
t = timeit(lambda: ed.x25519_batch_dev(out_t, sec_t, pub_t)); res["x25519"] = N / t * 1e3
t = timeit(lambda: ed.x25519_base_batch_dev(out_t, sec_t)); res["x25519_base"] = N / t * 1e3
print(json.dumps(res), flush=True)
