"""Tuning helper: device-resident throughput of every kernel for the library named by LIBEDDSA_B200_SO."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libeddsa_b200 as ed
dev = torch.device("cuda:0"); N = 1 << int(os.environ.get("LOG2N", "20"))
g = torch.Generator(device=dev); g.manual_seed(1)
sec = torch.randint(0, 256, (N, 32), dtype=torch.uint8, device=dev, generator=g)
msg = torch.randint(0, 256, (N, 64), dtype=torch.uint8, device=dev, generator=g)
pts = torch.randint(0, 256, (N, 32), dtype=torch.uint8, device=dev, generator=g)
pub = torch.empty((N, 32), dtype=torch.uint8, device=dev); sig = torch.empty((N, 64), dtype=torch.uint8, device=dev)
ok = torch.empty((N,), dtype=torch.uint8, device=dev); out = torch.empty((N, 32), dtype=torch.uint8, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return round(N / best / 1e3, 2)
r = {"lib": os.path.basename(ed.LIB_PATH)}
r["genpub"] = t(lambda: ed.ed25519_genpub_batch_dev(pub, sec))
r["sign"] = t(lambda: ed.ed25519_sign_batch_dev(sig, sec, pub, msg, fixed_len=64))
r["verify"] = t(lambda: ed.ed25519_verify_batch_dev(ok, sig, pub, msg, fixed_len=64)); r["ok"] = bool(ok.all().item())
r["x25519"] = t(lambda: ed.x25519_batch_dev(out, sec, pts))
r["x25519_base"] = t(lambda: ed.x25519_base_batch_dev(out, sec))
if os.environ.get("TUNE_1KB"):                    # 1 KB messages (BASELINE config 5's shape), 2^18 operations
    M = 1 << 18
    big = torch.randint(0, 256, (M, 1024), dtype=torch.uint8, device=dev, generator=g)
    N_, N = N, M
    r["sign_1kb"] = t(lambda: ed.ed25519_sign_batch_dev(sig[:M], sec[:M], pub[:M], big, fixed_len=1024))
    r["verify_1kb"] = t(lambda: ed.ed25519_verify_batch_dev(ok[:M], sig[:M], pub[:M], big, fixed_len=1024)); r["ok_1kb"] = bool(ok[:M].all().item())
    r["chk_1kb"] = int(sig[:M].to(torch.int64).sum().item())
    N = N_
    ed.ed25519_sign_batch_dev(sig, sec, pub, msg, fixed_len=64)
chk = int(out.to(torch.int64).sum().item()) ^ int(sig.to(torch.int64).sum().item())
r["chk"] = chk
print(json.dumps(r), flush=True)
