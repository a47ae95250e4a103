"""Secondary BASELINE.json configurations on one GPU (developer tool): verify / sign with 1 KB messages and
a 10 % corrupted mix (config 5 shape), keygen at 2^24 (config 4).  Device-resident, CUDA-event timed."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libeddsa_b200 as ed
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(7)
def t(fn, n, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return round(n / best / 1e3, 2)
res = {}
for log2n, mlen in ((20, 1024), (24, 1024), (20, 64)):       # (24, 1024) = BASELINE config 5 at full size: 16 GiB of messages
    n = 1 << log2n
    sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msg = torch.randint(0, 256, (n, mlen), dtype=torch.uint8, device=dev, generator=g)
    pub = torch.empty((n, 32), dtype=torch.uint8, device=dev); sig = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ok = torch.empty((n,), dtype=torch.uint8, device=dev)
    ed.ed25519_genpub_batch_dev(pub, sec)
    res[f"sign_{mlen}B_2^{log2n}"] = t(lambda: ed.ed25519_sign_batch_dev(sig, sec, pub, msg, fixed_len=mlen), n)
    res[f"verify_{mlen}B_2^{log2n}"] = t(lambda: ed.ed25519_verify_batch_dev(ok, sig, pub, msg, fixed_len=mlen), n)
    assert ok.all().item()
    # 10 % corrupted / non-canonical (config 5): R bit flips, S + L (still ACCEPTED: no range check on S), message
    # bit flips, public-key bit flips, R replaced by a non-canonical encoding
    bad = sig.clone(); msg2 = msg; pub2 = pub.clone()
    idx = torch.arange(0, n, 10, device=dev); k = idx // 10 % 4
    bad[idx[k == 0], 3] ^= 1
    bad[idx[k == 1], 40] ^= 0x20
    pub2[idx[k == 2], 5] ^= 4
    bad[idx[k == 3], :32] = 0xff                                                       # y >= p: non-canonical R
    res[f"verify_{mlen}B_2^{log2n}_10pct_bad"] = t(lambda: ed.ed25519_verify_batch_dev(ok, bad, pub2, msg2, fixed_len=mlen), n)
    assert int(ok.sum().item()) == n - len(idx)
    del sec, msg, pub, sig, ok, bad
n = 1 << 24
sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g); out = torch.empty_like(sec)
res["genpub_2^24"] = t(lambda: ed.ed25519_genpub_batch_dev(out, sec), n)
res["x25519_base_2^24"] = t(lambda: ed.x25519_base_batch_dev(out, sec), n)
print(json.dumps(res), flush=True)
