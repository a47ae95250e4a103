import os, sys, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import libeddsa_b200 as ed
dev = torch.device("cuda:0"); n = 1 << 20
g = torch.Generator(device=dev); g.manual_seed(3)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return round(n / best / 1e3, 2)
res = {}
for name, lens in (("fixed64_as_offsets", torch.full((n,), 64, dtype=torch.int64, device=dev)),
                   ("ragged_0_128", torch.randint(0, 129, (n,), device=dev, generator=g)),
                   ("ragged_0_128_mult16", torch.randint(0, 9, (n,), device=dev, generator=g) * 16),
                   ("ragged_0_1024", torch.randint(0, 1025, (n,), device=dev, generator=g)),
                   ("ragged_0_1024_sorted_by_length", torch.sort(torch.randint(0, 1025, (n,), device=dev, generator=g))[0]),
                   ("ragged_0_1024_mult16_sorted", torch.sort(torch.randint(0, 65, (n,), device=dev, generator=g) * 16)[0])):
    off = torch.zeros(n + 1, dtype=torch.int64, device=dev); off[1:] = torch.cumsum(lens, 0)
    total = int(off[-1].item())
    blob = torch.randint(0, 256, (total + 16,), dtype=torch.uint8, device=dev, generator=g)
    sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    pub = torch.empty((n, 32), dtype=torch.uint8, device=dev); sig = torch.empty((n, 64), dtype=torch.uint8, device=dev); ok = torch.empty((n,), dtype=torch.uint8, device=dev)
    ed.ed25519_genpub_batch_dev(pub, sec)
    res["sign_" + name] = t(lambda: ed.ed25519_sign_batch_dev(sig, sec, pub, blob, off=off))
    res["verify_" + name] = t(lambda: ed.ed25519_verify_batch_dev(ok, sig, pub, blob, off=off))
    assert ok.all().item()
print(json.dumps(res))
