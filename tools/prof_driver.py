"""Short driver for ncu captures: two launches of every kernel on 2^18 device-resident operations."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libeddsa_b200 as ed
dev = torch.device("cuda:0"); n = 1 << 18
sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
msg = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device=dev)
pts = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
pub = torch.empty((n, 32), dtype=torch.uint8, device=dev); sig = torch.empty((n, 64), dtype=torch.uint8, device=dev)
ok = torch.empty((n,), dtype=torch.uint8, device=dev); out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
for _ in range(2):
    ed.ed25519_genpub_batch_dev(pub, sec)
    ed.ed25519_sign_batch_dev(sig, sec, pub, msg, fixed_len=64)
    ed.ed25519_verify_batch_dev(ok, sig, pub, msg, fixed_len=64)
    ed.x25519_batch_dev(out, sec, pts)
    ed.x25519_base_batch_dev(out, sec)
torch.cuda.synchronize()
assert ok.all().item()
