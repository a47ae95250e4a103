"""Driver of tools/ct_dynamic.py: launches every secret-key kernel once on n operations whose SECRETS are all-zero,
all-one or random bytes (argv[1]) while every public input (messages, public keys, points) is the same in all three
runs.  Run under ncu; the kernels' instruction and memory-transaction counters must not depend on the class."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libeddsa_b200 as ed
cls, n = sys.argv[1], int(sys.argv[2])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1234)
msg = torch.randint(0, 256, (n, 96), dtype=torch.uint8, device=dev, generator=g)
pub = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
pts = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
lens = torch.randint(0, 200, (n,), device=dev, generator=g)
off = torch.zeros(n + 1, dtype=torch.int64, device=dev); off[1:] = torch.cumsum(lens, 0)
blob = torch.randint(0, 256, (int(off[-1].item()) + 16,), dtype=torch.uint8, device=dev, generator=g)
if cls == "zeros":
    sec = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
elif cls == "ones":
    sec = torch.full((n, 32), 255, dtype=torch.uint8, device=dev)
else:
    sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
out = torch.empty((n, 32), dtype=torch.uint8, device=dev); sig = torch.empty((n, 64), dtype=torch.uint8, device=dev)
ed.ed25519_genpub_batch_dev(out, sec)                       # k_expand_key, k_comb<0>
ed.ed25519_sign_batch_dev(sig, sec, pub, msg, fixed_len=96)  # k_sign_nonce<0>, k_comb<0>, k_sign_finish<0>
ed.ed25519_sign_batch_dev(sig, sec, pub, blob, off=off)      # k_sign_nonce<1>, k_comb<0>, k_sign_finish<1>
ed.x25519_batch_dev(out, sec, pts)                          # k_x25519
ed.x25519_base_batch_dev(out, sec)                          # k_comb<1>
ed.lib().sk_ed25519_to_x25519_batch_dev(n, out.data_ptr(), sec.data_ptr(), None)     # k_sk_convert
torch.cuda.synchronize()
