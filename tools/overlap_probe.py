"""Probe: does an ALU-bound hash kernel hide under the multiplier-bound comb kernel when both run at once (two streams)?
x25519_base (k_comb only) on one stream, sk_ed25519_to_x25519 (SHA-512 only) on another, sequential vs concurrent."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libeddsa_b200 as ed
dev = torch.device("cuda:0"); n = 1 << 20; reps_hash = int(os.environ.get("HASH_REPS", "6"))
sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
o1 = torch.empty((n, 32), dtype=torch.uint8, device=dev); o2 = torch.empty((n, 32), dtype=torch.uint8, device=dev)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
L = ed.lib()
def comb(): 
    with torch.cuda.stream(sa): ed.x25519_base_batch_dev(o1, sec)
def hashes():
    with torch.cuda.stream(sb):
        for _ in range(reps_hash): L.sk_ed25519_to_x25519_batch_dev(n, o2.data_ptr(), sec.data_ptr(), torch.cuda.current_stream().cuda_stream)
def t(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); 
        sa.wait_stream(torch.cuda.current_stream()); sb.wait_stream(torch.cuda.current_stream())
        fn()
        torch.cuda.current_stream().wait_stream(sa); torch.cuda.current_stream().wait_stream(sb)
        b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return round(best, 3)
r = {"lib": os.path.basename(ed.LIB_PATH), "comb_ms": t(comb), "hash_ms": t(hashes), "both_ms": t(lambda: (comb(), hashes())), "both_hash_first_ms": t(lambda: (hashes(), comb()))}
print(json.dumps(r))
