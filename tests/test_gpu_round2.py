"""GPU tests added in round 2 (run on the B200 box: `pytest -m gpu`), everything through the C-ABI:
device-level scalar arithmetic and shared inversion, the device-built comb table, long messages and over-budget
chunks, concurrent host threads, argument checks of the device-pointer API, many user streams, the secret-scrubbing
contract of the staging buffers, the caller's current device, the library's own multi-device sharding, and BASELINE
config 5 at full size under SURVEY.md §8(d)'s comparison protocol."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import golden_util as gu
from edmodel import L, P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand_bytes(rng, *shape):
    return rng.integers(0, 256, shape, dtype=np.uint8)


def to_rows(vals, width=32):
    return np.frombuffer(b"".join(v.to_bytes(width, "little") for v in vals), np.uint8).reshape(-1, width)


def from_rows(a):
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(len(a))]


# ------------------------------------------------------------------------------------------------ scalar layer (a13-a16)
def test_device_scalar_arithmetic(ed):
    """sc.cuh on the device against Python integers: Barrett reduction of 512-bit and 256-bit values (S + kL included:
    reduced, never rejected — Q1), a b + c mod L."""
    rng = np.random.default_rng(51)
    edge512 = [0, 1, L - 1, L, L + 1, 2 * L, 3 * L - 1, 2**252, 2**256 - 1, 2**512 - 1, (2**512 // L) * L, (2**512 // L) * L - 1, L * L,
               2**511, 2**504 - 1, (L - 1) * (L - 1)]
    xs = edge512 + [int.from_bytes(rng.bytes(64), "little") for _ in range(20000)] + [int.from_bytes(rng.bytes(64), "little") >> int(rng.integers(0, 511)) for _ in range(4000)]
    lo, hi = to_rows([x & (2**256 - 1) for x in xs]), to_rows([x >> 256 for x in xs])
    z = np.zeros_like(lo)
    got = from_rows(ed.sc_selftest(lo, hi, z, 0))
    assert all(g == x % L for g, x in zip(got, xs))
    ys = [12345 + k * L for k in range(16)] + [2**256 - 1, L - 1, L] + [int.from_bytes(rng.bytes(32), "little") for _ in range(5000)]
    ys = [y for y in ys if y < 2**256]
    a = to_rows(ys)
    got = from_rows(ed.sc_selftest(a, np.zeros_like(a), np.zeros_like(a), 1))
    assert all(g == y % L for g, y in zip(got, ys))
    n = 20000
    A = [int.from_bytes(rng.bytes(32), "little") for _ in range(n)] + [2**256 - 1, L - 1, 0]
    B = [int.from_bytes(rng.bytes(32), "little") >> 3 for _ in range(n)] + [2**253 - 1, L - 1, 0]
    C = [int.from_bytes(rng.bytes(32), "little") for _ in range(n)] + [2**256 - 1, L - 1, 0]
    got = from_rows(ed.sc_selftest(to_rows(A), to_rows(B), to_rows(C), 2))
    assert all(g == (x * y + c) % L for g, x, y, c in zip(got, A, B, C))


def test_device_shared_inversion_with_zeros(ed):
    """fe_batch_inv as shipped (EDG_BATCH = 32): groups of 32 consecutive items share one exponentiation; zeros in any
    representative (0, p, 2p) at any position — first, last, runs, whole groups — stay zero and leave their neighbours
    intact (inv(0) = 0, Q7).  The last group is short (n is not a multiple of 32)."""
    rng = np.random.default_rng(52)
    vals = []
    for g in range(300):
        grp = [int.from_bytes(rng.bytes(32), "little") for _ in range(32)]
        for j in range(32):
            if rng.random() < 0.2:
                grp[j] = [0, P, 2 * P][int(rng.integers(0, 3))]
        if g % 7 == 0:
            grp[0] = 0
        if g % 7 == 1:
            grp[31] = P
        if g % 7 == 2:
            grp[9:] = [0] * 23
        if g % 7 == 3:
            grp = [[0, P, 2 * P][j % 3] for j in range(32)]
        vals += grp
    vals += [5, 0, 7, P, 11]                                   # short tail group
    a = to_rows(vals)
    got = from_rows(ed.fe_selftest(a, a, 9))
    bad = [i for i, (g, v) in enumerate(zip(got, vals)) if g != pow(v % P, P - 2, P)]
    assert not bad, bad[:5]


def test_comb_table_built_on_the_device(ed):
    """Every entry of the fixed-base comb table the kernels stage in shared memory, against the big-integer model."""
    import edmodel as em
    tab = ed.comb_table()
    w = tab.shape[1].bit_length()
    assert tab.shape[0] == (255 + w - 1) // w
    em.check_comb_table(tab, w)


# ------------------------------------------------------------------------------------------------ long messages (f3)
@pytest.mark.parametrize("length", [16 * 1024, (1 << 20) + 3])
def test_long_messages(ed, cpu, length):
    """Many-block SHA-512 (up to 8 193 compressions per hash) in the sign and verify front-ends, fixed length and as
    part of a ragged batch next to empty and tiny messages."""
    rng = np.random.default_rng(length)
    n = 6
    sec = rand_bytes(rng, n, 32)
    pub = cpu.genpub(sec)
    msgs = rand_bytes(rng, n, length)
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=length)
    assert (sig == cpu.sign(sec, pub, msgs, fixed_len=length)).all()
    bad = msgs.copy()
    bad[1, length - 1] ^= 1                                    # last byte of a long message
    bad[2, length // 2] ^= 0x80
    assert (ed.ed25519_verify_batch(sig, pub, bad, fixed_len=length) == np.array([1, 0, 0, 1, 1, 1], np.uint8)).all()
    ragged = [b"", msgs[0].tobytes(), b"x", msgs[1, :length - 129].tobytes(), bytes(127), msgs[2, 5:].tobytes()]
    blob, off = gu.ragged(ragged)
    rsig = ed.ed25519_sign_batch(sec, pub, blob, off=off)
    assert (rsig == cpu.sign(sec, pub, blob, off=off)).all()
    assert ed.ed25519_verify_batch(rsig, pub, blob, off=off).all()


def test_message_larger_than_the_staging_budget():
    """EDDSA_B200_CHUNK_MB=1 with 3 MB messages: a single item exceeds the chunk budget, the staging buffers grow
    (host.c: run_shard) and chunks shrink to one item.  Own process: the budget is read once at initialisation."""
    code = r"""
import os, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import libeddsa_b200 as ed
from cpu_ref import best_cpu_impl
cpu = best_cpu_impl()
rng = np.random.default_rng(3)
n, length = 5, 3 * (1 << 20) + 17
sec = rng.integers(0, 256, (n, 32), dtype=np.uint8); pub = cpu.genpub(sec)
msgs = rng.integers(0, 256, (n, length), dtype=np.uint8)
sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=length)
assert (sig == cpu.sign(sec, pub, msgs, fixed_len=length)).all()
sig[3, 2] ^= 4
assert (ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=length) == np.array([1, 1, 1, 0, 1], np.uint8)).all()
lens = [0, length, 5, 70000, 1 << 20]
blob = np.concatenate([msgs[i, :l] for i, l in enumerate(lens)]); off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
rs = ed.ed25519_sign_batch(sec, pub, blob, off=off)
assert (rs == cpu.sign(sec, pub, blob, off=off)).all() and ed.ed25519_verify_batch(rs, pub, blob, off=off).all()
small = rng.integers(0, 256, (300000, 32), dtype=np.uint8)      # many chunks of a many-item batch under the same budget
assert (ed.x25519_base_batch(small)[::997] == cpu.x25519_base(small[::997])).all()
print("ok")
""" % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, EDDSA_B200_CHUNK_MB="1")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]


# ------------------------------------------------------------------------------------------------ threads, streams, arguments
def test_concurrent_host_threads(ed, cpu):
    """eddsa_batch.h promises that every entry point may be called concurrently: four host threads issue mixed batch
    and single-operation calls at the same time; every result is compared with the CPU checker."""
    rng = np.random.default_rng(77)
    n = 20000
    sec, pts, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 32), rand_bytes(rng, n, 96)
    pub = cpu.genpub(sec)
    sig = cpu.sign(sec, pub, msgs, fixed_len=96)
    bad = sig.copy()
    bad[::5, 1] ^= 2
    want_ok = cpu.verify(bad, pub, msgs, fixed_len=96)
    want_x, want_xb = cpu.x25519(sec, pts), cpu.x25519_base(sec)
    errors = []

    def worker(t):
        try:
            for rep in range(3):
                k = (t + rep) % 4
                if k == 0:
                    assert (ed.ed25519_verify_batch(bad, pub, msgs, fixed_len=96) == want_ok).all()
                elif k == 1:
                    assert (ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=96) == sig).all()
                elif k == 2:
                    assert (ed.x25519_batch(sec, pts) == want_x).all() and (ed.ed25519_genpub_batch(sec) == pub).all()
                else:
                    assert (ed.x25519_base_batch(sec) == want_xb).all()
                    for i in range(20):
                        assert ed.ed25519_verify(bad[i].tobytes(), pub[i].tobytes(), msgs[i].tobytes()) == bool(want_ok[i])
                        assert ed.ed25519_genpub(sec[i].tobytes()) == pub[i].tobytes()
        except BaseException as e:                             # noqa: BLE001 — reported by the main thread
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_device_pointer_api_argument_checks(ed):
    """Misaligned or NULL device arrays are refused with EDDSA_B200_EINVAL before anything is launched."""
    import torch
    dev = torch.device("cuda:0")
    n = 64
    buf = torch.zeros(n * 64 + 64, dtype=torch.uint8, device=dev)
    out = torch.zeros(n * 64 + 64, dtype=torch.uint8, device=dev)
    lib = ed.lib()
    vp = ctypes.c_void_p
    before = ed.launch_count()
    assert lib.x25519_base_batch_dev(n, vp(out.data_ptr()), vp(buf.data_ptr() + 4), None) == -1
    assert b"aligned" in lib.eddsa_b200_last_error()
    assert lib.x25519_base_batch_dev(n, vp(out.data_ptr() + 8), vp(buf.data_ptr()), None) == -1
    assert lib.ed25519_genpub_batch_dev(n, None, vp(buf.data_ptr()), None) == -1
    assert lib.x25519_batch_dev(n, vp(out.data_ptr()), vp(buf.data_ptr()), None, None) == -1
    assert lib.ed25519_verify_batch_dev(n, vp(out.data_ptr()), vp(buf.data_ptr() + 1), vp(buf.data_ptr()), vp(buf.data_ptr()), None, 8, None) == -1
    assert lib.ed25519_sign_batch_dev(n, vp(out.data_ptr()), vp(buf.data_ptr()), vp(buf.data_ptr()), None, None, 8, None) == -1
    assert ed.launch_count() == before
    # the accept flags may sit at any address
    sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
    msg = torch.randint(0, 256, (n, 8), dtype=torch.uint8, device=dev)
    pub = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    sig = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ed.ed25519_genpub_batch_dev(pub, sec)
    ed.ed25519_sign_batch_dev(sig, sec, pub, msg, fixed_len=8)
    ok = torch.zeros(n + 3, dtype=torch.uint8, device=dev)
    ed.ed25519_verify_batch_dev(ok[3:], sig, pub, msg, fixed_len=8)
    assert ok[3:].all().item() and not ok[:3].any().item()


def test_many_user_streams(ed):
    """Kernel scratch is stream-ordered: 24 distinct streams (round 1 failed on the 17th) verify and sign concurrently
    and never synchronise inside the calls."""
    import torch
    dev = torch.device("cuda:0")
    n = 3000
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msg = torch.randint(0, 256, (n, 40), dtype=torch.uint8, device=dev, generator=g)
    pub = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    ref = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ed.ed25519_genpub_batch_dev(pub, sec)
    ed.ed25519_sign_batch_dev(ref, sec, pub, msg, fixed_len=40)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(24)]
    sigs = [torch.empty((n, 64), dtype=torch.uint8, device=dev) for _ in streams]
    oks = [torch.zeros((n,), dtype=torch.uint8, device=dev) for _ in streams]
    for rep in range(2):
        for s, sg, ok in zip(streams, sigs, oks):
            with torch.cuda.stream(s):
                ed.ed25519_sign_batch_dev(sg, sec, pub, msg, fixed_len=40)
                ed.ed25519_verify_batch_dev(ok, sg, pub, msg, fixed_len=40)
    torch.cuda.synchronize()
    for sg, ok in zip(sigs, oks):
        assert torch.equal(sg, ref) and ok.all().item()


def test_verify_several_passes_per_call():
    """A device-API verify call of more signatures than one pass holds runs pass after pass over the same record slab
    (by default a pass is 16 waves = 1.2 M signatures; EDDSA_B200_VERIFY_WAVES=1 makes it 75 776, so that a 200 000-
    signature call is three passes with a ragged last one).  Fixed and ragged messages, corrupted rows rejected by
    position, same decisions as the host pipeline.  Own process: the setting is read once at initialisation."""
    code = r"""
import os, sys, numpy as np, torch
sys.path.insert(0, %r)
import libeddsa_b200 as ed
rng = np.random.default_rng(12)
n = 200_000 + 37
sec = rng.integers(0, 256, (n, 32), dtype=np.uint8); msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
pub = ed.ed25519_genpub_batch(sec); sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)
bad = np.zeros(n, bool); bad[::9] = True; bad[[75_775, 75_776, 151_551, 151_552, n - 1]] = True
sig[bad, 40] ^= 2
dev = torch.device("cuda:0"); t = lambda a: torch.from_numpy(a).to(dev)
ok = torch.empty(n, dtype=torch.uint8, device=dev)
before = ed.launch_count()
ed.ed25519_verify_batch_dev(ok, t(sig), t(pub), t(msgs), fixed_len=64); torch.cuda.synchronize()
assert ed.launch_count() - before == 9, ed.launch_count() - before          # three passes of three kernels
assert (ok.cpu().numpy().astype(bool) == ~bad).all()
assert (ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=64).astype(bool) == ~bad).all()
lens = rng.integers(0, 150, n); off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
blob = rng.integers(0, 256, int(off[-1]) + 16, dtype=np.uint8)
rsig = ed.ed25519_sign_batch(sec, pub, blob, off=off.astype(np.uint64)); rsig[bad, 3] ^= 1
ed.ed25519_verify_batch_dev(ok, t(rsig), t(pub), t(blob), off=t(off)); torch.cuda.synchronize()
assert (ok.cpu().numpy().astype(bool) == ~bad).all()
print("ok")
""" % (ROOT,)
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, EDDSA_B200_VERIFY_WAVES="1"), capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]


# ------------------------------------------------------------------------------------------------ secrets
def test_staging_buffers_are_scrubbed():
    """After a host-buffer call with secret inputs / outputs from ordinary (pageable) memory, the pinned host staging
    slots and the device staging slots hold zeros where the secrets were (reference: burn after every secret-key call,
    ed25519-sha512.c:77,136,255, x25519.c:208,221).  Controls: a verify leaves its public inputs in place (so the peek
    really reads the live buffers), and EDDSA_B200_DEBUG_NO_SCRUB=1 leaves the secrets behind."""
    code = r"""
import os, sys, numpy as np
sys.path.insert(0, %r)
import libeddsa_b200 as ed
rng = np.random.default_rng(9)
n = 150000                                                        # 3 chunks of >= 65536: every pipeline slot gets used
sec = rng.integers(1, 256, (n, 32), dtype=np.uint8); pts = rng.integers(1, 256, (n, 32), dtype=np.uint8)
msgs = rng.integers(1, 256, (n, 16), dtype=np.uint8)
chunks = [65536, 65536, n - 2 * 65536]                           # host.c: run_shard cuts n into chunks of max(n / 4, 65536); slot = chunk
def residue(width_in, width_out):
    # non-zero bytes left where each chunk's first input array (the secrets) and its outputs were staged
    r = []
    for slot, m in enumerate(chunks):
        for which, w in ((0, width_in), (1, width_in), (2, width_out), (3, width_out)):
            if w:
                b = ed.peek_staging(which, slot, m * w)
                assert len(b) == m * w
                r.append(int(np.count_nonzero(b)))
    return r
pub = ed.ed25519_genpub_batch(sec);                r_genpub = residue(32, 0)
sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=16); r_sign = residue(32, 0)
out = ed.x25519_batch(sec, pts);                   r_x = residue(32, 32)
xb = ed.x25519_base_batch(sec);                    r_xb = residue(32, 0)
sk = ed.sk_ed25519_to_x25519_batch(sec);           r_sk = residue(32, 32)
ok = ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=16); r_verify = residue(64, 0)
assert ok.all() and out.any() and sk.any()
print(sum(r_genpub), sum(r_sign), sum(r_x), sum(r_xb), sum(r_sk), sum(r_verify))
""" % (ROOT,)
    def run(extra):
        res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **extra), capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stderr[-2000:]
        return [int(x) for x in res.stdout.split()[-6:]]
    scrubbed = run({})
    assert scrubbed[:5] == [0, 0, 0, 0, 0], scrubbed              # nothing left of sec / scalar / shared secrets
    assert scrubbed[5] > 1_000_000, scrubbed                      # control: public signature bytes are still there
    dirty = run({"EDDSA_B200_DEBUG_NO_SCRUB": "1"})
    assert min(dirty[:5]) > 1_000_000, dirty                      # negative control: without the scrub the keys stay


def test_callers_current_device_is_kept(ed):
    """A host-buffer call (single operations included) must not change the calling thread's current CUDA device
    (run_shard switched it to the engine's first device)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(4)
    sec = rand_bytes(rng, 1000, 32)
    last = torch.cuda.device_count() - 1
    ed.set_device_count(1)
    try:
        torch.cuda.set_device(last)
        torch.zeros(1, device="cuda")                             # make the runtime's current device explicit
        ed.ed25519_genpub_batch(sec)
        ed.x25519_base(sec[0].tobytes())
        assert torch.cuda.current_device() == last
        t = torch.ones(4, device="cuda")
        assert t.device.index == last
    finally:
        ed.set_device_count(0)
        torch.cuda.set_device(0)


# ------------------------------------------------------------------------------------------------ multi-device (e)
def test_library_shards_over_all_devices(ed, cpu):
    """The in-library multi-GPU path a C caller gets (host.c: run_job): one call, contiguous index shards over all
    visible devices, long-lived worker thread per device.  Every operation kind, fixed and ragged messages; results
    must equal the single-device results bit for bit and a sample must equal the CPU checker."""
    if ed.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(2024)
    n = 100_003 * ed.device_count()                              # not a multiple of anything
    sec, pts, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 32), rand_bytes(rng, n, 64)
    lens = rng.integers(0, 200, n)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    blob = rand_bytes(rng, int(off[-1]) + 16)

    def run_all():
        pub = ed.ed25519_genpub_batch(sec)
        sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)
        rsig = ed.ed25519_sign_batch(sec, pub, blob, off=off)
        bad = sig.copy()
        bad[::9, 33] ^= 1
        return dict(pub=pub, sig=sig, rsig=rsig, ok=ed.ed25519_verify_batch(bad, pub, msgs, fixed_len=64),
                    rok=ed.ed25519_verify_batch(rsig, pub, blob, off=off), x=ed.x25519_batch(sec, pts), xb=ed.x25519_base_batch(sec),
                    pk=ed.pk_ed25519_to_x25519_batch(pub), sk=ed.sk_ed25519_to_x25519_batch(sec))

    multi = run_all()
    ed.set_device_count(1)
    try:
        single = run_all()
    finally:
        ed.set_device_count(0)
    for k in multi:
        assert (multi[k] == single[k]).all(), k
    assert multi["rok"].all() and multi["ok"].sum() == n - len(range(0, n, 9))
    sample = np.concatenate([np.arange(0, 300), np.arange(n - 300, n), rng.integers(0, n, 3000),
                             np.arange(n // ed.device_count() - 50, n // ed.device_count() + 50)])   # incl. a shard boundary
    assert (multi["pub"][sample] == cpu.genpub(sec[sample])).all()
    assert (multi["sig"][sample] == cpu.sign(sec[sample], multi["pub"][sample], msgs[sample], fixed_len=64)).all()
    assert (multi["x"][sample] == cpu.x25519(sec[sample], pts[sample])).all()
    assert (multi["xb"][sample] == cpu.x25519_base(sec[sample])).all()


# ------------------------------------------------------------------------------------------------ config 5 at full size
def mutate_config5(sig, pub, msg, idx, cls):
    """Deterministic corruption of rows idx (class cls[i] in 0..7) of torch CUDA tensors, in place.  Classes: 0 R bit,
    1 S bit, 2 message bit, 3 public-key bit, 4 non-canonical R (y >= p), 5 S + L (ACCEPTED by the reference: S is
    never range-checked, Q1), 6 all-zero signature, 7 sign bit of the public key."""
    import torch
    Lb = torch.tensor(list(L.to_bytes(32, "little")), dtype=torch.int64, device=sig.device)
    for c in range(8):
        r = idx[cls == c]
        if len(r) == 0:
            continue
        if c == 0:
            sig[r, 3] ^= 1
        elif c == 1:
            sig[r, 40] ^= 0x20
        elif c == 2:
            msg[r, msg.shape[1] - 1] ^= 0x10
        elif c == 3:
            pub[r, 5] ^= 4
        elif c == 4:
            sig[r, :32] = 0xFF
        elif c == 5:                                              # S += L byte-wise with carry (S < L, so S + L < 2^254)
            s = sig[r, 32:].to(torch.int64) + Lb
            for b in range(31):
                carry = s[:, b] >> 8
                s[:, b] &= 0xFF
                s[:, b + 1] += carry
            sig[r, 32:] = s.to(torch.uint8)
        elif c == 6:
            sig[r, :] = 0
        elif c == 7:
            pub[r, 31] ^= 0x80


@pytest.mark.slow
def test_config5_full_size_protocol(ed, cpu):
    """BASELINE config 5 exactly as SURVEY.md §8(d) states it: 2^24 signatures over 1024-byte messages (16 GiB, generated
    on the device), valid signatures from the parity-checked GPU sign path, a deterministic 10 % mutated (eight classes,
    accept-quirks such as S + L included), one verify call over the whole batch — then EVERY mutated row (~1.68 M) and
    2^16 random unmutated rows are compared with the CPU reference, decision by decision."""
    import torch
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info(dev)
    n = 1 << 24
    if free < 22 << 30:
        pytest.skip("needs 22 GB of free device memory")
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0005)
    sec = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    msg = torch.empty((n, 1024), dtype=torch.uint8, device=dev)
    for lo in range(0, n, 1 << 21):                               # (int64 randint temporaries: fill in slices)
        msg[lo:lo + (1 << 21)] = torch.randint(0, 256, (1 << 21, 1024), dtype=torch.uint8, device=dev, generator=g)
    pub = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    sig = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ok = torch.empty((n,), dtype=torch.uint8, device=dev)
    ed.ed25519_genpub_batch_dev(pub, sec)
    ed.ed25519_sign_batch_dev(sig, sec, pub, msg, fixed_len=1024)
    h = (torch.arange(n, device=dev, dtype=torch.int64) * 0x9E3779B1 + 0x5EED) & 0xFFFFFFFF
    h = (h ^ (h >> 15)) * 0x85EBCA6B & 0xFFFFFFFF
    h = h ^ (h >> 13)
    idx = torch.nonzero(h % 10 == 0).flatten()
    cls = (h[idx] // 10) % 8
    mutate_config5(sig, pub, msg, idx, cls)
    ed.ed25519_verify_batch_dev(ok, sig, pub, msg, fixed_len=1024)
    torch.cuda.synchronize()
    assert 0.095 * n < len(idx) < 0.105 * n
    keep = torch.ones(n, dtype=torch.bool, device=dev)
    keep[idx] = False
    rest = torch.nonzero(keep).flatten()
    rest = rest[torch.randperm(len(rest), device=dev, generator=g)[: 1 << 16]]
    rows = torch.cat([idx, rest])
    got = ok[rows].cpu().numpy()
    want = np.empty(len(rows), np.uint8)
    step = 1 << 18
    for lo in range(0, len(rows), step):                          # 256 MB of messages per slice through the CPU reference
        r = rows[lo:lo + step]
        want[lo:lo + step] = cpu.verify(sig[r].cpu().numpy(), pub[r].cpu().numpy(), msg[r].cpu().numpy(), fixed_len=1024)
    assert (got == want).all(), np.nonzero(got != want)[0][:10]
    acc = got[: len(idx)]
    c = cls.cpu().numpy()
    assert acc[c == 5].all() and not acc[c != 5].any()            # of the eight classes only S + L is accepted
    assert got[len(idx):].all() and int(ok.sum().item()) == n - len(idx) + int((c == 5).sum())
