"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C-ABI of
libeddsa_b200.so (include/eddsa.h, include/eddsa_batch.h) and is compared bit-for-bit with
  * the committed golden fixtures (the reference's own x25519 KAT table + reference-generated rows),
  * the CPU checker on fresh seeded inputs (compiled reference when oracle/_ref travelled, else the
    oracle port) at sizes it finishes in seconds,
  * size-independent properties at the BASELINE.json batch sizes (2^20): sign->verify round trips,
    corruption => reject, x25519 commutativity, x25519_base == x25519(., 9), sampled oracle rows.
"""
import numpy as np
import pytest

import golden_util as gu
from edmodel import L

pytestmark = pytest.mark.gpu


def rand_bytes(rng, *shape):
    return rng.integers(0, 256, shape, dtype=np.uint8)


# ------------------------------------------------------------------------------------------------ field arithmetic
def test_device_field_arithmetic(ed):
    """The PTX carry-chain field library (fe.cuh) against Python big integers, on every combination of
    the corner values (0, p, 2p, 2^256-1, values whose fold carries twice ...) and random 256-bit inputs."""
    from edmodel import P
    rng = np.random.default_rng(42)
    edge = [0, 1, 2, 19, 37, 38, 39, P - 1, P, P + 1, P + 18, 2 * P, 2 * P + 1, 2 * P + 37, 2**255 - 20, 2**255 - 1, 2**255, 2**255 + 18,
            2**256 - 39, 2**256 - 38, 2**256 - 1, 2**256 - 2**32, 2**224, 2**32 - 1, (2**256 - 1) // 3, 2**128 - 1, 2**255 - 19 + 2**32]
    vals_a = [x for x in edge for _ in edge] + [int.from_bytes(rng.bytes(32), "little") for _ in range(20000)]
    vals_b = [y for _ in edge for y in edge] + [int.from_bytes(rng.bytes(32), "little") for _ in range(20000)]
    a = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals_a), np.uint8).reshape(-1, 32)
    b = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals_b), np.uint8).reshape(-1, 32)

    def run(op):
        out = ed.fe_selftest(a, b, op)
        return [int.from_bytes(out[i].tobytes(), "little") for i in range(len(out))]

    for op, fn in ((0, lambda x, y: x * y), (1, lambda x, y: x * x), (2, lambda x, y: x + y), (3, lambda x, y: x - y),
                   (4, lambda x, y: x * 121665), (8, lambda x, y: -x)):
        got = run(op)
        bad = [i for i in range(len(got)) if got[i] % P != fn(vals_a[i], vals_b[i]) % P]
        assert not bad, (op, bad[:5])
    got = run(5)
    assert all(got[i] == vals_a[i] % P for i in range(len(got)))            # canonical form is THE value in [0, p)
    n_small = 3000
    a, b = a[:n_small], b[:n_small]
    got = run(6)
    assert all(got[i] % P == pow(vals_a[i] % P, P - 2, P) for i in range(n_small))     # inv(0) = 0 included
    got = run(7)
    assert all(got[i] % P == pow(vals_a[i] % P, (P - 5) // 8, P) for i in range(n_small))


# ------------------------------------------------------------------------------------------------ fixtures
def test_x25519_reference_kat_table(ed):
    point, scalar, result = gu.x25519_kat()
    assert (ed.x25519_batch(scalar, point) == result).all()


def test_x25519_edge_cases(ed):
    point, scalar, result = gu.x25519_edge()
    assert (ed.x25519_batch(scalar, point) == result).all()


def test_x25519_base_fixture(ed):
    bs, bo = gu.x25519_base_kat()
    assert (ed.x25519_base_batch(bs) == bo).all()
    nine = np.zeros((len(bs), 32), np.uint8)
    nine[:, 0] = 9
    assert (ed.x25519_batch(bs, nine) == bo).all()              # selftest-x25519_base.c:22-42


def test_ed25519_kat_all_message_lengths(ed):
    """genpub / sign / verify on 1024 reference rows, message length = row index 0..1023 (ragged batch)."""
    sec, pub, sig, msgs = gu.ed25519_kat()
    blob, off = gu.ragged(msgs)
    assert (ed.ed25519_genpub_batch(sec) == pub).all()
    assert (ed.ed25519_sign_batch(sec, pub, blob, off=off) == sig).all()
    assert ed.ed25519_verify_batch(sig, pub, blob, off=off).all()


def test_adversarial_verify_decisions(ed):
    """Every accept/reject decision of the reference on malformed / non-canonical / small-order /
    mixed-order / S+kL / off-curve inputs."""
    sig, pub, msgs, cls, expect = gu.verify_adv()
    blob, off = gu.ragged(msgs)
    got = ed.ed25519_verify_batch(sig, pub, blob, off=off)
    bad = np.nonzero(got != expect)[0]
    assert len(bad) == 0, [(int(i), int(cls[i]), int(expect[i])) for i in bad[:10]]


def test_sign_with_wrong_pub(ed):
    sec, pub, sig, msgs = gu.sign_wrongpub()
    blob, off = gu.ragged(msgs)
    assert (ed.ed25519_sign_batch(sec, pub, blob, off=off) == sig).all()


def test_key_conversion_fixture(ed):
    edsk, edpk, xsk, xpk = gu.convert_kat()
    assert (ed.sk_ed25519_to_x25519_batch(edsk) == xsk).all()
    assert (ed.pk_ed25519_to_x25519_batch(edpk) == xpk).all()
    # selftest-convert.c:21-45: x25519_base(sk_conv(sk)) == pk_conv(genpub(sk))
    assert (ed.x25519_base_batch(ed.sk_ed25519_to_x25519_batch(edsk)) == ed.pk_ed25519_to_x25519_batch(ed.ed25519_genpub_batch(edsk))).all()


# ------------------------------------------------------------------------------------------------ single-op API
def test_single_operation_api(ed, cpu):
    """eddsa.h entry points (batch of one), incl. the obsolete aliases, against the CPU checker."""
    import ctypes
    rng = np.random.default_rng(11)
    lib = ed.lib()
    for i in range(8):
        sk = rand_bytes(rng, 32)
        msg = rand_bytes(rng, int(rng.integers(0, 200)))
        pk = ed.ed25519_genpub(sk.tobytes())
        assert pk == cpu.genpub(sk).tobytes()
        sig = ed.ed25519_sign(sk.tobytes(), pk, msg.tobytes())
        assert sig == cpu.sign(sk, np.frombuffer(pk, np.uint8), msg, fixed_len=len(msg)).tobytes()
        assert ed.ed25519_verify(sig, pk, msg.tobytes()) is True
        assert ed.ed25519_verify(sig, pk, msg.tobytes() + b"x") is False
        pt = rand_bytes(rng, 32)
        assert ed.x25519(sk.tobytes(), pt.tobytes()) == cpu.x25519(sk, pt).tobytes()
        assert ed.x25519_base(sk.tobytes()) == cpu.x25519_base(sk).tobytes()
        out = ctypes.create_string_buffer(32)
        lib.DH(out, sk.tobytes(), pt.tobytes())
        assert out.raw == cpu.x25519(sk, pt).tobytes()
        lib.eddsa_genpub(out, sk.tobytes())
        assert out.raw == pk
        lib.eddsa_verify.restype = ctypes.c_bool
        assert lib.eddsa_verify(sig, pk, msg.tobytes(), ctypes.c_size_t(len(msg)))
        assert ed.sk_ed25519_to_x25519(sk.tobytes()) == cpu.sk_to_x25519(sk.tobytes())
        assert ed.pk_ed25519_to_x25519(pk) == cpu.pk_to_x25519(pk)


# ------------------------------------------------------------------------------------------------ differential vs CPU
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 128, 129, 1000, 4097])
def test_batch_boundaries(ed, cpu, n):
    """Batch sizes around warp / block multiples; n = 0 is legal."""
    rng = np.random.default_rng(100 + n)
    sec, pt, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 32), rand_bytes(rng, n, 64)
    pub = ed.ed25519_genpub_batch(sec)
    assert (pub == cpu.genpub(sec)).all()
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)
    assert (sig == cpu.sign(sec, pub, msgs, fixed_len=64)).all()
    if n:
        sig[::2, 7] ^= 0x40
    assert (ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=64) == cpu.verify(sig, pub, msgs, fixed_len=64)).all()
    assert (ed.x25519_batch(sec, pt) == cpu.x25519(sec, pt)).all()
    assert (ed.x25519_base_batch(sec) == cpu.x25519_base(sec)).all()


@pytest.mark.parametrize("length", [0, 1, 47, 48, 63, 64, 65, 111, 112, 127, 128, 175, 176, 239, 240, 1024, 1025])
def test_message_lengths_straddling_sha512_blocks(ed, cpu, length):
    """The hashed strings are prefix32||M and R||A||M, so block boundaries sit at 128k-32 / 128k-64."""
    rng = np.random.default_rng(200 + length)
    n = 96
    sec, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, max(length, 1))[:, :length]
    msgs = np.ascontiguousarray(msgs)
    pub = cpu.genpub(sec)
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=length)
    assert (sig == cpu.sign(sec, pub, msgs, fixed_len=length)).all()
    sig[1::3, 40] ^= 1
    assert (ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=length) == cpu.verify(sig, pub, msgs, fixed_len=length)).all()


def test_ragged_and_unaligned_messages(ed, cpu):
    rng = np.random.default_rng(300)
    n = 700
    lens = rng.integers(0, 300, n)
    msgs = [rng.integers(0, 256, int(l), dtype=np.uint8).tobytes() for l in lens]
    blob, off = gu.ragged(msgs)
    sec = rand_bytes(rng, n, 32)
    pub = cpu.genpub(sec)
    sig = ed.ed25519_sign_batch(sec, pub, blob, off=off)
    assert (sig == cpu.sign(sec, pub, blob, off=off)).all()
    assert ed.ed25519_verify_batch(sig, pub, blob, off=off).all()


def test_ragged_batch_across_sign_tiles(ed, cpu):
    """Ragged batches are signed in tiles of 1024 signatures visited in order of message length (full rounds of one
    tile per block, then an evenly split remainder): a batch larger than one round, lengths 0..400, every signature
    verified, a sample and both ends of the batch compared with the oracle."""
    rng = np.random.default_rng(350)
    n = 148 * 2 * 1024 + 5000 + 13
    lens = rng.integers(0, 401, n)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    blob = rng.integers(0, 256, int(off[-1]) + 16, dtype=np.uint8)
    sec = rand_bytes(rng, n, 32)
    pub = ed.ed25519_genpub_batch(sec)
    sig = ed.ed25519_sign_batch(sec, pub, blob, off=off)
    assert ed.ed25519_verify_batch(sig, pub, blob, off=off).all()
    sample = np.concatenate([np.arange(0, 600), np.arange(n - 600, n), rng.integers(0, n, 1200)])
    sub = [blob[int(off[i]):int(off[i + 1])].tobytes() for i in sample]
    sblob, soff = gu.ragged(sub)
    assert (sig[sample] == cpu.sign(sec[sample], pub[sample], sblob, off=soff)).all()


def test_random_differential_medium(ed, cpu):
    """2^14 fresh operations of every kind vs the CPU checker, with 1/8 of the signatures corrupted."""
    rng = np.random.default_rng(400)
    n = 1 << 14
    sec, pt, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 32), rand_bytes(rng, n, 64)
    pub = ed.ed25519_genpub_batch(sec)
    assert (pub == cpu.genpub(sec)).all()
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)
    assert (sig == cpu.sign(sec, pub, msgs, fixed_len=64)).all()
    idx = np.arange(0, n, 8)
    sig[idx, rng.integers(0, 64, len(idx))] ^= (1 << rng.integers(0, 8, len(idx))).astype(np.uint8)
    assert (ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=64) == cpu.verify(sig, pub, msgs, fixed_len=64)).all()
    assert (ed.x25519_batch(sec, pt) == cpu.x25519(sec, pt)).all()
    assert (ed.x25519_base_batch(sec) == cpu.x25519_base(sec)).all()


def test_mutation_fuzz_verify_large(ed, cpu):
    """2^17 signatures, three quarters of them mutated in one of twelve ways (bit flips in R / S / A / message, S + kL,
    S := 0 or L, non-canonical or small-order R, R of another signature, small-order / non-canonical / sign-flipped A,
    truncated hash input): every decision must equal the reference's (both accept-quirks and rejects occur)."""
    import edmodel as em
    rng = np.random.default_rng(1234)
    n = 1 << 17
    sec, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 48)
    pub = ed.ed25519_genpub_batch(sec)
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=48)
    L = em.L
    small = [np.frombuffer(em.enc(p), np.uint8) for p in em.small_order_points()]
    small_nc = [np.frombuffer(em.enc(p, noncanonical=True), np.uint8) for p in em.small_order_points() if p[1] < 19]
    cls = rng.integers(0, 16, n)                                # 12..15: left valid
    for i in np.nonzero(cls < 12)[0]:
        c = cls[i]
        if c == 0: sig[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
        elif c == 1: sig[i, 32 + rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
        elif c == 2: pub[i, rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
        elif c == 3: msgs[i, rng.integers(0, 48)] ^= 1 << rng.integers(0, 8)
        elif c == 4:
            v = int.from_bytes(sig[i, 32:].tobytes(), "little") + int(rng.integers(1, 16)) * L
            if v < 2**256: sig[i, 32:] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        elif c == 5: sig[i, 32:] = np.frombuffer((0 if rng.integers(0, 2) else L).to_bytes(32, "little"), np.uint8)
        elif c == 6: sig[i, :32] = small_nc[rng.integers(0, len(small_nc))] if rng.integers(0, 2) else 0xff
        elif c == 7: sig[i, :32] = small[rng.integers(0, 8)]
        elif c == 8: sig[i, :32] = sig[(i + 1) % n, :32]
        elif c == 9: pub[i] = small[rng.integers(0, 8)]
        elif c == 10: pub[i, 31] ^= 0x80
        elif c == 11: pub[i] = small_nc[rng.integers(0, len(small_nc))]
    got = ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=48)
    want = cpu.verify(sig, pub, msgs, fixed_len=48)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, [(int(i), int(cls[i]), int(got[i]), int(want[i])) for i in bad[:10]]
    assert want[cls >= 12].all() and want[cls == 4].all()        # valid rows and S + kL rows are accepted
    assert 0.25 * n < want.sum() < 0.45 * n


def test_s_plus_kl_is_accepted(ed):
    """SURVEY Q1 as a property: S + kL verifies for every k with S + kL < 2^256."""
    rng = np.random.default_rng(500)
    n = 256
    sec, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 64)
    pub = ed.ed25519_genpub_batch(sec)
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)
    for k in (1, 7, 15):
        s2 = sig.copy()
        for i in range(n):
            v = int.from_bytes(sig[i, 32:].tobytes(), "little") + k * L
            assert v < 2**256
            s2[i, 32:] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        assert ed.ed25519_verify_batch(s2, pub, msgs, fixed_len=64).all()


# ------------------------------------------------------------------------------------------------ full BASELINE sizes
def test_full_size_properties_2pow20(ed, cpu):
    """Config sizes of BASELINE.json (batch 2^20): properties that need no oracle at scale, plus a
    sampled oracle comparison."""
    rng = np.random.default_rng(600)
    n = 1 << 20
    sec, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 64)
    pub = ed.ed25519_genpub_batch(sec)
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)
    ok = ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=64)
    assert ok.all()                                             # sign -> verify round trip
    bad = sig.copy()
    idx = np.nonzero(rng.integers(0, 10, n) == 0)[0]            # ~10 % corrupted (config 5 shape)
    bad[idx, rng.integers(0, 64, len(idx))] ^= (1 << rng.integers(0, 8, len(idx))).astype(np.uint8)
    ok = ed.ed25519_verify_batch(bad, pub, msgs, fixed_len=64)
    sample = np.concatenate([idx[:3000], rng.integers(0, n, 3000)])
    assert (ok[sample] == cpu.verify(bad[sample], pub[sample], msgs[sample], fixed_len=64)).all()
    untouched = np.ones(n, bool)
    untouched[idx] = False
    assert ok[untouched].all()
    assert (pub[sample] == cpu.genpub(sec[sample])).all()
    assert (sig[sample] == cpu.sign(sec[sample], pub[sample], msgs[sample], fixed_len=64)).all()
    # X25519: Diffie-Hellman commutes; fixed base agrees with the ladder on u = 9
    a, b = rand_bytes(rng, n, 32), rand_bytes(rng, n, 32)
    pa, pb = ed.x25519_base_batch(a), ed.x25519_base_batch(b)
    assert (ed.x25519_batch(a, pb) == ed.x25519_batch(b, pa)).all()
    nine = np.zeros((n, 32), np.uint8)
    nine[:, 0] = 9
    assert (ed.x25519_batch(a, nine) == pa).all()
    pts = rand_bytes(rng, n, 32)                                # bit 255 left random, as in the KAT table (Q6)
    out = ed.x25519_batch(a, pts)
    assert (out[sample] == cpu.x25519(a[sample], pts[sample])).all()


def test_window_tables_built_on_the_device(ed):
    """The 2 x 32769-entry base-point tables k_wtab_build leaves in device memory: sampled entries against the
    big-integer model, every entry against its predecessor (entry e - entry e-1 must be the table's base point:
    checked through the affine coordinates recovered from (y+x, y-x))."""
    import random
    from test_host_sim import wtab_expected, WTAB_SAMPLE
    import edmodel as em
    tabs = ed.verify_tables()
    rng = random.Random(9)
    for m in range(2):
        for e in WTAB_SAMPLE + [rng.randrange(32769) for _ in range(60)]:
            assert tabs[m, e].tobytes() == wtab_expected(m, e), (m, e)
    # all entries: (y+x)(y-x) = y^2 - x^2 and 2dxy consistent with the curve equation  -x^2 + y^2 = 1 + d x^2 y^2
    P = em.P
    inv2 = pow(2, P - 2, P)
    for m in range(2):
        for e in range(0, 32769, 97):
            ypx, ymx, xy2d = (int.from_bytes(tabs[m, e, k].tobytes(), "little") for k in range(3))
            x, y = (ypx - ymx) * inv2 % P, (ypx + ymx) * inv2 % P
            assert (y * y - x * x - 1 - em.D * x * x * y * y) % P == 0, (m, e)
            assert xy2d == 2 * em.D * x * y % P, (m, e)


def test_verify_full_length_fallback_path_on_the_device(cpu):
    """The (rho, tau) = (1, t) fallback of the half-size-scalar verification (64 windows instead of ~33) can only be
    reached by challenge values nobody can construct, so a debug switch (EDDSA_B200_DEBUG_FULL_SCALARS=1, read once
    at library start-up) forces it for every signature: the adversarial fixtures and a corrupted random batch must be
    decided exactly as before.  Runs in a fresh interpreter because the switch is read at initialisation."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import os, sys, numpy as np
        sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
        import libeddsa_b200 as ed, golden_util as gu
        from cpu_ref import best_cpu_impl
        cpu = best_cpu_impl()
        sig, pub, msgs, cls, expect = gu.verify_adv()
        blob, off = gu.ragged(msgs)
        got = ed.ed25519_verify_batch(sig, pub, blob, off=off)
        assert (got == expect).all(), np.nonzero(got != expect)[0][:10]
        rng = np.random.default_rng(77); n = 40000
        sec = rng.integers(0, 256, (n, 32), dtype=np.uint8); m = rng.integers(0, 256, (n, 64), dtype=np.uint8)
        pk = ed.ed25519_genpub_batch(sec); sg = ed.ed25519_sign_batch(sec, pk, m, fixed_len=64)
        sg[::7, rng.integers(0, 64)] ^= 0x04
        ok = ed.ed25519_verify_batch(sg, pk, m, fixed_len=64)
        assert (ok == cpu.verify(sg, pk, m, fixed_len=64)).all()
        assert 0.8 * n < ok.sum() < n
        print("full-scalar path ok", int(ok.sum()))
    """)
    env = dict(os.environ, EDDSA_B200_DEBUG_FULL_SCALARS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "full-scalar path ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_verify_pass_boundaries(ed, cpu):
    """The verify kernels work in passes of whole waves of resident threads and visit the records in sorted order; the
    host pipeline cuts a batch into chunks of four waves (303 104 signatures on 148 SMs) after a one-wave first chunk:
    a batch one past a chunk, with a ragged last warp, through the device API and the host pipeline, every tenth
    signature corrupted, compared with the oracle on a sample and by position (the corrupted ones, and only they, are
    rejected).  (Several passes inside ONE device-API call: test_gpu_round2.py::test_verify_several_passes_per_call.)"""
    import torch
    rng = np.random.default_rng(650)
    n = 303_104 + 1 + 37
    sec, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 64)
    pub = ed.ed25519_genpub_batch(sec)
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)
    bad = np.zeros(n, bool)
    bad[::10] = True
    bad[[n - 1, n - 2, 303_103, 303_104]] = True
    sig[bad, 7] ^= 0x10                                        # R changes: still a curve point or not, never the right one
    dev = torch.device("cuda:0")
    ok_t = torch.empty((n,), dtype=torch.uint8, device=dev)
    ed.ed25519_verify_batch_dev(ok_t, torch.from_numpy(sig).to(dev), torch.from_numpy(pub).to(dev), torch.from_numpy(msgs).to(dev), fixed_len=64)
    ok_dev = ok_t.cpu().numpy()
    ok_host = ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=64)
    assert (ok_dev == ok_host).all()
    assert (ok_host.astype(bool) == ~bad).all()
    sample = np.concatenate([np.arange(n - 64, n), np.arange(303_040, n), rng.integers(0, n, 1500)])
    assert (ok_host[sample] == cpu.verify(sig[sample], pub[sample], msgs[sample], fixed_len=64)).all()


def test_verify_1kb_messages_with_corruption(ed, cpu):
    """Config 5 shape at a size the CPU can check: 1 KB messages, 10 % corrupted / non-canonical."""
    rng = np.random.default_rng(700)
    n = 1 << 13
    sec, msgs = rand_bytes(rng, n, 32), rand_bytes(rng, n, 1024)
    pub = ed.ed25519_genpub_batch(sec)
    sig = ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=1024)
    assert (sig[:512] == cpu.sign(sec[:512], pub[:512], msgs[:512], fixed_len=1024)).all()
    for i in range(0, n, 10):
        c = (i // 10) % 4
        if c == 0:
            sig[i, int(rng.integers(0, 64))] ^= 1
        elif c == 1:
            msgs[i, int(rng.integers(0, 1024))] ^= 1
        elif c == 2:                                            # S + L: still accepted (Q1)
            v = int.from_bytes(sig[i, 32:].tobytes(), "little") + L
            sig[i, 32:] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        else:
            pub[i, 31] ^= 0x80
    got = ed.ed25519_verify_batch(sig, pub, msgs, fixed_len=1024)
    assert (got == cpu.verify(sig, pub, msgs, fixed_len=1024)).all()
    assert 0.85 * n < got.sum() < n


# ------------------------------------------------------------------------------------------------ device-pointer API
def test_device_pointer_api_matches_host_api(ed):
    import torch
    rng = np.random.default_rng(800)
    n = 5000
    sec, msgs, pt = rand_bytes(rng, n, 32), rand_bytes(rng, n, 64), rand_bytes(rng, n, 32)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    sec_t, msg_t, pt_t = t(sec), t(msgs), t(pt)
    pub_t = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    sig_t = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ok_t = torch.empty((n,), dtype=torch.uint8, device=dev)
    out_t = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    before = ed.launch_count()
    ed.ed25519_genpub_batch_dev(pub_t, sec_t)
    ed.ed25519_sign_batch_dev(sig_t, sec_t, pub_t, msg_t, fixed_len=64)
    ed.ed25519_verify_batch_dev(ok_t, sig_t, pub_t, msg_t, fixed_len=64)
    torch.cuda.synchronize()
    assert ed.launch_count() - before == 8              # genpub: hash + comb; sign: nonce + comb + finish; verify: scalars + points + window loop
    pub = ed.ed25519_genpub_batch(sec)
    assert (pub_t.cpu().numpy() == pub).all()
    assert (sig_t.cpu().numpy() == ed.ed25519_sign_batch(sec, pub, msgs, fixed_len=64)).all()
    assert ok_t.all().item()
    ed.x25519_batch_dev(out_t, sec_t, pt_t)
    assert (out_t.cpu().numpy() == ed.x25519_batch(sec, pt)).all()
    ed.x25519_base_batch_dev(out_t, sec_t)
    assert (out_t.cpu().numpy() == ed.x25519_base_batch(sec)).all()


def test_pinned_host_buffers_and_chunking(ed, cpu):
    """Page-locked caller buffers are consumed directly; a tiny chunk budget forces many pipeline chunks."""
    import torch
    rng = np.random.default_rng(900)
    n = 200_000
    sec = torch.from_numpy(rand_bytes(rng, n, 32)).pin_memory().numpy()
    pub = ed.ed25519_genpub_batch(sec)
    sample = rng.integers(0, n, 2000)
    assert (pub[sample] == cpu.genpub(sec[sample])).all()
    assert (pub == ed.ed25519_genpub_batch(np.array(sec))).all()
