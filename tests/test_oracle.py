"""Pins the CPU oracle (oracle/oracle.c) — and, when built, the compiled reference — against the
golden fixtures: the reference's own x25519 KAT table, RFC 8032 vectors, and reference-generated
Ed25519 / adversarial / edge-case vectors.  No GPU needed."""
import hashlib
import os

import numpy as np
import pytest

import golden_util as gu
from edmodel import L

RFC8032 = [  # (secret, public, message, signature)  RFC 8032 section 7.1, tests 1-3
    ("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60",
     "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a", "",
     "e5564300c360ac729086e2cc806e828a84877f1eb8e5d974d873e065224901555fb8821590a33bacc61e39701cf9b46bd25bf5f0595bbe24655141438e7a100b"),
    ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb",
     "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c", "72",
     "92a009a9f0d4cab8720e820b5f642540a2b27b5416503f8fb3762223ebdb69da085ac1e43e15996e458f3613d0f11d8c387b2eaeb4302aeeb00d291612bb0c00"),
    ("c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7",
     "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025", "af82",
     "6291d657deec24024827e69c3abe01a30ce548a284743a445e3680d7db5ac3ac18ff9b538d16f290ae67f760984dc6594a7c15e9716ed28dc027beceea1ec40a"),
]


def impls(oracle):
    from cpu_ref import Reference, have_reference
    out = [oracle]
    if have_reference():
        out.append(Reference())
    return out


def test_sha512_matches_hashlib(oracle):
    rng = np.random.default_rng(1)
    for n in list(range(0, 300)) + [1023, 1024, 1025, 16384]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.sha512(d) == hashlib.sha512(d).digest()


def test_scalar_arithmetic(oracle):
    rng = np.random.default_rng(2)
    cases = [0, 1, L - 1, L, L + 1, 2 * L, 2**252, 2**256 - 1, 2**512 - 1, L * L]
    cases += [int.from_bytes(rng.bytes(64), "little") for _ in range(300)]
    for x in cases:
        nbytes = 64 if x >= 2**256 else 32
        assert int.from_bytes(oracle.sc_reduce(x.to_bytes(nbytes, "little")), "little") == x % L
    for _ in range(200):
        a, b, c = (int.from_bytes(rng.bytes(32), "little") for _ in range(3))
        got = oracle.sc_muladd(a.to_bytes(32, "little"), b.to_bytes(32, "little"), c.to_bytes(32, "little"))
        assert int.from_bytes(got, "little") == ((a % L) * (b % L) + c % L) % L


def test_x25519_reference_kat_table(oracle):
    """All 1024 rows of the reference's test/x25519-table.h (pins Q6/Q7: 508 rows have bit 255 set)."""
    point, scalar, result = gu.x25519_kat()
    assert len(point) == 1024 and int((point[:, 31] & 0x80 != 0).sum()) == 508
    for impl in impls(oracle):
        assert (impl.x25519(scalar, point) == result).all(), impl.kind


def test_x25519_edge_and_base(oracle):
    point, scalar, result = gu.x25519_edge()
    bs, bo = gu.x25519_base_kat()
    for impl in impls(oracle):
        assert (impl.x25519(scalar, point) == result).all(), impl.kind
        assert (impl.x25519_base(bs) == bo).all(), impl.kind


def test_rfc8032_vectors(oracle):
    for impl in impls(oracle):
        for sk, pk, msg, sig in RFC8032:
            sk, pk, msg, sig = (bytes.fromhex(x) for x in (sk, pk, msg, sig))
            m = np.frombuffer(msg, np.uint8)
            assert impl.genpub(np.frombuffer(sk, np.uint8)).tobytes() == pk
            assert impl.sign(np.frombuffer(sk, np.uint8), np.frombuffer(pk, np.uint8), m, fixed_len=len(msg)).tobytes() == sig
            assert impl.verify(np.frombuffer(sig, np.uint8), np.frombuffer(pk, np.uint8), m, fixed_len=len(msg))[0] == 1


def test_ed25519_kat(oracle):
    """1024 reference-generated rows {sec, pub, sig}, message length = row index (shape of selftest-ed25519.c)."""
    sec, pub, sig, msgs = gu.ed25519_kat()
    blob, off = gu.ragged(msgs)
    for impl in impls(oracle):
        assert (impl.genpub(sec) == pub).all(), impl.kind
        assert (impl.sign(sec, pub, blob, off=off) == sig).all(), impl.kind
        assert impl.verify(sig, pub, blob, off=off).all(), impl.kind


def test_adversarial_verify(oracle):
    """Accept/reject decisions of the compiled reference on the SURVEY Q1-Q5 classes."""
    sig, pub, msgs, cls, expect = gu.verify_adv()
    blob, off = gu.ragged(msgs)
    assert 0 < expect.sum() < len(expect)
    for impl in impls(oracle):
        got = impl.verify(sig, pub, blob, off=off)
        bad = np.nonzero(got != expect)[0]
        assert len(bad) == 0, (impl.kind, [(int(i), int(cls[i])) for i in bad[:10]])


def test_sign_with_wrong_pub(oracle):
    sec, pub, sig, msgs = gu.sign_wrongpub()
    blob, off = gu.ragged(msgs)
    for impl in impls(oracle):
        assert (impl.sign(sec, pub, blob, off=off) == sig).all(), impl.kind


def test_key_conversion(oracle):
    edsk, edpk, xsk, xpk = gu.convert_kat()
    for impl in impls(oracle):
        for i in range(len(edsk)):
            assert impl.sk_to_x25519(edsk[i].tobytes()) == xsk[i].tobytes()
            assert impl.pk_to_x25519(edpk[i].tobytes()) == xpk[i].tobytes()


def test_oracle_matches_reference_on_random_inputs(oracle, reference):
    rng = np.random.default_rng(3)
    n = 512
    sec = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    pt = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, 96), dtype=np.uint8)
    pub = reference.genpub(sec)
    assert (oracle.genpub(sec) == pub).all()
    sig = reference.sign(sec, pub, msgs, fixed_len=96)
    assert (oracle.sign(sec, pub, msgs, fixed_len=96) == sig).all()
    sig[::2, rng.integers(0, 64)] ^= 0x10
    assert (oracle.verify(sig, pub, msgs, fixed_len=96) == reference.verify(sig, pub, msgs, fixed_len=96)).all()
    assert (oracle.x25519(sec, pt) == reference.x25519(sec, pt)).all()
    assert (oracle.x25519_base(sec) == reference.x25519_base(sec)).all()


def test_reference_32_and_64_bit_builds_agree():
    """SURVEY Q11."""
    from cpu_ref import Reference, have_reference
    if not (have_reference(64) and have_reference(32)):
        pytest.skip("reference builds not present")
    a, b = Reference(64), Reference(32)
    rng = np.random.default_rng(4)
    sec = rng.integers(0, 256, (256, 32), dtype=np.uint8)
    pt = rng.integers(0, 256, (256, 32), dtype=np.uint8)
    assert (a.genpub(sec) == b.genpub(sec)).all()
    assert (a.x25519(sec, pt) == b.x25519(sec, pt)).all()
