"""Generates the committed golden fixtures in tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/Makefile):
    python tests/golden/make_golden.py
Fixtures (all little, all deterministic given the seeds below):
  x25519_kat.bin        1024 x {point32, scalar32, result32} parsed verbatim from the reference's
                        own KAT table test/x25519-table.h (struct layout: test/selftest-x25519.c:7-13)
  ed25519_kat.bin       1024 x {sec32, pub32, sig64}; row i signs kat_message(i) of length i, the
                        shape of test/selftest-ed25519.c:31-51 (whose table is a missing blob);
                        produced by the compiled reference, cross-checked against `cryptography`
  verify_adv.bin        adversarial verify rows {sig64, pub32, len_u16, class_u8, expect_u8, msg128}
                        expect = decision of the compiled reference (64-bit build; the 32-bit build
                        must agree or generation fails)
  sign_wrongpub.bin     256 x {sec32, pub32 (NOT the matching key), sig64}, 64-byte kat_message(i+5000) (Q8)
  x25519_edge.bin       edge-case {point32, scalar32, result32} rows from the compiled reference
  x25519_base_kat.bin   512 x {scalar32, out32} from the compiled reference
  convert_kat.bin       256 x {edsk32, edpk32, xsk32, xpk32} (sk/pk_ed25519_to_x25519)
"""
import hashlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import edmodel as em  # noqa: E402
from cpu_ref import Reference, build_oracle  # noqa: E402

REF_ROOT = "/root/reference"


def kat_message(i, length=None):
    """Deterministic message bytes for row i (SHA-512 in counter mode)."""
    length = i if length is None else length
    out = b""
    ctr = 0
    while len(out) < length:
        out += hashlib.sha512(b"libeddsa_b200/kat-msg" + i.to_bytes(4, "little") + ctr.to_bytes(4, "little")).digest()
        ctr += 1
    return out[:length]


def det_bytes(tag, i, n=32):
    return hashlib.sha512(tag + i.to_bytes(4, "little")).digest()[:n]


def main():
    build_oracle()
    ref, ref32 = Reference(64), Reference(32)

    # ---- x25519 KAT from the reference's own table -------------------------------------------
    text = open(os.path.join(REF_ROOT, "test", "x25519-table.h")).read()
    raw = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", text))
    assert len(raw) == 1024 * 96, len(raw)
    kat = np.frombuffer(raw, dtype=np.uint8).reshape(1024, 3, 32)
    assert (ref.x25519(kat[:, 1], kat[:, 0]) == kat[:, 2]).all()
    open(os.path.join(HERE, "x25519_kat.bin"), "wb").write(raw)

    # ---- ed25519 KAT --------------------------------------------------------------------------
    from cryptography.hazmat.primitives.asymmetric import ed25519 as ced
    from cryptography.hazmat.primitives import serialization as ser
    rows = []
    for i in range(1024):
        sk = det_bytes(b"libeddsa_b200/kat-sk", i)
        msg = kat_message(i)
        pk = ref.genpub(np.frombuffer(sk, np.uint8)).tobytes()
        sig = ref.sign(np.frombuffer(sk, np.uint8), np.frombuffer(pk, np.uint8), np.frombuffer(msg, np.uint8), fixed_len=len(msg)).tobytes()
        key = ced.Ed25519PrivateKey.from_private_bytes(sk)
        assert key.public_key().public_bytes(ser.Encoding.Raw, ser.PublicFormat.Raw) == pk
        assert key.sign(msg) == sig
        assert ref.verify(np.frombuffer(sig, np.uint8), np.frombuffer(pk, np.uint8), np.frombuffer(msg, np.uint8), fixed_len=len(msg))[0] == 1
        rows.append(sk + pk + sig)
    open(os.path.join(HERE, "ed25519_kat.bin"), "wb").write(b"".join(rows))
    # RFC 8032 7.1 test 1 sanity
    sk1 = bytes.fromhex("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60")
    assert ref.genpub(np.frombuffer(sk1, np.uint8)).tobytes().hex() == "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"

    # ---- adversarial verify rows ---------------------------------------------------------------
    adv = []

    def emit(cls, sig, pub, msg):
        assert len(msg) <= 128
        s = np.frombuffer(sig, np.uint8); p = np.frombuffer(pub, np.uint8); m = np.frombuffer(msg, np.uint8)
        e64 = int(ref.verify(s, p, m, fixed_len=len(msg))[0])
        e32 = int(ref32.verify(s, p, m, fixed_len=len(msg))[0])
        assert e64 == e32, ("reference builds disagree", cls, sig.hex(), pub.hex())
        adv.append(sig + pub + len(msg).to_bytes(2, "little") + bytes([cls, e64]) + msg.ljust(128, b"\0"))
        return e64

    rng = np.random.default_rng(0xADE25519)

    def valid(i, mlen):
        sk = det_bytes(b"libeddsa_b200/adv-sk", i)
        msg = kat_message(100000 + i, mlen)
        pk = ref.genpub(np.frombuffer(sk, np.uint8)).tobytes()
        sig = ref.sign(np.frombuffer(sk, np.uint8), np.frombuffer(pk, np.uint8), np.frombuffer(msg, np.uint8), fixed_len=mlen).tobytes()
        return sk, pk, sig, msg

    def flip(b, bit):
        b = bytearray(b); b[bit >> 3] ^= 1 << (bit & 7); return bytes(b)

    lens = [0, 1, 31, 32, 47, 48, 63, 64, 65, 111, 112, 127, 128]
    for i in range(64):
        sk, pk, sig, msg = valid(i, lens[i % len(lens)])
        assert emit(0, sig, pk, msg) == 1                                   # class 0: valid
        if msg:
            emit(1, sig, pk, flip(msg, int(rng.integers(0, 8 * len(msg)))))  # 1: message bit flip
        emit(2, flip(sig, int(rng.integers(0, 256))), pk, msg)              # 2: R bit flip
        emit(3, flip(sig, 256 + int(rng.integers(0, 256))), pk, msg)        # 3: S bit flip
        emit(4, sig, flip(pk, int(rng.integers(0, 255))), msg)              # 4: A bit flip (usually off-curve or other point)
        emit(12, sig, flip(pk, 255), msg)                                   # 12: sign-flipped A
        if len(msg) > 1:
            emit(14, sig, pk, msg[:-1])                                     # 14: truncated message
        if len(msg) < 128:
            emit(14, sig, pk, msg + b"\0")
        s = int.from_bytes(sig[32:], "little")
        for k in range(1, 17):                                              # 5: S + kL (Q1)
            v = s + k * em.L
            if v < 2**256:
                emit(5, sig[:32] + v.to_bytes(32, "little"), pk, msg)
        emit(5, sig[:32] + ((s - em.L) % 2**256).to_bytes(32, "little"), pk, msg)

    ident, ident_nc = em.enc((0, 1)), em.enc((0, 1), noncanonical=True)
    m0 = b"identity key"
    zero32, Lb = bytes(32), em.L.to_bytes(32, "little")
    for a_enc in (ident, ident_nc, em.enc((0, 1), flip_sign=True), em.enc((0, 1), noncanonical=True, flip_sign=True)):
        for r_enc in (ident, ident_nc, em.enc((0, 1), flip_sign=True)):
            for s_enc in (zero32, Lb, (2 * em.L).to_bytes(32, "little"), (1).to_bytes(32, "little")):
                emit(6, r_enc + s_enc, a_enc, m0)                           # 6-9: identity / non-canonical / negative zero (Q2, Q3)
    emit(6, bytes(64), bytes(32), b"")
    emit(6, bytes(64), bytes(32), m0)
    emit(6, b"\xff" * 64, b"\xff" * 32, m0)

    small = em.small_order_points()
    for ai, A in enumerate(small):                                          # 10: small-order A, all small-order R candidates
        encs = [em.enc(A)]
        if A[1] < 19:
            encs.append(em.enc(A, noncanonical=True))
        if A[0] == 0:
            encs.append(em.enc(A, flip_sign=True))
        for a_enc in encs:
            for trial in range(6):
                msg = kat_message(200000 + 8 * ai + trial, 16 + trial)
                for R in small:
                    emit(10, em.enc(R) + zero32, a_enc, msg)
                    if R[1] < 19:
                        emit(7, em.enc(R, noncanonical=True) + zero32, a_enc, msg)   # 7: non-canonical R

    t8 = em.torsion8()
    for i in range(96):                                                     # 11: mixed-order A, honest signer (Q4)
        sk = det_bytes(b"libeddsa_b200/mixed-sk", i)
        a, prefix = em.clamp_scalar(sk)
        k = 1 + (i % 7)
        A = em.add(em.mul(a, em.B), em.mul(k, t8))
        msg = kat_message(300000 + i, 10 + (i % 100))
        a_enc = em.enc(A)
        emit(11, em.sign_with(a, prefix, a_enc, msg), a_enc, msg)

    n_off = 0
    i = 0
    while n_off < 96:                                                       # 13: off-curve A (policy: reject)
        cand = det_bytes(b"libeddsa_b200/offcurve", i); i += 1
        if em.on_curve_y(cand):
            continue
        sk, pk, sig, msg = valid(1000 + n_off, 40)
        assert emit(13, sig, cand, msg) == 0
        n_off += 1
    for i in range(32):                                                     # 15: random / off-curve R
        sk, pk, sig, msg = valid(2000 + i, 33)
        emit(15, det_bytes(b"libeddsa_b200/randR", i) + sig[32:], pk, msg)
    open(os.path.join(HERE, "verify_adv.bin"), "wb").write(b"".join(adv))
    arr = np.frombuffer(b"".join(adv), np.uint8).reshape(-1, 228)
    print("verify_adv rows", len(arr), "accepted", int(arr[:, 99].sum()),
          {c: (int((arr[:, 98] == c).sum()), int(arr[arr[:, 98] == c][:, 99].sum())) for c in sorted(set(arr[:, 98]))})

    # ---- sign with a wrong public key (Q8) ------------------------------------------------------
    rows = []
    for i in range(256):
        sk = det_bytes(b"libeddsa_b200/wp-sk", i)
        pub = det_bytes(b"libeddsa_b200/wp-pub", i)          # arbitrary bytes, not the key of sk
        msg = kat_message(5000 + i, 64)
        sig = ref.sign(np.frombuffer(sk, np.uint8), np.frombuffer(pub, np.uint8), np.frombuffer(msg, np.uint8), fixed_len=64).tobytes()
        assert sig == ref32.sign(np.frombuffer(sk, np.uint8), np.frombuffer(pub, np.uint8), np.frombuffer(msg, np.uint8), fixed_len=64).tobytes()
        rows.append(sk + pub + sig)
    open(os.path.join(HERE, "sign_wrongpub.bin"), "wb").write(b"".join(rows))

    # ---- x25519 edge cases ----------------------------------------------------------------------
    P = em.P
    pts = [0, 1, 2, 9, P - 1, P, P + 1, P + 9, 2**255 - 1, 2**255, 2**255 + 9, 2**256 - 1, 2**255 - 19 + 2**254,
           325606250916557431795983626356110631294008115727848805560023387167927233504,     # low-order u (order 8)
           39382357235489614581723060781553021112529911719440698176882885853963445705823,
           P - 1 + 2**255, 19, 2**255 + 18, 2**255 - 20]
    scs = [bytes(32), b"\xff" * 32, (1).to_bytes(32, "little"), (8).to_bytes(32, "little"), det_bytes(b"libeddsa_b200/edge-sc", 0),
           det_bytes(b"libeddsa_b200/edge-sc", 1), (2**254).to_bytes(32, "little"), (2**255 - 8).to_bytes(32, "little")]
    rows = []
    for pt in pts:
        for sc in scs:
            pb = pt.to_bytes(32, "little")
            out = ref.x25519(np.frombuffer(sc, np.uint8), np.frombuffer(pb, np.uint8)).tobytes()
            assert out == ref32.x25519(np.frombuffer(sc, np.uint8), np.frombuffer(pb, np.uint8)).tobytes()
            rows.append(pb + sc + out)
    open(os.path.join(HERE, "x25519_edge.bin"), "wb").write(b"".join(rows))

    # ---- x25519_base ----------------------------------------------------------------------------
    sc = np.stack([np.frombuffer(det_bytes(b"libeddsa_b200/base-sc", i), np.uint8) for i in range(504)] +
                  [np.frombuffer(x, np.uint8) for x in scs])
    out = ref.x25519_base(sc)
    assert (out == ref32.x25519_base(sc)).all()
    assert (out == ref.x25519(sc, np.tile(np.frombuffer((9).to_bytes(32, "little"), np.uint8), (len(sc), 1)))).all()
    open(os.path.join(HERE, "x25519_base_kat.bin"), "wb").write(np.concatenate([sc, out], axis=1).tobytes())

    # ---- key conversion -------------------------------------------------------------------------
    rows = []
    for i in range(256):
        sk = det_bytes(b"libeddsa_b200/conv-sk", i)
        pk = ref.genpub(np.frombuffer(sk, np.uint8)).tobytes()
        if i >= 224:   # arbitrary (possibly off-curve / non-canonical) public-key bytes
            pk = det_bytes(b"libeddsa_b200/conv-pk", i)
        if i == 255:
            pk = em.enc((0, 1))          # y = 1 -> z - y = 0 -> inv(0) = 0 -> u = 0
        rows.append(sk + pk + ref.sk_to_x25519(sk) + ref.pk_to_x25519(pk))
    open(os.path.join(HERE, "convert_kat.bin"), "wb").write(b"".join(rows))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".bin"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
