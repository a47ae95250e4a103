"""Generates tests/golden/verify_torsion.bin FROM THE REFERENCE ITSELF: verify rows whose public key AND whose R
carry components of order 2, 4 or 8 (same row format as verify_adv.bin: {sig64, pub32, len_u16, class_u8,
expect_u8, msg128}; expect = decision of the compiled reference, 64- and 32-bit-limb builds must agree).

Why: the engine decides  encode(S*B - t*A) == R  through  rho*(S*B - t*A - R') = O  with half-size scalars
(libeddsa_b200/csrc/hgcd.cuh).  That is only exact when rho is odd and tau = rho*t holds modulo the FULL group
order 8L — exactly the property these rows pin: for A = a*B + kA*T8, R = r*B + kR*T8, S = r + t*a the residual
S*B - t*A - R equals -(t*kA + kR)*T8, so the reference accepts iff t*kA + kR = 0 (mod 8).
  class 16  residual 0: accepted although A and R are mixed-order points
  class 17  residual of order 2, 4 or 8: rejected (a verifier that multiplies through by an even number, or
            reduces the A-scalar modulo L only, gets these wrong)
  class 18  class-16 rows with S replaced by S + k*L (no range check on S)
  class 19  R replaced by another valid point (R + B, -R, 2R) or by the honest R of a different message
Run in the build container:  python tests/golden/make_golden_torsion.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import edmodel as em  # noqa: E402
from cpu_ref import Reference, build_oracle  # noqa: E402
from make_golden import det_bytes, kat_message  # noqa: E402


def main():
    build_oracle()
    ref, ref32 = Reference(64), Reference(32)
    rows = []

    def emit(cls, sig, pub, msg):
        s = np.frombuffer(sig, np.uint8); p = np.frombuffer(pub, np.uint8); m = np.frombuffer(msg, np.uint8)
        e64 = int(ref.verify(s, p, m, fixed_len=len(msg))[0])
        e32 = int(ref32.verify(s, p, m, fixed_len=len(msg))[0])
        assert e64 == e32, ("reference builds disagree", cls, sig.hex(), pub.hex())
        rows.append(sig + pub + len(msg).to_bytes(2, "little") + bytes([cls, e64]) + msg.ljust(128, b"\0"))
        return e64

    t8 = em.torsion8()
    tors = [em.mul(k, t8) for k in range(8)]
    n16 = 0
    for i in range(88):
        sk = det_bytes(b"libeddsa_b200/tors-sk", i)
        a, prefix = em.clamp_scalar(sk)
        if i < 72:                                                          # odd kA: every residual is reachable
            kA, kR, want = 1 + 2 * (i % 4), (3 * i + 1) % 8, [0, 0, 4, 2, 6, 1, 3, 5, 7][i % 9]
        else:                                                               # even kA (incl. 0): whatever residual comes
            kA, kR, want = 2 * (i % 4), i % 8, None
        A = em.add(em.mul(a, em.B), tors[kA])
        a_enc = em.enc(A)
        ctr = 0
        while True:
            msg = kat_message(400000 + 1000 * i + ctr, 20 + (i % 60))
            r = em.h_mod_l(prefix, msg)
            R = em.enc(em.add(em.mul(r, em.B), tors[kR]))
            t = em.h_mod_l(R, a_enc, msg)
            res = (t * kA + kR) % 8
            if want is None or res == want:
                break
            ctr += 1
        S = (r + t * a) % em.L
        sig = R + S.to_bytes(32, "little")
        got = emit(16 if res == 0 else 17, sig, a_enc, msg)
        assert got == (1 if res == 0 else 0), (i, kA, kR, res, got)
        if res == 0:
            n16 += 1
            for k in (1, 7, 15):
                v = S + k * em.L
                if v < 2**256:
                    assert emit(18, R + v.to_bytes(32, "little"), a_enc, msg) == 1
    assert n16 >= 8
    for i in range(24):                                                     # 19: R swapped for another valid point
        sk = det_bytes(b"libeddsa_b200/tors-sk", 100 + i)
        a, prefix = em.clamp_scalar(sk)
        a_enc = em.enc(em.mul(a, em.B))
        msg = kat_message(500000 + i, 30 + i)
        sig = em.sign_with(a, prefix, a_enc, msg)
        assert emit(0, sig, a_enc, msg) == 1
        x = em.recover_x(int.from_bytes(sig[:32], "little") & ((1 << 255) - 1), sig[31] >> 7)
        Rp = (x, int.from_bytes(sig[:32], "little") & ((1 << 255) - 1))
        for alt in (em.add(Rp, em.B), em.neg(Rp), em.add(Rp, Rp), em.add(Rp, tors[4]), em.add(Rp, tors[1])):
            assert emit(19, em.enc(alt) + sig[32:], a_enc, msg) == 0
        other = em.sign_with(a, prefix, a_enc, msg + b"!")
        assert emit(19, other[:32] + sig[32:], a_enc, msg) == 0
    open(os.path.join(HERE, "verify_torsion.bin"), "wb").write(b"".join(rows))
    arr = np.frombuffer(b"".join(rows), np.uint8).reshape(-1, 228)
    print("verify_torsion rows", len(arr), {int(c): (int((arr[:, 98] == c).sum()), int(arr[arr[:, 98] == c][:, 99].sum())) for c in sorted(set(arr[:, 98]))})


if __name__ == "__main__":
    main()
