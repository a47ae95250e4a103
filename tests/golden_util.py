"""Loaders for the committed fixtures in tests/golden/ (written by tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kat_message(i, length=None):
    """Deterministic message bytes for row i (same generator as make_golden.py)."""
    length = i if length is None else length
    out = b""
    ctr = 0
    while len(out) < length:
        out += hashlib.sha512(b"libeddsa_b200/kat-msg" + i.to_bytes(4, "little") + ctr.to_bytes(4, "little")).digest()
        ctr += 1
    return out[:length]


def _load(name, width):
    return np.fromfile(os.path.join(GOLDEN, name), dtype=np.uint8).reshape(-1, width)


def x25519_kat():
    """(point, scalar, result) — the reference's own test/x25519-table.h, 1024 rows."""
    a = _load("x25519_kat.bin", 96)
    return a[:, :32].copy(), a[:, 32:64].copy(), a[:, 64:].copy()


def x25519_edge():
    a = _load("x25519_edge.bin", 96)
    return a[:, :32].copy(), a[:, 32:64].copy(), a[:, 64:].copy()


def x25519_base_kat():
    a = _load("x25519_base_kat.bin", 64)
    return a[:, :32].copy(), a[:, 32:].copy()


def ragged(msgs):
    """list of bytes -> (blob uint8, offsets uint64[n+1])"""
    lens = np.array([len(m) for m in msgs], dtype=np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    blob = np.frombuffer(b"".join(msgs), dtype=np.uint8).copy() if off[-1] else np.zeros(0, np.uint8)
    return blob, off


def ed25519_kat():
    """(sec, pub, sig, [msg_i]) — 1024 rows, message length = row index."""
    a = _load("ed25519_kat.bin", 128)
    return a[:, :32].copy(), a[:, 32:64].copy(), a[:, 64:].copy(), [kat_message(i) for i in range(len(a))]


def verify_adv():
    """(sig, pub, [msg], cls, expect) adversarial rows; expect = decision of the compiled reference."""
    a = np.concatenate([_load("verify_adv.bin", 228), _load("verify_torsion.bin", 228)])   # + mixed-order A and R rows
    lens = a[:, 96].astype(np.int64) | (a[:, 97].astype(np.int64) << 8)
    msgs = [a[i, 100:100 + lens[i]].tobytes() for i in range(len(a))]
    return a[:, :64].copy(), a[:, 64:96].copy(), msgs, a[:, 98].copy(), a[:, 99].copy()


def sign_wrongpub():
    a = _load("sign_wrongpub.bin", 128)
    return a[:, :32].copy(), a[:, 32:64].copy(), a[:, 64:].copy(), [kat_message(5000 + i, 64) for i in range(len(a))]


def convert_kat():
    a = _load("convert_kat.bin", 128)
    return a[:, :32].copy(), a[:, 32:64].copy(), a[:, 64:96].copy(), a[:, 96:].copy()
