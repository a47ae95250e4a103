"""TEST INFRASTRUCTURE: ctypes access to the CPU checkers.

* ``Oracle``  — oracle/liboracle.so, the C restatement (oracle/oracle.c).
* ``Reference`` — oracle/_ref/libeddsa_ref.so, the unmodified reference compiled by
  oracle/Makefile (present when it was built in the container; travels to the GPU box).

Both expose the same batch interface (numpy uint8 arrays in / out), driven through the threaded
harness in oracle/harness.c.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
OP_GENPUB, OP_SIGN, OP_VERIFY, OP_X25519, OP_X25519_BASE = range(5)


def build_oracle():
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(ctypes.c_void_p)


class _CpuImpl:
    """Batch front-end over five eddsa.h-shaped function pointers."""

    kind = "?"

    def __init__(self, fns):
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build_oracle()
        self._h = ctypes.CDLL(so)
        self._h.harness_run.restype = ctypes.c_double
        self._h.harness_run.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        self._fns = fns
        self.threads = os.cpu_count() or 1

    def _run(self, op, n, out, a, b=None, msgs=None, off=None, fixed_len=0, seconds=0.0, threads=None):
        fn = ctypes.cast(self._fns[op], ctypes.c_void_p)
        done = ctypes.c_size_t(0)
        keep = []

        def ptr(x):
            if x is None:
                return None
            arr, p = _u8(x)
            keep.append(arr)
            return p

        offp = None
        if off is not None:
            off = np.ascontiguousarray(off, dtype=np.uint64)
            keep.append(off)
            offp = off.ctypes.data_as(ctypes.c_void_p)
        wall = self._h.harness_run(op, fn, threads or self.threads, float(seconds), n, out.ctypes.data_as(ctypes.c_void_p),
                                   ptr(a), ptr(b), ptr(msgs), offp, fixed_len, ctypes.byref(done))
        return wall, done.value

    # ---- batch API (same argument meaning as include/eddsa_batch.h) ----
    def genpub(self, sec):
        sec = np.ascontiguousarray(sec, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros_like(sec)
        self._run(OP_GENPUB, len(sec), out, sec)
        return out

    def sign(self, sec, pub, msgs, off=None, fixed_len=0):
        sec = np.ascontiguousarray(sec, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros((len(sec), 64), dtype=np.uint8)
        self._run(OP_SIGN, len(sec), out, sec, pub, msgs, off, fixed_len)
        return out

    def verify(self, sig, pub, msgs, off=None, fixed_len=0):
        sig = np.ascontiguousarray(sig, dtype=np.uint8).reshape(-1, 64)
        out = np.zeros(len(sig), dtype=np.uint8)
        self._run(OP_VERIFY, len(sig), out, sig, pub, msgs, off, fixed_len)
        return out

    def x25519(self, scalar, point):
        scalar = np.ascontiguousarray(scalar, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros_like(scalar)
        self._run(OP_X25519, len(scalar), out, scalar, point)
        return out

    def x25519_base(self, scalar):
        scalar = np.ascontiguousarray(scalar, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros_like(scalar)
        self._run(OP_X25519_BASE, len(scalar), out, scalar)
        return out

    def time_op(self, op, seconds, threads, n, a, b=None, msgs=None, fixed_len=0):
        """Timing mode: every thread loops over its slice for `seconds`; returns ops/s."""
        width = {OP_GENPUB: 32, OP_SIGN: 64, OP_VERIFY: 1, OP_X25519: 32, OP_X25519_BASE: 32}[op]
        out = np.zeros((n, width), dtype=np.uint8)
        wall, done = self._run(op, n, out, a, b, msgs, None, fixed_len, seconds=seconds, threads=threads)
        return done / wall


class Oracle(_CpuImpl):
    kind = "port"

    def __init__(self):
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            build_oracle()
        lib = ctypes.CDLL(so)
        self.lib = lib
        lib.oracle_ed25519_verify.restype = ctypes.c_int
        super().__init__({OP_GENPUB: lib.oracle_ed25519_genpub, OP_SIGN: lib.oracle_ed25519_sign,
                          OP_VERIFY: lib.oracle_ed25519_verify, OP_X25519: lib.oracle_x25519,
                          OP_X25519_BASE: lib.oracle_x25519_base})

    def sha512(self, data: bytes) -> bytes:
        out = ctypes.create_string_buffer(64)
        self.lib.oracle_sha512(out, data, ctypes.c_size_t(len(data)))
        return out.raw

    def sc_reduce(self, data: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.oracle_sc_reduce(out, data, ctypes.c_size_t(len(data)))
        return out.raw

    def sc_muladd(self, a: bytes, b: bytes, c: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.oracle_sc_muladd(out, a, b, c)
        return out.raw

    def pk_to_x25519(self, pk: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.oracle_pk_ed25519_to_x25519(out, pk)
        return out.raw

    def sk_to_x25519(self, sk: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.oracle_sk_ed25519_to_x25519(out, sk)
        return out.raw


def reference_path(bits=64):
    return os.path.join(ORACLE_DIR, "_ref", "libeddsa_ref.so" if bits == 64 else "libeddsa_ref32.so")


def have_reference(bits=64):
    return os.path.exists(reference_path(bits))


class Reference(_CpuImpl):
    """The unmodified reference library (eddsa.h API: /root/reference/lib/eddsa.h:44-81)."""
    kind = "reference"

    def __init__(self, bits=64):
        lib = ctypes.CDLL(reference_path(bits))
        self.lib = lib
        lib.ed25519_verify.restype = ctypes.c_bool
        super().__init__({OP_GENPUB: lib.ed25519_genpub, OP_SIGN: lib.ed25519_sign, OP_VERIFY: lib.ed25519_verify,
                          OP_X25519: lib.x25519, OP_X25519_BASE: lib.x25519_base})

    def pk_to_x25519(self, pk: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.pk_ed25519_to_x25519(out, pk)
        return out.raw

    def sk_to_x25519(self, sk: bytes) -> bytes:
        out = ctypes.create_string_buffer(32)
        self.lib.sk_ed25519_to_x25519(out, sk)
        return out.raw


def best_cpu_impl():
    """Reference when its compiled library is available, else the oracle port."""
    return Reference() if have_reference() else Oracle()
