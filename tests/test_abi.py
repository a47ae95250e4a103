"""The drop-in boundary: libeddsa_b200.so builds, loads, and exports exactly the symbols declared
in include/eddsa.h and include/eddsa_batch.h (no compute calls here — those need a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    import libeddsa_b200
    if not os.path.exists(libeddsa_b200.LIB_PATH):
        libeddsa_b200.build()
    return libeddsa_b200.LIB_PATH


def declared_symbols():
    names = []
    for hdr in ("eddsa.h", "eddsa_batch.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"EDDSA_DECL\s+(?:const\s+)?[A-Za-z_ ]+?\**\s*\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    return sorted(set(names))


def test_headers_declare_reference_api():
    names = declared_symbols()
    # the 13 symbols the reference exports ...
    for n in ["ed25519_genpub", "ed25519_sign", "ed25519_verify", "x25519_base", "x25519", "pk_ed25519_to_x25519",
              "sk_ed25519_to_x25519", "eddsa_genpub", "eddsa_sign", "eddsa_verify", "DH", "eddsa_pk_eddsa_to_dh",
              "eddsa_sk_eddsa_to_dh"]:
        assert n in names
    # ... and the batch C-ABI
    for n in ["ed25519_genpub_batch", "ed25519_sign_batch", "ed25519_verify_batch", "x25519_batch", "x25519_base_batch"]:
        assert n in names and n + "_dev" in names


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for n in declared_symbols():
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared_symbols()


def test_headers_compile_as_c_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "eddsa_batch.h"\nint main(void){return ED25519_KEY_LEN+ED25519_SIG_LEN+X25519_KEY_LEN==128?0:1;}\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)
    subprocess.run(["g++", "-x", "c++", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "t2.o")], check=True)


def test_no_cpu_fallback_without_device(libpath):
    """Without a CUDA device every batch entry point reports an error instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import libeddsa_b200 as ed
    with pytest.raises(ed.EddsaB200Error):
        ed.x25519_base_batch(np.zeros((4, 32), np.uint8))
    assert ed.device_count() == 0


def test_product_does_not_reference_oracle():
    """Nothing under libeddsa_b200/ may import, include or link the oracle — or the test-only host builds and the CUDA
    runtime simulator of tests/host_sim — and the shipped library contains no simulator symbol."""
    pkg = os.path.join(ROOT, "libeddsa_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h", ".inc", "Makefile", ".map")):
                text = open(os.path.join(dp, f), errors="replace").read().lower()
                assert "oracle" not in text, (dp, f)
                assert "cudasim" not in text and "libeddsa_sim" not in text, (dp, f)
    so = os.path.join(pkg, "libeddsa_b200.so")
    if os.path.exists(so):
        syms = subprocess.run(["nm", "-D", "--defined-only", so], stdout=subprocess.PIPE, text=True, check=True).stdout
        assert "cudasim" not in syms
