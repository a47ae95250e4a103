"""The reference's own acceptance tests, UNMODIFIED, against this library ("the cheapest drop-in
acceptance test there is").

oracle/Makefile compiles /root/reference/test/selftest-{x25519,x25519_base,convert,ed25519}.c (registered at
/root/reference/test/CMakeLists.txt:17-20) from the sources where they lie, against include/eddsa.h (the product's
header), linked against the REFERENCE's shared library libeddsa.so.0 (soname as lib/CMakeLists.txt:43-44).  The
binaries travel to the GPU box in oracle/_ref/selftests/.  Here they run
  * with the reference library on the library path (CPU; shows the binaries and the regenerated ed25519-table.h are good),
  * with libeddsa.so.0 -> libeddsa_b200.so on the library path (GPU): same binaries, no relink — libeddsa_b200.so carries
    the reference's soname, so the dynamic loader accepts it in the reference's place.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ST = os.path.join(ROOT, "oracle", "_ref", "selftests")
TESTS = ["x25519", "x25519_base", "convert", "ed25519"]


def _need_binaries():
    missing = [t for t in TESTS if not os.path.exists(os.path.join(ST, "selftest-" + t))]
    if missing or not os.path.exists(os.path.join(ST, "libeddsa.so.0")):
        pytest.skip("oracle/_ref/selftests not built (needs /root/reference at build time)")


def _run(test, libdir):
    env = dict(os.environ, LD_LIBRARY_PATH=libdir)
    exe = os.path.join(ST, "selftest-" + test)
    trace = subprocess.run([exe], env=dict(env, LD_TRACE_LOADED_OBJECTS="1"), capture_output=True, text=True).stdout
    line = next(l for l in trace.splitlines() if "libeddsa.so.0" in l)
    resolved = os.path.realpath(line.split("=>")[1].split("(")[0].strip())
    res = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=600)
    return resolved, res


@pytest.mark.parametrize("test", TESTS)
def test_reference_selftests_with_reference_library(test):
    _need_binaries()
    resolved, res = _run(test, ST)
    assert resolved == os.path.realpath(os.path.join(ST, "libeddsa.so.0"))
    assert res.returncode == 0, res.stderr


def test_library_carries_the_reference_soname():
    import libeddsa_b200
    out = subprocess.run(["readelf", "-d", libeddsa_b200.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "Library soname: [libeddsa.so.0]" in out


@pytest.mark.parametrize("test", TESTS)
def test_reference_selftests_with_the_host_layer_on_the_simulator(test, tmp_path):
    """No GPU: the same unmodified binaries with libeddsa.so.0 -> tests/host_sim/libeddsa_sim.so — the product's host layer
    (host.c, every call a batch of one through its whole pipeline) on the CUDA runtime simulator, whose kernels run the per-thread
    operation bodies of ops.cuh.  Test infrastructure; the product itself has no CPU path."""
    _need_binaries()
    hs = os.path.join(ROOT, "tests", "host_sim")
    subprocess.run(["make", "-s", "-C", hs, "libeddsa_sim.so"], check=True)
    os.symlink(os.path.join(hs, "libeddsa_sim.so"), tmp_path / "libeddsa.so.0")
    env_keep = {k: os.environ.pop(k) for k in list(os.environ) if k.startswith(("CUDASIM_", "EDDSA_B200_"))}
    os.environ["CUDASIM_DEVICES"] = "1"
    try:
        resolved, res = _run(test, str(tmp_path))
    finally:
        del os.environ["CUDASIM_DEVICES"]
        os.environ.update(env_keep)
    assert resolved == os.path.realpath(os.path.join(hs, "libeddsa_sim.so"))
    assert res.returncode == 0, res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("test", TESTS)
def test_reference_selftests_with_b200_library(test, tmp_path):
    """1024 rows each: x25519 against the reference's own KAT table, x25519_base == x25519(., 9), key conversion commutes
    (new and obsolete API), genpub / sign / verify of the Ed25519 table with message lengths 0..1023 (new and obsolete API)
    — every call a batch of one on the GPU."""
    _need_binaries()
    import libeddsa_b200
    os.symlink(libeddsa_b200.LIB_PATH, tmp_path / "libeddsa.so.0")
    resolved, res = _run(test, str(tmp_path))
    assert resolved == os.path.realpath(libeddsa_b200.LIB_PATH)
    assert res.returncode == 0, res.stderr
