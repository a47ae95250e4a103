"""Constant-time audit of the secret-key kernels (north_star: "checked by a SASS audit").
Static taint analysis of the shipped SASS (tools/ct_audit.py): no branch, memory address or
variable-latency instruction may depend on data loaded through a secret pointer parameter.
A negative control proves the audit actually detects leaks.  No GPU needed (cuobjdump only)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ct_audit  # noqa: E402


@pytest.fixture(scope="module")
def libpath():
    import libeddsa_b200
    if not os.path.exists(libeddsa_b200.LIB_PATH):
        libeddsa_b200.build()
    return libeddsa_b200.LIB_PATH


def test_secret_kernels_are_constant_time(libpath):
    results = ct_audit.run(libpath)
    assert {r["kernel"] for r in results} == {"k_x25519", "k_combILi0", "k_combILi1", "k_expand_key", "k_sign_nonceILb0", "k_sign_nonceILb1",
                                              "k_sign_finishILb0", "k_sign_finishILb1", "k_sk_convert"}
    for r in results:
        assert "error" not in r, r
        assert r["secret_loads"] > 0, f"{r['kernel']}: the audit saw no secret loads (taint source not found)"
        # (k_sign_finish hashes public data; the secrets a, r only enter at S = r + t a)
        assert r["instructions_on_secret_data"] > 0.1 * r["instructions"], r["kernel"]
        assert r["reached"] > 0.95 * r["instructions"], r["kernel"]
        # the batched kernels keep each thread's projective results in per-thread local arrays (scrubbed
        # before exit); the audit then treats EVERY local-memory load as secret, so a PASS also shows that
        # no branch or address was derived from anything that went through the stack
        assert r["violations"] == [], (r["kernel"], r["violations"][:5])


def test_audit_detects_leaks(tmp_path):
    """Negative control: a secret-indexed table lookup and a secret-dependent branch must be flagged;
    the masked-select version of the same lookup must pass."""
    obj = tmp_path / "leaky.o"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-c",
                    os.path.join(ROOT, "tests", "ct_negative", "leaky.cu"), "-o", str(obj)], check=True)
    spec = {name: {"params": ["n", "out", "sec", "table"], "secret": ["sec"]} for name in ("k_leaky_index", "k_leaky_branch", "k_naive_select", "k_clean_select", "k_leaky_smem", "k_leaky_shfl")}
    res = {r["kernel"]: r for r in ct_audit.run(str(obj), specs=spec)}
    kinds = lambda k: {v["kind"] for v in res[k]["violations"]}
    assert "secret-dependent memory address / predicate" in kinds("k_leaky_index")
    assert "secret-dependent control flow" in kinds("k_leaky_branch")
    # nvcc compiles the naive mask idiom into secret-predicated loads; the audit must notice
    assert all(v["kind"] == "secret-dependent memory address / predicate" for v in res["k_naive_select"]["violations"])
    assert res["k_clean_select"]["violations"] == [] and res["k_clean_select"]["secret_loads"] > 0
    # secrets may travel through shared memory and shuffles as data, never as an index or a lane number
    assert "secret-dependent memory address / predicate" in kinds("k_leaky_smem") and res["k_leaky_smem"]["shared_memory_holds_secrets"]
    assert "secret-dependent shuffle lane" in kinds("k_leaky_shfl")
