"""Dynamic constant-time check on the GPU (SURVEY.md §7 hard part 7): the secret-key kernels execute the same number
of warp and thread instructions and issue the same global / local / shared memory requests, sectors and wavefronts
whether the secret keys are all-zero, all-one or random (tools/ct_dynamic.py, ncu counters)."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.gpu
def test_secret_kernels_have_secret_independent_counters(tmp_path):
    if not shutil.which("ncu"):
        pytest.skip("ncu not installed")
    import ct_dynamic
    try:
        res = ct_dynamic.run(n=4096, workdir=str(tmp_path))
    except RuntimeError as e:
        if "ERR_NVGPUCTRPERM" in str(e) or "permission" in str(e).lower():
            pytest.skip("no permission to read GPU performance counters")
        raise
    kernels = {r["kernel"] for r in res["launches"]}
    for k in ("k_comb<0>", "k_comb<1>", "k_x25519", "k_expand_key", "k_sign_nonce<0>", "k_sign_nonce<1>", "k_sign_finish<0>", "k_sign_finish<1>", "k_sk_convert"):
        assert any(k in name for name in kernels), (k, kernels)
    bad = [(r["kernel"], m.split("__")[1], v) for r in res["launches"] if not r["identical"] for m, v in r["counters"].items() if isinstance(v, dict)]
    assert not bad, bad
