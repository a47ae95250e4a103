import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: full-size configuration (tens of seconds on a B200; part of `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from cpu_ref import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from cpu_ref import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libeddsa_ref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def cpu():
    """Best available CPU checker: the compiled reference if present, else the oracle port."""
    from cpu_ref import best_cpu_impl
    return best_cpu_impl()


@pytest.fixture(scope="session")
def ed():
    """The product binding; requires the built CUDA library (no fallback)."""
    import libeddsa_b200
    libeddsa_b200.lib()
    return libeddsa_b200
