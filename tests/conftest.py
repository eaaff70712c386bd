import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pth in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if pth not in sys.path:
        sys.path.insert(0, pth)


# the test suite JIT-compiles ~60 models: keep their cubins out of the in-tree cache that ships with the library
os.environ.setdefault("B200ENS_CACHE_DIR", os.path.join(os.environ.get("TMPDIR", "/tmp"), "b200ens_test_cubin_cache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand from oracle/."""
    import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def B():
    import b200ens

    return b200ens


@pytest.fixture(scope="session")
def gpu_lib(B):
    """libb200ens with a usable device; GPU tests fail (not skip) if the extension is missing."""
    n = B._lib.lib().b200ens_device_count()
    assert n > 0, "no CUDA device visible to libb200ens: " + B._lib.lib().b200ens_last_error().decode()
    return B._lib
