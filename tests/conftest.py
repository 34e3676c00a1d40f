import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
# the library's cubin cache (NVRTC-compiled generated models) stays inside the repository during the tests
os.environ.setdefault("SDEB200_CACHE_DIR", os.path.join(ROOT, ".pytest_cache", "sdeb200_cubins"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def S():
    """The product package (requires the in-tree libsdeb200.so; built by __graft_entry__.build())."""
    import _pkg
    return _pkg.load()


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu(S):
    import ctypes as C
    n = C.c_int()
    rc = S._lib.lib().sdeb200_device_count(C.byref(n))
    assert rc == 0 and n.value > 0, "GPU tests need a CUDA device"
    S._lib.check(S._lib.lib().sdeb200_init(0))
    return S


BS_P = (0.01, 0.3)
HESTON_P = (0.01, 10.0, 0.3, 0.1, 0.4)
