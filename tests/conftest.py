import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    """True when the CUDA driver reports at least one device (no torch import: keeps collection fast)."""
    import ctypes

    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a GPU skips the gpu-marked tests instead of failing in them."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests run on the B200 box (pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def libmpx():
    """Build (if needed) and load the C-ABI library."""
    import __graft_entry__ as ge

    ge.build()
    from mpopt_b200 import _lib

    return _lib.lib()
