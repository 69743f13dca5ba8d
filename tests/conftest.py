import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    """Number of CUDA devices according to the driver (no torch import, no context created)."""
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        if cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return int(n.value)
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    # A plain `pytest` run on a machine without a GPU skips the gpu-marked tests instead of failing in
    # bpomp_create (there is no CPU fallback to run them on).  `-m gpu` on the B200 box runs them all.
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests need the B200 box (no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle
