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
    import subprocess
    from oracle import oracle
    oracle.build()
    # the reference build (oracle/_ref) is (re)made where /root/reference exists; elsewhere the copy that
    # travelled with the snapshot is used, and the tests that need it skip when there is none
    if os.path.exists("/root/reference/src/BatchingPOMP.cpp"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    return oracle
