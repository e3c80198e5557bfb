import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_cuda():
    """True when a CUDA device is usable (driver library present and >= 1 device)."""
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return False
    n = ctypes.c_int(0)
    if cuda.cuInit(0) != 0 or cuda.cuDeviceGetCount(ctypes.byref(n)) != 0:
        return False
    return n.value > 0


def pytest_collection_modifyitems(config, items):
    # a plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of
    # erroring in them; on a GPU box nothing is skipped (a missing extension still fails loudly)
    if _have_cuda():
        return
    import pytest
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
