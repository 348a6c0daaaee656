import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "galileo-sdr-sim_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    try:
        import ctypes
        n = ctypes.c_int(0)
        rt = ctypes.CDLL("libcudart.so")
        return rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a CUDA device skips the gpu-marked tests instead of failing
    them (`-m gpu` on such a machine reports them as skipped, never as passed)."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the gpu-marked tests run on the B200 box (python -m pytest tests -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
