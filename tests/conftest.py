import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    """devices the CUDA driver sees, without creating a context (the device library aborts, by design, when asked for a device that is
    not there: a plain `pytest tests` on a CPU-only machine must skip the GPU tests, not die in create_device)"""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if cu.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: GPU parity tests run on the B200 box (-m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def device():
    """One b200 device for the whole GPU session; created through luisa_compute_lib_interface like the Rust frontend would."""
    import luisa_compute_rs_b200 as lc
    ctx = lc.Context()
    dev = ctx.create_device("b200")
    yield dev
    dev.close()
