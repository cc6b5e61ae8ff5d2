import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def device():
    """One b200 device for the whole GPU session; created through luisa_compute_lib_interface like the Rust frontend would."""
    import luisa_compute_rs_b200 as lc
    ctx = lc.Context()
    dev = ctx.create_device("b200")
    yield dev
    dev.close()
