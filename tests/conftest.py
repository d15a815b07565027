import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _gpu_count():
    try:
        import nextgp.jl_b200 as ngp
        return ngp._lib.lib().ngp_device_count()
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu():
    if _gpu_count() == 0:
        pytest.fail("test marked gpu but no CUDA device / libngp.so available (no CPU fallback exists)")
    return 0
