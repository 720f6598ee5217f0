import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def b3d():
    """The product package (name starts with a digit => importlib)."""
    return importlib.import_module("3dvision_b200")


@pytest.fixture(scope="session")
def oracle():
    """CPU parity oracle (test infrastructure; never imported by the product)."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx(b3d):
    if not b3d.cuda_available():
        pytest.fail("GPU test selected but no sm_100 CUDA device is usable (no CPU fallback exists)")
    c = b3d.Context(0)
    yield c
    c.close()
