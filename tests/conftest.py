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
def cuda_lib():
    """Build (if needed) and load libprn_b200.so; GPU tests fail loudly if it is missing."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from planerecnet_b200 import _lib
    return _lib.lib()
