import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (no-op when up to date) and load the C-ABI library; GPU tests must never fall back."""
    import torch
    assert torch.cuda.is_available(), "GPU test collected without a CUDA device"
    import fldr_vfi_b200._lib as L
    return L.lib()
