import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def tko():
    from oracle import tko as mod
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def tk():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tiktaalik_b200 as mod
    mod._lib.lib()
    return mod
