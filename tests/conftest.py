import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu

    return cpu.get()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through the C ABI.  GPU tests must not silently pass without it."""
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    from capycrypt_b200 import Engine

    eng = Engine()
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def kat():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "sha3_kat.json")) as f:
        return json.load(f)
