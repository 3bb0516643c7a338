import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def nd():
    import ndb200
    return ndb200


@pytest.fixture(scope="session")
def cuda():
    """torch + a CUDA device + the built engine library; anything missing is a hard failure on a GPU box."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    import ndb200
    ndb200._cabi.lib()  # raises loudly if libnd_b200.so is missing
    return torch


@pytest.fixture(params=["split", "fused", "jag"])
def kernel_mode(request, monkeypatch):
    """run a test once per evaluation mode of the engine (edge-once split passes / fused tile kernel / jagged
    warp-slice kernel = default)"""
    monkeypatch.setenv("ND_B200_KERNEL", request.param)
    return request.param
