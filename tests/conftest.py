import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


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


class GpuBackend:
    """device vectors = torch CUDA tensors, engine = libnd_b200.so on the B200"""
    name = "gpu"
    scale = 1.0          # problem-size factor of the parity cases
    rk4_steps = 1000     # north_star: 1000 fixed-step RK4 steps

    def __init__(self, torch, device="cuda"):
        self.torch, self.device = torch, device

    def dev(self, a):
        if a is None or (hasattr(a, "size") and a.size == 0):
            return None
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy()).to(self.device)

    def nan(self, n):
        return self.torch.full((int(n),), float("nan"), dtype=self.torch.float64, device=self.device)

    def host(self, x):
        self.sync()
        return x.cpu().numpy().copy()

    def sync(self):
        if self.device == "cuda":
            self.torch.cuda.synchronize()

    def fill_(self, x, value):
        x.fill_(value)


class SimBackend:
    """device vectors = numpy arrays, engine = the product's CUDA sources executed by the CPU SIMT emulator
    (tests/cusim; test infrastructure, never loaded by the package itself)"""
    name = "sim"
    scale = 0.1
    rk4_steps = 100

    def dev(self, a):
        import cusim
        if a is None or (hasattr(a, "size") and a.size == 0):
            return None
        return cusim.dev(a)

    def nan(self, n):
        import cusim
        return cusim.empty(n)

    def host(self, x):
        return x.numpy().copy()

    def sync(self):
        pass

    def fill_(self, x, value):
        x.a[:] = value


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "sim"])
def backend(request):
    """run a parity test on the B200 (`-m gpu`) and, at reduced size, on the CPU emulation of the same kernel sources"""
    if request.param == "gpu":
        import torch
        assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
        import ndb200
        ndb200._cabi.lib()
        yield GpuBackend(torch)
    else:
        import cusim
        with cusim.use():
            yield SimBackend()


@pytest.fixture
def gpu_backend(cuda):
    return GpuBackend(cuda)
