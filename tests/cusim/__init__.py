"""CPU SIMT emulation of the engine's CUDA sources -- TEST INFRASTRUCTURE ONLY.

`build()` compiles csrc/nd_b200.cu (+ nd_b200_kernels.cuh) with g++ against the emulated CUDA runtime in this directory
into tests/cusim/_build/libnd_b200_sim.so, which exports the same C ABI as libnd_b200.so.  `use()` is a context manager
that points the package's ctypes binding at that library for the duration of a test, so that the CPU test suite runs the
real kernel code (tile kernel, jagged kernel, long rows, RK4 epilogues, get_buffers, run-time compiled kinds, the NVLink
publish / wait protocol with emulated ranks as threads) against the oracle.

What the emulated runtime checks beyond results (cusim_rt.cpp): warp collectives / barriers complete only with every named
lane present (divergence is an error, not a hang); non-null streams are deferred, so a missing event dependency between
streams gives a wrong result; device allocations end at a guard page (out-of-bounds accesses fault); CUSIM_ORDER=reverse|
shuffle permutes the order in which a block's threads run between scheduling points (data races change results).

The product never loads this library: `networkdynamics.jl_b200._cabi.lib()` only ever opens libnd_b200.so and fails
loudly without it.  "Device" vectors of the emulated engine are numpy arrays wrapped in `DeviceArray`.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_DIR))
LIB_PATH = os.path.join(_DIR, "_build", "libnd_b200_sim.so")


def build(force: bool = False) -> str:
    import ndb200 as nd
    cabi = nd._cabi
    srcs = cabi.SOURCES + [os.path.join(_DIR, "cusim_rt.cpp")]
    deps = srcs + cabi.HEADERS + [os.path.join(_DIR, n) for n in ("cusim_device.h", "cuda_runtime.h", "nvrtc.h")]
    newest = max(os.path.getmtime(f) for f in deps)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    cabi.write_embedded_header()
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-DND_CUSIM=1",
           f'-DCUSIM_INCLUDE_DIR="{_DIR}"', "-I", _DIR, "-I", os.path.join(_ROOT, "include"), "-x", "c++"] + srcs + \
          ["-o", LIB_PATH + f".tmp{os.getpid()}", "-ldl", "-lpthread"]
    subprocess.check_call(cmd)
    os.replace(LIB_PATH + f".tmp{os.getpid()}", LIB_PATH)     # atomic: concurrent test workers never load a half-written library
    return LIB_PATH


_sim = None


def lib():
    global _sim
    if _sim is None:
        import ndb200 as nd
        build()
        _sim = nd._cabi.bind(C.CDLL(LIB_PATH))
    return _sim


@contextlib.contextmanager
def use():
    """route the package's C-ABI calls to the emulated library inside the `with` block"""
    import ndb200 as nd
    cabi = nd._cabi
    prev = cabi._lib
    cabi._lib = lib()
    try:
        yield cabi._lib
    finally:
        cabi._lib = prev


class _Owner:
    """frees a guarded allocation of the emulated library when the last view of it goes away"""

    def __init__(self, L, ptr):
        self.L, self.ptr = L, ptr

    def __del__(self):
        try:
            self.L.nd_b200_host_free(self.ptr)
        except Exception:
            pass


class _Arr(np.ndarray):
    pass


def _guarded(n: int) -> np.ndarray:
    """float64 vector inside one of the emulator's guarded allocations (it ends at an inaccessible page), so that a kernel
    reading or writing past the end of u / du / p faults at the access"""
    L = lib()
    n = int(n)
    ptr = L.nd_b200_host_alloc(max(n, 1) * 8)
    if not ptr:
        raise MemoryError("emulated allocation failed")
    buf = (C.c_double * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float64, count=max(n, 1))[:n].view(_Arr)
    arr._owner = _Owner(L, ptr)
    return arr


class DeviceArray:
    """a float64 vector presented as device memory (`__cuda_array_interface__`) to the emulated engine"""

    def __init__(self, a):
        a = np.asarray(a, dtype=np.float64).ravel()
        self.a = _guarded(a.size)
        self.a[:] = a

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.a.shape, "typestr": "<f8", "data": (self.a.ctypes.data, False), "version": 3}

    def numpy(self):
        return self.a


def dev(a) -> DeviceArray:
    return DeviceArray(a)


def empty(n) -> DeviceArray:
    return DeviceArray(np.full(int(n), np.nan))
