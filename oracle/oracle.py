"""ctypes binding of the CPU oracle (oracle/nd_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity pinning: see the header of nd_oracle.h (Julia cannot run here; pinned against the
reference's Julia-free known answers).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libnd_oracle.so")

# kind ids (mirror nd_oracle.h)
V_DIFFUSION, V_KURAMOTO_FIRST, V_KURAMOTO_SECOND, V_KURAMOTO_SECOND_BENCH, V_SWING_DQ = range(5)
V_OPAQUE = 100
E_DIFFUSION, E_DIFFUSION_NOP, E_KURAMOTO, E_LINE_DQ, E_DIFFUSION_ODE, E_RELAX_ODE, E_DIFFUSION_FID, E_LOOPBACK = range(8)
E_OPAQUE = 100
ANTISYMMETRIC, SYMMETRIC, DIRECTED, FIDUCIAL = range(4)


class _VSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("pdim", C.c_int32), ("outdim", C.c_int32)]


class _ESpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("coupling", C.c_int32), ("dim", C.c_int32), ("pdim", C.c_int32),
                ("outdim_src", C.c_int32), ("outdim_dst", C.c_int32), ("mask_src", C.c_int32), ("mask_dst", C.c_int32)]


@dataclass(frozen=True)
class VSpec:
    kind: int
    dim: int
    pdim: int
    outdim: int
    extdim: int = 0        # external inputs (Python twin only; the C restatement rejects them)
    ff: bool = False       # feed-forward g(v, ins, p, t) of an injector leaf (Python twin only)


@dataclass(frozen=True)
class ESpec:
    kind: int
    coupling: int
    dim: int
    pdim: int
    outdim_src: int
    outdim_dst: int
    mask_src: int = 0      # edges with states: 1-based first state of the src / dst output StateMask
    mask_dst: int = 0
    extdim: int = 0        # external inputs (Python twin only)


# the model zoo of test/ComponentLibrary.jl and benchmark/benchmark_models.jl (dims as declared there)
VSPECS = {
    "diffusion_vertex": VSpec(V_DIFFUSION, 1, 0, 1),
    "kuramoto_first": VSpec(V_KURAMOTO_FIRST, 1, 1, 1),
    "kuramoto_second": VSpec(V_KURAMOTO_SECOND, 2, 3, 1),
    "kuramoto_second_bench": VSpec(V_KURAMOTO_SECOND_BENCH, 2, 1, 1),
    "swing_dq": VSpec(V_SWING_DQ, 2, 4, 2),
}
ESPECS = {
    "diffusion_edge": ESpec(E_DIFFUSION, ANTISYMMETRIC, 0, 1, 1, 1),
    "diffusion_edge_nop": ESpec(E_DIFFUSION_NOP, ANTISYMMETRIC, 0, 0, 1, 1),
    "kuramoto_edge": ESpec(E_KURAMOTO, ANTISYMMETRIC, 0, 1, 1, 1),
    "line_dq": ESpec(E_LINE_DQ, ANTISYMMETRIC, 0, 3, 2, 2),
    # test/ComponentLibrary.jl:30-40: f=diffusion_dedge!, dim=2, g=Fiducial(dst=1:1, src=2:2)
    "diffusion_odeedge": ESpec(E_DIFFUSION_ODE, FIDUCIAL, 2, 1, 1, 1, 2, 1),
    # test/diffusion_test.jl:96-101: f=real_ode_edge!, dim=2, g=Fiducial(2,1) = Fiducial(src=2, dst=1)
    "relax_odeedge": ESpec(E_RELAX_ODE, FIDUCIAL, 2, 0, 1, 1, 2, 1),
    # test/ComponentLibrary.jl:22-28: two-sided static g
    "diffusion_edge_fid": ESpec(E_DIFFUSION_FID, FIDUCIAL, 0, 1, 1, 1),
}


def build_library(force: bool = False) -> str:
    """Compile oracle/libnd_oracle.so (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "nd_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libnd_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_library()
        L = C.CDLL(_LIB)
        i64p = C.POINTER(C.c_int64)
        i32p = C.POINTER(C.c_int32)
        dp = C.POINTER(C.c_double)
        L.ndo_build.restype = C.c_void_p
        L.ndo_build.argtypes = [C.c_int64, C.c_int64, i64p, i64p, C.c_int32, C.POINTER(_VSpec), i32p,
                                C.c_int32, C.POINTER(_ESpec), i32p]
        L.ndo_free.argtypes = [C.c_void_p]
        L.ndo_last_error.restype = C.c_char_p
        L.ndo_size.restype = C.c_int64
        L.ndo_size.argtypes = [C.c_void_p, C.c_int]
        L.ndo_table.restype = i64p
        L.ndo_table.argtypes = [C.c_void_p, C.c_int]
        L.ndo_batch_len.restype = C.c_int64
        L.ndo_batch_len.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ndo_batch_indices.restype = i64p
        L.ndo_batch_indices.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ndo_batch_spec.restype = C.c_int32
        L.ndo_batch_spec.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ndo_rhs_sequential.argtypes = [C.c_void_p, dp, dp, dp, C.c_double]
        L.ndo_rhs_sequential_bufs.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, dp, dp]
        L.ndo_rhs_threaded.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_int]
        L.ndo_rk4.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_int64, C.c_int]
        L.ndo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class OracleNetwork:
    """`Network(g, vertexm, edgem; execution=SequentialExecution{true}(), aggregator=SequentialAggregator(+))`
    restated (src/construction.jl:31-236).  `esrc`/`edst` are 1-based in `edges(g)` order; `vtype`/`etype`
    map every component to an entry of `vspecs`/`especs` (object identity of the model)."""

    def __init__(self, nv, esrc, edst, vspecs, vtype, especs, etype):
        L = lib()
        self.nv = int(nv)
        self.esrc = np.ascontiguousarray(esrc, dtype=np.int64)
        self.edst = np.ascontiguousarray(edst, dtype=np.int64)
        self.ne = int(self.esrc.size)
        self.vspecs = list(vspecs)
        self.especs = list(especs)
        vt = np.ascontiguousarray(vtype, dtype=np.int32)
        et = np.ascontiguousarray(etype, dtype=np.int32)
        assert vt.size == self.nv and et.size == self.ne
        if any(getattr(s, "extdim", 0) or getattr(s, "ff", False) for s in self.vspecs + self.especs):
            raise ValueError("external inputs / feed-forward vertices are not restated in the C oracle (use the Python twin)")
        vs = (_VSpec * max(1, len(self.vspecs)))(*[_VSpec(s.kind, s.dim, s.pdim, s.outdim) for s in self.vspecs])
        es = (_ESpec * max(1, len(self.especs)))(*[_ESpec(s.kind, s.coupling, s.dim, s.pdim, s.outdim_src, s.outdim_dst,
                                                          getattr(s, "mask_src", 0), getattr(s, "mask_dst", 0))
                                                   for s in self.especs])
        i64p = C.POINTER(C.c_int64)
        i32p = C.POINTER(C.c_int32)
        self._h = L.ndo_build(self.nv, self.ne, self.esrc.ctypes.data_as(i64p), self.edst.ctypes.data_as(i64p),
                              len(self.vspecs), vs, vt.ctypes.data_as(i32p), len(self.especs), es,
                              et.ctypes.data_as(i32p))
        if not self._h:
            raise ValueError(L.ndo_last_error().decode())
        sz = lambda w: int(L.ndo_size(self._h, w))
        self.lastidx_dynamic, self.lastidx_p, self.lastidx_out = sz(0), sz(1), sz(2)
        self.lastidx_aggr, self.lastidx_gbuf = sz(3), sz(4)
        self.n_vbatches, self.n_ebatches, self.vdepth, self.edepth = sz(5), sz(6), sz(7), sz(8)
        self.aggmap_first, self.aggmap_len = sz(9), sz(10)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().ndo_free(h)
            self._h = None

    # -- tables (copies, 1-based int64) -------------------------------------------------------
    def _tab(self, which, n):
        ptr = lib().ndo_table(self._h, which)
        return np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].copy()

    def table(self, name):
        names = ["v_data", "v_out", "v_para", "v_aggr", "e_data", "e_out_src", "e_out_dst", "e_para",
                 "e_gbuf_src", "e_gbuf_dst", "gbufmap", "aggmap"]
        w = names.index(name)
        n = self.nv if w < 4 else self.ne if w < 10 else self.lastidx_gbuf if w == 10 else self.aggmap_len
        return self._tab(w, n)

    def batches(self, kind):
        k = 0 if kind == "vertex" else 1
        L = lib()
        out = []
        for b in range(self.n_vbatches if k == 0 else self.n_ebatches):
            n = int(L.ndo_batch_len(self._h, k, b))
            idx = np.ctypeslib.as_array(L.ndo_batch_indices(self._h, k, b), shape=(n,)).copy()
            out.append((int(L.ndo_batch_spec(self._h, k, b)), idx))
        return out

    # -- evaluation -----------------------------------------------------------------------------
    def _chk(self, u, p):
        u = np.ascontiguousarray(u, dtype=np.float64)
        if u.size != self.lastidx_dynamic:
            raise ValueError(f"du or u does not have expected size {self.lastidx_dynamic}")
        if self.lastidx_p > 0:
            p = np.ascontiguousarray(p, dtype=np.float64)
            if p.size != self.lastidx_p:
                raise ValueError(f"p does not has expecte size {self.lastidx_p}")
        else:
            p = None
        return u, p

    def rhs(self, u, p=None, t=0.0, threads=1, return_bufs=False):
        u, p = self._chk(u, p)
        du = np.empty_like(u)
        L = lib()
        if return_bufs:
            o = np.empty(self.lastidx_out)
            agg = np.empty(self.lastidx_aggr)
            rc = L.ndo_rhs_sequential_bufs(self._h, _dp(du), _dp(u), _dp(p), float(t), _dp(o), _dp(agg))
        elif threads == 1:
            rc = L.ndo_rhs_sequential(self._h, _dp(du), _dp(u), _dp(p), float(t))
        else:
            rc = L.ndo_rhs_threaded(self._h, _dp(du), _dp(u), _dp(p), float(t), int(threads))
        if rc:
            raise ValueError(L.ndo_last_error().decode())
        return (du, o, agg) if return_bufs else du

    def rhs_into(self, du, u, p, t=0.0, threads=1):
        """no-allocation variant for timing"""
        L = lib()
        if threads == 1:
            return L.ndo_rhs_sequential(self._h, _dp(du), _dp(u), _dp(p), float(t))
        return L.ndo_rhs_threaded(self._h, _dp(du), _dp(u), _dp(p), float(t), int(threads))

    def rk4(self, u, p, t0, dt, nsteps, threads=1):
        u, p = self._chk(u, p)
        u = u.copy()
        rc = lib().ndo_rk4(self._h, _dp(u), _dp(p), float(t0), float(dt), int(nsteps), int(threads))
        if rc:
            raise ValueError(lib().ndo_last_error().decode())
        return u


def max_threads() -> int:
    return int(lib().ndo_max_threads())
