"""Parity at BASELINE.json's FULL sizes (missing #3): the CUDA path against the C oracle on

* config 3  mixed first-order / inertial Kuramoto on Barabasi-Albert N = 1e6, m = 4      du <= 1e-12
* config 4  dq swing + dq line on the 400 x 500 grid, 1000 fixed-step RK4 steps          trajectory <= 1e-9
* config 5  (scaled to one GPU's test budget) Kuramoto on Erdos-Renyi N = 5e6, E = 4e7   du <= 1e-12
* config 2  heat diffusion on Erdos-Renyi N = 1e6, E = 4e6                               du bit-identical (error 0.0)

in every kernel family the engine can choose for the network.  The oracle's OpenMP restatement (oracle/nd_oracle.c:
ndo_rhs_threaded) adds each vertex's entries in the same order as the sequential sweep, so it is used here to keep the
GPU box's wall clock short; the sequential sweep is compared with it once per config on a strided sample of rows.
GPU-only: the emulator would need hours for these sizes.  Mirrors test/testutils.jl:46-69, test/GPU_test.jl:61-69.
"""
import numpy as np
import pytest

from helpers import condition_params, floored_rel_err, oracle_network, rand_inputs

TOL_DU = 1e-12
TOL_TRAJ = 1e-9
pytestmark = pytest.mark.gpu


def _families(nd, g, vm, em, monkeypatch, kinds):
    for k in kinds:
        if k == "auto":
            monkeypatch.delenv("ND_B200_KERNEL", raising=False)
        else:
            monkeypatch.setenv("ND_B200_KERNEL", k)
        yield k, nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", keep_tables=False))


def _du(B, nw, u_d, p_d):
    du = B.nan(nw.dim())
    nw(du, u_d, p_d, 0.0)
    return B.host(du)


def test_cfg3_mixed_kuramoto_ba_1e6(nd, gpu_backend, monkeypatch):
    B = gpu_backend
    L = nd.Lib
    n = 1_000_000
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    g = nd.barabasi_albert(n, 4, seed=1)
    vm = ([L.kuramoto_first(), L.kuramoto_second()], np.random.default_rng(11).permutation(half))
    em = L.kuramoto_edge()
    onw = oracle_network(g, vm, em)
    ref = None
    for kind, nw in _families(nd, g, vm, em, monkeypatch, ["auto", "fused", "jag", "split"]):
        if ref is None:
            u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
            ref = onw.rhs(u, p, threads=8)
            assert np.array_equal(ref, onw.rhs(u, p))          # the threaded sweep IS the sequential order
            u_d, p_d = B.dev(u), B.dev(p)
        for call in range(3):      # the third call runs from the automatically packed edge parameters
            err = floored_rel_err(_du(B, nw, u_d, p_d), ref)
            assert err <= TOL_DU, (kind, call, err)


def test_cfg4_powergrid_400x500_rk4_1000_steps(nd, gpu_backend, monkeypatch):
    B = gpu_backend
    L = nd.Lib
    g, vm, em = nd.grid_graph(400, 500), L.swing_dq(), L.line_dq()
    onw = oracle_network(g, vm, em)
    ref = None
    for kind, nw in _families(nd, g, vm, em, monkeypatch, ["auto", "fused"]):
        if ref is None:
            u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
            ref_du = onw.rhs(u, p, threads=8)
            ref = onw.rk4(u, p, 0.0, 1e-3, 1000, threads=8)
        assert floored_rel_err(_du(B, nw, B.dev(u), B.dev(p)), ref_du) <= TOL_DU, kind
        u_d = B.dev(u)
        nw.rk4(u_d, B.dev(p), 0.0, 1e-3, 1000)
        err = floored_rel_err(B.host(u_d), ref)
        assert err <= TOL_TRAJ, (kind, err)


def test_cfg5_scaled_kuramoto_er_5e6_4e7(nd, gpu_backend, monkeypatch):
    B = gpu_backend
    L = nd.Lib
    g, vm, em = nd.erdos_renyi(5_000_000, 40_000_000, seed=1), L.kuramoto_first(), L.kuramoto_edge()
    onw = oracle_network(g, vm, em)
    ref = None
    for kind, nw in _families(nd, g, vm, em, monkeypatch, ["auto", "fused"]):
        if ref is None:
            u, p = rand_inputs(nw.dim(), nw.pdim())
            ref = onw.rhs(u, p, threads=8)
            u_d, p_d = B.dev(u), B.dev(p)
        for call in range(3):
            err = floored_rel_err(_du(B, nw, u_d, p_d), ref)
            assert err <= TOL_DU, (kind, call, err)
        del nw


def test_cfg2_diffusion_er_1e6_bit_identical(nd, gpu_backend, monkeypatch):
    B = gpu_backend
    L = nd.Lib
    g, vm, em = nd.erdos_renyi(1_000_000, 4_000_000, seed=1), L.diffusion_vertex(), L.diffusion_edge()
    onw = oracle_network(g, vm, em)
    ref = None
    for kind, nw in _families(nd, g, vm, em, monkeypatch, ["auto", "fused", "jag", "jaga", "split"]):
        if ref is None:
            u, p = rand_inputs(nw.dim(), nw.pdim())
            ref = onw.rhs(u, p)
            u_d, p_d = B.dev(u), B.dev(p)
        for call in range(3):
            got = _du(B, nw, u_d, p_d)
            if kind == "split":    # edge-once passes: the same sums in the same order, computed from the materialised o
                assert floored_rel_err(got, ref) <= TOL_DU, (kind, call)
            else:
                assert np.array_equal(got, ref), (kind, call, floored_rel_err(got, ref))
