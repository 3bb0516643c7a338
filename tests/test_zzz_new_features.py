"""Parity tests of the features written after the round's last GPU run (edges with states, the verbatim test/GPU_test.jl
network, packed edge parameters, parameter-free multi-batch networks): green on the CPU emulation of the kernels
(tests/cusim), first B200 run pending -- kept in a file that sorts last so that an unforeseen GPU-only failure here cannot
mask the established parity tests under `pytest -x`."""
import numpy as np
import pytest

from helpers import condition_params, floored_rel_err, oracle_network, rand_inputs
from test_gpu_parity import TOL_DU, TOL_TRAJ, _run_gpu


def test_reference_gpu_test_network_verbatim(nd, backend, kernel_mode):
    """test/GPU_test.jl:12-69 verbatim: complete_graph(4), vertices [kuramoto_second, diffusion_vertex, ...], edges
    [diffusion_odeedge, kuramoto_edge, kuramoto_edge, diffusion_edge_fid, diffusion_odeedge, diffusion_edge_fid] -- an
    edge batch WITH states (StateMask outputs Fiducial(dst=1:1, src=2:2)), a static AntiSymmetric batch and a static
    two-sided (Fiducial) batch in one network; `du`, `get_buffers` and RK4."""
    B = backend
    L = nd.Lib
    g = nd.complete_graph(4)
    vm = [L.kuramoto_second(), L.diffusion_vertex(), L.kuramoto_second(), L.diffusion_vertex()]
    em = [L.diffusion_odeedge(), L.kuramoto_edge(), L.kuramoto_edge(), L.diffusion_edge_fid(), L.diffusion_odeedge(), L.diffusion_edge_fid()]
    nw = nd.Network(g, vm, em)
    onw = oracle_network(g, vm, em)
    assert nw.dim() == 10 and nw.pdim() == 12          # 2*2 + 2*1 vertex states + 2*2 edge states; 2*3 + 6 parameters
    u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: 0.5 + condition_params(nw, q))
    du = _run_gpu(B, nw, u, p)
    assert floored_rel_err(du, onw.rhs(u, p)) <= TOL_DU
    o, agg = B.nan(nw.im.lastidx_out), B.nan(nw.im.lastidx_aggr)
    nw.get_buffers(o, agg, B.dev(u), B.dev(p), 0.0)
    _, o_ref, agg_ref = onw.rhs(u, p, return_bufs=True)
    assert floored_rel_err(B.host(o), o_ref) <= TOL_DU and floored_rel_err(B.host(agg), agg_ref) <= TOL_DU
    ud = B.dev(u)
    nw.rk4(ud, B.dev(p), 0.0, 1e-3, 100)
    assert floored_rel_err(B.host(ud), onw.rk4(u, p, 0.0, 1e-3, 100)) <= TOL_TRAJ


def test_edges_with_states(nd, backend, kernel_mode):
    """Edges with states ("ODE edges", src/coreloop.jl:41,76): PASS 2 reads their StateMask outputs, PASS 4 evaluates
    their f.  Known answer (test/diffusion_test.jl:96-129, the relaxation edge de = (vs - vd) - e): with every edge state
    on its constraint the vertex part of du is -L*x and the edge part is exactly 0.  Then random states against the
    oracle, a mixed static / stateful network on a power-law graph, and 1000 (emulator: 100) RK4 steps."""
    B = backend
    L = nd.Lib
    g = nd.erdos_renyi(int(20_000 * B.scale), int(80_000 * B.scale), seed=3)
    nw = nd.Network(g, L.diffusion_vertex(), L.relax_odeedge())
    assert nw.dim() == g.nv + 2 * g.ne and nw.pdim() == 0
    x = np.random.default_rng(0).standard_normal(g.nv)
    e = np.stack([x[g.src - 1] - x[g.dst - 1], x[g.dst - 1] - x[g.src - 1]], axis=1).ravel()
    du = _run_gpu(B, nw, np.concatenate([x, e]), None)
    s, d = g.src - 1, g.dst - 1
    Lx = np.zeros(g.nv)
    np.add.at(Lx, s, x[s] - x[d])
    np.add.at(Lx, d, x[d] - x[s])
    assert np.allclose(du[:g.nv], -Lx, rtol=1e-12, atol=1e-12) and np.all(du[g.nv:] == 0.0)
    onw = oracle_network(g, L.diffusion_vertex(), L.relax_odeedge())
    u, _ = rand_inputs(nw.dim(), 0)
    assert np.array_equal(_run_gpu(B, nw, u, None), onw.rhs(u, None))          # no transcendental: bit-identical
    # mixed: two vertex batches, stateful + static + two-sided static edge batches, hubs
    rng = np.random.default_rng(5)
    n = int(20_000 * B.scale)
    g = nd.barabasi_albert(n, 4, seed=4)
    vm = ([L.kuramoto_first(), L.kuramoto_second(), L.diffusion_vertex()], rng.integers(0, 3, g.nv))
    em = ([L.diffusion_odeedge(), L.kuramoto_edge(), L.relax_odeedge(), L.diffusion_edge_fid(), L.diffusion_edge_nop()], rng.integers(0, 5, g.ne))
    nw = nd.Network(g, vm, em)
    onw = oracle_network(g, vm, em)
    u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: 0.5 + condition_params(nw, q))
    assert floored_rel_err(_run_gpu(B, nw, u, p), onw.rhs(u, p)) <= TOL_DU
    assert nw.engine_sizes()["launches_per_rhs"] == 3                            # row kernel + one f kernel per stateful batch
    hdu = np.empty_like(u)
    nw(hdu, u, p, 0.0)                                                           # host-buffer path (unpipelined form)
    assert floored_rel_err(hdu, onw.rhs(u, p)) <= TOL_DU
    ud = B.dev(u)
    nw.rk4(ud, B.dev(p), 0.0, 1e-3, B.rk4_steps)
    assert floored_rel_err(B.host(ud), onw.rk4(u, p, 0.0, 1e-3, B.rk4_steps, threads=4)) <= TOL_TRAJ
    # row-partitioned engines (all-gather exchange): the owner of a row range also evaluates f for the same proportion of
    # every stateful batch; two ranges write disjoint states that tile du.  The packed-halo layout refuses them.
    ref, out = onw.rhs(u, p), np.full(nw.dim(), np.nan)
    for a, b in ((0, n // 3), (n // 3, n)):
        part = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", row_range=(a, b), keep_tables=False))
        dp = B.nan(nw.dim())
        part(dp, B.dev(u), B.dev(p), 0.0)
        w = ~np.isnan(B.host(dp))
        assert not np.any(w & ~np.isnan(out)), "two row ranges wrote the same state"
        out[w] = B.host(dp)[w]
    assert floored_rel_err(out, ref) <= TOL_DU
    with pytest.raises(nd.ArgumentError):
        nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", row_range=(0, n // 2), gather_offset=np.arange(g.nv), gather_len=nw.dim()))


def test_parameter_free_multi_batch_network(nd, backend, kernel_mode):
    """several edge batches, none with parameters: the generic kernels still read their parameter-offset stream"""
    L = nd.Lib
    g = nd.watts_strogatz(int(5000 * backend.scale) + 100, 4, 0.5, seed=2, directed=True)
    em = ([L.diffusion_edge_nop(), nd.EdgeModel(g=nd.Directed(L.diffusionedge_nop), outdim=1, pdim=0, name="dir_diff"),
           nd.EdgeModel(g=nd.Symmetric(L.diffusionedge_nop), outdim=1, pdim=0, name="sym_diff")], np.random.default_rng(1).integers(0, 3, g.ne))
    nw = nd.Network(g, L.diffusion_vertex(), em)
    onw = oracle_network(g, L.diffusion_vertex(), em)
    u, _ = rand_inputs(nw.dim(), 0)
    assert np.array_equal(_run_gpu(backend, nw, u, None), onw.rhs(u, None))


@pytest.mark.parametrize("mode", ["fused", "jag"])
def test_packed_edge_parameters(nd, backend, monkeypatch, mode):
    """nd_b200_pack_params: the engine's per-entry copy of the edge parameters (coalesced reads instead of one scattered
    read per entry).  Packed evaluation is bit-identical to the default one; the contract is visible -- edge parameters
    changed behind the engine's back are NOT seen until the next pack, vertex parameters always are; nd_b200_rk4 with
    ND_B200_RK4_PACK=1 packs per call and gives the same trajectory bit for bit."""
    B = backend
    L = nd.Lib
    monkeypatch.setenv("ND_B200_KERNEL", mode)
    n = int(20_000 * B.scale)
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    cases = [(nd.erdos_renyi(n, 4 * n, seed=3), L.diffusion_vertex(), L.diffusion_edge()),
             (nd.barabasi_albert(n, 4, seed=3), ([L.kuramoto_first(), L.kuramoto_second()], np.random.default_rng(1).permutation(half)), L.kuramoto_edge()),
             (nd.grid_graph(30, 40), L.swing_dq(), L.line_dq())]
    for g, vm, em in cases:
        nw = nd.Network(g, vm, em)
        onw = oracle_network(g, vm, em)
        u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
        ref = _run_gpu(B, nw, u, p)
        p_d = B.dev(p)
        nw.pack_params(p_d)
        du = B.nan(nw.dim())
        nw(du, B.dev(u), p_d, 0.0)
        assert np.array_equal(B.host(du), ref)
        assert floored_rel_err(ref, onw.rhs(u, p)) <= TOL_DU
        # edge parameters changed without a new pack: not seen (that is the contract); vertex parameters: seen
        eb = nw.layer.edgebatches[0]
        p2 = p.copy()
        p2[eb.p_first - 1:] *= 1.5
        nw(du, B.dev(u), B.dev(p2), 0.0)
        assert np.array_equal(B.host(du), ref)
        if eb.p_first > 1:
            p3 = p.copy()
            p3[:eb.p_first - 1] *= 1.25
            nw(du, B.dev(u), B.dev(p3), 0.0)
            assert floored_rel_err(B.host(du), onw.rhs(u, p3)) <= TOL_DU
        nw.pack_params(B.dev(p2))
        nw(du, B.dev(u), B.dev(p2), 0.0)
        assert floored_rel_err(B.host(du), onw.rhs(u, p2)) <= TOL_DU
        o, agg = B.nan(nw.im.lastidx_out), B.nan(nw.im.lastidx_aggr)
        nw.get_buffers(o, agg, B.dev(u), B.dev(p2), 0.0)
        assert floored_rel_err(B.host(agg), onw.rhs(u, p2, return_bufs=True)[2]) <= TOL_DU
        nw.pack_params(None)                                   # back to the reference's semantics
        nw(du, B.dev(u), p_d, 0.0)
        assert np.array_equal(B.host(du), ref)
        # RK4: per-call packing
        ua, ub = B.dev(u), B.dev(u)
        nw.rk4(ua, p_d, 0.0, 1e-3, 12)
        monkeypatch.setenv("ND_B200_RK4_PACK", "1")
        nw.rk4(ub, p_d, 0.0, 1e-3, 12)
        monkeypatch.delenv("ND_B200_RK4_PACK")
        assert np.array_equal(B.host(ua), B.host(ub))
        nw(du, B.dev(u), B.dev(p2), 0.0)                       # ... and the engine is unpacked again afterwards
        assert floored_rel_err(B.host(du), onw.rhs(u, p2)) <= TOL_DU
    # several edge batches / no edge parameters: no packed kernels, the request is refused, nothing falls back silently
    g = nd.erdos_renyi(500, 2000, seed=1)
    em = ([L.diffusion_edge(), L.kuramoto_edge()], np.random.default_rng(2).integers(0, 2, g.ne))
    nw = nd.Network(g, L.kuramoto_first(), em)
    with pytest.raises(nd.ArgumentError):
        nw.pack_params(B.dev(np.ones(nw.pdim())))
    with pytest.raises(nd.ArgumentError):
        nd.Network(g, L.diffusion_vertex(), L.diffusion_edge_nop()).pack_params(B.dev(np.ones(1)))


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["fused", "jag"])
def test_two_gpu_fused_rk4(nd, cuda, tmp_path, kernel):
    """nd_b200_rk4_exchange (written after the round's last GPU run; green on emulated ranks, tests/test_cusim_multirank.py):
    five fused RK4 steps equal five host-driven ones on both GPUs"""
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    from test_gpu_multi import _worker
    mp.spawn(_worker, args=(2, port, str(tmp_path), "p2p", kernel, True), nprocs=2, join=True)
