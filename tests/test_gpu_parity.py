"""Parity: the CUDA path (through the C ABI) against the CPU oracle on identical graphs, parameters and states.

Every test takes the `backend` fixture and runs twice: on the B200 (`-m gpu`: libnd_b200.so, torch CUDA tensors) and, at
reduced size, in the CPU suite on tests/cusim -- the SAME kernel sources and engine code compiled with g++ against an
emulated CUDA runtime (threads as fibres, real warp collectives / block barriers).  The emulator shares libm with the
oracle, so there the sequential-order results are bit-identical and the 1e-12 bar is met with zero error.

Bar (BASELINE.json north_star): index/CSR construction bit-exact; `du` within 1e-12 (floored relative)
per RHS evaluation in fp64; trajectories within 1e-9 after 1000 fixed-step RK4 steps.
Mirrors the reference's cross-style equivalence harness (test/testutils.jl:46-69), test/GPU_test.jl:61-69 and
test/aggregators_test.jl:59-67.
"""
import numpy as np
import pytest

from helpers import condition_params, floored_rel_err, oracle_network, rand_inputs

TOL_DU = 1e-12      # north_star: du within 1e-12 relative per RHS evaluation
TOL_TRAJ = 1e-9     # north_star: trajectories within 1e-9 after 1000 RK4 steps


def _configs(nd, scale=1.0):
    """the five BASELINE.json configs at sizes the oracle finishes in seconds"""
    L = nd.Lib
    rng = np.random.default_rng(11)
    n3 = int(20000 * scale) // 2 * 2
    half = np.array([0] * (n3 // 2) + [1] * (n3 // 2))
    return {
        "cfg1_kuramoto_ws": (nd.watts_strogatz(int(10_000 * scale), 10, 0.1, seed=1), L.kuramoto_first(), L.kuramoto_edge()),
        "cfg2_diffusion_er": (nd.erdos_renyi(int(50_000 * scale), int(200_000 * scale), seed=1), L.diffusion_vertex(), L.diffusion_edge()),
        "cfg2_diffusion_er_nop": (nd.erdos_renyi(int(50_000 * scale), int(200_000 * scale), seed=2), L.diffusion_vertex(), L.diffusion_edge_nop()),
        "cfg3_mixed_kuramoto_ba": (nd.barabasi_albert(n3, 4, seed=1), ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(half)), L.kuramoto_edge()),
        "cfg3b_bench_inertia_ba": (nd.barabasi_albert(n3, 4, seed=3), ([L.kuramoto_second_bench(), L.kuramoto_first()], rng.permutation(half)), L.kuramoto_edge()),
        "cfg4_powergrid_grid": (nd.grid_graph(40, 50), L.swing_dq(), L.line_dq()),
        "cfg5_kuramoto_er": (nd.erdos_renyi(int(20_000 * scale), int(160_000 * scale), seed=5), L.kuramoto_first(), L.kuramoto_edge()),
    }


def _run_gpu(B, nw, u, p, t=0.0):
    du_d = B.nan(nw.dim())
    nw(du_d, B.dev(u), B.dev(p), t)
    return B.host(du_d)


@pytest.mark.parametrize("name", ["cfg1_kuramoto_ws", "cfg2_diffusion_er", "cfg2_diffusion_er_nop", "cfg3_mixed_kuramoto_ba",
                                  "cfg3b_bench_inertia_ba", "cfg4_powergrid_grid", "cfg5_kuramoto_er"])
def test_rhs_matches_sequential_oracle(nd, backend, kernel_mode, name):
    g, vm, em = _configs(nd, backend.scale)[name]
    nw = nd.Network(g, vm, em, execution=nd.B200Execution(), aggregator=nd.B200Aggregator("+"))
    onw = oracle_network(g, vm, em)
    for seed in (1, 2):
        u, p = rand_inputs(nw.dim(), nw.pdim(), seed=seed, layout=lambda q: condition_params(nw, q))
        ref = onw.rhs(u, p)
        du = _run_gpu(backend, nw, u, p)
        assert not np.isnan(du).any(), "du not fully defined"
        err = floored_rel_err(du, ref)
        assert err <= TOL_DU, (name, err)


@pytest.mark.parametrize("name", ["cfg1_kuramoto_ws", "cfg3_mixed_kuramoto_ba", "cfg4_powergrid_grid"])
def test_csr_is_bit_exact(nd, backend, name):
    """The engine's destination-sorted CSR equals the one implied by the reference's AggregationMap + gather map:
    for every aggregation slot, the contributing `o` indices in ascending order (= SequentialAggregator order)."""
    g, vm, em = _configs(nd, scale=0.2)[name]
    nw = nd.Network(g, vm, em)
    onw = oracle_network(g, vm, em)
    rowptr, nbr, eid, side = nw.export_tables()
    ed = onw.edepth
    amap, first = onw.table("aggmap"), onw.aggmap_first
    e_out_src, e_out_dst = onw.table("e_out_src"), onw.table("e_out_dst")
    # expected entry list per row from the aggregation map: scalar k of o (1-based first+k) feeds slot amap[k]
    o_idx = np.arange(amap.size) + first
    lead = (amap - 1) % ed == 0                         # first scalar of every edepth-wide output block
    rows = (amap[lead] - 1) // ed
    o_first = o_idx[lead]
    order = np.lexsort((o_first, rows))
    rows, o_first = rows[order], o_first[order]
    exp_rowptr = np.zeros(g.nv + 1, dtype=np.int64)
    np.add.at(exp_rowptr, rows + 1, 1)
    exp_rowptr = np.cumsum(exp_rowptr)
    assert np.array_equal(rowptr, exp_rowptr)
    # every exported entry must sit exactly at the o position the oracle says
    got_o = np.where(side == 1, e_out_src[eid - 1], e_out_dst[eid - 1])
    assert np.array_equal(got_o, o_first)
    exp_nbr = np.where(side == 1, g.dst[eid - 1], g.src[eid - 1])
    assert np.array_equal(nbr, exp_nbr)
    # and the row of each entry is the slot of the vertex at the other end
    v_aggr = onw.table("v_aggr")
    own = np.where(side == 1, g.src[eid - 1], g.dst[eid - 1])
    assert np.array_equal((v_aggr[own - 1] - 1) // ed, np.repeat(np.arange(g.nv), np.diff(rowptr)))


def test_mixed_edge_batches_directed_graph(nd, backend, kernel_mode):
    """test/aggregators_test.jl:14-67 restated for the registry: directed WS graph, random mix of vertex and edge types
    with AntiSymmetric / Symmetric / Directed wrappers -> generic multi-batch kernel."""
    L = nd.Lib
    rng = np.random.default_rng(0)
    g = nd.watts_strogatz(int(10_000 * backend.scale), 4, 0.8, seed=1, directed=True)
    vtypes = [L.kuramoto_first(), L.kuramoto_second(), L.diffusion_vertex()]
    etypes = [L.kuramoto_edge(), L.diffusion_edge(), L.diffusion_edge_nop(),
              nd.EdgeModel(g=nd.Symmetric(L.diffusionedge), outdim=1, pdim=1, name="sym_diff"),
              nd.EdgeModel(g=nd.Symmetric(L.kuramoto_edge_f), outdim=1, pdim=1, name="sym_kura"),
              nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura"),
              nd.EdgeModel(g=nd.Directed(L.diffusionedge_nop), outdim=1, pdim=0, name="dir_diff")]
    vm = (vtypes, rng.integers(0, len(vtypes), g.nv))
    em = (etypes, rng.integers(0, len(etypes), g.ne))
    nw = nd.Network(g, vm, em)
    onw = oracle_network(g, vm, em)
    u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
    err = floored_rel_err(_run_gpu(backend, nw, u, p), onw.rhs(u, p))
    assert err <= TOL_DU, err


def test_reference_gpu_test_network(nd, backend, kernel_mode):
    """test/GPU_test.jl:12-69 restated for the registry's STATIC models: complete_graph(4), two vertex types, several edge
    types (the verbatim network with its stateful and two-sided edges is tests/test_zzz_new_features.py)."""
    L = nd.Lib
    g = nd.complete_graph(4)
    vm = [L.kuramoto_second(), L.diffusion_vertex(), L.kuramoto_second(), L.diffusion_vertex()]
    em = [L.diffusion_edge(), L.kuramoto_edge(), L.kuramoto_edge(), L.diffusion_edge(), L.kuramoto_edge(), L.diffusion_edge()]
    nw = nd.Network(g, vm, em)
    onw = oracle_network(g, vm, em)
    u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
    du = _run_gpu(backend, nw, u, p)
    assert floored_rel_err(du, onw.rhs(u, p)) <= TOL_DU


def test_edge_cases(nd, backend, kernel_mode):
    L = nd.Lib
    # no edges at all: du = f_v(u, 0, p)
    g = nd.SimpleGraph(5, [], [])
    nw = nd.Network(g, L.kuramoto_first(), L.kuramoto_edge())
    u, p = rand_inputs(nw.dim(), nw.pdim())
    assert np.array_equal(_run_gpu(backend, nw, u, p), p)
    # isolated vertices keep aggregation 0.0 (Appendix A.10)
    g = nd.SimpleGraph(6, [1, 2], [2, 3])
    nw = nd.Network(g, L.diffusion_vertex(), L.diffusion_edge_nop())
    onw = oracle_network(g, L.diffusion_vertex(), L.diffusion_edge_nop())
    u, _ = rand_inputs(nw.dim(), 0)
    du = _run_gpu(backend, nw, u, None)
    assert np.array_equal(du, onw.rhs(u, None)) and np.all(du[3:] == 0.0)
    # a star: one hub row far above the long-row threshold, leaves of degree 1
    n = 5000
    g = nd.SimpleGraph(n, np.ones(n - 1, dtype=np.int64), np.arange(2, n + 1))
    for thr in (0, 2**31 - 1):  # block-tree reduction of the hub vs. strictly sequential accumulation
        if thr and n - 1 > 2048:
            with pytest.raises(nd.ArgumentError):
                nd.Network(g, L.kuramoto_first(), L.kuramoto_edge(), aggregator=nd.B200Aggregator("+", long_row_threshold=thr))
            continue
        nw = nd.Network(g, L.kuramoto_first(), L.kuramoto_edge(), aggregator=nd.B200Aggregator("+", long_row_threshold=thr))
        onw = oracle_network(g, L.kuramoto_first(), L.kuramoto_edge())
        u, p = rand_inputs(nw.dim(), nw.pdim())
        assert floored_rel_err(_run_gpu(backend, nw, u, p), onw.rhs(u, p)) <= TOL_DU
        assert nw.engine_sizes()["n_long_rows"] == 1


def test_diffusion_bit_exact_and_laplacian(nd, backend, kernel_mode):
    """diffusion has no transcendental: with the reference's accumulation order the result is BIT-identical to the
    sequential oracle (rows below the long-row threshold), and equals -L*x (test/diffusion_test.jl:80-90)."""
    g = nd.erdos_renyi(2000, 8000, seed=9)
    for em in (nd.Lib.diffusion_edge(), nd.Lib.diffusion_edge_nop()):
        nw = nd.Network(g, nd.Lib.diffusion_vertex(), em)
        onw = oracle_network(g, nd.Lib.diffusion_vertex(), em)
        assert nw.engine_sizes()["n_long_rows"] == 0
        for seed in range(5):
            x = np.random.default_rng(seed).standard_normal(g.nv)
            p = np.random.default_rng(seed + 50).random(nw.pdim())
            du = _run_gpu(backend, nw, x, p)
            assert np.array_equal(du, onw.rhs(x, p))
    Lm = g.laplacian()
    x = np.random.default_rng(0).standard_normal(g.nv)
    du = _run_gpu(backend, nw, x, None)
    assert np.allclose(du, -Lm @ x, rtol=1e-12, atol=1e-12)


def test_get_buffers_matches_oracle(nd, backend, kernel_mode):
    """get_buffers / RET=:buf_init (src/coreloop.jl:103-109): o and aggbuf in the reference layout"""
    B = backend
    for name in ("cfg3_mixed_kuramoto_ba", "cfg4_powergrid_grid"):
        g, vm, em = _configs(nd, scale=0.1)[name]
        nw = nd.Network(g, vm, em)
        onw = oracle_network(g, vm, em)
        u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
        _, o_ref, agg_ref = onw.rhs(u, p, return_bufs=True)
        o, agg = B.nan(nw.im.lastidx_out), B.nan(nw.im.lastidx_aggr)
        nw.get_buffers(o, agg, B.dev(u), B.dev(p), 0.0)
        assert floored_rel_err(B.host(o), o_ref) <= TOL_DU
        assert floored_rel_err(B.host(agg), agg_ref) <= TOL_DU


def test_host_buffer_path_and_errors(nd, backend):
    B = backend
    g, vm, em = _configs(nd, scale=0.1)["cfg2_diffusion_er"]
    nw = nd.Network(g, vm, em)
    onw = oracle_network(g, vm, em)
    u, p = rand_inputs(nw.dim(), nw.pdim())
    hu, hp, hdu = nd.pinned_empty(u.size), nd.pinned_empty(p.size), nd.pinned_empty(u.size)
    hu[:], hp[:] = u, p
    nw(hdu, hu, hp, 0.0)                       # host path: H2D + RHS + D2H + sync
    assert np.array_equal(hdu, onw.rhs(u, p))
    du2 = np.empty_like(u)
    nw(du2, u, p, 0.0)                         # pageable host memory works too
    assert np.array_equal(du2, hdu)
    # size / type checks raise ArgumentError like src/coreloop.jl:2-7 and test/GPU_test.jl:63,70
    with pytest.raises(nd.ArgumentError):
        nw(np.empty(u.size + 1), u, p, 0.0)
    with pytest.raises(nd.ArgumentError):
        nw(du2, u, p[:-1].copy(), 0.0)
    with pytest.raises(nd.ArgumentError):
        nw(du2, u.astype(np.float32), p, 0.0)
    with pytest.raises(nd.ArgumentError):
        nw(B.nan(u.size), u, p, 0.0)   # mixed host/device
    with pytest.raises(nd.ArgumentError):
        nw(du2, u, p, 0.0, RET="buf_init")
    with pytest.raises(nd.ArgumentError):
        nw(du2, u, None, 0.0)
    # p is re-read on every call (callbacks mutate it, docs/examples/cascading_failure.jl:107-110)
    p_d, u_d, du_d = B.dev(p), B.dev(u), B.nan(u.size)
    nw(du_d, u_d, p_d, 0.0)
    assert np.count_nonzero(B.host(du_d)) > 0
    B.fill_(p_d, 0.0)                         # same pointer, new content
    nw(du_d, u_d, p_d, 0.0)
    assert np.count_nonzero(B.host(du_d)) == 0


@pytest.mark.parametrize("chunks", ["8", "3", "1"])
def test_host_buffer_pipeline(nd, backend, kernel_mode, monkeypatch, chunks):
    """nd_b200_rhs_host on graphs large enough for the pipelined form (H2D of p in pieces, row groups start when their
    parameters have landed, D2H of finished rows overlaps): identical to the device-resident call, bit for bit."""
    monkeypatch.setenv("ND_B200_HOST_CHUNKS", chunks)
    L = nd.Lib
    n = int(200_000 * min(1.0, 2 * backend.scale))
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    cases = [(nd.erdos_renyi(n, 4 * n, seed=3), L.diffusion_vertex(), L.diffusion_edge()),
             (nd.barabasi_albert(n, 4, seed=3), ([L.kuramoto_first(), L.kuramoto_second()], np.random.default_rng(1).permutation(half)), L.kuramoto_edge()),
             (nd.watts_strogatz(n, 6, 0.2, seed=3, directed=True), L.kuramoto_first(), nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura"))]
    for g, vm, em in cases:
        nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", keep_tables=False))
        u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
        ref = _run_gpu(backend, nw, u, p)
        hu, hp, hdu = nd.pinned_empty(u.size), nd.pinned_empty(p.size), nd.pinned_empty(u.size)
        hu[:], hp[:] = u, p
        for _ in range(3):
            hdu[:] = np.nan
            nw(hdu, hu, hp, 0.0)
            assert np.array_equal(hdu, ref)


@pytest.mark.parametrize("name", ["cfg4_powergrid_grid", "cfg1_kuramoto_ws", "cfg3_mixed_kuramoto_ba"])
def test_rk4_trajectory(nd, backend, kernel_mode, name):
    """north_star: trajectories agree within 1e-9 after 1000 fixed-step RK4 steps (dt = 1e-3); the emulated run takes
    backend.rk4_steps (100) steps of a smaller network."""
    B = backend
    nsteps = B.rk4_steps
    g, vm, em = _configs(nd, scale=0.1 if B.name == "gpu" else 0.02)[name]
    if name == "cfg1_kuramoto_ws":
        g = nd.watts_strogatz(2000 if B.name == "gpu" else 500, 10, 0.1, seed=1)
    if name == "cfg4_powergrid_grid" and B.name != "gpu":
        g = nd.grid_graph(20, 25)
    nw = nd.Network(g, vm, em)
    onw = oracle_network(g, vm, em)
    u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
    ref = onw.rk4(u, p, 0.0, 1e-3, nsteps, threads=4)
    u_d, p_d = B.dev(u), B.dev(p)
    nw.rk4(u_d, p_d, 0.0, 1e-3, nsteps)
    err = floored_rel_err(B.host(u_d), ref)
    assert err <= TOL_TRAJ, (name, err)
    # the fused-stage RK4 equals four plain RHS calls + host-side stage algebra for one step
    u1 = B.dev(u)
    nw.rk4(u1, p_d, 0.0, 1e-3, 1)
    dt = 1e-3

    def rhs(x):
        k = B.nan(x.size)
        nw(k, B.dev(x), p_d, 0.0)
        return B.host(k)
    k1 = rhs(u)
    k2 = rhs(u + (0.5 * dt) * k1)
    k3 = rhs(u + (0.5 * dt) * k2)
    k4 = rhs(u + dt * k3)
    un = u + (dt / 6.0) * (((k1 + 2.0 * k2) + 2.0 * k3) + k4)
    assert floored_rel_err(B.host(u1), un) <= 1e-14


def test_launch_shapes_agree(nd, backend, monkeypatch):
    """every compiled launch shape of the fused kernel gives the same answer"""
    g, vm, em = _configs(nd, scale=0.2)["cfg3_mixed_kuramoto_ba"]
    onw = oracle_network(g, vm, em)
    outs = []
    for kernel, block, ept in (("split", 256, 8), ("split", 256, 4), ("split", 128, 8), ("split", 128, 4), ("fused", 256, 8),
                               ("fused", 256, 4), ("fused", 128, 8), ("fused", 128, 4)):
        monkeypatch.setenv("ND_B200_KERNEL", kernel)
        monkeypatch.setenv("ND_B200_BLOCK", str(block))
        monkeypatch.setenv("ND_B200_EPT", str(ept))
        nw = nd.Network(g, vm, em)
        u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
        outs.append(_run_gpu(backend, nw, u, p))
        assert floored_rel_err(outs[-1], onw.rhs(u, p)) <= TOL_DU
    # jagged kernel: columns per iteration x occupancy cap x row-split width (7 forces many split rows on BA hubs)
    monkeypatch.setenv("ND_B200_KERNEL", "jag")
    for unroll, wps, split in ((2, 32, 32), (2, 48, 32), (2, 64, 32), (4, 32, 32), (4, 48, 32), (4, 64, 32), (2, 48, 7), (4, 64, 63)):
        monkeypatch.setenv("ND_B200_JAG_U", str(unroll))
        monkeypatch.setenv("ND_B200_JAG_WPS", str(wps))
        monkeypatch.setenv("ND_B200_JAG_SPLIT", str(split))
        nw = nd.Network(g, vm, em)
        u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
        assert floored_rel_err(_run_gpu(backend, nw, u, p), onw.rhs(u, p)) <= TOL_DU, (unroll, wps, split)
    # degree-bucketed slices: rows of a 64 / 128-row window dealt to the lanes by decreasing degree
    for window, split in ((64, 32), (128, 32), (128, 7)):
        monkeypatch.setenv("ND_B200_JAG_WINDOW", str(window))
        monkeypatch.setenv("ND_B200_JAG_SPLIT", str(split))
        nw = nd.Network(g, vm, em)
        u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
        assert floored_rel_err(_run_gpu(backend, nw, u, p), onw.rhs(u, p)) <= TOL_DU, (window, split)


@pytest.mark.gpu
def test_full_size_properties_cfg2(nd, gpu_backend):
    """BASELINE config 2 at full size (N=1e6, E=4e6) through size-independent properties: linearity of the diffusion
    RHS, conservation (sum of du == 0 for antisymmetric coupling, up to rounding), and parity on a sampled subgraph
    of rows recomputed on the host in the reference's order."""
    torch = gpu_backend
    g = nd.erdos_renyi(1_000_000, 4_000_000, seed=1)
    nw = nd.Network(g, nd.Lib.diffusion_vertex(), nd.Lib.diffusion_edge(), aggregator=nd.B200Aggregator("+", keep_tables=False))
    rng = np.random.default_rng(1)
    x, y, p = rng.random(g.nv), rng.random(g.nv), rng.random(g.ne)
    fx, fy, fxy = _run_gpu(torch, nw, x, p), _run_gpu(torch, nw, y, p), _run_gpu(torch, nw, x + y, p)
    assert np.max(np.abs(fxy - (fx + fy))) <= 1e-13 * np.max(np.abs(fxy))
    assert abs(np.sum(fx)) <= 1e-9 * np.sum(np.abs(fx))
    # host recomputation of 2000 sampled rows in ascending-neighbour order (SequentialAggregator order)
    rows = rng.choice(g.nv, 2000, replace=False) + 1
    s, d = g.src, g.dst
    for v in rows[:200]:
        as_dst = np.nonzero(d == v)[0]
        as_src = np.nonzero(s == v)[0]
        terms = [(e, p[e] * (x[s[e] - 1] - x[v - 1])) for e in as_dst] + [(e, -(p[e] * (x[v - 1] - x[d[e] - 1]))) for e in as_src]
        # o order: edge id ascending; for one undirected batch that is ascending neighbour id
        acc = 0.0
        for _, val in sorted(terms, key=lambda t: t[0]):
            acc = acc + val
        assert fx[v - 1] == acc
