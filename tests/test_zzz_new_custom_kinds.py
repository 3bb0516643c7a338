"""User-supplied kinds, the features written after the round's last GPU run (edges with states, external inputs, loopback
connections with feed-forward injectors): green on the CPU emulation (tests/cusim; "NVRTC" = g++ of the same generated
source), real NVRTC compilation for sm_100a checked in test_custom_kinds.py, first B200 run pending -- kept in a file that
sorts last (see test_zzz_new_features.py)."""
import math

import numpy as np
import pytest

from helpers import floored_rel_err
from oracle import oracle as O
from oracle import oracle_np as ONP
from test_custom_kinds import _cases_stateful, _check_cases_against_twin, _check_rk4_and_host_buffers, _models


@pytest.mark.parametrize("mode", ["fused", "jag"])
def test_stateful_custom_kinds_match_python_twin(nd, backend, monkeypatch, mode):
    _check_cases_against_twin(nd, backend, monkeypatch, mode, _cases_stateful(nd))


def test_stateful_custom_kinds_rk4_and_host_buffers(nd, backend):
    _check_rk4_and_host_buffers(nd, backend, stateful=True)


def test_stateful_custom_kinds_on_row_partitioned_engines(nd, backend):
    """user-supplied edges with states under a row partition (all-gather exchange): three row ranges write disjoint states
    that tile du and equal the unpartitioned engine's result bit for bit"""
    B = backend
    for name, g, vms, vt, ems, et in _cases_stateful(nd):
        vm, em = (vms, vt), (ems, et)
        nw = nd.Network(g, vm, em)
        rng = np.random.default_rng(2)
        u, p = rng.random(nw.dim()), 0.5 + rng.random(nw.pdim())
        full = B.nan(nw.dim())
        nw(full, B.dev(u), B.dev(p), 0.3)
        full, out = B.host(full), np.full(nw.dim(), np.nan)
        cuts = [0, g.nv // 4, g.nv // 2 + 1, g.nv]
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", row_range=(a, b), keep_tables=False))
            dp = B.nan(nw.dim())
            part(dp, B.dev(u), B.dev(p), 0.3)
            dp = B.host(dp)
            w = ~np.isnan(dp)
            assert not np.any(w & ~np.isnan(out)), name
            out[w] = dp[w]
        assert np.array_equal(out, full), name


def _tiles(nd, B, g, vm, em, u, p, t, full, cuts):
    """row-partitioned engines over `cuts` write disjoint states that tile du and equal the unpartitioned result"""
    out = np.full(full.size, np.nan)
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", row_range=(a, b), keep_tables=False))
        dp = B.nan(full.size)
        part(dp, B.dev(u), B.dev(p), t)
        dp = B.host(dp)
        w = ~np.isnan(dp)
        assert not np.any(w & ~np.isnan(out))
        out[w] = dp[w]
    assert np.array_equal(out, full)


def test_external_inputs(nd, backend, monkeypatch):
    """External inputs (src/external_inputs.jl, src/coreloop.jl:61): a component's f reads states / outputs of OTHER
    components -- vertices (f(dv, v, esum, ext, p, t)) and edges with states (f(de, e, vs, vd, ext, p, t)).  Sources: a
    state (by symbol or index), a StateMask vertex output, a computed vertex output, the output of an edge with states (both
    sides, AntiSymmetric sign).  Components that differ only in WHAT they refer to share a batch.  Against the Python twin,
    `du`, RK4, both kernel families; outputs of static (feed-forward) edges are refused like in the reference (:42-44)."""
    B = backend
    C = nd.CudaFunction
    M = _models(nd)
    rng = np.random.default_rng(8)

    def ctrl(refs):       # a controller vertex: follows the average of two remote quantities, coupled through its first state
        return nd.VertexModel(f=C("ctrl_f", "vertex_f", "dv[0] = p[0]*(0.5*(ext[0] + ext[1]) - v[0]) + esum[0]; dv[1] = ext[1]*t - v[1];",
                                  py=lambda v, e, x, p, t: [p[0] * (0.5 * (x[0] + x[1]) - v[0]) + e[0], x[1] * t - v[1]]),
                              g=nd.StateMask((1,)), dim=2, pdim=1, sym=("x", "y"), extin=tuple(refs), name="ctrl")

    def obs_edge(refs):   # an edge with states that integrates a remote state
        return nd.EdgeModel(f=C("obs_f", "edge_f", "de[0] = ext[0] - e[0] + p[0]*(v_src[0] - v_dst[0]);",
                                py=lambda e, vs, vd, x, p, t: [x[0] - e[0] + p[0] * (vs[0] - vd[0])]),
                            g=nd.AntiSymmetric(1), dim=1, pdim=1, outdim=1, sym=("q",), extin=tuple(refs), name="obs")
    for mode in ("fused", "jag"):
        monkeypatch.setenv("ND_B200_KERNEL", mode)
        g = nd.barabasi_albert(60, 2, seed=3)
        fhn = nd.VertexModel(f=M["fhn"].f, g=M["fhn"].g, dim=2, pdim=3, sym=("a", "b"), name="fhn")
        vms = [fhn] * g.nv
        # three controllers with different references: a state by symbol, a state by index, a vertex output, an edge output
        vms[4] = ctrl([nd.VIndex(2, "b"), nd.VIndex(7, 1)])
        vms[11] = ctrl([nd.VIndex(5, ("out", 1)), nd.EIndex(3, ("out", 1))])
        vms[30] = ctrl([nd.EIndex(3, ("out", 2)), nd.EIndex(3, "q")])
        ems = [M["wsin"]] * g.ne
        ems[2] = obs_edge([nd.VIndex(9, "a")])           # edge 3 (1-based): referenced above through both of its outputs
        ems[10] = obs_edge([nd.VIndex(12, 2)])
        nw = nd.Network(g, vms, ems)
        assert len(nw.vertexbatches) == 2 and len(nw.layer.edgebatches) == 2 and nw.im.lastidx_extbuf == 3 * 2 + 2
        # twin: same batching (hash = model without the references), per-component ExtMap
        vspec = {id(fhn): O.VSpec(ONP.PyKind(f=fhn.f.py), 2, 3, 1)}
        vs, vt = [], []
        for m in vms:
            key = "ctrl" if m.extdim else "fhn"
            if key not in [k for k, _ in vs]:
                vs.append((key, O.VSpec(ONP.PyKind(f=m.f.py), m.dim, m.pdim, m.outdim, m.extdim)))
            vt.append([k for k, _ in vs].index(key))
        es, et = [], []
        for m in ems:
            key = "obs" if m.dim else "wsin"
            if key not in [k for k, _ in es]:
                kind = ONP.PyKind(f=m.f.py) if m.dim else ONP.PyKind(g=m.g.g.py)
                es.append((key, O.ESpec(kind, m.coupling, m.dim, m.pdim, m.outdim_src, m.outdim_dst, *(m.state_masks() or (0, 0)), m.extdim)))
            et.append([k for k, _ in es].index(key))
        im = ONP.IndexManager(g.nv, g.src, g.dst, [s for _, s in vs], vt, [s for _, s in es], et)
        assert (im.last["dynamic"], im.last["p"], im.last["ext"]) == (nw.dim(), nw.pdim(), nw.im.lastidx_extbuf)
        from networkdynamics_jl_b200.network import resolve_extin
        extmap = [0] * im.last["ext"]
        for i, m in enumerate(vms, start=1):
            for k, ref in enumerate(m.extin):
                extmap[im.v_ext[i].first - 1 + k] = resolve_extin(nw.im, ref)
        for i, m in enumerate(ems, start=1):
            for k, ref in enumerate(m.extin):
                extmap[im.e_ext[i].first - 1 + k] = resolve_extin(nw.im, ref)
        u, p = rng.uniform(-1, 1, nw.dim()), 0.25 + rng.random(nw.pdim())
        for t in (0.0, 0.6):
            ref = ONP.rhs(im, u, p, t, extmap)[0]
            du = B.nan(nw.dim())
            nw(du, B.dev(u), B.dev(p), t)
            assert floored_rel_err(B.host(du), ref) <= 1e-12, (mode, t)
            _tiles(nd, B, g, vms, ems, u, p, t, B.host(du), [0, g.nv // 3, g.nv // 2, g.nv])
        dt, x, t0 = 1e-2, u.copy(), 0.1
        for s in range(3):
            t = t0 + s * dt
            k1 = ONP.rhs(im, x, p, t, extmap)[0]
            k2 = ONP.rhs(im, x + 0.5 * dt * k1, p, t + 0.5 * dt, extmap)[0]
            k3 = ONP.rhs(im, x + 0.5 * dt * k2, p, t + 0.5 * dt, extmap)[0]
            k4 = ONP.rhs(im, x + dt * k3, p, t + dt, extmap)[0]
            x = x + (dt / 6.0) * (((k1 + 2.0 * k2) + 2.0 * k3) + k4)
        ud = B.dev(u)
        nw.rk4(ud, B.dev(p), t0, dt, 3)
        assert floored_rel_err(B.host(ud), x) <= 1e-11, mode
    # computed (non-StateMask) vertex outputs as sources: read from the materialised output block of the same stage
    monkeypatch.setenv("ND_B200_KERNEL", "fused")
    g = nd.grid_graph(6, 5)
    osc = nd.VertexModel(f=M["osc"].f, g=M["osc"].g, dim=2, pdim=1, outdim=2, sym=("th", "r"), name="osc")
    watcher = nd.VertexModel(f=C("watch_f", "vertex_f", "dv[0] = ext[0]*ext[1] - v[0] + esum[0]; dv[1] = ext[2] + esum[1] - v[1];",
                                 py=lambda v, e, x, p, t: [x[0] * x[1] - v[0] + e[0], x[2] + e[1] - v[1]]),
                             g=M["osc"].g, dim=2, pdim=1, outdim=2, sym=("th", "r"),
                             extin=(nd.VIndex(3, ("out", 1)), nd.VIndex(3, ("out", 2)), nd.VIndex(8, "r")), name="watcher")
    vms = [osc] * g.nv
    vms[20] = watcher
    nw = nd.Network(g, vms, M["line2"])
    vs = [O.VSpec(ONP.PyKind(f=osc.f.py, g=osc.g.py), 2, 1, 2), O.VSpec(ONP.PyKind(f=watcher.f.py, g=watcher.g.py), 2, 1, 2, 3)]
    vt = [1 if m.extdim else 0 for m in vms]
    es = [O.ESpec(ONP.PyKind(g=M["line2"].g.g.py), M["line2"].coupling, 0, 2, 2, 2)]
    im = ONP.IndexManager(g.nv, g.src, g.dst, vs, vt, es, [0] * g.ne)
    from networkdynamics_jl_b200.network import resolve_extin
    extmap = [resolve_extin(nw.im, ref) for ref in watcher.extin]
    u, p = rng.uniform(-1, 1, nw.dim()), 0.25 + rng.random(nw.pdim())
    du = B.nan(nw.dim())
    nw(du, B.dev(u), B.dev(p), 0.3)
    assert floored_rel_err(B.host(du), ONP.rhs(im, u, p, 0.3, extmap)[0]) <= 1e-12
    _tiles(nd, B, g, vms, M["line2"], u, p, 0.3, B.host(du), [0, 7, 21, g.nv])
    # refusals: output of a static (feed-forward) edge; unknown symbol; registry kinds do not take external inputs
    g = nd.complete_graph(4)
    with pytest.raises(nd.ArgumentError, match="feed-forward"):
        nd.Network(g, [ctrl([nd.EIndex(1, ("out", 1)), nd.VIndex(2, 1)])] + [fhn] * 3, M["wsin"])
    with pytest.raises(nd.ArgumentError, match="not a state symbol"):
        nd.Network(g, [ctrl([nd.VIndex(2, "nope"), nd.VIndex(2, 1)])] + [fhn] * 3, M["wsin"])
    L = nd.Lib
    kf = L.kuramoto_first()
    bad = nd.VertexModel(f=kf.f, g=kf.g, dim=1, pdim=1, extin=(nd.VIndex(2, 1),), name="k_ext")
    with pytest.raises(nd.ArgumentError):
        nd.Network(g, [bad] + [kf] * 3, L.kuramoto_edge())


def test_loopback_connections_and_feed_forward_injectors(nd, backend, monkeypatch):
    """LoopbackConnection + injector vertices (src/post_utils.jl:105-234, src/coreloop.jl:47,55; test/loopback_test.jl part A:
    a capacitor hub with a resistor injector (pure feed forward, no states), an inductor injector (state output) and a
    voltage source behind a resistor edge).  The injector's input is its hub's output, the hub receives minus the injector's
    output; feed-forward g runs after the loopback copy.  Closed form, Python twin, RK4, get_buffers; then a registry-only
    network against the C oracle; then the reference's topology errors."""
    B = backend
    C = nd.CudaFunction
    L = nd.Lib
    hub = nd.VertexModel(f=C("cap_f", "vertex_f", "dv[0] = esum[0] / p[0];", py=lambda v, e, p, t: [e[0] / p[0]]),
                         g=nd.StateMask((1,)), dim=1, pdim=1, sym=("v",), name="hub")
    rinj = nd.VertexModel(f=C("rinj_f", "vertex_f", "", py=lambda v, e, p, t: []),
                          g=C("rinj_g", "vertex_gff", "out[0] = ins[0] / p[0];", py=lambda v, ins, p, t: [ins[0] / p[0]]),
                          dim=0, pdim=1, outdim=1, name="R_injector")
    linj = nd.VertexModel(f=C("linj_f", "vertex_f", "dv[0] = esum[0] / p[0];", py=lambda v, e, p, t: [e[0] / p[0]]),
                          g=nd.StateMask((1,)), dim=1, pdim=1, sym=("i",), name="L_injector")
    vsrc = nd.VertexModel(f=C("vs_f", "vertex_f", "", py=lambda v, e, p, t: []),
                          g=C("vs_g", "vertex_g", "out[0] = p[0];", py=lambda v, p, t: [p[0]]), dim=0, pdim=1, outdim=1, name="vs")
    res = L.diffusion_edge()                                  # i_dst = (1/R) (v_src - v_dst), i_src = -i_dst
    g = nd.SimpleDiGraph(4, [2, 3, 4], [1, 1, 1])             # R_injector -> hub, L_injector -> hub, vs -> hub
    vms, ems = [hub, rinj, linj, vsrc], [L.loopback(), L.loopback(), res]
    for mode in ("fused", "jag"):
        monkeypatch.setenv("ND_B200_KERNEL", mode)
        nw = nd.Network(g, vms, ems)
        assert nw.dim() == 2 and nw.pdim() == 5
        # layout: u = [v_hub, i_L]; p = [C, R, Lind, V, 1/R_edge]
        v, iL = 0.3, -0.2
        Cc, R, Lind, V, ginv = 2.0, 100.0, 0.1, 1.0, 0.5
        u = np.zeros(2); p = np.zeros(5)
        for b in nw.vertexbatches:
            nm = b.model.name
            if nm == "hub": u[b.state_first - 1] = v; p[b.p_first - 1] = Cc
            if nm == "L_injector": u[b.state_first - 1] = iL; p[b.p_first - 1] = Lind
            if nm == "R_injector": p[b.p_first - 1] = R
            if nm == "vs": p[b.p_first - 1] = V
        eb = [b for b in nw.layer.edgebatches if b.model.pdim][0]
        p[eb.p_first - 1] = ginv
        du = B.nan(2)
        nw(du, B.dev(u), B.dev(p), 0.0)
        du = B.host(du)
        hub_b = [b for b in nw.vertexbatches if b.model.name == "hub"][0]
        l_b = [b for b in nw.vertexbatches if b.model.name == "L_injector"][0]
        want_hub = (ginv * (V - v) + -1.0 * (v / R) + -1.0 * iL) / Cc
        assert abs(du[hub_b.state_first - 1] - want_hub) <= 1e-15 and du[l_b.state_first - 1] == v / Lind, mode
        # row-partitioned engines with the complete u (all-gather exchange): the ranges tile du, injectors included
        out = np.full(2, np.nan)
        for a, b in ((0, 1), (1, 3), (3, 4)):
            part = nd.Network(g, vms, ems, aggregator=nd.B200Aggregator("+", row_range=(a, b), keep_tables=False))
            dp = B.nan(2)
            part(dp, B.dev(u), B.dev(p), 0.0)
            dp = B.host(dp)
            w = ~np.isnan(dp)
            assert not np.any(w & ~np.isnan(out)), mode
            out[w] = dp[w]
        assert np.array_equal(out, du), mode
        # twin (same batching / layout), get_buffers and RK4
        from helpers import model_types
        um, vt = model_types(vms, g.nv)
        uem, et = model_types(ems, g.ne)

        def vspec(m):
            if m.custom_spec() is not None:
                return O.VSpec(ONP.PyKind(f=m.f.py, g=(m.g.py if isinstance(m.g, C) else None)), m.dim, m.pdim, m.outdim, 0, m.hasff)
            return O.VSpec(m.kernel_kind(), m.dim, m.pdim, m.outdim)
        im = ONP.IndexManager(g.nv, g.src, g.dst, [vspec(m) for m in um], list(vt),
                              [O.ESpec(m.kernel_kind(), m.coupling, m.dim, m.pdim, m.outdim_src, m.outdim_dst) for m in uem], list(et))
        ref_du, ref_o, ref_agg = ONP.rhs(im, u, p, 0.0)
        assert floored_rel_err(du, ref_du) <= 1e-14
        o, agg = B.nan(nw.im.lastidx_out), B.nan(nw.im.lastidx_aggr)
        nw.get_buffers(o, agg, B.dev(u), B.dev(p), 0.0)
        assert floored_rel_err(B.host(o), ref_o) <= 1e-14 and floored_rel_err(B.host(agg), ref_agg) <= 1e-14
        dt, x = 1e-3, u.copy()
        for s in range(20):
            k1 = ONP.rhs(im, x, p)[0]
            k2 = ONP.rhs(im, x + 0.5 * dt * k1, p)[0]
            k3 = ONP.rhs(im, x + 0.5 * dt * k2, p)[0]
            k4 = ONP.rhs(im, x + dt * k3, p)[0]
            x = x + (dt / 6.0) * (((k1 + 2.0 * k2) + 2.0 * k3) + k4)
        ud = B.dev(u)
        nw.rk4(ud, B.dev(p), 0.0, dt, 20)
        assert floored_rel_err(B.host(ud), x) <= 1e-12, mode
        # registry-only: a ring of Kuramoto hubs, each with an inertial injector leaf behind a loopback edge (C oracle)
        n = 40
        ring_s, ring_d = np.arange(1, n + 1), np.roll(np.arange(1, n + 1), -1)
        gs = np.concatenate([ring_s, np.arange(n + 1, 2 * n + 1)])
        gd = np.concatenate([ring_d, np.arange(1, n + 1)])
        g2 = nd.SimpleDiGraph(2 * n, gs, gd)
        vm2 = [L.kuramoto_first()] * n + [L.kuramoto_second()] * n
        is_loop = g2.src > n
        em2 = ([L.kuramoto_edge(), L.loopback()], is_loop.astype(np.int64))
        nw2 = nd.Network(g2, vm2, em2)
        from helpers import condition_params, oracle_network
        onw = oracle_network(g2, vm2, em2)
        u2 = np.random.default_rng(2).random(nw2.dim())
        p2 = 0.5 + condition_params(nw2, np.random.default_rng(3).random(nw2.pdim()))
        d2 = B.nan(nw2.dim())
        nw2(d2, B.dev(u2), B.dev(p2), 0.0)
        assert floored_rel_err(B.host(d2), onw.rhs(u2, p2)) <= 1e-12, mode
        ud = B.dev(u2)
        nw2.rk4(ud, B.dev(p2), 0.0, 1e-3, 17)
        assert floored_rel_err(B.host(ud), onw.rk4(u2, p2, 0.0, 1e-3, 17)) <= 1e-11, mode
    # the reference's topology rules (src/construction.jl:52-80)
    with pytest.raises(nd.ArgumentError, match="leaf"):      # loopback from a non-leaf
        nd.Network(nd.SimpleDiGraph(3, [1, 1, 2], [2, 3, 3]), [hub, hub, hub], [L.loopback(), res, res])
    with pytest.raises(nd.ArgumentError, match="[Ff]eed.forward"):   # feed-forward vertex without a loopback edge
        nd.Network(nd.SimpleDiGraph(2, [2], [1]), [hub, rinj], [res])
