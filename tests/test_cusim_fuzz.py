"""Seeded fuzzing of the engine on the CPU emulation of its kernels (tests/cusim): random small graphs (directed and
undirected, isolated vertices, hubs), random mixes of vertex / edge batch types (all wrappers, the two-sided static edge,
edges with states), random kernel family and launch / layout switches (row split width, degree-bucket window, unroll,
block shape, long-row threshold) -- `du`, `get_buffers`, a few RK4 steps and the host-buffer call against the sequential
oracle.  The B200 runs the same engine code on the fixed parity cases (`-m gpu`); this widens the CPU suite's coverage of
corner cases (empty rows and batches, rows cut into lane parts, single-row tiles, hub rows reduced by a block)."""
import numpy as np
import pytest

from helpers import condition_params, floored_rel_err, oracle_network


def _random_graph(nd, rng):
    n = int(rng.integers(1, 260))
    kind = rng.integers(0, 5)
    directed = bool(rng.integers(0, 2))
    if n < 3 or kind == 0:                       # sparse random pairs, possibly none
        m = int(rng.integers(0, 3 * n + 1))
        s, d = rng.integers(1, n + 1, m), rng.integers(1, n + 1, m)
    elif kind == 1:                              # a hub (or two) + random edges: rows far above the split width
        hubs = rng.integers(1, n + 1, int(rng.integers(1, 3)))
        s = np.concatenate([np.repeat(hubs, n), rng.integers(1, n + 1, n)])
        d = np.concatenate([np.tile(np.arange(1, n + 1), hubs.size), rng.integers(1, n + 1, n)])
    elif kind == 2:                              # dense: every row long
        m = int(rng.integers(n, min(40 * n, n * n) + 1))
        s, d = rng.integers(1, n + 1, m), rng.integers(1, n + 1, m)
    elif kind == 3:                              # ring + chords
        s = np.concatenate([np.arange(1, n + 1), rng.integers(1, n + 1, n // 2)])
        d = np.concatenate([np.roll(np.arange(1, n + 1), -1), rng.integers(1, n + 1, n // 2)])
    else:                                        # many isolated vertices
        m = int(rng.integers(0, max(2, n // 4)))
        s, d = rng.integers(1, n // 2 + 2, m), rng.integers(1, n // 2 + 2, m)
        s, d = np.minimum(s, n), np.minimum(d, n)
    return (nd.SimpleDiGraph if directed else nd.SimpleGraph)(n, s, d)


def _random_models(nd, rng, g):
    L = nd.Lib
    vpool = [L.kuramoto_first(), L.kuramoto_second(), L.kuramoto_second_bench(), L.diffusion_vertex()]
    epool = [L.diffusion_edge(), L.diffusion_edge_nop(), L.kuramoto_edge(), L.diffusion_edge_fid(), L.diffusion_odeedge(), L.relax_odeedge(),
             nd.EdgeModel(g=nd.Symmetric(L.diffusionedge), outdim=1, pdim=1, name="sym_diff"),
             nd.EdgeModel(g=nd.Symmetric(L.kuramoto_edge_f), outdim=1, pdim=1, name="sym_kura"),
             nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura"),
             nd.EdgeModel(g=nd.Directed(L.diffusionedge_nop), outdim=1, pdim=0, name="dir_diff")]
    nvt, net = int(rng.integers(1, 4)), int(rng.integers(1, 5))
    vsel = [vpool[i] for i in rng.choice(len(vpool), nvt, replace=False)]
    esel = [epool[i] for i in rng.choice(len(epool), net, replace=False)]
    vm = vsel[0] if nvt == 1 and rng.integers(0, 2) else (vsel, rng.integers(0, nvt, g.nv))
    em = esel[0] if net == 1 and rng.integers(0, 2) else (esel, rng.integers(0, net, g.ne))
    return vm, em


def _random_switches(rng):
    env = {"ND_B200_KERNEL": ["fused", "jag", "jag", "split"][int(rng.integers(0, 4))]}
    if env["ND_B200_KERNEL"] == "jag":
        env["ND_B200_JAG_SPLIT"] = str([1, 3, 7, 32, 63][int(rng.integers(0, 5))])
        env["ND_B200_JAG_WINDOW"] = str([32, 64, 128][int(rng.integers(0, 3))])
        env["ND_B200_JAG_U"] = str([2, 4][int(rng.integers(0, 2))])
        env["ND_B200_JAG_WPS"] = str([32, 48, 64][int(rng.integers(0, 3))])
    else:
        env["ND_B200_BLOCK"] = str([128, 256][int(rng.integers(0, 2))])
        env["ND_B200_EPT"] = str([4, 8][int(rng.integers(0, 2))])
    thr = [0, 0, 2, 5, 40][int(rng.integers(0, 5))]
    return env, thr


@pytest.mark.parametrize("seed", range(48))
def test_fuzz_registry_networks_on_the_emulator(nd, monkeypatch, seed):
    import cusim
    rng = np.random.default_rng(1000 + seed)
    g = _random_graph(nd, rng)
    vm, em = _random_models(nd, rng, g)
    env, thr = _random_switches(rng)
    if g.directed and rng.integers(0, 3) == 0:
        # LoopbackConnections: a few injector leaves (new vertices, any registry vertex kind) behind loopback edges to
        # random hubs; the new edges sort behind the old ones (their src ids are the largest)
        L = nd.Lib
        k = int(rng.integers(1, 6))
        hubs = rng.integers(1, g.nv + 1, k)
        vlist, vtypes = (list(vm[0]), np.asarray(vm[1])) if isinstance(vm, tuple) else ([vm], np.zeros(g.nv, dtype=np.int64))
        elist, etypes = (list(em[0]), np.asarray(em[1])) if isinstance(em, tuple) else ([em], np.zeros(g.ne, dtype=np.int64))
        vm = (vlist, np.concatenate([vtypes, rng.integers(0, len(vlist), k)]))
        em = (elist + [L.loopback()], np.concatenate([etypes, np.full(k, len(elist))]))
        g = nd.SimpleDiGraph(g.nv + k, np.concatenate([g.src, g.nv + 1 + np.arange(k)]), np.concatenate([g.dst, hubs]))
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    with cusim.use():
        try:
            nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", long_row_threshold=thr))
        except nd.ArgumentError as ex:           # the one legitimate refusal: a short row that no tile can hold
            assert "long rows disabled" in str(ex) or "exceeds" in str(ex), (seed, str(ex))
            return
        onw = oracle_network(g, vm, em)
        assert (nw.dim(), nw.pdim()) == (onw.lastidx_dynamic, onw.lastidx_p)
        u = rng.standard_normal(nw.dim())
        p = 0.5 + condition_params(nw, rng.random(nw.pdim()))
        pd = cusim.dev(p) if p.size else None
        ref, o_ref, agg_ref = onw.rhs(u, p, return_bufs=True)
        du = cusim.empty(nw.dim())
        nw(du, cusim.dev(u), pd, 0.0)
        assert not np.isnan(du.numpy()).any(), (seed, env)
        assert floored_rel_err(du.numpy(), ref) <= 1e-12, (seed, env, thr)
        o, agg = cusim.empty(nw.im.lastidx_out), cusim.empty(nw.im.lastidx_aggr)
        nw.get_buffers(o, agg, cusim.dev(u), pd, 0.0)
        assert floored_rel_err(agg.numpy(), agg_ref) <= 1e-12, (seed, env, thr)
        assert floored_rel_err(o.numpy(), o_ref) <= 1e-12 and np.array_equal(np.isnan(o.numpy()), np.isnan(o_ref)), (seed, env)
        # the aggregator's own aggregate!(a, aggbuf, o): the sequential sweep over the oracle's AggregationMap, added to
        # what aggbuf holds -- one thread per slot, same order, bit-identical
        o_r, a_r = rng.standard_normal(nw.im.lastidx_out), rng.standard_normal(nw.im.lastidx_aggr)
        amap, first = onw.table("aggmap"), onw.aggmap_first
        want = a_r.copy()
        for k in np.nonzero(amap > 0)[0]:
            want[amap[k] - 1] = want[amap[k] - 1] + o_r[first - 1 + k]
        a_d = cusim.dev(a_r)
        nw.layer.aggregator.aggregate(a_d, cusim.dev(o_r))
        assert np.array_equal(a_d.numpy(), want), (seed, env)
        ud = cusim.dev(u)
        nw.rk4(ud, pd, 0.0, 1e-3, 3)
        assert floored_rel_err(ud.numpy(), onw.rk4(u, p, 0.0, 1e-3, 3)) <= 1e-11, (seed, env, thr)
        hdu = np.full(nw.dim(), np.nan)
        nw(hdu, u, p if p.size else None, 0.0)
        assert np.array_equal(hdu, du.numpy()), (seed, env)
        try:
            nw.pack_params(pd)
        except nd.ArgumentError:
            pass
        else:
            du2 = cusim.empty(nw.dim())
            nw(du2, cusim.dev(u), pd, 0.0)
            assert np.array_equal(du2.numpy(), du.numpy()), (seed, env)
            nw.pack_params(None)
        # RK4 step counts around the 8-step graph unroll (whole graphs + remainder steps)
        nsteps = int(rng.integers(1, 20))
        ud = cusim.dev(u)
        nw.rk4(ud, pd, 0.0, 1e-3, nsteps)
        assert floored_rel_err(ud.numpy(), onw.rk4(u, p, 0.0, 1e-3, nsteps)) <= 1e-11, (seed, env, nsteps)
        # row-partitioned engines without a halo layout (the all-gather exchange path): every rank evaluates its own row
        # range from a complete state vector; the ranges tile du
        has_states = any(getattr(m, "dim", 0) > 0 for m in (em[0] if isinstance(em, tuple) else [em]))
        has_loopback = any(m.name == "loopback" for m in (em[0] if isinstance(em, tuple) else [em]))
        if g.nv >= 2:      # (edges with states: every range also evaluates its chunk of each stateful batch)
            cuts = sorted(set([0, g.nv] + [int(c) for c in rng.integers(0, g.nv + 1, int(rng.integers(1, 4)))]))
            out = np.full(nw.dim(), np.nan)
            for a, b in zip(cuts[:-1], cuts[1:]):
                part = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", long_row_threshold=thr, row_range=(a, b), keep_tables=False))
                dp = cusim.empty(nw.dim())
                part(dp, cusim.dev(u), pd, 0.0)
                w = ~np.isnan(dp.numpy())
                assert not np.any(w & ~np.isnan(out)), "two row ranges wrote the same state"
                out[w] = dp.numpy()[w]
            assert np.array_equal(out, du.numpy()), (seed, env, cuts)
        # homogeneous registry networks: the engine built straight from the edge list is the same engine
        if not isinstance(vm, tuple) and not isinstance(em, tuple) and not has_states and em.coupling != 3 and g.ne > 0:
            el = nd.Network.from_edgelist(g, vm, em)
            d3 = cusim.empty(nw.dim())
            el(d3, cusim.dev(u), pd, 0.0)
            assert floored_rel_err(d3.numpy(), ref) <= 1e-12, (seed, env)


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_dq_networks_on_the_emulator(nd, monkeypatch, seed):
    """vdepth = edepth = 2 family (computed vertex outputs, vertex_out pre-pass, ping-pong outputs in RK4)"""
    import cusim
    rng = np.random.default_rng(5000 + seed)
    g = _random_graph(nd, rng)
    if g.directed:
        g = nd.SimpleGraph(g.nv, g.src, g.dst)
    env, thr = _random_switches(rng)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    L = nd.Lib
    with cusim.use():
        try:
            nw = nd.Network(g, L.swing_dq(), L.line_dq(), aggregator=nd.B200Aggregator("+", long_row_threshold=thr))
        except nd.ArgumentError as ex:
            assert "long rows disabled" in str(ex) or "exceeds" in str(ex), (seed, str(ex))
            return
        onw = oracle_network(g, L.swing_dq(), L.line_dq())
        u = rng.standard_normal(nw.dim())
        p = condition_params(nw, rng.random(nw.pdim()))
        ref, o_ref, agg_ref = onw.rhs(u, p, return_bufs=True)
        du = cusim.empty(nw.dim())
        nw(du, cusim.dev(u), cusim.dev(p), 0.0)
        assert floored_rel_err(du.numpy(), ref) <= 1e-12, (seed, env, thr)
        o, agg = cusim.empty(nw.im.lastidx_out), cusim.empty(nw.im.lastidx_aggr)
        nw.get_buffers(o, agg, cusim.dev(u), cusim.dev(p), 0.0)
        assert floored_rel_err(agg.numpy(), agg_ref) <= 1e-12 and floored_rel_err(o.numpy(), o_ref) <= 1e-12, (seed, env)
        ud = cusim.dev(u)
        nw.rk4(ud, cusim.dev(p), 0.0, 1e-3, 3)
        assert floored_rel_err(ud.numpy(), onw.rk4(u, p, 0.0, 1e-3, 3)) <= 1e-11, (seed, env, thr)


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_emulated_ranks(nd, monkeypatch, seed):
    """random graphs / model mixes / layout switches on 2..5 emulated ranks (threads): row partition, packed halo plan,
    interior-first order, publish / wait protocol; states after a few exchanging Euler steps against the oracle"""
    from test_cusim_multirank import _run_world
    rng = np.random.default_rng(9000 + seed)
    g = _random_graph(nd, rng)
    L = nd.Lib
    vpool = [L.kuramoto_first(), L.kuramoto_second(), L.kuramoto_second_bench(), L.diffusion_vertex()]
    epool = [L.diffusion_edge(), L.diffusion_edge_nop(), L.kuramoto_edge(), L.diffusion_edge_fid(),
             nd.EdgeModel(g=nd.Symmetric(L.kuramoto_edge_f), outdim=1, pdim=1, name="sym_kura"),
             nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura")]
    nvt, net = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    vm = ([vpool[i] for i in rng.choice(len(vpool), nvt, replace=False)], rng.integers(0, nvt, g.nv))
    em = ([epool[i] for i in rng.choice(len(epool), net, replace=False)], rng.integers(0, net, g.ne))
    env, _thr = _random_switches(rng)
    if env["ND_B200_KERNEL"] == "split":
        env = {"ND_B200_KERNEL": "fused"}           # halo engines: default tile shape only
    elif env["ND_B200_KERNEL"] == "fused":
        env = {"ND_B200_KERNEL": "fused"}
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    world = int(rng.integers(2, 6))
    out, ref, _plans, _sizes, _k = _run_world(nd, g, vm, em, world, ncalls=4, h=0.01)
    assert not np.isnan(out).any(), (seed, env, world)
    assert floored_rel_err(out, ref) <= 1e-12, (seed, env, world)
    out, ref, _plans, _sizes, _k = _run_world(nd, g, vm, em, world, ncalls=3, rk4=1e-2)     # nd_b200_rk4_exchange
    assert not np.isnan(out).any() and floored_rel_err(out, ref) <= 1e-11, (seed, env, world)


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_host_buffer_pipeline(nd, monkeypatch, seed):
    """nd_b200_rhs_host's pipelined form (parameter vector uploaded in 2..10 pieces, row groups released as their
    parameters land, D2H of finished rows overlapped) on mid-size random networks.  The emulator's streams are deferred,
    so a row group launched before the last parameter it reads has been copied reads poisoned memory."""
    import cusim
    rng = np.random.default_rng(7000 + seed)
    L = nd.Lib
    n = int(rng.integers(6000, 20000))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        g = nd.erdos_renyi(n, int(rng.integers(2, 6)) * n, seed=seed)
    elif kind == 1:
        g = nd.barabasi_albert(n, int(rng.integers(2, 5)), seed=seed)
    elif kind == 2:
        g = nd.watts_strogatz(n, 6, 0.3, seed=seed, directed=True)
    else:
        g = nd.grid_graph(100, n // 100)
    vpool = [L.kuramoto_first(), L.kuramoto_second(), L.kuramoto_second_bench(), L.diffusion_vertex()]
    nvt = int(rng.integers(1, 4))
    vm = ([vpool[i] for i in rng.choice(len(vpool), nvt, replace=False)], rng.integers(0, nvt, g.nv))
    if g.directed:
        em = nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura")
    else:
        em = [L.diffusion_edge(), L.kuramoto_edge(), ([L.diffusion_edge(), L.kuramoto_edge()], rng.integers(0, 2, g.ne))][int(rng.integers(0, 3))]
    monkeypatch.setenv("ND_B200_KERNEL", ["fused", "jag"][int(rng.integers(0, 2))])
    if rng.integers(0, 2):
        monkeypatch.setenv("ND_B200_JAG_WINDOW", "128")
    monkeypatch.setenv("ND_B200_HOST_CHUNKS", str(int(rng.integers(2, 11))))
    with cusim.use():
        nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", keep_tables=False))
        onw = oracle_network(g, vm, em)
        u = rng.standard_normal(nw.dim())
        p = 0.5 + condition_params(nw, rng.random(nw.pdim()))
        ref = onw.rhs(u, p)
        for _ in range(2):
            hdu = np.full(nw.dim(), np.nan)
            nw(hdu, u, p, 0.0)
            assert floored_rel_err(hdu, ref) <= 1e-12, seed


def _expr(rng, terms):
    """a random expression over the given operand strings, valid (and evaluated in the same order) in C and Python"""
    k = int(rng.integers(1, 4))
    parts = []
    for _ in range(k):
        a, b = terms[int(rng.integers(0, len(terms)))], terms[int(rng.integers(0, len(terms)))]
        c = round(float(rng.uniform(-1.5, 1.5)), 3)
        parts.append([f"{c}*{a}", f"{a}*{b}", f"sin({a} - {b})", f"({a} + {c})*{b}"][int(rng.integers(0, 4))])
    return " + ".join(parts)


def _pyfun(args, outs):
    import math
    src = f"lambda {args}: [" + ", ".join(outs) + "]"
    return eval(src, {"sin": math.sin})


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_user_supplied_kinds(nd, monkeypatch, seed):
    """run-time compiled kinds with random dimensions (vdepth != edepth, up to 5 states, 0..3 parameters), random
    bodies generated as the same expression text for CUDA and for the Python twin, computed or StateMask vertex outputs,
    all wrappers, two-sided static edges, edges with states and per-side masks -- on the emulator ("NVRTC" = g++)"""
    import cusim
    from oracle import oracle as O
    from oracle import oracle_np as ONP
    rng = np.random.default_rng(11000 + seed)
    C = nd.CudaFunction
    vdepth, edepth = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    g = _random_graph(nd, rng)
    if g.nv < 3 or g.ne == 0:       # the bodies read esum[0..edepth-1]: the network needs at least one edge
        g = nd.complete_graph(4)

    def vertex(k):
        dim, pdim = int(rng.integers(vdepth, 6)), int(rng.integers(0, 4))
        # external inputs (src/external_inputs.jl): the first state / the first output of random vertices
        extin = tuple(nd.VIndex(int(rng.integers(1, g.nv + 1)), [1, ("out", 1)][int(rng.integers(0, 2))])
                      for _ in range(int(rng.integers(0, 3)) if rng.integers(0, 2) else 0))
        ops = [f"v[{i}]" for i in range(dim)] + [f"esum[{i}]" for i in range(edepth)] + [f"p[{i}]" for i in range(pdim)] + ["t"] + \
              [f"ext[{i}]" for i in range(len(extin))]
        fo = [_expr(rng, ops) for _ in range(dim)]
        f = C(f"vf{k}", "vertex_f", " ".join(f"dv[{i}] = {e};" for i, e in enumerate(fo)),
              py=_pyfun("v, esum, ext, p, t" if extin else "v, esum, p, t", fo))
        if rng.integers(0, 2):
            gops = [f"v[{i}]" for i in range(dim)] + [f"p[{i}]" for i in range(pdim)]
            go = [_expr(rng, gops) for _ in range(vdepth)]
            gfun = C(f"vg{k}", "vertex_g", " ".join(f"out[{i}] = {e};" for i, e in enumerate(go)), py=_pyfun("v, p, t", go))
        else:
            gfun = nd.StateMask(tuple(range(1, vdepth + 1)))
        return nd.VertexModel(f=f, g=gfun, dim=dim, pdim=pdim, outdim=vdepth, name=f"v{k}", extin=extin)

    def edge(k):
        pdim = int(rng.integers(0, 4))
        ops = [f"v_src[{i}]" for i in range(vdepth)] + [f"v_dst[{i}]" for i in range(vdepth)] + [f"p[{i}]" for i in range(pdim)]
        style = int(rng.integers(0, 3))
        if style == 0:      # one-sided g under a wrapper
            eo = [_expr(rng, ops) for _ in range(edepth)]
            body = C(f"eg{k}", "edge_g", " ".join(f"e_dst[{i}] = {e};" for i, e in enumerate(eo)), py=_pyfun("v_src, v_dst, p, t", eo))
            w = [nd.AntiSymmetric, nd.Symmetric, nd.Directed][int(rng.integers(0, 3))]
            return nd.EdgeModel(g=w(body), outdim=edepth, pdim=pdim, name=f"e{k}")
        if style == 1:      # two-sided static g
            es, ed = [_expr(rng, ops) for _ in range(edepth)], [_expr(rng, ops) for _ in range(edepth)]
            txt = " ".join(f"e_src[{i}] = {e};" for i, e in enumerate(es)) + " " + " ".join(f"e_dst[{i}] = {e};" for i, e in enumerate(ed))
            py_s, py_d = _pyfun("v_src, v_dst, p, t", es), _pyfun("v_src, v_dst, p, t", ed)
            body = C(f"eg2{k}", "edge_g2", txt, py=lambda vs, vd, p, t, a=py_s, b=py_d: (a(vs, vd, p, t), b(vs, vd, p, t)))
            return nd.EdgeModel(g=nd.Fiducial(body), outdim=edepth, pdim=pdim, name=f"e{k}")
        dim = int(rng.integers(edepth, 2 * edepth + 2))      # edge with states
        fops = ops + [f"e[{i}]" for i in range(dim)] + ["t"]
        fo = [_expr(rng, fops) for _ in range(dim)]
        f = C(f"ef{k}", "edge_f", " ".join(f"de[{i}] = {e};" for i, e in enumerate(fo)), py=_pyfun("e, v_src, v_dst, p, t", fo))
        md = int(rng.integers(1, dim - edepth + 2))
        ms = int(rng.integers(1, dim - edepth + 2))
        mask = lambda a: tuple(range(a, a + edepth))
        w = [nd.AntiSymmetric(mask(md)), nd.Symmetric(mask(md)), nd.Directed(mask(md)), nd.Fiducial(src=mask(ms), dst=mask(md))][int(rng.integers(0, 4))]
        return nd.EdgeModel(f=f, g=w, dim=dim, pdim=pdim, outdim=edepth, name=f"e{k}")

    vms = [vertex(k) for k in range(int(rng.integers(1, 3)))]
    ems = [edge(k) for k in range(int(rng.integers(1, 4)))]
    vt, et = rng.integers(0, len(vms), g.nv), rng.integers(0, len(ems), g.ne)
    monkeypatch.setenv("ND_B200_KERNEL", ["fused", "jag"][int(rng.integers(0, 2))])
    monkeypatch.setenv("ND_B200_JAG_SPLIT", str([3, 32][int(rng.integers(0, 2))]))

    def kind(m):
        if hasattr(m, "outdim_dst"):
            if m.dim > 0:
                return ONP.PyKind(f=m.f.py)
            inner = m.g.g
            return ONP.PyKind(g=inner.py)
        return ONP.PyKind(f=m.f.py, g=(m.g.py if isinstance(m.g, C) else None))
    vs = [O.VSpec(kind(m), m.dim, m.pdim, m.outdim, m.extdim) for m in vms]
    es = [O.ESpec(kind(m), m.coupling, m.dim, m.pdim, m.outdim_src, m.outdim_dst, *(m.state_masks() or (0, 0))) for m in ems]
    im = ONP.IndexManager(g.nv, g.src, g.dst, vs, list(vt), es, list(et))
    with cusim.use():
        nw = nd.Network(g, (vms, vt), (ems, et), aggregator=nd.B200Aggregator("+", long_row_threshold=[0, 4][int(rng.integers(0, 2))]))
        assert (nw.dim(), nw.pdim()) == (im.last["dynamic"], im.last["p"])
        u, p = rng.uniform(-1, 1, nw.dim()), rng.uniform(0.2, 1.2, nw.pdim())
        pd = cusim.dev(p) if p.size else None
        from networkdynamics_jl_b200.network import resolve_extin
        extmap = [0] * im.last["ext"]
        for i in range(1, g.nv + 1):
            for k, ref in enumerate(vms[vt[i - 1]].extin):
                extmap[im.v_ext[i].first - 1 + k] = resolve_extin(nw.im, ref)
        for t in (0.0, 0.4):
            ref, o_ref, agg_ref = ONP.rhs(im, u, p, t, extmap)
            du = cusim.empty(nw.dim())
            nw(du, cusim.dev(u), pd, t)
            assert floored_rel_err(du.numpy(), ref) <= 1e-12, (seed, t)
        o, agg = cusim.empty(nw.im.lastidx_out), cusim.empty(nw.im.lastidx_aggr)
        nw.get_buffers(o, agg, cusim.dev(u), pd, 0.4)
        assert floored_rel_err(agg.numpy(), agg_ref) <= 1e-12 and floored_rel_err(o.numpy(), o_ref) <= 1e-12, seed
        # row-partitioned engines (complete u): random row ranges write disjoint states -- vertex rows plus a chunk of every
        # stateful edge batch -- that tile du; rows below the long-row threshold sum in the same order (the thresholds agree)
        thr = nw.layer.aggregator._opts["long_row_threshold"]
        cuts = sorted(set([0, g.nv] + [int(c) for c in rng.integers(0, g.nv + 1, int(rng.integers(1, 4)))]))
        out = np.full(nw.dim(), np.nan)
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = nd.Network(g, (vms, vt), (ems, et), aggregator=nd.B200Aggregator("+", long_row_threshold=thr, row_range=(a, b), keep_tables=False))
            dp = cusim.empty(nw.dim())
            part(dp, cusim.dev(u), pd, 0.4)
            w = ~np.isnan(dp.numpy())
            assert not np.any(w & ~np.isnan(out)), (seed, cuts)
            out[w] = dp.numpy()[w]
        assert np.array_equal(out, du.numpy()), (seed, cuts)


@pytest.mark.parametrize("order,select", [("reverse", "dq_networks or registry_networks"),
                                          ("shuffle", "rhs_matches_sequential_oracle or edge_cases")])
def test_results_do_not_depend_on_thread_order(order, select):
    """the emulator's racecheck: between two barriers / warp collectives the threads of a block may run in any order on real
    hardware.  Re-run the single-engine fuzz (reverse order) and the fixed tile / jagged / split parity cases (shuffled order)
    in a fresh process -- a missing __syncthreads between a shared-memory write and another thread's read would change a
    result.  (The whole suite has been run once in both orders.)"""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CUSIM_ORDER=order)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider",
           os.path.join(root, "tests", "test_cusim_fuzz.py"), os.path.join(root, "tests", "test_gpu_parity.py"),
           "-k", f"({select}) and not thread_order"]
    r = subprocess.run(cmd, env=env, cwd=root, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
