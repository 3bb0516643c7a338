"""`aggregate!(aggregator, aggbuf, o)` as a stand-alone call on the B200Aggregator (nd_b200_aggregate).

Mirrors test/aggregators_test.jl:14-80: a directed Watts-Strogatz graph with a random mix of vertex and edge types (all
four output wrappers), every aggregator must give the SequentialAggregator's sums (:59-67) and must ADD to what aggbuf
already holds (:69-79).  The expected sums are the reference's sweep written out in numpy over the oracle's
AggregationMap (src/aggregators.jl:140-151); one thread per slot adds in the same order, so equality is bit-exact.
Runs on the B200 (`-m gpu`) and on tests/cusim.
"""
import numpy as np
import pytest

from helpers import condition_params, floored_rel_err, oracle_network, rand_inputs
from test_gpu_parity import TOL_DU


def _sweep(onw, aggbuf, o):
    """for (dat, dst) in zip(view(o, range), map): aggbuf[dst] += dat   (dst = 0: no slot)"""
    amap, first = onw.table("aggmap"), onw.aggmap_first
    out = aggbuf.copy()
    blk = o[first - 1:first - 1 + amap.size]
    for k in np.nonzero(amap > 0)[0]:
        out[amap[k] - 1] = out[amap[k] - 1] + blk[k]
    return out


def _networks(nd, scale):
    L = nd.Lib
    rng = np.random.default_rng(0)
    g = nd.watts_strogatz(int(4000 * scale) + 40, 4, 0.8, seed=1, directed=True)
    vtypes = [L.kuramoto_first(), L.kuramoto_second(), L.diffusion_vertex()]
    etypes = [L.kuramoto_edge(), L.diffusion_edge(), L.diffusion_edge_fid(), L.relax_odeedge(),
              nd.EdgeModel(g=nd.Symmetric(L.diffusionedge), outdim=1, pdim=1, name="sym_diff"),
              nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura")]
    yield "mixed-directed", g, (vtypes, rng.integers(0, len(vtypes), g.nv)), (etypes, rng.integers(0, len(etypes), g.ne))
    yield "dq-depth2", nd.grid_graph(12, 9), L.swing_dq(), L.line_dq()
    n = 600
    yield "star", nd.SimpleGraph(n, np.ones(n - 1, dtype=np.int64), np.arange(2, n + 1)), L.kuramoto_first(), L.kuramoto_edge()
    yield "no-edges", nd.SimpleGraph(5, [], []), L.kuramoto_first(), L.kuramoto_edge()
    # loopback edges have no src-side slot in `o`; the engine's hub -> injector entry must not be summed
    m = 30
    gs = np.concatenate([np.arange(1, m + 1), np.arange(m + 1, 2 * m + 1)])
    gd = np.concatenate([np.roll(np.arange(1, m + 1), -1), np.arange(1, m + 1)])
    g2 = nd.SimpleDiGraph(2 * m, gs, gd)
    yield "loopback", g2, [L.kuramoto_first()] * m + [L.kuramoto_second()] * m, ([L.kuramoto_edge(), L.loopback()], (g2.src > m).astype(np.int64))


def test_standalone_aggregate_is_sequential_and_additive(nd, backend, kernel_mode):
    B = backend
    for name, g, vm, em in _networks(nd, B.scale):
        nw = nd.Network(g, vm, em)
        onw = oracle_network(g, vm, em)
        agg = nw.layer.aggregator
        rng = np.random.default_rng(3)
        o = rng.standard_normal(nw.im.lastidx_out)
        a0 = rng.standard_normal(nw.im.lastidx_aggr)
        a_d, o_d = B.dev(a0), B.dev(o)
        agg.aggregate(a_d, o_d)
        once = _sweep(onw, a0, o)
        assert np.array_equal(B.host(a_d), once), name
        agg.aggregate(a_d, o_d)                                   # adds to what is there
        assert np.array_equal(B.host(a_d), _sweep(onw, once, o)), name
        # the materialised outputs of get_buffers, aggregated on their own, give get_buffers' aggregation buffer
        # (the loopback copy into the injector slots is apply_loopback!, not aggregate!: compare the other slots)
        u, p = rand_inputs(nw.dim(), nw.pdim(), layout=lambda q: condition_params(nw, q))
        ob, ab = B.nan(nw.im.lastidx_out), B.nan(nw.im.lastidx_aggr)
        nw.get_buffers(ob, ab, B.dev(u), B.dev(p), 0.0)
        z = B.dev(np.zeros(nw.im.lastidx_aggr))
        agg.aggregate(z, ob)
        got, want = B.host(z), B.host(ab)
        if name == "loopback":
            ed = onw.edepth
            keep = np.ones(want.size, dtype=bool)
            v_aggr = onw.table("v_aggr")
            for v in g.src[g.src > g.nv // 2]:
                keep[v_aggr[v - 1] - 1:v_aggr[v - 1] - 1 + ed] = False
            got, want = got[keep], want[keep]
        assert floored_rel_err(got, want) <= TOL_DU, name
    with pytest.raises(nd.ArgumentError):
        agg.aggregate(B.dev(np.zeros(3)), o_d)
    with pytest.raises(nd.ArgumentError):
        agg.aggregate(np.zeros(nw.im.lastidx_aggr), np.zeros(nw.im.lastidx_out))   # host vectors


def test_aggregate_refused_on_partitioned_engines(nd, backend):
    L = nd.Lib
    g = nd.erdos_renyi(300, 900, seed=2)
    nw = nd.Network(g, L.diffusion_vertex(), L.diffusion_edge(), aggregator=nd.B200Aggregator("+", row_range=(0, 150)))
    with pytest.raises(nd.ArgumentError):
        nw.layer.aggregator.aggregate(backend.dev(np.zeros(nw.im.lastidx_aggr)), backend.dev(np.zeros(nw.im.lastidx_out)))


def test_aggregate_on_an_edgelist_engine(nd, backend):
    """nd_b200_create_from_edgelist builds the same tables: same sums as the table-driven constructor, bit for bit"""
    B, L = backend, nd.Lib
    g = nd.erdos_renyi(500, 2000, seed=4)
    a = nd.Network(g, L.kuramoto_first(), L.kuramoto_edge())
    b = nd.Network.from_edgelist(g, L.kuramoto_first(), L.kuramoto_edge(), keep_tables=True)
    lean = nd.Network.from_edgelist(g, L.kuramoto_first(), L.kuramoto_edge())       # default: no host tables are kept
    with pytest.raises(nd.ArgumentError, match="NO_EXPORT"):
        lean.layer.aggregator.aggregate(B.dev(np.zeros(a.im.lastidx_aggr)), B.dev(np.zeros(a.im.lastidx_out)))
    rng = np.random.default_rng(1)
    o, a0 = rng.standard_normal(a.im.lastidx_out), rng.standard_normal(a.im.lastidx_aggr)
    assert (b.im.lastidx_out, b.im.lastidx_aggr) == (a.im.lastidx_out, a.im.lastidx_aggr)
    xa, xb = B.dev(a0), B.dev(a0)
    a.layer.aggregator.aggregate(xa, B.dev(o))
    b.layer.aggregator.aggregate(xb, B.dev(o))
    assert np.array_equal(B.host(xa), B.host(xb)) and not np.array_equal(B.host(xa), a0)
