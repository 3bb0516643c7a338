"""Host-side mirror of the reference's index construction (networkdynamics.jl_b200/network.py) against the oracle's
independent restatement: bit-exact equality of every table (no GPU needed: the aggregator closure is a no-op)."""
import numpy as np
import pytest

from helpers import null_aggregator, oracle_network


def _cases(nd):
    L = nd.Lib
    rng = np.random.default_rng(5)
    n = 300
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    yield "cfg1-like", nd.watts_strogatz(200, 10, 0.1, seed=1), L.kuramoto_first(), L.kuramoto_edge()
    yield "cfg2-like", nd.erdos_renyi(400, 1600, seed=1), L.diffusion_vertex(), L.diffusion_edge()
    yield "cfg3-like", nd.barabasi_albert(n, 4, seed=1), ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(half)), L.kuramoto_edge()
    yield "cfg4-like", nd.grid_graph(8, 7), L.swing_dq(), L.line_dq()
    g = nd.watts_strogatz(60, 4, 0.5, seed=3, directed=True)
    em = [L.kuramoto_edge(), L.diffusion_edge(), nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir")]
    yield "directed-mixed", g, [L.kuramoto_first(), L.kuramoto_second_bench()] * 30, [em[k] for k in rng.integers(0, 3, g.ne)]


def test_tables_match_oracle(nd):
    for name, g, vm, em in _cases(nd):
        nw = nd.Network(g, vm, em, aggregator=null_aggregator)
        onw = oracle_network(g, vm, em)
        im = nw.im
        for t in ["v_data", "v_out", "v_para", "v_aggr", "e_data", "e_out_src", "e_out_dst", "e_para", "e_gbuf_src",
                  "e_gbuf_dst"]:
            assert np.array_equal(getattr(im, t), onw.table(t)), (name, t)
        for a in ["lastidx_dynamic", "lastidx_p", "lastidx_out", "lastidx_aggr", "lastidx_gbuf", "vdepth", "edepth"]:
            assert getattr(im, a) == getattr(onw, a), (name, a)
        ob = onw.batches("vertex") + onw.batches("edge")
        pb = nw.vertexbatches + nw.layer.edgebatches
        assert len(ob) == len(pb)
        for (_, oi), b in zip(ob, pb):
            assert np.array_equal(oi, b.indices), name
        assert nd.dim(nw) == onw.lastidx_dynamic and nd.pdim(nw) == onw.lastidx_p


def test_graph_canonical_edge_order(nd):
    """edges(g) order: src<dst, sorted by (src,dst); docs/src/mathematical_model.md:108-115"""
    g = nd.SimpleGraph(5, [5, 3, 2, 1, 4, 2], [1, 1, 3, 2, 2, 1])
    assert list(zip(g.src, g.dst)) == [(1, 2), (1, 3), (1, 5), (2, 3), (2, 4)]
    d = nd.SimpleDiGraph(3, [3, 1, 2, 1], [1, 3, 1, 2])
    assert list(zip(d.src, d.dst)) == [(1, 2), (1, 3), (2, 1), (3, 1)]
    for g in (nd.erdos_renyi(1000, 4000, seed=2), nd.barabasi_albert(1000, 4, seed=2), nd.watts_strogatz(1000, 10, 0.1)):
        key = g.src * (g.nv + 1) + g.dst
        assert np.all(np.diff(key) > 0) and np.all(g.src < g.dst)
    assert nd.erdos_renyi(1000, 4000, seed=2).ne == 4000
    assert nd.grid_graph(400, 500).ne == 399100


def test_constructor_argument_errors(nd):
    L = nd.Lib
    g = nd.complete_graph(4)
    with pytest.raises(nd.ArgumentError):        # src/construction.jl:48-51
        nd.Network(g, [L.kuramoto_first()] * 3, L.kuramoto_edge(), aggregator=null_aggregator)
    with pytest.raises(nd.ArgumentError):        # src/construction.jl:99-102 (different vertex outdim)
        nd.Network(g, [L.kuramoto_first(), L.swing_dq()] * 2, L.kuramoto_edge(), aggregator=null_aggregator)
    with pytest.raises(nd.ArgumentError):        # src/construction.jl:96
        nd.Network(g, L.kuramoto_first(), L.kuramoto_edge(), execution="sequential", aggregator=null_aggregator)
    with pytest.raises(nd.ArgumentError):        # only + is supported, no fallback
        nd.B200Aggregator(max)


def test_unregistered_component_is_rejected_before_any_device_work(nd):
    """north_star: unsupported component types raise an error instead of silently running elsewhere"""
    L = nd.Lib
    g = nd.complete_graph(3)
    custom = nd.VertexModel(f=lambda dv, v, acc, p, t: None, g=nd.StateMask(1), dim=1, pdim=0, name="custom")
    with pytest.raises(nd.ArgumentError, match="no kernel in the B200 registry"):
        nd.Network(g, custom, L.kuramoto_edge())
    ode_edge = nd.EdgeModel(g=nd.AntiSymmetric(L.kuramoto_edge_f), f=L.kuramoto_vertex, dim=1, outdim=1, pdim=1, name="ode")
    with pytest.raises(nd.ArgumentError, match="no kernel in the B200 registry"):
        nd.Network(g, L.kuramoto_first(), ode_edge)
    masked = nd.VertexModel(f=L.kuramoto_inertia, g=nd.StateMask(2), dim=2, pdim=3, name="mask2")
    with pytest.raises(nd.ArgumentError):
        nd.Network(g, masked, L.kuramoto_edge())


def test_create_from_edgelist_builds_the_same_engine(nd, monkeypatch):
    """nd_b200_create_from_edgelist (homogeneous network from a bare edge list) against the table-driven
    constructor: same sizes, same CSR (rowptr, neighbour, edge id, side), same jagged layout -- on host-only engines."""
    L = nd.Lib
    cases = [(nd.erdos_renyi(3000, 12000, seed=2), L.kuramoto_first(), L.kuramoto_edge()),
             (nd.barabasi_albert(2000, 4, seed=2), L.kuramoto_second(), L.kuramoto_edge()),
             (nd.grid_graph(31, 17), L.swing_dq(), L.line_dq()),
             (nd.watts_strogatz(1500, 6, 0.2, seed=1, directed=True), L.diffusion_vertex(),
              nd.EdgeModel(g=nd.Directed(L.diffusionedge_nop), outdim=1, pdim=0, name="dir_diff")),
             (nd.SimpleGraph(7, [], []), L.kuramoto_first(), L.kuramoto_edge())]
    for mode in ("fused", "jag"):
        monkeypatch.setenv("ND_B200_KERNEL", mode)
        for g, vm, em in cases:
            for rr in (None, (g.nv // 3, g.nv - 2)):
                a = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", host_only=True, row_range=rr))
                b = nd.Network.from_edgelist(g, vm, em, host_only=True, keep_tables=True, row_range=rr)
                assert (a.dim(), a.pdim(), a.im.lastidx_out, a.im.lastidx_aggr) == (b.dim(), b.pdim(), b.im.lastidx_out, b.im.lastidx_aggr)
                assert a.engine_sizes() == b.engine_sizes()
                for x, y in zip(a.export_tables(), b.export_tables()):
                    assert np.array_equal(x, y)
                if mode == "jag":
                    ja, jb = a.export_jag(), b.export_jag()
                    for k in ("slices", "lanes", "longs", "order"):
                        assert np.array_equal(ja[k], jb[k]), k
    with pytest.raises(nd.ArgumentError):      # user-supplied kinds go through the table-driven constructor
        nd.Network.from_edgelist(cases[0][0], nd.VertexModel(f=nd.CudaFunction("f", "vertex_f", "dv[0]=0;"), g=nd.StateMask((1,)), dim=1),
                                 L.kuramoto_edge(), host_only=True)


def test_create_rejects_inconsistent_descriptors(nd):
    """nd_b200_create validates that the descriptor describes the layout register_vertices! / register_edges! produce
    (src/network_structure.jl:224-258); every single-field corruption of a valid descriptor is refused with a status code
    and a message -- no crash, no engine (host-only build of the real library: no GPU needed)."""
    import copy
    import ctypes as C
    cabi = nd._cabi
    L = nd.Lib
    rng = np.random.default_rng(5)
    g = nd.barabasi_albert(300, 3, seed=1)
    vm = ([L.kuramoto_first(), L.kuramoto_second()], rng.integers(0, 2, g.nv))
    em = ([L.kuramoto_edge(), L.diffusion_edge_nop(), L.diffusion_odeedge()], rng.integers(0, 3, g.ne))
    nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", host_only=True))
    agg = nw.layer.aggregator
    lib = cabi.lib()

    def create(desc):
        h = C.c_void_p()
        rc = lib.nd_b200_create(C.byref(desc), C.byref(h))
        if rc == cabi.OK:
            lib.nd_b200_destroy(h)
        else:
            assert not h.value and lib.nd_b200_last_error(None), "a refusal carries a message and no engine"
        return rc

    def clone():
        d = cabi.Desc.from_buffer_copy(agg._desc)
        vb = (cabi.VBatch * d.n_vbatches).from_buffer_copy((cabi.VBatch * d.n_vbatches).from_address(C.addressof(d.vbatches.contents)))
        eb = (cabi.EBatch * d.n_ebatches).from_buffer_copy((cabi.EBatch * d.n_ebatches).from_address(C.addressof(d.ebatches.contents)))
        d.vbatches, d.ebatches = vb, eb
        return d, vb, eb

    d, _, _ = clone()
    assert create(d) == cabi.OK                                 # the untouched copy is fine
    top = ["abi_version", "nv", "ne", "vdepth", "edepth", "n_vbatches", "n_ebatches", "lastidx_dynamic", "lastidx_p", "lastidx_out"]
    refused = 0
    for name in top:
        for delta in (1, -1, 1000):
            d, _, _ = clone()
            setattr(d, name, getattr(d, name) + delta)
            if name in ("n_vbatches", "n_ebatches", "nv", "ne") and delta > 0:
                continue                                        # would make the engine read past the caller's arrays
            assert create(d) in (cabi.EINVAL, cabi.EUNSUPPORTED), (name, delta)
            refused += 1
    for k in range(2):
        for name in ("kind", "dim", "pdim", "outdim", "count", "state_first", "p_first", "out_first", "aggr_first"):
            d, vb, _ = clone()
            delta = -1 if name == "count" else 1
            setattr(vb[k], name, getattr(vb[k], name) + delta)
            if name == "kind" and k == 0:
                setattr(vb[k], name, 77)
            rc = create(d)
            if name == "p_first" and vb[k].pdim == 0:
                continue
            assert rc in (cabi.EINVAL, cabi.EUNSUPPORTED), ("vbatch", k, name)
            refused += 1
    for k in range(3):
        for name in ("kind", "coupling", "dim", "pdim", "outdim_src", "outdim_dst", "count", "state_first", "p_first", "out_first",
                     "mask_src_first", "mask_dst_first"):
            d, _, eb = clone()
            old = getattr(eb[k], name)
            setattr(eb[k], name, old - 1 if name == "count" else (9 if name == "coupling" else old + (5 if name.startswith("mask") else 1)))
            rc = create(d)
            irrelevant = (name == "p_first" and eb[k].pdim == 0) or (name == "state_first" and eb[k].dim == 0) or \
                         (name.startswith("mask") and eb[k].dim == 0) or (name == "mask_src_first" and eb[k].coupling != cabi.FIDUCIAL)
            if irrelevant:
                continue
            assert rc in (cabi.EINVAL, cabi.EUNSUPPORTED), ("ebatch", k, name)
            refused += 1
    # null tables, duplicated / out-of-range component ids, edges with bad endpoints
    for field in ("vbatches", "ebatches", "edge_src", "edge_dst"):
        d, _, _ = clone()
        setattr(d, field, None)
        assert create(d) == cabi.EINVAL, field
    d, vb, _ = clone()
    idx = np.ctypeslib.as_array(vb[0].indices, shape=(vb[0].count,)).copy()
    idx[1] = idx[0]
    vb[0].indices = idx.ctypes.data_as(cabi.i64p)
    assert create(d) == cabi.EINVAL
    d, _, eb = clone()
    idx = np.ctypeslib.as_array(eb[0].indices, shape=(eb[0].count,)).copy()
    idx[-1] = g.ne + 5
    eb[0].indices = idx.ctypes.data_as(cabi.i64p)
    assert create(d) == cabi.EINVAL
    d, _, _ = clone()
    bad = np.ascontiguousarray(g.src).copy()
    bad[3] = g.nv + 1
    d.edge_src = bad.ctypes.data_as(cabi.i64p)
    assert create(d) == cabi.EINVAL
    assert refused > 60
