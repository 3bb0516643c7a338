"""Pin the CPU oracle against the reference's own Julia-free known answers.

The reference is Julia and cannot run here, and it ships no golden vectors; what its tests DO hold for this path
are closed-form answers and layout facts.  Each test names the reference test it restates.
"""
import numpy as np
import pytest

from helpers import espec_of, oracle_network, vspec_of
from oracle import oracle as O
from oracle import oracle_np as ONP


def _np_im(g, vm, vt, em, et):
    return ONP.IndexManager(g.nv, g.src, g.dst, [vspec_of(m) for m in vm], list(vt), [espec_of(m) for m in em], list(et))


def test_find_identical_order(nd):
    """test/utils_test.jl:43-61: [v1,v2,v3,v2,v2,v1,v1,v3] -> [[1,6,7],[2,4,5],[3,8]]"""
    assert ONP.find_identical(["v1", "v2", "v3", "v2", "v2", "v1", "v1", "v3"]) == [[1, 6, 7], [2, 4, 5], [3, 8]]
    L = nd.Lib
    vs = [L.kuramoto_second(), L.diffusion_vertex(), L.kuramoto_first(), L.diffusion_vertex(), L.diffusion_vertex(),
          L.kuramoto_second(), L.kuramoto_second(), L.kuramoto_first()]
    g = nd.path_graph(8)
    onw = oracle_network(g, vs, L.kuramoto_edge())
    assert [list(idx) for _, idx in onw.batches("vertex")] == [[1, 6, 7], [2, 4, 5], [3, 8]]
    assert [list(i) for i in nd.find_identical(np.array([0, 1, 2, 1, 1, 0, 0, 2]))] == [[1, 6, 7], [2, 4, 5], [3, 8]]


def test_handworked_three_vertex_layout(nd):
    """complete_graph(3), [kuramoto_first, kuramoto_second, kuramoto_first], kuramoto_edge."""
    L = nd.Lib
    g = nd.complete_graph(3)
    assert list(zip(g.src, g.dst)) == [(1, 2), (1, 3), (2, 3)]
    onw = oracle_network(g, [L.kuramoto_first(), L.kuramoto_second(), L.kuramoto_first()], L.kuramoto_edge())
    assert [list(i) for _, i in onw.batches("vertex")] == [[1, 3], [2]]
    assert list(onw.table("v_data")) == [1, 3, 2]
    assert list(onw.table("v_out")) == [1, 3, 2]
    assert list(onw.table("v_para")) == [1, 3, 2]
    assert list(onw.table("v_aggr")) == [1, 3, 2]
    assert list(onw.table("e_out_src")) == [4, 6, 8] and list(onw.table("e_out_dst")) == [5, 7, 9]
    assert list(onw.table("e_para")) == [6, 7, 8]
    assert list(onw.table("e_gbuf_src")) == [1, 3, 5] and list(onw.table("e_gbuf_dst")) == [2, 4, 6]
    assert list(onw.table("gbufmap")) == [1, 3, 1, 2, 3, 2]
    assert (onw.aggmap_first, onw.aggmap_len) == (4, 6)
    assert list(onw.table("aggmap")) == [1, 3, 1, 2, 3, 2]
    assert (onw.lastidx_dynamic, onw.lastidx_p, onw.lastidx_out, onw.lastidx_aggr, onw.lastidx_gbuf) == (4, 8, 9, 3, 6)
    # u = [th1, th3, d2, w2], p = [w1, w3, M2, D2, Pm2, K1, K2, K3]
    th1, th3, d2, w2 = 0.3, -0.7, 1.1, 0.25
    w1, w3, M2, D2, Pm2, K1, K2, K3 = 0.5, -0.2, 1.7, 0.1, 0.9, 2.0, 3.0, 5.0
    du, o, agg = onw.rhs([th1, th3, d2, w2], [w1, w3, M2, D2, Pm2, K1, K2, K3], return_bufs=True)
    e12, e13, e23 = K1 * np.sin(th1 - d2), K2 * np.sin(th1 - th3), K3 * np.sin(d2 - th3)
    assert np.array_equal(o, [th1, th3, d2, -e12, e12, -e13, e13, -e23, e23])
    assert agg[0] == (0.0 + -e12) + -e13 and agg[2] == e12 + -e23 and agg[1] == e13 + e23
    assert du[0] == w1 + agg[0] and du[1] == w3 + agg[1]
    assert du[2] == w2 and du[3] == 1.0 / M2 * (Pm2 - D2 * w2 + agg[2])


def test_homogeneous_two_state_layout(nd):
    """test/symbolicindexing_test.jl:27-29,42-62: 3 kuramoto_second vertices: VIndex(3,:ω) <-> u[6]; parameters
    are all vertices (M,D,Pm each) then all edges (K)."""
    L = nd.Lib
    g = nd.complete_graph(3)
    onw = oracle_network(g, L.kuramoto_second(), L.kuramoto_edge())
    assert onw.table("v_data")[2] + 1 == 6                       # ω is state 2 of vertex 3
    assert list(onw.table("v_para")) == [1, 4, 7] and list(onw.table("e_para")) == [10, 11, 12]
    # observed order [EIndex(1,:₋P), EIndex(1,:P), ...]: src output precedes dst output (:46)
    assert list(onw.table("e_out_src")) == [4, 6, 8] and list(onw.table("e_out_dst")) == [5, 7, 9]


def test_heterogeneous_complete4_layout(nd):
    """test/symbolicindexing_test.jl:84-111: complete_graph(4), vf=[k2,diff,k2,diff], ef=[ode,kura,kura,fid,ode,fid]:
    variable_index(nw, EIndex(1,:e_dst)) == 7 (all vertex batches precede edge states, edge 1 is in edge batch 1)."""
    g = nd.complete_graph(4)
    vs = [O.VSpec(O.V_KURAMOTO_SECOND, 2, 3, 1), O.VSpec(O.V_DIFFUSION, 1, 0, 1)]
    es = [O.ESpec(O.E_OPAQUE, O.FIDUCIAL, 2, 1, 1, 1),      # diffusion_odeedge: dim 2, pdim 1
          O.ESpec(O.E_KURAMOTO, O.ANTISYMMETRIC, 0, 1, 1, 1),
          O.ESpec(O.E_OPAQUE, O.FIDUCIAL, 0, 1, 1, 1)]      # diffusion_edge_fid
    onw = O.OracleNetwork(4, g.src, g.dst, vs, [0, 1, 0, 1], es, [0, 1, 1, 2, 0, 2])
    assert onw.table("e_data")[0] == 7
    assert list(onw.table("v_data")) == [1, 5, 3, 6]
    assert [list(i) for _, i in onw.batches("edge")] == [[1, 5], [2, 3], [4, 6]]
    assert onw.lastidx_dynamic == 10
    with pytest.raises(ValueError):
        onw.rhs(np.zeros(10), np.zeros(onw.lastidx_p))       # opaque kinds have no RHS in the oracle


def _ba_like(nd, n, m, seed):
    return nd.barabasi_albert(n, m, seed=seed)


@pytest.mark.parametrize("edge", ["diffusion_edge", "diffusion_edge_nop"])
def test_diffusion_equals_minus_laplacian(nd, edge):
    """test/diffusion_test.jl:80-90: nw(dx, x, nothing, 0) ≈ -L*x for 30 random x on barabasi_albert(10,5)."""
    g = _ba_like(nd, 10, 5, seed=42)
    Lm = g.laplacian()
    em = getattr(nd.Lib, edge)()
    onw = oracle_network(g, nd.Lib.diffusion_vertex(), em)
    p = np.ones(onw.lastidx_p)
    rng = np.random.default_rng(42)
    for _ in range(30):
        x = rng.standard_normal(g.nv)
        dx = onw.rhs(x, p)
        assert np.allclose(dx, -Lm @ x, rtol=1e-13, atol=1e-13)
        assert np.array_equal(dx, onw.rhs(x, p, threads=3))


def test_diffusion_trajectory_vs_matrix_exponential(nd):
    """test/diffusion_test.jl:119-129 restated with the oracle's fixed-step RK4: |x(t) - exp(-tL) x0| < 1e-6."""
    from scipy.linalg import expm
    g = _ba_like(nd, 10, 5, seed=7)
    Lm = g.laplacian()
    onw = oracle_network(g, nd.Lib.diffusion_vertex(), nd.Lib.diffusion_edge_nop())
    x0 = np.random.default_rng(3).random(g.nv)
    x = onw.rk4(x0, None, 0.0, 1e-3, 1000)
    assert np.max(np.abs(x - expm(-1.0 * Lm) @ x0)) < 1e-6


def _mixed_case(nd, seed):
    L = nd.Lib
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 40))
    g = nd.watts_strogatz(n, 4, 0.3, seed=seed)
    choices = [L.kuramoto_first(), L.kuramoto_second(), L.kuramoto_second_bench(), L.diffusion_vertex()]
    vm = [choices[k] for k in rng.integers(0, len(choices), size=n)]
    echoices = [L.kuramoto_edge(), L.diffusion_edge(), L.diffusion_edge_nop(),
                nd.EdgeModel(g=nd.Symmetric(L.diffusionedge), outdim=1, pdim=1, name="sym"),
                nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir")]
    em = [echoices[k] for k in rng.integers(0, len(echoices), size=g.ne)]
    return g, vm, em


@pytest.mark.parametrize("seed", range(8))
def test_c_oracle_equals_numpy_twin(nd, seed):
    """Two independent restatements (C and pure Python) agree bit for bit: tables, o, aggbuf and du."""
    from helpers import model_types
    g, vm, em = _mixed_case(nd, seed)
    onw = oracle_network(g, vm, em)
    uv, vt = model_types(vm, g.nv)
    ue, et = model_types(em, g.ne)
    im = _np_im(g, uv, vt, ue, et)
    assert [list(i) for _, i in onw.batches("vertex")] == [i for _, i in im.vbatches]
    assert [list(i) for _, i in onw.batches("edge")] == [i for _, i in im.ebatches]
    assert list(onw.table("v_data")) == [im.v_data[i + 1].first for i in range(g.nv)]
    assert list(onw.table("e_out_dst")) == [im.e_out[i + 1][1].first for i in range(g.ne)]
    assert np.array_equal(onw.table("gbufmap"), im.gbuf_map())
    first, amap = im.aggregation_map()
    assert onw.aggmap_first == first and np.array_equal(onw.table("aggmap"), amap)
    rng = np.random.default_rng(seed)
    u, p = rng.random(onw.lastidx_dynamic), rng.random(onw.lastidx_p) + 0.5
    du, o, agg = onw.rhs(u, p, return_bufs=True)
    du2, o2, agg2 = ONP.rhs(im, u, p)
    assert np.array_equal(o, o2, equal_nan=True) and np.array_equal(agg, agg2) and np.array_equal(du, du2)
    assert np.array_equal(du, onw.rhs(u, p, threads=4))


def test_dq_powergrid_oracle_matches_twin(nd):
    L = nd.Lib
    g = nd.grid_graph(5, 4)
    onw = oracle_network(g, L.swing_dq(), L.line_dq())
    im = _np_im(g, [L.swing_dq()], np.zeros(g.nv, int), [L.line_dq()], np.zeros(g.ne, int))
    rng = np.random.default_rng(0)
    u, p = rng.random(onw.lastidx_dynamic), rng.random(onw.lastidx_p) + 0.5
    du, o, agg = onw.rhs(u, p, return_bufs=True)
    du2, o2, agg2 = ONP.rhs(im, u, p)
    assert np.array_equal(o, o2) and np.array_equal(agg, agg2) and np.array_equal(du, du2)
    assert (onw.vdepth, onw.edepth) == (2, 2)
    assert onw.lastidx_out == 2 * g.nv + 4 * g.ne and onw.lastidx_aggr == 2 * g.nv


def test_rk4_threaded_equals_sequential(nd):
    g, vm, em = _mixed_case(nd, 3)
    onw = oracle_network(g, vm, em)
    rng = np.random.default_rng(1)
    u, p = rng.random(onw.lastidx_dynamic), rng.random(onw.lastidx_p) + 0.5
    assert np.array_equal(onw.rk4(u, p, 0.0, 1e-3, 20), onw.rk4(u, p, 0.0, 1e-3, 20, threads=3))


def test_size_checks_raise(nd):
    """src/coreloop.jl:2-7: wrong-size u / p raise ArgumentError"""
    g = nd.complete_graph(3)
    onw = oracle_network(g, nd.Lib.kuramoto_first(), nd.Lib.kuramoto_edge())
    with pytest.raises(ValueError):
        onw.rhs(np.zeros(4), np.zeros(6))
    with pytest.raises(ValueError):
        onw.rhs(np.zeros(3), np.zeros(5))


def test_edges_with_states_known_answers(nd):
    """Edges with states (src/coreloop.jl:41,76).  Layout: test/GPU_test.jl:12-22's network has dim 10 = 2*2 + 2*1 vertex
    states followed by 2*2 edge states, pdim 12.  Known answer (test/diffusion_test.jl:96-129, real_ode_edge! with
    g=Fiducial(2,1)): with every edge state on its constraint e = (vs-vd, vd-vs) the vertex part of du is -L*x and the
    edge part is exactly 0; off the constraint the edge part is the residual.  C oracle == Python twin bit for bit."""
    L = nd.Lib
    g = nd.complete_graph(4)
    vm = [L.kuramoto_second(), L.diffusion_vertex(), L.kuramoto_second(), L.diffusion_vertex()]
    em = [L.diffusion_odeedge(), L.kuramoto_edge(), L.kuramoto_edge(), L.diffusion_edge_fid(), L.diffusion_odeedge(), L.diffusion_edge_fid()]
    onw = oracle_network(g, vm, em)
    assert (onw.lastidx_dynamic, onw.lastidx_p, onw.lastidx_out) == (10, 12, 4 + 2 * 6)
    assert [list(i) for _, i in onw.batches("edge")] == [[1, 5], [2, 3], [4, 6]]
    assert list(onw.table("e_data")) == [7, 11, 11, 11, 9, 11]          # stateful batch first: 7:8, 9:10; empty ranges after
    from helpers import model_types
    ems, et = model_types(em, g.ne)
    vms, vt = model_types(vm, g.nv)
    im = _np_im(g, vms, vt, ems, et)
    rng = np.random.default_rng(3)
    u, p = rng.random(10), 0.5 + rng.random(12)
    du, o, agg = onw.rhs(u, p, return_bufs=True)
    du2, o2, agg2 = ONP.rhs(im, u, p)
    assert np.array_equal(du, du2) and np.array_equal(o, o2) and np.array_equal(agg, agg2)
    assert np.array_equal(onw.rhs(u, p, threads=3), du)
    # edge 1 = (1,2) is stateful: its dst output is its first state, its src output the second (Fiducial(dst=1:1, src=2:2))
    assert o[onw.table("e_out_dst")[0] - 1] == u[6] and o[onw.table("e_out_src")[0] - 1] == u[7]
    tau = p[onw.table("e_para")[0] - 1]
    th1, x2 = u[0], u[4]                                                  # vertex 1 = kuramoto_second (state 1), vertex 2 = diffusion
    assert du[6] == 1.0 / tau * (np.sin(th1 - x2) - u[6]) and du[7] == 1.0 / tau * (np.sin(x2 - th1) - u[7])

    g = nd.erdos_renyi(200, 800, seed=3)
    onw = oracle_network(g, L.diffusion_vertex(), L.relax_odeedge())
    x = rng.standard_normal(g.nv)
    e = np.stack([x[g.src - 1] - x[g.dst - 1], x[g.dst - 1] - x[g.src - 1]], axis=1).ravel()
    du = onw.rhs(np.concatenate([x, e]), None)
    assert np.allclose(du[:g.nv], -g.laplacian() @ x, rtol=1e-12, atol=1e-12) and np.all(du[g.nv:] == 0.0)
    e2 = e + 0.25
    du = onw.rhs(np.concatenate([x, e2]), None)
    assert np.allclose(du[g.nv:], e - e2, rtol=0, atol=1e-15)


def test_loopback_known_answers(nd):
    """LoopbackConnection (src/post_utils.jl:105-234, src/coreloop.jl:47): after the RHS the aggregation slot of an injector
    leaf holds its hub's output (apply_loopback!), the hub's slot holds the sum of its ordinary edges plus MINUS the injector's
    output (LOOPBACK_G: outdst = -1 .* insrc), and the loopback edge has no src output.  C oracle == Python twin."""
    L = nd.Lib
    n = 6
    ring_s, ring_d = np.arange(1, n + 1), np.roll(np.arange(1, n + 1), -1)
    g = nd.SimpleDiGraph(2 * n, np.concatenate([ring_s, np.arange(n + 1, 2 * n + 1)]), np.concatenate([ring_d, np.arange(1, n + 1)]))
    vm = [L.kuramoto_first()] * n + [L.kuramoto_second()] * n
    em = ([L.kuramoto_edge(), L.loopback()], (g.src > n).astype(np.int64))
    onw = oracle_network(g, vm, em)
    rng = np.random.default_rng(4)
    u, p = rng.random(onw.lastidx_dynamic), 0.5 + rng.random(onw.lastidx_p)
    du, o, agg = onw.rhs(u, p, return_bufs=True)
    assert onw.lastidx_out == 2 * n + 2 * n + n                   # vertex outputs, ring edges (src+dst), loopback edges (dst only)
    v_out, v_aggr = onw.table("v_out"), onw.table("v_aggr")
    e_dst_out = onw.table("e_out_dst")
    for k in range(n):
        hub, inj = k + 1, n + k + 1
        assert agg[v_aggr[inj - 1] - 1] == o[v_out[hub - 1] - 1] == u[k]            # injector input = hub output = theta_hub
        loop_edge = int(np.nonzero((g.src == inj) & (g.dst == hub))[0][0])
        assert o[e_dst_out[loop_edge] - 1] == -1.0 * o[v_out[inj - 1] - 1]           # hub receives minus the injector's output
    from helpers import model_types
    um, vt = model_types(vm, g.nv)
    uem, et = model_types(em, g.ne)
    im = _np_im(g, um, vt, uem, et)
    du2, o2, agg2 = ONP.rhs(im, u, p)
    assert np.array_equal(du, du2) and np.array_equal(o, o2, equal_nan=True) and np.array_equal(agg, agg2)
    assert np.array_equal(onw.rhs(u, p, threads=3), du)
    # the inertial injector: dv2 = 1/M (Pm - D v2 + theta_hub)
    inj_states = onw.table("v_data")[n] - 1
    M, D, Pm = p[onw.table("v_para")[n] - 1: onw.table("v_para")[n] + 2]
    assert du[inj_states + 1] == 1.0 / M * (Pm - D * u[inj_states + 1] + u[0])


def _kuramoto_path4_fixpoint(nd):
    """test/linear_analysis_test.jl:42-61 ("Kuramoto system test"): path_graph(4), Lib.kuramoto_second() vertices with
    Pm = [1, -0.5, -0.5, 0], M = 1, D = 0.2, Lib.kuramoto_edge() with K = 2; the reference finds a fixpoint of it and asserts
    `isfixpoint(s0)`.  With the reference's conventions (test/ComponentLibrary.jl:51-67: e = K sin(theta_src - theta_dst) enters
    dst with +, src with - (AntiSymmetric); dv2 = 1/M (Pm - D w + acc); p = (M, D, Pm)) the fixpoint is closed form:
        node 4: sin(th3 - th4) = 0;  node 1: 1 - 2 sin(th1 - th2) = 0;  node 3: -0.5 + 2 sin(th2 - th3) = 0;  node 2: balanced
    => theta = (pi/6 + asin(1/4), asin(1/4), 0, 0) + c, omega = 0.  Any sign / parameter-order slip leaves |du| = O(1)."""
    g = nd.path_graph(4)
    vm, em = nd.Lib.kuramoto_second(), nd.Lib.kuramoto_edge()
    nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", host_only=True))
    c = 0.37
    theta = np.array([np.pi / 6 + np.arcsin(0.25), np.arcsin(0.25), 0.0, 0.0]) + c
    u = np.zeros(nw.dim())
    u[0::2] = theta                                   # (delta_i, omega_i) interleaved: one homogeneous batch
    p = np.zeros(nw.pdim())
    p[0:12:3], p[1:12:3], p[2:12:3] = 1.0, 0.2, [1.0, -0.5, -0.5, 0.0]   # (M, D, Pm) per vertex
    p[12:15] = 2.0                                    # K per edge
    return g, vm, em, u, p


def test_kuramoto_fixpoint_of_the_reference_linear_analysis_test(nd):
    """the oracle (C and Python twin) evaluates du = 0 at the closed-form fixpoint of test/linear_analysis_test.jl:42-61 --
    a reference-held pin of the Kuramoto edge + inertial vertex arithmetic (signs, AntiSymmetric direction, parameter order)"""
    g, vm, em, u, p = _kuramoto_path4_fixpoint(nd)
    onw = oracle_network(g, vm, em)
    du = onw.rhs(u, p)
    assert np.max(np.abs(du)) <= 4e-16, du
    from helpers import model_types
    uv, vt = model_types(vm, g.nv)
    ue, et = model_types(em, g.ne)
    du_twin, _, _ = ONP.rhs(_np_im(g, uv, vt, ue, et), u, p)
    assert np.max(np.abs(du_twin)) <= 4e-16
    # off the fixpoint the residual is what the conventions say: more mechanical power on node 1 accelerates node 1 only
    p2 = p.copy()
    p2[2] += 0.25
    du2 = onw.rhs(u, p2)
    assert abs(du2[1] - 0.25) <= 1e-15 and np.max(np.abs(np.delete(du2, 1))) <= 4e-16
    # relabelling every edge's direction is the same physics under AntiSymmetric coupling; a Symmetric wrapper is not
    flipped = oracle_network(nd.SimpleGraph(4, g.dst, g.src), vm, em)
    assert np.max(np.abs(flipped.rhs(u, p))) <= 4e-16
    sym = nd.EdgeModel(g=nd.Symmetric(nd.Lib.kuramoto_edge_f), outdim=1, pdim=1, psym=("K",), name="kuramoto_sym")
    assert np.max(np.abs(oracle_network(g, vm, sym).rhs(u, p))) > 0.5


def _dq_two_bus_fixpoint(nd, R=0.0):
    """Two `SwingDQ` buses (test/ComponentLibrary.jl:139-160: u = V (cos th, sin th), Pel = u_r i_r + u_i i_i,
    dw = 1/M (Pmech - D w + Pel)) joined by one `StaticPowerLineDQ` (:212-245: idst = active / (R + jX) (Vsrc - Vdst), isrc = -idst).
    Closed form from those equations: the power delivered to the dst bus is Re(Vdst conj(idst)), to the src bus Re(Vsrc conj(isrc));
    for R = 0 that is the textbook power-angle relation +-(V^2 / X) sin(th_src - th_dst).  Choosing Pmech = minus the delivered power
    makes (th, w = 0) a fixpoint."""
    g = nd.SimpleGraph(2, [1], [2])
    vm, em = nd.Lib.swing_dq(), nd.Lib.line_dq()
    nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", host_only=True))
    V, X, d = 1.05, 0.8, np.pi / 6
    ths, thd = 0.4 + d, 0.4
    Vs, Vd = V * np.exp(1j * ths), V * np.exp(1j * thd)
    idst = (Vs - Vd) / complex(R, X)
    Pdst, Psrc = (Vd * np.conj(idst)).real, (Vs * np.conj(-idst)).real
    u = np.array([ths, 0.0, thd, 0.0])
    p = np.array([2.0, 0.3, -Psrc, V,   1.5, 0.1, -Pdst, V,   R, X, 1.0])      # (M, D, Pmech, V) per bus, (R, X, active)
    assert nw.dim() == 4 and nw.pdim() == 11
    return g, vm, em, u, p, (Psrc, Pdst, V, X, d)


def test_dq_swing_and_line_fixpoint_closed_form(nd):
    """the hand-written equivalents of the MTK dq models (oracle, C and Python twin) reproduce the closed-form fixpoint of the
    equations they restate; for R = 0 the delivered power is the power-angle relation"""
    from helpers import model_types
    for R in (0.0, 0.25):
        g, vm, em, u, p, (Psrc, Pdst, V, X, d) = _dq_two_bus_fixpoint(nd, R)
        if R == 0.0:
            assert abs(Pdst - V * V / X * np.sin(d)) <= 1e-15 and abs(Psrc + V * V / X * np.sin(d)) <= 1e-15
        else:
            assert Psrc + Pdst < 0.0            # a resistive line dissipates
        onw = oracle_network(g, vm, em)
        du = onw.rhs(u, p)
        assert np.max(np.abs(du)) <= 1e-15, (R, du)
        uv, vt = model_types(vm, g.nv)
        ue, et = model_types(em, g.ne)
        du_twin, _, _ = ONP.rhs(_np_im(g, uv, vt, ue, et), u, p)
        assert np.max(np.abs(du_twin)) <= 1e-15
        # an opened line (active = 0) leaves each machine with its mechanical power only: dw = Pmech / M
        p0 = p.copy()
        p0[10] = 0.0
        du0 = onw.rhs(u, p0)
        assert np.allclose(du0, [0.0, p[2] / p[0], 0.0, p[6] / p[4]], rtol=0, atol=1e-15)
