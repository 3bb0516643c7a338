"""User-supplied component kinds: component functions given as CUDA C++ bodies are spliced into the
fused kernels and compiled for sm_100a when the network is built (NVRTC), lifting the fixed registry.

CPU tests: the generated source compiles (host-only engines compile but do not load), errors surface as ArgumentError
with the compiler log.  Parity tests (`backend` fixture: the B200, and the CPU emulation of the same generated source in
tests/cusim): `du`, `get_buffers` and RK4 against the pure-Python oracle twin evaluating the SAME models from host
callables (oracle/oracle_np.py: PyKind), on small graphs the twin finishes in seconds.
"""
import math

import numpy as np
import pytest

from helpers import floored_rel_err
from oracle import oracle as O
from oracle import oracle_np as ONP


def _models(nd):
    """FitzHugh-Nagumo vertices (dim 2, coupling through the first state), Stuart-Landau-like vertices with a computed
    2-dim output (non-StateMask g), weighted sine / cubic edges, a two-sided (Fiducial) edge, a 2-dim edge."""
    C = nd.CudaFunction
    fhn = nd.VertexModel(
        f=C("fhn_f", "vertex_f", "dv[0] = v[0] - v[0]*v[0]*v[0]/3.0 - v[1] + esum[0]; dv[1] = p[0]*(v[0] + p[1] - p[2]*v[1]);",
            py=lambda v, e, p, t: [v[0] - v[0] * v[0] * v[0] / 3.0 - v[1] + e[0], p[0] * (v[0] + p[1] - p[2] * v[1])]),
        g=nd.StateMask((1,)), dim=2, pdim=3, name="fhn")
    relax3 = nd.VertexModel(      # dim 3, time dependent forcing
        f=C("relax3_f", "vertex_f", "dv[0] = -p[0]*v[0] + esum[0] + sin(t); dv[1] = v[0] - v[1]; dv[2] = v[1] - v[2]*v[2];",
            py=lambda v, e, p, t: [-p[0] * v[0] + e[0] + math.sin(t), v[0] - v[1], v[1] - v[2] * v[2]]),
        g=nd.StateMask((1,)), dim=3, pdim=1, name="relax3")
    wsin = nd.EdgeModel(g=nd.AntiSymmetric(C("wsin", "edge_g", "e_dst[0] = p[0] * sin(v_src[0] - v_dst[0] - p[1]);",
                                             py=lambda vs, vd, p, t: [p[0] * math.sin(vs[0] - vd[0] - p[1])])),
                        outdim=1, pdim=2, name="wsin")
    cubic = nd.EdgeModel(g=nd.Symmetric(C("cubic", "edge_g", "double d = v_src[0] - v_dst[0]; e_dst[0] = d*d*d;",
                                          py=lambda vs, vd, p, t: [(vs[0] - vd[0]) * (vs[0] - vd[0]) * (vs[0] - vd[0])])),
                         outdim=1, pdim=0, name="cubic")
    fid = nd.EdgeModel(g=nd.Fiducial(C("fid", "edge_g2", "e_dst[0] = p[0]*(v_src[0] - v_dst[0]); e_src[0] = -0.5*e_dst[0] + t;",
                                       py=lambda vs, vd, p, t: ([-0.5 * (p[0] * (vs[0] - vd[0])) + t], [p[0] * (vs[0] - vd[0])]))),
                       outdim=1, pdim=1, name="fid")
    # vertices with a computed 2-dim output and 2-dim edges (vdepth = edepth = 2, like the dq models but user-supplied)
    osc = nd.VertexModel(
        f=C("osc_f", "vertex_f", "dv[0] = p[0] + esum[0]*cos(v[0]) - esum[1]*sin(v[0]); dv[1] = -v[1] + esum[0]*esum[1];",
            py=lambda v, e, p, t: [p[0] + e[0] * math.cos(v[0]) - e[1] * math.sin(v[0]), -v[1] + e[0] * e[1]]),
        g=C("osc_g", "vertex_g", "out[0] = (1.0 + v[1]) * cos(v[0]); out[1] = (1.0 + v[1]) * sin(v[0]);",
            py=lambda v, p, t: [(1.0 + v[1]) * math.cos(v[0]), (1.0 + v[1]) * math.sin(v[0])]),
        dim=2, pdim=1, outdim=2, name="osc")
    line2 = nd.EdgeModel(g=nd.AntiSymmetric(C("line2", "edge_g", "e_dst[0] = p[0]*(v_src[0]-v_dst[0]) - p[1]*(v_src[1]-v_dst[1]); e_dst[1] = p[1]*(v_src[0]-v_dst[0]) + p[0]*(v_src[1]-v_dst[1]);",
                                              py=lambda vs, vd, p, t: [p[0] * (vs[0] - vd[0]) - p[1] * (vs[1] - vd[1]), p[1] * (vs[0] - vd[0]) + p[0] * (vs[1] - vd[1])])),
                         outdim=2, pdim=2, name="line2")
    # edges WITH states (user-supplied f, StateMask outputs): an RL line between the 2-dim outputs of `osc` (states = the
    # two current components, output AntiSymmetric(1:2)), and a 3-state edge whose two sides show different states
    dynline = nd.EdgeModel(f=C("dynline_f", "edge_f", "de[0] = (v_src[0] - v_dst[0] - p[0]*e[0] + p[1]*e[1]) / p[1]; de[1] = (v_src[1] - v_dst[1] - p[0]*e[1] - p[1]*e[0]) / p[1];",
                                py=lambda e, vs, vd, p, t: [(vs[0] - vd[0] - p[0] * e[0] + p[1] * e[1]) / p[1], (vs[1] - vd[1] - p[0] * e[1] - p[1] * e[0]) / p[1]]),
                           g=nd.AntiSymmetric((1, 2)), dim=2, pdim=2, outdim=2, name="dynline")
    lag3 = nd.EdgeModel(f=C("lag3_f", "edge_f", "de[0] = sin(t) - e[0]; de[1] = p[0]*(v_src[0] - v_dst[0]) - e[1]; de[2] = -e[1] - e[2]*e[0];",
                             py=lambda e, vs, vd, p, t: [math.sin(t) - e[0], p[0] * (vs[0] - vd[0]) - e[1], -e[1] - e[2] * e[0]]),
                        g=nd.Fiducial(dst=2, src=3), dim=3, pdim=1, outdim=1, name="lag3")
    return dict(fhn=fhn, relax3=relax3, wsin=wsin, cubic=cubic, fid=fid, osc=osc, line2=line2, dynline=dynline, lag3=lag3)


def _py_kind(m):
    """registry kinds keep their id; user-supplied kinds become host callables for the Python twin"""
    kk = m.kernel_kind()
    if kk is not None:
        return kk
    is_cuda = lambda x: x.__class__.__name__ == "CudaFunction"
    if hasattr(m, "outdim_dst"):     # edge: wrapper(CudaFunction) or an unwrapped two-sided CudaFunction; with states: f
        if m.dim > 0:
            return ONP.PyKind(f=m.f.py)
        return ONP.PyKind(g=(m.g if is_cuda(m.g) else m.g.g).py)
    return ONP.PyKind(f=m.f.py, g=(m.g.py if is_cuda(m.g) else None))


def _twin(g, vms, vtypes, ems, etypes):
    vs = [O.VSpec(_py_kind(m), m.dim, m.pdim, m.outdim) for m in vms]
    es = [O.ESpec(_py_kind(m), m.coupling, m.dim, m.pdim, m.outdim_src, m.outdim_dst, *(m.state_masks() or (0, 0))) for m in ems]
    return ONP.IndexManager(g.nv, g.src, g.dst, vs, list(vtypes), es, list(etypes))


def _cases(nd):
    M = _models(nd)
    L = nd.Lib
    rng = np.random.default_rng(3)
    g1 = nd.erdos_renyi(600, 2400, seed=1)
    yield "fhn+wsin", g1, [M["fhn"]], np.zeros(g1.nv, int), [M["wsin"]], np.zeros(g1.ne, int)
    g2 = nd.barabasi_albert(500, 3, seed=2)
    # user-supplied and registry kinds mixed in one network, all four wrappers, hubs above the row-split width
    yield ("mixed", g2, [M["fhn"], L.kuramoto_first(), M["relax3"], L.kuramoto_second()], rng.integers(0, 4, g2.nv),
           [M["wsin"], M["cubic"], M["fid"], L.kuramoto_edge(), L.diffusion_edge_nop()], rng.integers(0, 5, g2.ne))
    g3 = nd.grid_graph(20, 15)
    yield "osc+line2", g3, [M["osc"]], np.zeros(g3.nv, int), [M["line2"]], np.zeros(g3.ne, int)
    g4 = nd.watts_strogatz(400, 4, 0.3, seed=1, directed=True)
    dirw = nd.EdgeModel(g=nd.Directed(M["wsin"].g.g), outdim=1, pdim=2, name="dir_wsin")
    yield "directed", g4, [M["relax3"]], np.zeros(g4.nv, int), [dirw, M["fid"]], rng.integers(0, 2, g4.ne)


def _cases_stateful(nd):
    """edges with states: all edges dynamic lines on computed 2-dim vertex outputs; static + stateful + registry stateful
    mixed (first B200 run pending: used by tests/test_zzz_new_custom_kinds.py)"""
    M = _models(nd)
    L = nd.Lib
    rng = np.random.default_rng(13)
    g5 = nd.grid_graph(12, 10)
    yield "osc+dynline", g5, [M["osc"]], np.zeros(g5.nv, int), [M["dynline"], M["line2"]], rng.integers(0, 2, g5.ne)
    g6 = nd.barabasi_albert(300, 3, seed=5)
    yield ("stateful_mix", g6, [M["fhn"], L.kuramoto_first()], rng.integers(0, 2, g6.nv),
           [M["lag3"], M["wsin"], L.diffusion_odeedge(), L.diffusion_edge_fid()], rng.integers(0, 4, g6.ne))


def test_custom_kinds_compile_and_report_errors(nd):
    M = _models(nd)
    g = nd.erdos_renyi(300, 900, seed=1)
    nw = nd.Network(g, M["fhn"], M["wsin"], aggregator=nd.B200Aggregator("+", host_only=True))
    src = nw.custom_source()
    assert "vertex_f_1000" in src and "edge_g_1001" in src and "#define ND_MAX_VDIM 2" in src
    assert "rhs_fused_kernel" in src                        # the same kernel templates as the precompiled path
    # registry-only networks generate nothing
    assert nd.Network(g, nd.Lib.kuramoto_first(), nd.Lib.kuramoto_edge(), aggregator=nd.B200Aggregator("+", host_only=True)).custom_source() is None
    # a body that does not compile: ArgumentError carrying the compiler's message, never a fallback
    bad = nd.VertexModel(f=nd.CudaFunction("bad", "vertex_f", "dv[0] = no_such_symbol;"), g=nd.StateMask((1,)), dim=1, name="bad")
    with pytest.raises(nd.ArgumentError, match="no_such_symbol"):
        nd.Network(g, bad, M["wsin"], aggregator=nd.B200Aggregator("+", host_only=True))
    # a two-sided body needs the Fiducial wrapper (or no wrapper); a one-sided one cannot be Fiducial
    with pytest.raises(nd.ArgumentError):
        nd.Network(g, M["fhn"], nd.EdgeModel(g=nd.AntiSymmetric(M["fid"].g.g), outdim=1, pdim=1), aggregator=nd.B200Aggregator("+", host_only=True))
    # vdepth = edepth = 2 through user kinds, and the mixed network (4 vertex + 5 edge batches), compile as well
    for name, g, vms, vt, ems, et in list(_cases(nd)) + list(_cases_stateful(nd)):
        nw = nd.Network(g, (vms, vt), (ems, et), aggregator=nd.B200Aggregator("+", host_only=True))
        assert nw.custom_source() is not None, name


def _check_cases_against_twin(nd, B, monkeypatch, mode, cases):
    monkeypatch.setenv("ND_B200_KERNEL", mode)
    for name, g, vms, vt, ems, et in cases:
        nw = nd.Network(g, (vms, vt), (ems, et))
        im = _twin(g, vms, vt, ems, et)
        assert nw.dim() == im.last["dynamic"] and nw.pdim() == im.last["p"]
        rng = np.random.default_rng(5)
        u, p = rng.random(nw.dim()), 0.25 + rng.random(nw.pdim())
        for t in (0.0, 0.7):
            ref_du, ref_o, ref_agg = ONP.rhs(im, u, p, t)
            du = B.nan(nw.dim())
            nw(du, B.dev(u), B.dev(p), t)
            assert floored_rel_err(B.host(du), ref_du) <= 1e-12, (name, mode, t)
        o, agg = B.nan(nw.im.lastidx_out), B.nan(nw.im.lastidx_aggr)
        nw.get_buffers(o, agg, B.dev(u), B.dev(p), 0.7)
        assert floored_rel_err(B.host(o), ref_o) <= 1e-12, name
        assert floored_rel_err(B.host(agg), ref_agg) <= 1e-12, name


@pytest.mark.parametrize("mode", ["fused", "jag"])
def test_custom_kinds_match_python_twin(nd, backend, monkeypatch, mode):
    _check_cases_against_twin(nd, backend, monkeypatch, mode, _cases(nd))


def _check_rk4_and_host_buffers(nd, B, stateful):
    """fused-stage RK4 (vertex and edge states; user-supplied kinds may read t, so their steps are enqueued one by one
    with the stage times) and the host-buffer call work for user-supplied kinds"""
    M = _models(nd)
    L = nd.Lib
    g = nd.erdos_renyi(400, 1600, seed=4)
    rng = np.random.default_rng(1)
    nets = [([M["fhn"], L.kuramoto_first()], rng.integers(0, 2, g.nv), [M["lag3"], M["wsin"], L.diffusion_odeedge()], rng.integers(0, 3, g.ne))] \
        if stateful else [([M["fhn"]], np.zeros(g.nv, int), [M["wsin"]], np.zeros(g.ne, int))]
    for vms, vt, ems, et in nets:
        nw = nd.Network(g, (vms, vt), (ems, et))
        im = _twin(g, vms, vt, ems, et)
        u, p = rng.random(nw.dim()), 0.25 + rng.random(nw.pdim())
        dt, x, t0 = 1e-2, u.copy(), 0.3
        for s in range(5):
            t = t0 + s * dt
            k1 = ONP.rhs(im, x, p, t)[0]
            k2 = ONP.rhs(im, x + 0.5 * dt * k1, p, t + 0.5 * dt)[0]
            k3 = ONP.rhs(im, x + 0.5 * dt * k2, p, t + 0.5 * dt)[0]
            k4 = ONP.rhs(im, x + dt * k3, p, t + dt)[0]
            x = x + (dt / 6.0) * (((k1 + 2.0 * k2) + 2.0 * k3) + k4)
        ud = B.dev(u)
        nw.rk4(ud, B.dev(p), t0, dt, 5)
        assert floored_rel_err(B.host(ud), x) <= 1e-11
        hdu = np.empty_like(u)
        nw(hdu, u, p, 0.0)
        assert floored_rel_err(hdu, ONP.rhs(im, u, p)[0]) <= 1e-12


def test_custom_kinds_rk4_and_host_buffers(nd, backend):
    _check_rk4_and_host_buffers(nd, backend, stateful=False)
