"""Known-answer trajectories of the reference's own tests, run through the engine's fused RK4.

* test/inhomogeneous_test.jl:12-54 -- Barabasi-Albert(10, 5), vertex 1 is STATIC with output pi, the others are diffusion
  vertices, two edge functions of identical arithmetic under different names (hence two edge batches): "these dynamics
  should flow to pi", |x(500) - pi| < 1e-7.  The reference makes the static vertex a constraint (`ff_to_constraint`) for
  its mass-matrix solver; a state-free vertex with a computed output is the same network for an explicit stepper.
* test/diffusion_test.jl:119-129 -- |x(t) - exp(-tL) x0| < 1e-6 on the device (the oracle-side form is in
  tests/test_oracle_pins.py).
* test/linear_analysis_test.jl:42-61 -- the Kuramoto system whose fixpoint the reference finds and asserts: the engine evaluates
  du = 0 at the closed-form fixpoint and stays on it through the fused RK4.
"""
import math

import numpy as np

from helpers import floored_rel_err


def test_dynamics_flow_to_pi(nd, backend):
    B, C, L = backend, nd.CudaFunction, nd.Lib
    g = nd.barabasi_albert(10, 5, seed=42)
    stat = nd.VertexModel(f=C("stat_f", "vertex_f", "", py=lambda v, e, p, t: []),
                          g=C("stat_g", "vertex_g", "out[0] = 3.141592653589793;", py=lambda v, p, t: [math.pi]),
                          dim=0, pdim=0, outdim=1, name="statvertex")
    ode = L.diffusion_vertex()
    e1 = L.diffusion_edge_nop()
    e2 = nd.EdgeModel(g=nd.AntiSymmetric(C("diffusion_edge2", "edge_g", "e_dst[0] = v_src[0] - v_dst[0];",
                                           py=lambda vs, vd, p, t: [vs[0] - vd[0]])), outdim=1, pdim=0, name="staticedge2")
    em = [e1] * g.ne
    em[1] = e2
    nw = nd.Network(g, [stat] + [ode] * (g.nv - 1), em)
    assert nw.dim() == g.nv - 1 and nw.pdim() == 0 and len(nw.layer.edgebatches) == 2
    x0 = np.random.default_rng(42).random(nw.dim())
    # one RHS by hand: du_i = sum_j (x_j - x_i), x_1 = pi
    x = np.concatenate([[math.pi], x0])
    want = np.zeros(g.nv)
    for s, d in zip(g.src - 1, g.dst - 1):
        want[d] += x[s] - x[d]
        want[s] += x[d] - x[s]
    du = B.nan(nw.dim())
    nw(du, B.dev(x0), None, 0.0)
    assert np.allclose(B.host(du), want[1:], rtol=1e-13, atol=1e-13)
    # explicit RK4 is stable for dt * lambda_max < 2.78; lambda_max <= 2 * max degree
    deg = np.bincount(np.concatenate([g.src, g.dst]))[1:]
    dt = 1.0 / (2.0 * deg.max())
    ud = B.dev(x0)
    nw.rk4(ud, None, 0.0, dt, int(60.0 / dt))
    assert np.max(np.abs(B.host(ud) - math.pi)) < 1e-7


def test_diffusion_trajectory_matches_matrix_exponential(nd, backend):
    from scipy.linalg import expm
    B = backend
    g = nd.barabasi_albert(20, 5, seed=1)
    nw = nd.Network(g, nd.Lib.diffusion_vertex(), nd.Lib.diffusion_edge_nop())
    x0 = np.random.default_rng(1).random(g.nv)
    ud = B.dev(x0)
    nw.rk4(ud, None, 0.0, 1e-3, 1000)
    want = expm(-1.0 * g.laplacian().toarray() if hasattr(g.laplacian(), "toarray") else -1.0 * np.asarray(g.laplacian())) @ x0
    assert np.max(np.abs(B.host(ud) - want)) < 1e-6
    assert floored_rel_err(B.host(ud), want) < 1e-6


def test_kuramoto_fixpoint_of_the_reference_linear_analysis_test(nd, backend):
    """closed-form fixpoint of test/linear_analysis_test.jl:42-61 (derivation: tests/test_oracle_pins.py): du = 0 on the engine in
    both kernel families, and 200 fused RK4 steps do not leave it"""
    from test_oracle_pins import _kuramoto_path4_fixpoint
    B = backend
    g, vm, em, u, p = _kuramoto_path4_fixpoint(nd)
    nw = nd.Network(g, vm, em)
    du = B.nan(nw.dim())
    nw(du, B.dev(u), B.dev(p), 0.0)
    assert np.max(np.abs(B.host(du))) <= 1e-15
    ud = B.dev(u)
    nw.rk4(ud, B.dev(p), 0.0, 1e-2, 200)
    assert np.max(np.abs(B.host(ud) - u)) <= 1e-13
    # a perturbed state relaxes back (D > 0: the fixpoint is linearly stable, `is_linear_stable(s0; marginally_stable=true)`)
    u1 = u.copy()
    u1[0] += 0.05
    ud = B.dev(u1)
    nw.rk4(ud, B.dev(p), 0.0, 5e-2, 4000)          # 200 time units: e^(-D t / 2M) = 2e-9
    got = B.host(ud)
    assert np.max(np.abs(got[1::2])) <= 1e-6                       # omega -> 0
    th = got[0::2]
    assert abs((th[0] - th[1]) - np.pi / 6) <= 1e-6 and abs((th[1] - th[2]) - np.arcsin(0.25)) <= 1e-6 and abs(th[2] - th[3]) <= 1e-6


def test_dq_swing_and_line_fixpoint_closed_form(nd, backend):
    """closed-form fixpoint of the dq swing buses + static dq line (derivation: tests/test_oracle_pins.py) on the engine"""
    from test_oracle_pins import _dq_two_bus_fixpoint
    B = backend
    for R in (0.0, 0.25):
        g, vm, em, u, p, _ = _dq_two_bus_fixpoint(nd, R)
        nw = nd.Network(g, vm, em)
        du = B.nan(nw.dim())
        nw(du, B.dev(u), B.dev(p), 0.0)
        assert np.max(np.abs(B.host(du))) <= 2e-15, R
        ud = B.dev(u)
        nw.rk4(ud, B.dev(p), 0.0, 1e-2, 100)
        assert np.max(np.abs(B.host(ud) - u)) <= 1e-12, R
