"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): the C oracle reproduces them bit for
bit, the independent pure-Python twin agrees, the graph generators are stable, and -- on a GPU -- the CUDA path (and, in the CPU suite, its emulation in tests/cusim) matches
the stored numbers within the north_star tolerances (du 1e-12 per RHS; these 50-step trajectories 1e-11)."""
import importlib.util
import os

import numpy as np
import pytest

from helpers import condition_params, floored_rel_err, null_aggregator, oracle_network, rand_inputs

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
NAMES = ["cfg1_kuramoto_ws", "cfg2_diffusion_er", "cfg3_mixed_kuramoto_ba", "cfg4_powergrid_grid", "cfg5_kuramoto_er_deg16"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_committed_fixtures(nd, name):
    from oracle import oracle_np as ONP
    g, vm, em = make_golden.cases(nd)[name]
    z = np.load(os.path.join(HERE, name + ".npz"))
    assert np.array_equal(g.src, z["src"]) and np.array_equal(g.dst, z["dst"]), "graph generator changed"
    nw = nd.Network(g, vm, em, aggregator=null_aggregator)
    u, p = rand_inputs(nw.dim(), nw.pdim(), seed=7, layout=lambda q: condition_params(nw, q))
    assert np.array_equal(u, z["u"]) and np.array_equal(p, z["p"])
    onw = oracle_network(g, vm, em)
    assert list(z["lastidx"]) == [onw.lastidx_dynamic, onw.lastidx_p, onw.lastidx_out, onw.lastidx_aggr]
    du, o, agg = onw.rhs(u, p, return_bufs=True)
    assert np.array_equal(du, z["du"]) and np.array_equal(o, z["o"], equal_nan=True) and np.array_equal(agg, z["aggbuf"])
    assert np.array_equal(onw.rk4(u, p, 0.0, 1e-3, 50), z["rk4_50"])
    # the independent Python twin (different code, same reference semantics)
    from helpers import model_types, vspec_of, espec_of
    vms, vt = model_types(vm, g.nv)
    ems, et = model_types(em, g.ne)
    im = ONP.IndexManager(g.nv, g.src, g.dst, [vspec_of(m) for m in vms], list(vt), [espec_of(m) for m in ems], list(et))
    du2, o2, agg2 = ONP.rhs(im, u, p)
    assert np.array_equal(du2, z["du"]) and np.array_equal(agg2, z["aggbuf"])


@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_committed_fixtures(nd, backend, name):
    B = backend
    g, vm, em = make_golden.cases(nd)[name]
    z = np.load(os.path.join(HERE, name + ".npz"))
    nw = nd.Network(g, vm, em)
    u, p = B.dev(z["u"]), B.dev(z["p"])
    du = B.nan(nw.dim())
    nw(du, u, p, 0.0)
    o, agg = B.nan(nw.im.lastidx_out), B.nan(nw.im.lastidx_aggr)
    nw.get_buffers(o, agg, u, p, 0.0)
    ur = B.dev(z["u"])
    nw.rk4(ur, p, 0.0, 1e-3, 50)
    du, o, agg, ur = B.host(du), B.host(o), B.host(agg), B.host(ur)
    assert floored_rel_err(du, z["du"]) <= 1e-12
    assert floored_rel_err(o, z["o"]) <= 1e-12 and floored_rel_err(agg, z["aggbuf"]) <= 1e-12
    assert floored_rel_err(ur, z["rk4_50"]) <= 1e-11
    if name == "cfg2_diffusion_er":          # no transcendental, reference order: bit-identical
        assert np.array_equal(du, z["du"])
