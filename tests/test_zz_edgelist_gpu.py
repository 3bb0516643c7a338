"""GPU check of nd_b200_create_from_edgelist: the engine built from a bare edge list evaluates exactly like the one built
from the reference's tables (same kernels, same layout; bit-identical `du`)."""
import numpy as np
import pytest


def test_from_edgelist_matches_table_driven_network(nd, backend):
    B = backend
    L = nd.Lib
    for g, vm, em in [(nd.erdos_renyi(int(20_000 * B.scale), int(160_000 * B.scale), seed=5), L.kuramoto_first(), L.kuramoto_edge()),
                      (nd.grid_graph(40, 50), L.swing_dq(), L.line_dq())]:
        a = nd.Network(g, vm, em)
        b = nd.Network.from_edgelist(g, vm, em)
        rng = np.random.default_rng(3)
        u = B.dev(rng.random(a.dim()))
        p = B.dev(0.25 + rng.random(a.pdim()))
        da, db = B.nan(a.dim()), B.nan(a.dim())
        a(da, u, p, 0.0)
        b(db, u, p, 0.0)
        da, db = B.host(da), B.host(db)
        assert not np.isnan(db).any() and np.array_equal(da, db)
