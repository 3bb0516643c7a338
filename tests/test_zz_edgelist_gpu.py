"""GPU check of nd_b200_create_from_edgelist: the engine built from a bare edge list evaluates exactly like the one built
from the reference's tables (same kernels, same layout; bit-identical `du`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_from_edgelist_matches_table_driven_network(nd, cuda):
    torch = cuda
    L = nd.Lib
    for g, vm, em in [(nd.erdos_renyi(20_000, 160_000, seed=5), L.kuramoto_first(), L.kuramoto_edge()),
                      (nd.grid_graph(40, 50), L.swing_dq(), L.line_dq())]:
        a = nd.Network(g, vm, em)
        b = nd.Network.from_edgelist(g, vm, em)
        rng = np.random.default_rng(3)
        u = torch.from_numpy(rng.random(a.dim())).cuda()
        p = torch.from_numpy(0.25 + rng.random(a.pdim())).cuda()
        da, db = torch.full_like(u, float("nan")), torch.full_like(u, float("nan"))
        a(da, u, p, 0.0)
        b(db, u, p, 0.0)
        torch.cuda.synchronize()
        assert not torch.isnan(db).any() and torch.equal(da, db)
