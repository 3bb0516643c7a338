#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference (Julia) cannot run in this image and ships no stored vectors, so these fixtures are NOT
reference outputs: they are outputs of the CPU oracle (oracle/nd_oracle.c, pinned to the reference's Julia-free known
answers by tests/test_oracle_pins.py) on small seeded networks, committed so that (a) a change of the oracle's
arithmetic or index construction shows up as a diff of committed data, and (b) the CUDA path is compared with stored
numbers, not only with an oracle running in the same process.   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases(nd):
    """name -> (graph, vertex models, edge models); every BASELINE.json config family at a size of a few thousand"""
    L = nd.Lib
    rng = np.random.default_rng(17)
    n = 1200
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    return {
        "cfg1_kuramoto_ws": (nd.watts_strogatz(1000, 10, 0.1, seed=1), L.kuramoto_first(), L.kuramoto_edge()),
        "cfg2_diffusion_er": (nd.erdos_renyi(2000, 8000, seed=1), L.diffusion_vertex(), L.diffusion_edge()),
        "cfg3_mixed_kuramoto_ba": (nd.barabasi_albert(n, 4, seed=1), ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(half)), L.kuramoto_edge()),
        "cfg4_powergrid_grid": (nd.grid_graph(24, 20), L.swing_dq(), L.line_dq()),
        "cfg5_kuramoto_er_deg16": (nd.erdos_renyi(1500, 12000, seed=5), L.kuramoto_first(), L.kuramoto_edge()),
    }


def inputs(nw, helpers):
    u, p = helpers.rand_inputs(nw.dim(), nw.pdim(), seed=7, layout=lambda q: helpers.condition_params(nw, q))
    return u, p


def main():
    import ndb200 as nd
    import helpers
    for name, (g, vm, em) in cases(nd).items():
        nw = nd.Network(g, vm, em, aggregator=helpers.null_aggregator)
        onw = helpers.oracle_network(g, vm, em)
        u, p = inputs(nw, helpers)
        du, o, agg = onw.rhs(u, p, return_bufs=True)
        traj = onw.rk4(u, p, 0.0, 1e-3, 50)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), src=g.src, dst=g.dst, u=u, p=p, du=du, o=o, aggbuf=agg,
                            rk4_50=traj, lastidx=np.array([onw.lastidx_dynamic, onw.lastidx_p, onw.lastidx_out, onw.lastidx_aggr]))
        print(name, "written", du.size)


if __name__ == "__main__":
    main()
