#!/usr/bin/env python
"""Static cost model of the jagged (warp-slice) layout, from the engine's own tables (host-only engines: no GPU needed).

For every 32-lane slice the walk of rhs_jag_kernel costs
  iterations      = the longest lane of the slice (every column is one warp-wide load + gather, idle lanes included)
  gather sectors  = distinct 32-byte sectors among the neighbour outputs a column gathers (what the L1TEX pipe counts)
  self sectors    = distinct sectors of the lanes' OWN u / du entries (coalesced when the slice holds consecutive rows)
The degree-bucketed layout (ND_B200_JAG_WINDOW) trades self sectors for iterations; this prints both so that a measured
time can be set against them.   python tools/jag_model.py [er|ba|ws|grid] [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ndb200 as nd  # noqa: E402


def model(g, window):
    os.environ["ND_B200_KERNEL"] = "jag"
    os.environ["ND_B200_JAG_WINDOW"] = str(window)
    L = nd.Lib
    nw = nd.Network(g, L.diffusion_vertex(), L.diffusion_edge(), aggregator=nd.B200Aggregator("+", host_only=True))
    jag = nw.export_jag()
    _rowptr, nbr, _eid, _side = nw.export_tables()
    lanes = jag["lanes"].astype(np.int64)
    ln, rowrel, valid = lanes & 63, (lanes >> 6) & 127, (lanes >> 14) & 1
    order = jag["order"]
    iters = int(ln.max(axis=1).sum())
    entries = int(ln.sum())
    self_sectors = gather_sectors = 0
    for s, (e0, row0, _b, _mp) in enumerate(jag["slices"]):
        rows = row0 + rowrel[s][valid[s] == 1]
        self_sectors += np.unique(rows // 4).size            # 8-byte states: 4 per 32-byte sector
        base, j = int(e0), 0
        while True:
            act = ln[s] > j
            k = int(act.sum())
            if k == 0:
                break
            cols = nbr[order[base:base + k]] - 1             # neighbour vertex ids (1 state each)
            gather_sectors += np.unique(cols // 4).size
            base += k
            j += 1
    return dict(window=window, slices=len(jag["slices"]), entries=entries, iterations=iters,
                lane_utilisation=round(entries / (32.0 * iters), 3), gather_sectors=int(gather_sectors),
                self_sectors_per_pass=int(self_sectors))


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "er"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    g = {"er": lambda: nd.erdos_renyi(n, 4 * n, seed=1), "ba": lambda: nd.barabasi_albert(n, 4, seed=1),
         "ws": lambda: nd.watts_strogatz(n, 10, 0.1, seed=1), "grid": lambda: nd.grid_graph(int(n ** 0.5), int(n ** 0.5))}[kind]()
    print(f"{kind}: {g.nv} vertices, {g.ne} edges")
    base = None
    for w in (32, 64, 128):
        m = model(g, w)
        base = base or m
        m["iterations_vs_32"] = round(m["iterations"] / base["iterations"], 3)
        # u read + du write per row pass, relative to the gather sectors that dominate the L1TEX pipe
        m["self_overhead_vs_gathers"] = round(2 * (m["self_sectors_per_pass"] - base["self_sectors_per_pass"]) / m["gather_sectors"], 3)
        print(m)


if __name__ == "__main__":
    main()
