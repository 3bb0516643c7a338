#!/usr/bin/env python
"""End-to-end (host buffers) time of nw(du,u,p,t) on cfg2 for several ND_B200_HOST_CHUNKS; one JSON line each."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ndb200 as nd

g = nd.erdos_renyi(1_000_000, 4_000_000, seed=1)
nw = nd.Network(g, nd.Lib.diffusion_vertex(), nd.Lib.diffusion_edge(), aggregator=nd.B200Aggregator("+", keep_tables=False))
hu, hp, hdu = nd.pinned_empty(nw.dim()), nd.pinned_empty(nw.pdim()), nd.pinned_empty(nw.dim())
hu[:] = np.random.default_rng(1).random(nw.dim()); hp[:] = np.random.default_rng(2).random(nw.pdim())
ref = None
for chunks in sys.argv[1:] or ["1", "2", "3", "4", "5", "6", "8"]:
    os.environ["ND_B200_HOST_CHUNKS"] = chunks
    for _ in range(5):
        nw(hdu, hu, hp, 0.0)
    ts = []
    for _ in range(100):
        t0 = time.perf_counter(); nw(hdu, hu, hp, 0.0); ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = hdu.copy()
    print(json.dumps({"chunks": int(chunks), "e2e_ms_mean": 1e3 * float(np.mean(ts)), "e2e_ms_min": 1e3 * float(np.min(ts)),
                      "edge_evals_per_s": g.ne / float(np.mean(ts)), "same_as_unpipelined": bool(np.array_equal(hdu, ref))}), flush=True)
