#!/usr/bin/env python
"""parity of every launch shape of the streamed jagged kernel on a mid-size graph with few warps (long per-warp streams:
the shared-memory rings wrap many times).  python tools/js_check.py [n_vertices]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ndb200 as nd
from helpers import oracle_network, floored_rel_err
L = nd.Lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
g = nd.erdos_renyi(n, 4 * n, seed=3)
vm, em = L.kuramoto_first(), L.kuramoto_edge()
onw = oracle_network(g, vm, em)
os.environ["ND_B200_KERNEL"] = "js"
bad = 0
for u_ in (4, 8):
    for nst in (4, 8):
        for wps in (1, 4, 0):
            os.environ["ND_B200_JS_U"] = str(u_); os.environ["ND_B200_JS_NST"] = str(nst)
            if wps: os.environ["ND_B200_JS_WPS"] = str(wps)
            else: os.environ.pop("ND_B200_JS_WPS", None)
            nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", keep_tables=False))
            uh = np.random.default_rng(1).random(nw.dim()); ph = np.random.default_rng(2).random(nw.pdim())
            ref = onw.rhs(uh, ph, threads=8)
            u, p = torch.from_numpy(uh).cuda(), torch.from_numpy(ph).cuda()
            for packed in (False, True):
                nw.pack_params(p if packed else None)
                errs = []
                for rep in range(5):
                    du = torch.full_like(u, float("nan"))
                    nw(du, u, p, 0.0); torch.cuda.synchronize()
                    errs.append(floored_rel_err(du.cpu().numpy(), ref))
                ok = max(errs) <= 1e-12
                bad += not ok
                print(f"U={u_} NST={nst} WPS={wps or 'auto'} packed={packed}: max err {max(errs):.2e} {'ok' if ok else 'FAIL'}", flush=True)
print("failures:", bad)
sys.exit(1 if bad else 0)
