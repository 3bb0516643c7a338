#!/usr/bin/env python
"""One GPU, Kuramoto on an Erdos-Renyi graph whose vertex outputs exceed the L2 (default 2e7 vertices / 1.6e8 edges = 160 MB of
outputs), built from the edge list: the single-pass RHS against the column-blocked one (ND_B200_L2_BLOCKS).
python tools/profile_l2_blocks.py [nv ne] [k ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ndb200 as nd

args = [a for a in sys.argv[1:]]
nv, ne = (int(args[0]), int(args[1])) if len(args) >= 2 else (20_000_000, 160_000_000)
ks = args[2:] or ["auto"]
t0 = time.time()
g = nd.erdos_renyi(nv, ne, seed=1)
print(f"graph {time.time() - t0:.1f} s", flush=True)
u = torch.from_numpy(np.random.default_rng(1).random(nv)).cuda()
p = torch.from_numpy(np.random.default_rng(2).random(nv + ne)).cuda()
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
ref = None
for k in [None] + ks:
    os.environ.pop("ND_B200_L2_BLOCKS", None)
    if k is not None:
        os.environ["ND_B200_L2_BLOCKS"] = k
    t1 = time.time()
    nw = nd.Network.from_edgelist(g, nd.Lib.kuramoto_first(), nd.Lib.kuramoto_edge())
    tb = time.time() - t1
    du = torch.full_like(u, float("nan"))
    for _ in range(5):          # from call 3 on the packed parameter copies (edge_parameters="auto")
        nw(du, u, p, 0.0)
    torch.cuda.synchronize()
    n0 = nw.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        flush.zero_(); a.record(); nw(du, u, p, 0.0); b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    launches = (nw.launch_count() - n0) / 5
    if ref is None:
        ref = du.clone()
        same = "reference"
    else:
        same = "bit-identical to the single pass" if torch.equal(ref, du) else f"max abs diff {float((ref - du).abs().max()):.3e}"
    print(f"ND_B200_L2_BLOCKS={k}: engine {tb:.1f} s, {launches:.0f} launches per RHS, {ms:.3f} ms per RHS (L2 flushed), {same}", flush=True)
    del nw
