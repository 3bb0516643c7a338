#!/usr/bin/env python
"""One GPU, the full BASELINE configs[4] graph (Kuramoto on Erdos-Renyi 5e7 vertices / 4e8 edges) built straight from the edge
list: a few RHS evaluations for `ncu` (tools/gpu_round_r2_45.sh).  python tools/profile_cfg5_full.py [nv ne]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ndb200 as nd

nv, ne = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50_000_000, 400_000_000)
t0 = time.time()
g = nd.erdos_renyi(nv, ne, seed=1)
t1 = time.time()
nw = nd.Network.from_edgelist(g, nd.Lib.kuramoto_first(), nd.Lib.kuramoto_edge())
t2 = time.time()
u = torch.from_numpy(np.random.default_rng(1).random(nw.dim())).cuda()
p = torch.from_numpy(np.random.default_rng(2).random(nw.pdim())).cuda()
du = torch.empty_like(u)
for _ in range(6):          # calls 1-2 live parameters, from call 3 on the packed copy (edge_parameters="auto")
    nw(du, u, p, 0.0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    nw(du, u, p, 0.0)
b.record(); torch.cuda.synchronize()
print(f"graph {t1 - t0:.1f} s, engine {t2 - t1:.1f} s, kernel {nw.kernel_name()}, {a.elapsed_time(b) / 3:.3f} ms per RHS", flush=True)
if os.environ.get("ND_PROFILE_L2_GRANULARITY"):
    # cudaLimitMaxL2FetchGranularity (0x05): how many bytes an L2 miss fetches from DRAM (a hint; 32 / 64 / 128)
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    for gran in (128, 64, 32, 128):
        cur = ctypes.c_size_t(0)
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))
        rt.cudaDeviceGetLimit(ctypes.byref(cur), 5)
        for _ in range(2):
            nw(du, u, p, 0.0)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            nw(du, u, p, 0.0)
        b.record(); torch.cuda.synchronize()
        print(f"cudaLimitMaxL2FetchGranularity {gran} (rc {rc}, reads back {cur.value}): {a.elapsed_time(b) / 5:.3f} ms per RHS", flush=True)
