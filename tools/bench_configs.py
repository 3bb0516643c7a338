#!/usr/bin/env python
"""Time the RHS (and RK4) on all BASELINE.json configs that fit one GPU; one JSON line per (config, mode).
   python tools/bench_configs.py [cfg1 cfg2 cfg2nop cfg3 cfg4 cfg5s] [--check] [--quick]
                                 [--modes=name:ENV=VAL,ENV=VAL;name2:...]     engine env switches per mode"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ndb200 as nd
from helpers import condition_params, floored_rel_err, oracle_network

L = nd.Lib

def make(name):
    rng = np.random.default_rng(11)
    if name == "cfg1": return nd.watts_strogatz(10_000, 10, 0.1, seed=1), L.kuramoto_first(), L.kuramoto_edge()
    if name == "cfg2": return nd.erdos_renyi(1_000_000, 4_000_000, seed=1), L.diffusion_vertex(), L.diffusion_edge()
    if name == "cfg2nop": return nd.erdos_renyi(1_000_000, 4_000_000, seed=1), L.diffusion_vertex(), L.diffusion_edge_nop()
    if name == "cfg2kura": return nd.erdos_renyi(1_000_000, 4_000_000, seed=1), L.kuramoto_first(), L.kuramoto_edge()
    if name == "cfg3":
        n = 1_000_000
        half = np.array([0] * (n // 2) + [1] * (n // 2))
        return nd.barabasi_albert(n, 4, seed=1), ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(half)), L.kuramoto_edge()
    if name == "cfg4": return nd.grid_graph(400, 500), L.swing_dq(), L.line_dq()
    if name == "cfg5s": return nd.erdos_renyi(5_000_000, 40_000_000, seed=1), L.kuramoto_first(), L.kuramoto_edge()
    raise SystemExit(name)

def parse_modes():
    for a in sys.argv[1:]:
        if a.startswith("--modes="):
            out = []
            for m in a[len("--modes="):].split(";"):
                name, _, envs = m.partition(":")
                out.append((name, dict(kv.split("=") for kv in envs.split(",") if kv)))
            return out
    return [("default", {})]


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg1", "cfg2", "cfg2nop", "cfg3", "cfg4"]
    check = "--check" in sys.argv
    quick = "--quick" in sys.argv
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for name in names:
        g, vm, em = make(name)
        onw = oracle_network(g, vm, em) if check else None
        for mode, envs in parse_modes():
            for k in [k for k in os.environ if k.startswith("ND_B200_") and k != "ND_B200_LIB"]:
                del os.environ[k]
            os.environ.update(envs)
            t0 = time.time()
            nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", keep_tables=False))
            tb = time.time() - t0
            sz = nw.engine_sizes()
            u_h = np.random.default_rng(1).random(nw.dim())
            p_h = condition_params(nw, np.random.default_rng(2).random(nw.pdim()))
            u, p = torch.from_numpy(u_h).cuda(), torch.from_numpy(p_h).cuda()
            du = torch.empty_like(u)
            if os.environ.get("ND_B200_PACK_P") == "1":     # tool-level switch: evaluate from the engine's packed edge parameters
                try:
                    nw.pack_params(p)
                except nd.ArgumentError as ex:
                    print(json.dumps({"config": name, "mode": mode, "skipped": str(ex)}), flush=True)
                    continue
            for _ in range(5 if quick else 50): nw(du, u, p, 0.0)
            torch.cuda.synchronize()
            K = 5 if quick else 100
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            for a, b in ev:
                flush.zero_(); a.record(); nw(du, u, p, 0.0); b.record()
            torch.cuda.synchronize()
            cold = np.array([a.elapsed_time(b) for a, b in ev])
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(K): nw(du, u, p, 0.0)
            b.record(); torch.cuda.synchronize()
            warm = a.elapsed_time(b) / K
            b_alg = 16 * nw.dim() + 8 * nw.pdim() + 4 * (g.nv + 1) + 4 * sz["nentries"]
            line = {"config": name, "mode": mode, "nv": g.nv, "ne": g.ne, "entries": sz["nentries"], "blocks": sz["nblocks"], "long_rows": sz["n_long_rows"],
                    "rhs_us_cold_mean": float(cold.mean() * 1e3), "rhs_us_cold_min": float(cold.min() * 1e3), "rhs_us_warm": warm * 1e3,
                    "b_alg_MB": b_alg / 1e6, "GBs_cold": b_alg / cold.mean() / 1e6, "GBs_warm": b_alg / warm / 1e6,
                    "edge_evals_per_s_cold": g.ne / (cold.mean() * 1e-3), "build_s": round(tb, 2)}
            if not quick:   # RK4: 200 steps
                ur = u.clone()
                nw.rk4(ur, p, 0.0, 1e-3, 8); torch.cuda.synchronize()
                a.record(); nw.rk4(ur, p, 0.0, 1e-3, 200); b.record(); torch.cuda.synchronize()
                line["rk4_us_per_step"] = a.elapsed_time(b) / 200 * 1e3
            if check:
                line["parity"] = floored_rel_err(du.cpu().numpy(), onw.rhs(u_h, p_h, threads=8))
            print(json.dumps(line), flush=True)
            del nw


if __name__ == "__main__":
    main()
