"""2 GPUs, edges WITH states: row-partitioned engines with the all-gather exchange (NCCL) -- every rank evaluates its rows and
a contiguous chunk of each stateful edge batch; the result equals the oracle.  Skipped with fewer than 2 GPUs; the same
flow runs on 2 gloo ranks over the emulator in tests/test_distributed_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(nd):
    L = nd.Lib
    rng = np.random.default_rng(3)
    g = nd.barabasi_albert(20000, 3, seed=2)
    vm = ([L.kuramoto_first(), L.diffusion_vertex()], rng.integers(0, 2, g.nv))
    em = ([L.relax_odeedge(), L.kuramoto_edge(), L.diffusion_odeedge()], rng.integers(0, 3, g.ne))
    return g, vm, em


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import ndb200 as nd
    from networkdynamics_jl_b200.distributed import PartitionedNetwork
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g, vm, em = _case(nd)
    pn = PartitionedNetwork(g, vm, em, rank=rank, world=world, exchange="auto")
    assert pn.exchange_kind == "nccl", "the packed NVLink halo does not carry edge states"
    u0 = np.random.default_rng(1).random(pn.dim())
    p = 0.5 + np.random.default_rng(2).random(pn.pdim())
    u = torch.full((pn.dim(),), float("nan"), dtype=torch.float64, device="cuda")
    for a, b in pn.owned_segments:
        u[a:b] = torch.from_numpy(u0[a:b]).cuda()
    pd = torch.from_numpy(p).cuda()
    du = torch.full_like(u, float("nan"))
    pn.rhs(du, u, pd, 0.0)
    pn.exchange(du)
    pn.rk4(u, pd, 0.0, 1e-3, 5)
    pn.exchange(u)
    torch.cuda.synchronize()
    pn.close()
    if rank == 0:
        np.save(os.path.join(out_dir, "du.npy"), du.cpu().numpy())
        np.save(os.path.join(out_dir, "u.npy"), u.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_edges_with_states(nd, cuda, tmp_path):
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from helpers import floored_rel_err, oracle_network
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g, vm, em = _case(nd)
    onw = oracle_network(g, vm, em)
    u0 = np.random.default_rng(1).random(onw.lastidx_dynamic)
    p = 0.5 + np.random.default_rng(2).random(onw.lastidx_p)
    du = np.load(tmp_path / "du.npy")
    assert not np.isnan(du).any(), "the ranks' segments tile the state vector"
    assert floored_rel_err(du, onw.rhs(u0, p)) <= 1e-12
    assert floored_rel_err(np.load(tmp_path / "u.npy"), onw.rk4(u0, p, 0.0, 1e-3, 5)) <= 1e-12
