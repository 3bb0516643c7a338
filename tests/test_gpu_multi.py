"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the vertex-partitioned RHS over NCCL equals the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, exchange, kernel, fused_rk4=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ND_B200_KERNEL=kernel)
    import torch
    import torch.distributed as dist
    import ndb200 as nd
    from networkdynamics_jl_b200.distributed import PartitionedNetwork
    from helpers import condition_params
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    L = nd.Lib
    n = 20000
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    vm = ([L.kuramoto_first(), L.kuramoto_second()], np.random.default_rng(4).permutation(half))
    g = nd.barabasi_albert(n, 4, seed=2)
    pn = PartitionedNetwork(g, vm, L.kuramoto_edge(), rank=rank, world=world, exchange=exchange)
    assert pn.exchange_kind == exchange
    u0 = np.random.default_rng(1).random(pn.dim())
    p = condition_params(pn.nw, np.random.default_rng(2).random(pn.pdim()))
    u = torch.full((pn.dim(),), float("nan"), dtype=torch.float64, device="cuda")
    for a, b in pn.owned_segments:
        u[a:b] = torch.from_numpy(u0[a:b]).cuda()
    pd = torch.from_numpy(p).cuda()
    du = torch.full_like(u, float("nan"))
    pn.rhs(du, u, pd, 0.0)
    pn.exchange(du)                       # collect everybody's rows for the comparison
    u2 = u.clone()
    work = {}
    for s in range(5):
        pn.rk4_step(u, pd, s * 1e-3, 1e-3, work)
    if fused_rk4:
        pn.rk4(u2, pd, 0.0, 1e-3, 5)      # p2p: nd_b200_rk4_exchange (stage updates fused into the exchanging kernels)
        for a, b in pn.owned_segments:
            assert torch.allclose(u2[a:b], u[a:b], rtol=1e-13, atol=1e-15), "fused multi-GPU RK4 differs from host-driven stages"
    pn.exchange(u)
    torch.cuda.synchronize()
    assert not pn.comm_timed_out()
    pn.close()
    if rank == 0:
        np.save(os.path.join(out_dir, "du.npy"), du.cpu().numpy())
        np.save(os.path.join(out_dir, "u.npy"), u.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange,kernel", [("p2p", "fused"), ("p2p", "jag"), ("nccl", "fused")])
def test_two_gpu_partitioned_rhs(nd, cuda, tmp_path, exchange, kernel):
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from helpers import condition_params, floored_rel_err, null_aggregator, oracle_network
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), exchange, kernel), nprocs=2, join=True)
    L = nd.Lib
    n = 20000
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    vm = ([L.kuramoto_first(), L.kuramoto_second()], np.random.default_rng(4).permutation(half))
    g = nd.barabasi_albert(n, 4, seed=2)
    nw = nd.Network(g, vm, L.kuramoto_edge(), aggregator=null_aggregator)
    onw = oracle_network(g, vm, L.kuramoto_edge())
    u0 = np.random.default_rng(1).random(nw.dim())
    p = condition_params(nw, np.random.default_rng(2).random(nw.pdim()))
    assert floored_rel_err(np.load(tmp_path / "du.npy"), onw.rhs(u0, p)) <= 1e-12
    assert floored_rel_err(np.load(tmp_path / "u.npy"), onw.rk4(u0, p, 0.0, 1e-3, 5)) <= 1e-12
