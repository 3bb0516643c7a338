"""N>1 host logic on CPU: world_size-2 gloo run of the vertex partition + state exchange.
The device kernel is replaced by the oracle restricted to the owned rows, so what is tested is exactly the multi-rank
plumbing the GPU path uses: partition_rows / state_segments / exchange_states and the stage algebra of rk4_step."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(nd):
    L = nd.Lib
    n = 400
    rng = np.random.default_rng(4)
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    g = nd.barabasi_albert(n, 4, seed=2)
    return g, ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(half)), L.kuramoto_edge()


def test_partition_and_segments(nd):
    from networkdynamics_jl_b200 import distributed as D
    from helpers import null_aggregator
    g, vm, em = _case(nd)
    nw = nd.Network(g, vm, em, aggregator=null_aggregator)
    cnt = D.row_entry_counts(nw.im, nw.layer.edgebatches)
    assert cnt.sum() == 2 * g.ne
    deg = np.bincount(np.concatenate([g.src, g.dst]) - 1, minlength=g.nv)
    assert np.array_equal(cnt[D.row_of_vertex(nw.im)], deg)
    for world in (1, 2, 3, 8):
        rr = D.partition_rows(cnt, world)
        assert rr[0][0] == 0 and rr[-1][1] == g.nv and all(rr[i][1] == rr[i + 1][0] for i in range(world - 1))
        loads = [cnt[a:b].sum() + (b - a) for a, b in rr]
        assert max(loads) - min(loads) <= cnt.max() + 2          # balanced up to one row
        segs = [D.state_segments(nw.vertexbatches, a, b) for a, b in rr]
        owner = np.full(nw.dim(), -1)
        for r, ss in enumerate(segs):
            for a, b in ss:
                assert np.all(owner[a:b] == -1)
                owner[a:b] = r
        assert np.all(owner >= 0)                                   # every state has exactly one owner
        rov = D.row_of_vertex(nw.im)
        for v in range(g.nv):                                       # the owner of a state is the owner of its row
            r = next(i for i, (a, b) in enumerate(rr) if a <= rov[v] < b)
            assert owner[nw.im.v_data[v] - 1] == r


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import ndb200 as nd
    from networkdynamics_jl_b200 import distributed as D
    from helpers import condition_params, null_aggregator, oracle_network
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, vm, em = _case(nd)
    nw = nd.Network(g, vm, em, aggregator=null_aggregator)
    onw = oracle_network(g, vm, em)
    cnt = D.row_entry_counts(nw.im, nw.layer.edgebatches)
    rr = D.partition_rows(cnt, world)
    segs = [D.state_segments(nw.vertexbatches, a, b) for a, b in rr]
    u0 = np.random.default_rng(1).random(nw.dim())
    p = condition_params(nw, np.random.default_rng(2).random(nw.pdim()))

    class FakeLocal(D.PartitionedNetwork):       # same plumbing, oracle instead of the device kernel
        def __init__(self):
            self.rank, self.world, self.group, self.segments, self.row_ranges = rank, world, None, segs, rr

        def rhs(self, du, u, p_, t, *, exchange=True):
            if exchange:
                self.exchange(u)
            full = onw.rhs(u.numpy(), p_)
            du.fill_(float("nan"))
            for a, b in self.owned_segments:
                du[a:b] = torch.from_numpy(full[a:b])

    pn = FakeLocal()
    # (1) one RHS: start from a vector where only the owned states are valid
    u = torch.full((nw.dim(),), float("nan"), dtype=torch.float64)
    for a, b in pn.owned_segments:
        u[a:b] = torch.from_numpy(u0[a:b])
    du = torch.empty_like(u)
    pn.rhs(du, u, p, 0.0)
    assert np.array_equal(u.numpy(), u0)                      # exchange rebuilt the full state everywhere
    ref = onw.rhs(u0, p)
    for a, b in pn.owned_segments:
        assert np.array_equal(du[a:b].numpy(), ref[a:b])
    # (2) ten RK4 steps
    work = {}
    for s in range(10):
        pn.rk4_step(u, p, s * 1e-3, 1e-3, work)
    pn.exchange(u)
    np.save(os.path.join(out_dir, f"u_rank{rank}.npy"), u.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_rhs_and_rk4(nd, tmp_path):
    import torch.multiprocessing as mp
    from helpers import condition_params, null_aggregator, oracle_network
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g, vm, em = _case(nd)
    nw = nd.Network(g, vm, em, aggregator=null_aggregator)
    onw = oracle_network(g, vm, em)
    u0 = np.random.default_rng(1).random(nw.dim())
    p = condition_params(nw, np.random.default_rng(2).random(nw.pdim()))
    ref = onw.rk4(u0, p, 0.0, 1e-3, 10)
    got = [np.load(tmp_path / f"u_rank{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1])
    assert np.max(np.abs(got[0] - ref)) <= 1e-14 * max(1.0, np.max(np.abs(ref)))


def _worker_uniform(rank, world, port, out_dir):
    """single vertex batch, equal ranges -> the exchange is ONE in-place all_gather_into_tensor"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import ndb200 as nd
    from networkdynamics_jl_b200 import distributed as D
    from helpers import null_aggregator
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = nd.erdos_renyi(600, 2400, seed=3)
    nw = nd.Network(g, nd.Lib.kuramoto_first(), nd.Lib.kuramoto_edge(), aggregator=null_aggregator)
    cnt = D.row_entry_counts(nw.im, nw.layer.edgebatches)
    rr = D.partition_rows(cnt, world, prefer_equal_rows=0.5)
    segs = [D.state_segments(nw.vertexbatches, a, b) for a, b in rr]
    assert rr == [(r * 200, (r + 1) * 200) for r in range(world)] and D._uniform_allgather_layout(segs, nw.dim())
    u = torch.full((nw.dim(),), float("nan"), dtype=torch.float64)
    u[rr[rank][0]:rr[rank][1]] = torch.arange(rr[rank][0], rr[rank][1], dtype=torch.float64)
    D.exchange_states(u, segs)
    assert torch.equal(u, torch.arange(nw.dim(), dtype=torch.float64))
    dist.barrier()
    dist.destroy_process_group()


def test_three_rank_gloo_allgather_exchange(nd, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker_uniform, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)


def _worker_shared_graph(rank, world, port, out_dir):
    """bench.py's shared graph generation: rank 0 generates, the others map the edge arrays from shared memory; the
    edge-list based partition planning runs on the mapped arrays"""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import ndb200 as nd
    import bench
    from networkdynamics_jl_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def make():
        calls.append(1)
        return nd.erdos_renyi(4000, 16000, seed=1)
    g = bench.shared_graph(nd, make, "gloo_test", rank, world, dist.barrier)
    ref = nd.erdos_renyi(4000, 16000, seed=1)
    assert g.nv == ref.nv and np.array_equal(g.src, ref.src) and np.array_equal(g.dst, ref.dst)
    assert len(calls) == (1 if rank == 0 else 0)
    probe = nd.Network.from_edgelist(g, nd.Lib.kuramoto_first(), nd.Lib.kuramoto_edge(), layout_only=True)
    rr = D.partition_rows(D.row_entry_counts(probe.im, probe.layer.edgebatches), world)
    plan = D.halo_plan(probe.im, probe.layer.edgebatches, rr, rank)
    assert plan["halo_lens"][rank] > 0 and plan["gather_len"] == probe.dim() + plan["halo_lens"][rank]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shared_graph_generation(nd, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker_shared_graph, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("ndb200_gloo_test")] if os.path.isdir("/dev/shm") else True


def _worker_partitioned_class(rank, world, port, out_dir):
    """the real PartitionedNetwork class (all-gather exchange over gloo) with every rank's engine running on the CPU
    emulation of the kernels (tests/cusim): table-driven and edge-list construction, rhs, host-driven RK4"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import cusim
    import ndb200 as nd
    from networkdynamics_jl_b200.distributed import PartitionedNetwork
    from helpers import oracle_network
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = nd.Lib
    g = nd.erdos_renyi(1500, 6000, seed=5)
    vm, em = L.kuramoto_second(), L.kuramoto_edge()
    onw = oracle_network(g, vm, em)
    u0 = np.random.default_rng(1).random(onw.lastidx_dynamic)
    p = 0.5 + np.random.default_rng(2).random(onw.lastidx_p)
    with cusim.use():
        for from_edgelist in (False, True):
            pn = PartitionedNetwork(g, vm, em, rank=rank, world=world, exchange="nccl", from_edgelist=from_edgelist)
            assert pn.exchange_kind == "nccl" and pn.comm is None and (pn.dim(), pn.pdim()) == (u0.size, p.size)
            u = torch.full((pn.dim(),), float("nan"), dtype=torch.float64)
            for a, b in pn.owned_segments:
                u[a:b] = torch.from_numpy(u0[a:b])
            pt = torch.from_numpy(p)
            du = torch.full_like(u, float("nan"))
            pn.rhs(du, u, pt, 0.0)
            ref = onw.rhs(u0, p)
            for a, b in pn.owned_segments:
                assert np.array_equal(du[a:b].numpy(), ref[a:b]), from_edgelist
            pn.rk4(u, pt, 0.0, 1e-3, 3)
            pn.exchange(u)
            assert np.max(np.abs(u.numpy() - onw.rk4(u0, p, 0.0, 1e-3, 3))) <= 1e-14
            assert not pn.comm_timed_out()
            pn.close()
            # local parameters: the rank's engine on the subgraph of the edges incident to its rows; its
            # parameter vector = vertex parameters + those edges' parameters; same du, bit for bit
            pl = PartitionedNetwork(g, vm, em, rank=rank, world=world, exchange="nccl", from_edgelist=from_edgelist, local_parameters=True)
            assert pl.local_parameters and pl.global_pdim == p.size and pl.pdim() < p.size
            ploc = pl.localize_parameters(p)
            assert ploc.size == pl.pdim() and np.array_equal(ploc[:g.nv * vm.pdim], p[:g.nv * vm.pdim])
            u = torch.full((pl.dim(),), float("nan"), dtype=torch.float64)
            for a, b in pl.owned_segments:
                u[a:b] = torch.from_numpy(u0[a:b])
            # parameters outside parameter_segments() are never read: poison them
            pmask = np.zeros(ploc.size, dtype=bool)
            for a, b in pl.parameter_segments():
                pmask[a:b] = True
            ploc = np.where(pmask, ploc, np.nan)
            du = torch.full_like(u, float("nan"))
            pl.rhs(du, u, torch.from_numpy(ploc), 0.0)
            for a, b in pl.owned_segments:
                assert np.array_equal(du[a:b].numpy(), ref[a:b]), ("local parameters", from_edgelist)
            pl.close()
        # edges WITH states: every rank evaluates f for a contiguous chunk of each stateful batch; the chunks' states travel
        # with the all-gather like vertex states (the packed NVLink halo carries vertex outputs only and is refused)
        rng = np.random.default_rng(3)
        g2 = nd.barabasi_albert(900, 3, seed=2)
        vm2 = ([L.kuramoto_first(), L.diffusion_vertex()], rng.integers(0, 2, g2.nv))
        em2 = ([L.relax_odeedge(), L.kuramoto_edge(), L.diffusion_odeedge()], rng.integers(0, 3, g2.ne))
        onw2 = oracle_network(g2, vm2, em2)
        u0 = rng.random(onw2.lastidx_dynamic)
        p = 0.5 + rng.random(onw2.lastidx_p)
        try:
            PartitionedNetwork(g2, vm2, em2, rank=rank, world=world, exchange="p2p")
            raise AssertionError("p2p exchange accepted edges with states")
        except RuntimeError:
            pass
        pn = PartitionedNetwork(g2, vm2, em2, rank=rank, world=world, exchange="auto")
        assert pn.exchange_kind == "nccl"
        owned = np.zeros(u0.size, dtype=np.int64)
        for a, b in pn.owned_segments:
            owned[a:b] += 1
        cover = torch.from_numpy(owned.copy())
        dist.all_reduce(cover)
        assert np.all(cover.numpy() == 1), "the ranks' segments tile the state vector"
        assert owned[g2.nv:].any(), "this rank owns some edge states"
        u = torch.full((pn.dim(),), float("nan"), dtype=torch.float64)
        for a, b in pn.owned_segments:
            u[a:b] = torch.from_numpy(u0[a:b])
        pt = torch.from_numpy(p)
        du = torch.full_like(u, float("nan"))
        pn.rhs(du, u, pt, 0.0)
        ref = onw2.rhs(u0, p)
        assert np.max(np.abs(du.numpy()[owned == 1] - ref[owned == 1])) <= 1e-13
        pn.rk4(u, pt, 0.0, 1e-3, 3)
        pn.exchange(u)
        assert np.max(np.abs(u.numpy() - onw2.rk4(u0, p, 0.0, 1e-3, 3))) <= 1e-13
        pn.close()
        # LoopbackConnections: a ring of hubs, each with an injector leaf behind a loopback edge; the injector's row may
        # live on another rank than its hub -- it reads the hub's output from the all-gathered u
        m = 200
        gs = np.concatenate([np.arange(1, m + 1), np.arange(m + 1, 2 * m + 1)])
        gd = np.concatenate([np.roll(np.arange(1, m + 1), -1), np.arange(1, m + 1)])
        g3 = nd.SimpleDiGraph(2 * m, gs, gd)
        vm3 = [L.kuramoto_first()] * m + [L.kuramoto_second()] * m
        em3 = ([L.kuramoto_edge(), L.loopback()], (g3.src > m).astype(np.int64))
        onw3 = oracle_network(g3, vm3, em3)
        u0 = rng.random(onw3.lastidx_dynamic)
        p = 0.5 + rng.random(onw3.lastidx_p)
        pn = PartitionedNetwork(g3, vm3, em3, rank=rank, world=world, exchange="auto")
        assert pn.exchange_kind == "nccl"
        u = torch.full((pn.dim(),), float("nan"), dtype=torch.float64)
        for a, b in pn.owned_segments:
            u[a:b] = torch.from_numpy(u0[a:b])
        pt = torch.from_numpy(p)
        du = torch.full_like(u, float("nan"))
        pn.rhs(du, u, pt, 0.0)
        ref = onw3.rhs(u0, p)
        for a, b in pn.owned_segments:
            assert np.max(np.abs(du[a:b].numpy() - ref[a:b])) <= 1e-13
        pn.rk4(u, pt, 0.0, 1e-3, 3)
        pn.exchange(u)
        assert np.max(np.abs(u.numpy() - onw3.rk4(u0, p, 0.0, 1e-3, 3))) <= 1e-13
        pn.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partitioned_network_class_on_the_emulator(nd, tmp_path):
    import torch.multiprocessing as mp
    import cusim
    cusim.build()                      # once, before the ranks race to build it
    mp.spawn(_worker_partitioned_class, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
