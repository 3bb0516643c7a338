"""The NVLink peer-memory halo exchange (nd_b200_comm_* + nd_b200_rhs_exchange: publishing blocks, arrival flags,
interior-first tile order, double-buffered halo) with 2..8 EMULATED ranks in the CPU suite.

Each rank is a Python thread that owns an engine + comm object of the emulated library (tests/cusim: the product's CUDA
sources on a CPU SIMT emulator, CUDA IPC handles = raw pointers inside one process) and calls nd_b200_rhs_exchange
concurrently with the others (ctypes releases the GIL), so the publish / wait protocol of the HALO kernel variants runs
under real races: ranks drift apart by up to one call, halo buffers alternate by sequence parity, flags are awaited
with acquire loads.  Every rank advances its owned states with an explicit Euler step between calls; after K steps the
states must equal K Euler steps of the sequential oracle bit for bit -- any stale, early or torn halo read shows up.
The real two-GPU run of the same path is tests/test_gpu_multi.py.
"""
import ctypes as C
import threading

import numpy as np
import pytest

from helpers import condition_params, null_aggregator, oracle_network


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _run_world(nd, g, vm, em, world, ncalls, h=0.01, pack=False, rk4=None, edgelist=False):
    import cusim
    from networkdynamics_jl_b200 import distributed as D
    with cusim.use() as L:
        probe = nd.Network.from_edgelist(g, vm, em, layout_only=True) if edgelist else nd.Network(g, vm, em, aggregator=null_aggregator)
        rr = D.partition_rows(D.row_entry_counts(probe.im, probe.layer.edgebatches), world)
        segs = [D.state_segments(probe.vertexbatches, a, b) for a, b in rr]
        plans = [D.halo_plan(probe.im, probe.layer.edgebatches, rr, r) for r in range(world)]
        if edgelist:     # partition, halo plan and engines straight from the bare edge list (no per-component tables)
            nws = [nd.Network.from_edgelist(g, vm, em, device=r, row_range=rr[r], gather_offset=plans[r]["gather_offset"],
                                            gather_len=plans[r]["gather_len"]) for r in range(world)]
        else:
            nws = [nd.Network(g, vm, em, aggregator=nd.B200Aggregator(
                "+", device=r, row_range=rr[r], keep_tables=False, gather_offset=plans[r]["gather_offset"],
                gather_len=plans[r]["gather_len"])) for r in range(world)]
        comms, handles = [], []
        for r in range(world):
            c = C.c_void_p()
            assert L.nd_b200_comm_create(r, r, world, plans[r]["halo_lens"][r], max(plans[r]["halo_lens"]), C.byref(c)) == 0
            hb = C.create_string_buffer(nd._cabi.IPC_HANDLE_BYTES)
            assert L.nd_b200_comm_export(c, hb) == 0
            for peer, (offs, start) in plans[r]["sends"].items():
                offs = np.ascontiguousarray(offs, dtype=np.int64)
                assert L.nd_b200_comm_set_send(c, peer, offs.ctypes.data_as(nd._cabi.i64p), offs.size, start) == 0
            comms.append(c)
            handles.append(hb)
        for r in range(world):
            for q in range(world):
                assert L.nd_b200_comm_open_peer(comms[r], q, handles[q].raw) == 0
        n = probe.dim()
        u0 = np.random.default_rng(1).random(n)
        p = condition_params(probe, np.random.default_rng(2).random(probe.pdim()))
        us = []
        for r in range(world):
            u = np.full(n, np.nan)                 # only the owned states are valid on a rank
            for a, b in segs[r]:
                u[a:b] = u0[a:b]
            us.append(u)
        errors = []
        if pack:     # per-rank packed copy of the edge parameters (HALO + PK kernel variants)
            for r in range(world):
                assert L.nd_b200_pack_params(nws[r].handle, _ptr(p), None) == 0, L.nd_b200_last_error(nws[r].handle).decode()

        def rank_main(r):
            try:
                if rk4 is not None:        # fused multi-rank RK4: four exchanging launches per step, nothing else
                    rc = L.nd_b200_rk4_exchange(nws[r].handle, comms[r], _ptr(us[r]), _ptr(p) if p.size else None, 0.0, rk4, ncalls, None)
                    if rc:
                        raise RuntimeError(L.nd_b200_last_error(nws[r].handle).decode())
                    return
                du = np.full(n, np.nan)
                for k in range(ncalls):
                    rc = L.nd_b200_rhs_exchange(nws[r].handle, comms[r], _ptr(du), _ptr(us[r]), _ptr(p) if p.size else None, 0.0, None)
                    if rc:
                        raise RuntimeError(L.nd_b200_last_error(nws[r].handle).decode())
                    for a, b in segs[r]:
                        us[r][a:b] = us[r][a:b] + h * du[a:b]
            except Exception as e:      # noqa: BLE001 -- reported by the main thread
                errors.append((r, repr(e)))

        th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=300)
        assert not any(t.is_alive() for t in th), "emulated ranks hung"
        assert not errors, errors
        for r in range(world):
            v = C.c_int32(0)
            assert L.nd_b200_comm_status(comms[r], C.byref(v)) == 0 and v.value == 0, f"rank {r} timed out waiting for a peer"
        out = np.full(n, np.nan)
        for r in range(world):
            for a, b in segs[r]:
                out[a:b] = us[r][a:b]
        sizes = [nw.engine_sizes() for nw in nws]
        kernel = nws[0].kernel_name()
        for c in comms:
            L.nd_b200_comm_destroy(c)
        del nws
    onw = oracle_network(g, vm, em)
    if rk4 is not None:
        return out, onw.rk4(u0, p, 0.0, rk4, ncalls), plans, sizes, kernel
    ref = u0.copy()
    for _ in range(ncalls):
        ref = ref + h * onw.rhs(ref, p)
    return out, ref, plans, sizes, kernel


def _cases(nd):
    L = nd.Lib
    n = 3000
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    return {
        # no locality: nearly every row reads remote outputs, every rank needs most of every peer's outputs
        "er_diffusion": (nd.erdos_renyi(4000, 16000, seed=1), L.diffusion_vertex(), L.diffusion_edge()),
        # locality: the halo is one lattice row per neighbour, most tiles are interior and run before the halo arrives
        "grid_kuramoto": (nd.grid_graph(40, 120), L.kuramoto_first(), L.kuramoto_edge()),
        # two vertex batches (two owned state ranges per rank), hubs above the long-row threshold
        "ba_mixed": (nd.barabasi_albert(n, 4, seed=2), ([L.kuramoto_first(), L.kuramoto_second()], np.random.default_rng(4).permutation(half)), L.kuramoto_edge()),
    }


@pytest.mark.parametrize("kernel", ["fused", "jag"])
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["er_diffusion", "grid_kuramoto", "ba_mixed"])
def test_emulated_ranks_match_sequential_oracle(nd, monkeypatch, name, world, kernel):
    monkeypatch.setenv("ND_B200_KERNEL", kernel)
    g, vm, em = _cases(nd)[name]
    out, ref, plans, sizes, kname = _run_world(nd, g, vm, em, world, ncalls=12)
    assert kname == ("rhs_jag_kernel" if kernel == "jag" else "rhs_fused_kernel")
    assert not np.isnan(out).any()
    if name == "ba_mixed":
        # hub rows are reduced by a block tree / by lane parts: same terms, other association
        assert np.max(np.abs(out - ref)) <= 1e-13 * np.max(np.abs(ref))
    else:
        assert np.array_equal(out, ref)
    # the plan is what the survey asks for: only boundary outputs travel
    if name == "grid_kuramoto":
        assert max(max(pl["halo_lens"]) for pl in plans) <= 2 * 48       # 40 outputs per neighbour, blocks padded to 16


@pytest.mark.parametrize("kernel", ["fused", "jag"])
@pytest.mark.parametrize("name", ["er_diffusion", "grid_kuramoto"])
def test_emulated_ranks_with_packed_edge_parameters(nd, monkeypatch, name, kernel):
    """the same protocol with every rank evaluating from its packed per-entry copy of the edge parameters
    (nd_b200_pack_params on a halo engine)"""
    monkeypatch.setenv("ND_B200_KERNEL", kernel)
    g, vm, em = _cases(nd)[name]
    out, ref, _plans, _sizes, _k = _run_world(nd, g, vm, em, 3, ncalls=6, pack=True)
    assert np.array_equal(out, ref)


def test_locality_ordering_shrinks_the_halo(nd, monkeypatch):
    """ranks own contiguous ranges of a LOCALITY-ORDERED graph.  A lattice whose vertices carry random labels
    has no locality in its ids: every rank needs most of every peer's outputs.  nd.locality_order (reverse Cuthill-McKee)
    + nd.permute_graph relabel it; per-vertex / per-edge data follow the returned permutations; the halo shrinks by an order
    of magnitude and the emulated ranks still reproduce the oracle of the ORIGINAL network (same terms per vertex, summed in
    the order of the new ids)."""
    from networkdynamics_jl_b200 import distributed as D
    monkeypatch.setenv("ND_B200_KERNEL", "jag")
    L = nd.Lib
    rng = np.random.default_rng(3)
    g0 = nd.grid_graph(40, 60)
    shuffle = rng.permutation(g0.nv)
    g, _ = nd.permute_graph(g0, shuffle)                       # the "application" graph: a lattice with scrambled labels
    order = nd.locality_order(g)
    g2, edge_order = nd.permute_graph(g, order)
    world = 4

    def halo(gr):
        probe = nd.Network(gr, L.kuramoto_first(), L.kuramoto_edge(), aggregator=null_aggregator)
        rr = D.partition_rows(D.row_entry_counts(probe.im, probe.layer.edgebatches), world)
        return max(D.halo_plan(probe.im, probe.layer.edgebatches, rr, 0)["halo_lens"])
    h_scrambled, h_ordered = halo(g), halo(g2)
    assert h_ordered * 8 <= h_scrambled, (h_scrambled, h_ordered)
    # same dynamics: u, vertex parameters and edge parameters carried over with the permutations
    onw = oracle_network(g, L.kuramoto_first(), L.kuramoto_edge())
    u = rng.random(g.nv)
    omega, K = rng.random(g.nv), rng.random(g.ne)
    ref = onw.rhs(u, np.concatenate([omega, K]))
    probe2 = nd.Network(g2, L.kuramoto_first(), L.kuramoto_edge(), aggregator=null_aggregator)
    u2, p2 = u[order], np.concatenate([omega[order], K[edge_order]])
    onw2 = oracle_network(g2, L.kuramoto_first(), L.kuramoto_edge())
    du2 = onw2.rhs(u2, p2)
    back = np.empty_like(du2)
    back[order] = du2
    assert np.max(np.abs(back - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))
    # ... and the emulated 4-rank engine on the ordered graph agrees with ITS sequential oracle bit for bit
    out, ref2, plans, _sizes, _k = _run_world(nd, g2, L.kuramoto_first(), L.kuramoto_edge(), world, ncalls=3)
    assert np.array_equal(out, ref2)
    assert max(plans[0]["halo_lens"]) == h_ordered


@pytest.mark.parametrize("kernel", ["fused", "jag"])
@pytest.mark.parametrize("name,world", [("er_diffusion", 4), ("grid_kuramoto", 3), ("ba_mixed", 2)])
def test_emulated_ranks_fused_rk4(nd, monkeypatch, name, world, kernel):
    """nd_b200_rk4_exchange: classical RK4 on a row-partitioned network with the stage updates fused into the exchanging
    kernels (each stage publishes the boundary outputs of its own input vector); 10 steps against the oracle's RK4"""
    monkeypatch.setenv("ND_B200_KERNEL", kernel)
    g, vm, em = _cases(nd)[name]
    out, ref, _plans, _sizes, _k = _run_world(nd, g, vm, em, world, ncalls=10, rk4=1e-2)
    assert not np.isnan(out).any()
    if name == "ba_mixed":
        assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
    else:
        assert np.array_equal(out, ref)


def test_partition_from_the_bare_edge_list(nd, monkeypatch):
    """BASELINE config 5 path (5e7 vertices, 4e8 edges, 8 ranks on one host): partition, halo plan and engines built from
    the bare edge list (`Network.from_edgelist(layout_only=True)`, `PartitionedNetwork(from_edgelist=True)`) without any
    per-component table -- the same plan as the table-driven construction, and the same states on emulated ranks."""
    from networkdynamics_jl_b200 import distributed as D
    L = nd.Lib
    dirk = nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura")
    cases = [(nd.erdos_renyi(3000, 12000, seed=2), L.kuramoto_first(), L.kuramoto_edge()),
             (nd.barabasi_albert(2000, 3, seed=2), L.kuramoto_second(), L.kuramoto_edge()),       # dim 2: output = first state
             (nd.watts_strogatz(2000, 4, 0.3, seed=1, directed=True), L.diffusion_vertex(), dirk)]
    for g, vm, em in cases:
        a = nd.Network(g, vm, em, aggregator=null_aggregator)
        b = nd.Network.from_edgelist(g, vm, em, layout_only=True)
        assert (a.dim(), a.pdim()) == (b.dim(), b.pdim()) and b.handle is None
        ca, cb = D.row_entry_counts(a.im, a.layer.edgebatches), D.row_entry_counts(b.im, b.layer.edgebatches)
        assert np.array_equal(ca, cb)
        for world in (2, 5):
            rr = D.partition_rows(ca, world)
            assert [D.state_segments(a.vertexbatches, *r) for r in rr] == [D.state_segments(b.vertexbatches, *r) for r in rr]
            for r in range(world):
                pa, pb = D.halo_plan(a.im, a.layer.edgebatches, rr, r), D.halo_plan(b.im, b.layer.edgebatches, rr, r)
                assert np.array_equal(pa["gather_offset"], pb["gather_offset"]) and pa["halo_lens"] == pb["halo_lens"]
                assert pa["sends"].keys() == pb["sends"].keys()
                for q in pa["sends"]:
                    assert np.array_equal(pa["sends"][q][0], pb["sends"][q][0]) and pa["sends"][q][1] == pb["sends"][q][1]
    for kernel in ("fused", "jag"):
        monkeypatch.setenv("ND_B200_KERNEL", kernel)
        for g, vm, em in cases[:2]:
            out, ref, _p, _s, _k = _run_world(nd, g, vm, em, 3, ncalls=4, edgelist=True)
            assert np.max(np.abs(out - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_planning_in_chunks(nd, monkeypatch):
    """the partition / halo planning walks the edges in bounded chunks (config-5 scale: 4e8 edges); same plan for any chunk size"""
    from networkdynamics_jl_b200 import distributed as D
    L = nd.Lib
    g = nd.barabasi_albert(3000, 4, seed=7)
    em = ([L.kuramoto_edge(), nd.EdgeModel(g=nd.Directed(L.kuramoto_edge_f), outdim=1, pdim=1, name="dir_kura")], np.random.default_rng(1).integers(0, 2, g.ne))
    probes = [nd.Network(g, L.kuramoto_first(), em, aggregator=null_aggregator), nd.Network.from_edgelist(g, L.kuramoto_first(), L.kuramoto_edge(), layout_only=True)]
    for probe in probes:
        ref_cnt = D.row_entry_counts(probe.im, probe.layer.edgebatches)
        rr = D.partition_rows(ref_cnt, 4)
        ref = [D.halo_plan(probe.im, probe.layer.edgebatches, rr, r) for r in range(4)]
        monkeypatch.setattr(D, "_CHUNK", 777)
        assert np.array_equal(D.row_entry_counts(probe.im, probe.layer.edgebatches), ref_cnt)
        for r in range(4):
            pl = D.halo_plan(probe.im, probe.layer.edgebatches, rr, r)
            assert np.array_equal(pl["gather_offset"], ref[r]["gather_offset"]) and pl["halo_lens"] == ref[r]["halo_lens"]
            assert all(np.array_equal(pl["sends"][q][0], ref[r]["sends"][q][0]) for q in pl["sends"])
        monkeypatch.undo()


def test_hub_first_star_gives_an_empty_row_range_that_stays_empty(nd):
    """partition_rows can hand a rank the EMPTY range (hub first); the engine must read (a, a) as "no rows", not as
    "all rows" (ND_B200_FLAG_ROW_RANGE): results still equal the oracle's and the empty rank writes nothing"""
    from networkdynamics_jl_b200 import distributed as D
    assert D.partition_rows(np.array([100] + [1] * 100), 4)[0] == (0, 0)
    n = 400
    g = nd.SimpleGraph(n, np.ones(n - 1, dtype=np.int64), np.arange(2, n + 1))       # hub = vertex 1
    out, ref, plans, sizes, _ = _run_world(nd, g, nd.Lib.kuramoto_first(), nd.Lib.kuramoto_edge(), 4, 6)
    assert sizes[0]["nrows"] == 0 and sum(s["nrows"] for s in sizes) == n
    assert np.max(np.abs(out - ref)) <= 1e-13          # the hub row is reduced by a block tree: association differs


def test_halo_timeout_is_sticky_poisons_the_result_and_is_reported(nd, monkeypatch):
    """a rank whose peer never publishes gives up after the (configurable) spin budget, writes NaN instead
    of results computed from a stale halo, and every later exchange on that comm is refused with ND_B200_ETIMEOUT"""
    import cusim
    from networkdynamics_jl_b200 import distributed as D
    monkeypatch.setenv("ND_B200_HALO_TIMEOUT_MS", "0.0005")       # ~1000 emulator "ticks" (ns)
    g, vm, em = nd.erdos_renyi(600, 2400, seed=3), nd.Lib.diffusion_vertex(), nd.Lib.diffusion_edge()
    with cusim.use() as L:
        probe = nd.Network(g, vm, em, aggregator=null_aggregator)
        rr = D.partition_rows(D.row_entry_counts(probe.im, probe.layer.edgebatches), 2)
        plans = [D.halo_plan(probe.im, probe.layer.edgebatches, rr, r) for r in range(2)]
        nws = [nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", device=r, row_range=rr[r], keep_tables=False,
                                                                   gather_offset=plans[r]["gather_offset"], gather_len=plans[r]["gather_len"]))
               for r in range(2)]
        comms, handles = [], []
        for r in range(2):
            c = C.c_void_p()
            assert L.nd_b200_comm_create(r, r, 2, plans[r]["halo_lens"][r], max(plans[r]["halo_lens"]), C.byref(c)) == 0
            hb = C.create_string_buffer(nd._cabi.IPC_HANDLE_BYTES)
            assert L.nd_b200_comm_export(c, hb) == 0
            for peer, (offs, start) in plans[r]["sends"].items():
                offs = np.ascontiguousarray(offs, dtype=np.int64)
                assert L.nd_b200_comm_set_send(c, peer, offs.ctypes.data_as(nd._cabi.i64p), offs.size, start) == 0
            comms.append(c); handles.append(hb)
        for r in range(2):
            for q in range(2):
                assert L.nd_b200_comm_open_peer(comms[r], q, handles[q].raw) == 0
        n = probe.dim()
        u = np.random.default_rng(1).random(n)
        p = np.random.default_rng(2).random(probe.pdim())
        du = np.zeros(n)
        # only rank 0 calls: rank 1 never raises its arrival flag
        assert L.nd_b200_rhs_exchange(nws[0].handle, comms[0], _ptr(du), _ptr(u), _ptr(p), 0.0, None) == 0
        v = C.c_int32(0)
        assert L.nd_b200_comm_status(comms[0], C.byref(v)) == 0 and v.value == 1
        a, b = D.state_segments(probe.vertexbatches, *rr[0])[0]
        assert np.isnan(du[a:b]).any(), "rows that read the missing halo must not look valid"
        rc = L.nd_b200_rhs_exchange(nws[0].handle, comms[0], _ptr(du), _ptr(u), _ptr(p), 0.0, None)
        assert rc == nd._cabi.ETIMEOUT and "timed out" in L.nd_b200_last_error(nws[0].handle).decode()
        for c in comms:
            L.nd_b200_comm_destroy(c)
        del nws


def test_local_parameter_vectors_of_a_partition(nd):
    """SURVEY.md 8e "each rank owns ... the params of edges incident to its rows": PartitionedNetwork(local_parameters=True)
    builds every rank's engine on the subgraph of its incident edges.  Mixed vertex batches (table-driven mapping): the
    ranks' du tile the oracle's bit for bit, every rank's parameter vector is a gather of the global one, cut-edge
    parameters appear on both sides, and together the ranks hold far less than world copies of p."""
    import cusim
    import torch
    from networkdynamics_jl_b200.distributed import PartitionedNetwork
    L = nd.Lib
    n, world = 3000, 3
    rng = np.random.default_rng(8)
    g = nd.erdos_renyi(n, 4 * n, seed=6)
    vm = ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(np.array([0] * (n // 2) + [1] * (n // 2))))
    em = L.kuramoto_edge()
    onw = oracle_network(g, vm, em)
    u0 = rng.random(onw.lastidx_dynamic)
    p = 0.5 + rng.random(onw.lastidx_p)
    ref = onw.rhs(u0, p)
    covered = np.zeros(u0.size, dtype=np.int64)
    total_local = 0
    with cusim.use():
        for rank in range(world):
            pn = PartitionedNetwork(g, vm, em, rank=rank, world=world, exchange="nccl", local_parameters=True)
            assert pn.local_parameters and pn.global_pdim == p.size
            ploc = pn.localize_parameters(p)
            assert np.array_equal(ploc, p[pn.p_index]) and ploc.size < p.size
            total_local += sum(b - a for a, b in pn.parameter_segments())
            du = torch.full((pn.dim(),), float("nan"), dtype=torch.float64)
            pn.rhs(du, torch.from_numpy(u0.copy()), torch.from_numpy(ploc), 0.0, exchange=False)
            for a, b in pn.owned_segments:
                assert np.array_equal(du[a:b].numpy(), ref[a:b]), rank
                covered[a:b] += 1
    assert np.all(covered == 1)
    nvp = p.size - g.ne
    assert total_local < nvp + 2 * g.ne and total_local > nvp + g.ne       # cut edges twice, nothing world times
