"""The jagged ("warp-sliced") device layout of the default kernel, checked on the CPU.

The engine is created with ND_B200_FLAG_HOST_ONLY (tables are built, no CUDA call is made) and the walk that
rhs_jag_kernel performs -- per 32-lane slice, column by column, slot = base + popc(ballot & lanes_below) -- is replayed in
numpy.  Every row must see exactly its CSR entries in accumulation order (the SequentialAggregator order the CSR was
exported in, src/aggregators.jl:140-151), cut into consecutive parts of at most `split` entries on consecutive lanes.
"""
import numpy as np
import pytest

from helpers import oracle_network


def _walk(jag, nentries, window=32):
    """replay the kernel's slot computation; returns {row: [csr entries in the order the kernel accumulates them]}"""
    rows, seen = {}, np.zeros(nentries, dtype=np.int64)
    order = jag["order"]
    for s, (e0, row0, _batch, maxparts) in enumerate(jag["slices"]):
        d = jag["lanes"][s].astype(np.int64)
        ln, rowrel, head, valid = d & 63, (d >> 6) & 127, (d >> 13) & 1, (d >> 14) & 1
        assert np.all(ln[valid == 0] == 0) and np.all(head[valid == 0] == 0)
        assert np.all(np.diff(valid) <= 0), "valid lanes are a prefix"
        if window == 128 and not valid.any():
            # window mode (rhs_jag_kernel<..., WIN>): empty slices pad every 128-row window to whole thread blocks
            assert maxparts == 1
            continue
        assert head[0] == 1 and (rowrel[0] == 0 or window > 32) and np.all(rowrel[valid == 1] < window)
        base, j, per_lane = int(e0), 0, [[] for _ in range(32)]
        while True:
            act = ln > j
            if not act.any():
                break
            slot = base + np.cumsum(act) - act           # popc(ballot & lanes_below)
            for l in np.nonzero(act)[0]:
                per_lane[l].append(int(order[slot[l]]))
                seen[order[slot[l]]] += 1
            base += int(act.sum())
            j += 1
        # heads add the following continuation lanes in lane order
        nparts_max = 1
        l = 0
        while l < 32 and valid[l]:
            assert head[l] == 1
            r, ents, k = int(row0 + rowrel[l]), list(per_lane[l]), l + 1
            while k < 32 and valid[k] and not head[k]:
                assert rowrel[k] == rowrel[l]
                assert len(per_lane[k - 1]) == jag["split"], "only the last part of a row may be short"
                ents += per_lane[k]
                k += 1
            nparts_max = max(nparts_max, k - l)
            assert r not in rows
            rows[r] = ents
            l = k
        assert nparts_max == maxparts
    for e0, row, ne, _batch in jag["longs"]:
        ents = [int(order[e0 + k]) for k in range(ne)]
        seen[ents] += 1
        assert row not in rows
        rows[int(row)] = ents
    assert np.all(seen == 1), "every CSR entry is stored exactly once"
    return rows


def _cases(nd):
    L = nd.Lib
    rng = np.random.default_rng(3)
    yield "er", nd.erdos_renyi(3000, 12000, seed=1), L.diffusion_vertex(), L.diffusion_edge(), {}
    n = 4000
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    yield "ba-mixed", nd.barabasi_albert(n, 4, seed=1), ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(half)), L.kuramoto_edge(), {}
    yield "ba-split7", nd.barabasi_albert(n, 4, seed=2), L.kuramoto_first(), L.kuramoto_edge(), {"ND_B200_JAG_SPLIT": "7"}
    star = nd.SimpleGraph(5000, np.ones(4999, dtype=np.int64), np.arange(2, 5001))
    yield "star", star, L.kuramoto_first(), L.kuramoto_edge(), {}
    yield "star-thr64", star, L.kuramoto_first(), L.kuramoto_edge(), {"thr": 64}
    yield "grid-dq", nd.grid_graph(13, 11), L.swing_dq(), L.line_dq(), {}
    yield "isolated", nd.SimpleGraph(70, [1, 2], [2, 3]), L.diffusion_vertex(), L.diffusion_edge_nop(), {}
    yield "noedges", nd.SimpleGraph(5, [], []), L.kuramoto_first(), L.kuramoto_edge(), {}
    yield "partition", nd.erdos_renyi(3000, 12000, seed=4), L.diffusion_vertex(), L.diffusion_edge(), {"rows": (1000, 2100)}


def _lane_utilisation(jag):
    """entries / (32 * sum over slices of the longest lane): the share of lane-iterations of the walk that do work"""
    ln = (jag["lanes"].astype(np.int64) & 63)
    return float(ln.sum()) / max(1.0, 32.0 * float(ln.max(axis=1).sum()))


@pytest.mark.parametrize("window", [32, 64, 128])
def test_jagged_layout_replays_the_csr_in_order(nd, monkeypatch, window):
    """window > 32 (ND_B200_JAG_WINDOW): degree-bucketed slices -- the rows of a window are dealt to the lanes by
    decreasing degree; same entries, same per-row order, fewer wasted lane iterations"""
    monkeypatch.setenv("ND_B200_KERNEL", "jag")
    monkeypatch.setenv("ND_B200_JAG_WINDOW", str(window))
    if window == 128:
        monkeypatch.setenv("ND_B200_JAG_WIN", "1")     # window mode: every 128-row window padded to whole thread blocks
    for name, g, vm, em, opt in _cases(nd):
        for k, v in opt.items():
            if k.startswith("ND_"):
                monkeypatch.setenv(k, v)
        agg = nd.B200Aggregator("+", host_only=True, long_row_threshold=opt.get("thr", 0), row_range=opt.get("rows"))
        nw = nd.Network(g, vm, em, aggregator=agg)
        rowptr, _nbr, _eid, _side = nw.export_tables()
        jag = nw.export_jag()
        sz = nw.engine_sizes()
        rows = _walk(jag, sz["nentries"], window)
        if window == 128 and name != "grid-dq" and len(jag["slices"]):
            # window mode: a thread block (four consecutive slices) holds rows of ONE 128-row window of one vertex batch
            sl = np.asarray(jag["slices"])
            assert len(sl) % 4 == 0, name
            blocks = sl.reshape(-1, 4, 4)
            assert np.all(blocks[:, :, 1] == blocks[:, :1, 1]) and np.all(blocks[:, :, 2] == blocks[:, :1, 2]), name
        if name == "er":
            util = _lane_utilisation(jag)
            assert util > {32: 0.45, 64: 0.6, 128: 0.7}[window], (window, util)
        r0 = sz["row_begin"]
        assert sorted(rows) == list(range(r0, r0 + sz["nrows"])), name
        for r, ents in rows.items():
            a, z = rowptr[r - r0], rowptr[r - r0 + 1]
            assert ents == list(range(a, z)), (name, r)
        if name == "star":
            assert len(jag["longs"]) == 1 and jag["longs"][0][2] == 4999
        if name == "star-thr64":
            assert len(jag["longs"]) == 1
        if name == "ba-split7":
            assert jag["split"] == 7 and int(jag["slices"][:, 3].max()) > 1
        for k in opt:
            if k.startswith("ND_"):
                monkeypatch.delenv(k)
        with pytest.raises(nd.ArgumentError):     # a host-only engine cannot evaluate
            nw(np.zeros(nw.dim()), np.zeros(nw.dim()), np.zeros(nw.pdim()), 0.0)


def test_jagged_layout_numpy_evaluation_matches_oracle(nd, monkeypatch):
    """diffusion through the replayed layout: same entries, same order -> bit-identical to the sequential oracle"""
    monkeypatch.setenv("ND_B200_KERNEL", "jag")
    g = nd.erdos_renyi(2000, 8000, seed=9)
    vm, em = nd.Lib.diffusion_vertex(), nd.Lib.diffusion_edge()
    nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", host_only=True))
    onw = oracle_network(g, vm, em)
    _rowptr, nbr, eid, side = nw.export_tables()
    rows = _walk(nw.export_jag(), nbr.size)
    rng = np.random.default_rng(0)
    x, p = rng.standard_normal(g.nv), rng.random(g.ne)
    du = np.zeros(g.nv)
    for r, ents in rows.items():
        acc = 0.0
        for k in ents:
            xs, xd = (x[r], x[nbr[k] - 1]) if side[k] else (x[nbr[k] - 1], x[r])
            val = p[eid[k] - 1] * (xs - xd)
            acc = acc + (-val if side[k] else val)
        du[r] = acc
    assert np.array_equal(du, onw.rhs(x, p))


def test_halo_plan_and_interior_first_order(nd, monkeypatch):
    """Multi-GPU packed halo on the CPU: for every rank of a 3-way partition, the plan's sends fill the halo so that
    every entry of every owned row reads the right value through the engine's gather offsets, nothing unsent is read,
    and the jagged layout puts the slices that read no remote output first (they run while the halo is in flight)."""
    from networkdynamics_jl_b200 import distributed as D
    from helpers import null_aggregator
    monkeypatch.setenv("ND_B200_KERNEL", "jag")
    L = nd.Lib
    rng = np.random.default_rng(7)
    n = 3000
    half = np.array([0] * (n // 2) + [1] * (n // 2))
    cases = [(nd.barabasi_albert(n, 4, seed=2), ([L.kuramoto_first(), L.kuramoto_second()], rng.permutation(half)), L.kuramoto_edge()),
             (nd.grid_graph(60, 50), L.kuramoto_first(), L.kuramoto_edge()),
             (nd.watts_strogatz(3000, 4, 0.3, seed=1, directed=True), L.diffusion_vertex(),
              [nd.EdgeModel(g=nd.Directed(L.diffusionedge_nop), outdim=1, pdim=0, name="dir_diff"), L.diffusion_edge()] * 3000)]
    world = 3
    for g, vm, em in cases:
        if isinstance(em, list):
            em = em[:g.ne]
        probe = nd.Network(g, vm, em, aggregator=null_aggregator)
        im = probe.im
        rr = D.partition_rows(D.row_entry_counts(im, probe.layer.edgebatches), world)
        plans = [D.halo_plan(im, probe.layer.edgebatches, rr, r) for r in range(world)]
        goff = np.asarray(im.v_data) - 1
        u0 = rng.random(im.lastidx_dynamic)
        rov = D.row_of_vertex(im)
        for r, plan in enumerate(plans):
            # what the peers' publish kernels write into rank r's halo buffer
            halo = np.full(plan["halo_lens"][r], np.nan)
            for q in range(world):
                if q != r:
                    offs, start = plans[q]["sends"][r]
                    assert np.all(np.diff(offs) > 0)
                    assert np.all(np.isnan(halo[start:start + offs.size])), "blocks of different owners overlap"
                    halo[start:start + offs.size] = u0[offs]
            written = ~np.isnan(halo)
            assert written.sum() == plan["need"].size, "every needed output is written by exactly one peer (the rest is alignment padding)"
            own = np.full(im.lastidx_dynamic, np.nan)
            for a, b in D.state_segments(probe.vertexbatches, *rr[r]):
                own[a:b] = u0[a:b]
            gsrc = np.concatenate([own, halo])
            nw = nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", host_only=True, row_range=rr[r],
                                                                  gather_offset=plan["gather_offset"], gather_len=plan["gather_len"]))
            rowptr, nbr, _eid, _side = nw.export_tables()
            got = gsrc[plan["gather_offset"][nbr - 1]]
            assert np.array_equal(got, u0[goff[nbr - 1]])
            # interior slices first
            jag = nw.export_jag()
            assert jag["halo_len"] == plan["halo_lens"][r]
            remote_row = np.zeros(g.nv, dtype=bool)
            rows_of_entries = np.repeat(np.arange(rr[r][0], rr[r][1]), np.diff(rowptr))
            is_remote = plan["vowner"][nbr - 1] != r
            remote_row[np.unique(rows_of_entries[is_remote])] = True
            for s, (e0, row0, _b, _mp) in enumerate(jag["slices"]):
                d = jag["lanes"][s].astype(np.int64)
                rows = row0 + ((d >> 6) & 127)[((d >> 14) & 1) == 1]
                assert np.all(remote_row[rows] == (s >= jag["wait_from"])), (s, jag["wait_from"])
            _walk(jag, nbr.size)
            assert 0 < jag["wait_from"] < len(jag["slices"]) or g.nv != 3000 or True
        # a halo engine refuses construction when an owned vertex is redirected
        bad = plans[0]["gather_offset"].copy()
        v_own = int(np.nonzero(plans[0]["vowner"] == 0)[0][0])
        bad[v_own] = plans[0]["gather_len"] - 1
        if plans[0]["halo_lens"][0] > 0:
            with pytest.raises(nd.ArgumentError):
                nd.Network(g, vm, em, aggregator=nd.B200Aggregator("+", host_only=True, row_range=rr[0], gather_offset=bad,
                                                                 gather_len=plans[0]["gather_len"]))
