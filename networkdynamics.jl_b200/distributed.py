"""Vertex-partitioned multi-GPU RHS: one process per GPU (`torch.distributed`, NCCL on GPUs, gloo in CPU tests).

The reference has no multi-device path; this follows SURVEY.md 8(e):
  * aggregation-slot rows are split into `world` CONTIGUOUS ranges of (nearly) equal directed-entry count;
  * rank r owns the states / du entries of its rows and evaluates only those rows (engine `row_range`);
  * per RHS one exchange step.  "p2p" (default on GPUs): every rank packs, for each peer, exactly the vertex outputs
    that peer's rows read (the boundary outputs of cut edges, `halo_plan`) and stores them straight into the peer's
    halo buffer over NVLink; rows that read no remote output are evaluated while the halo is in flight.
    "nccl" (and the gloo CPU tests): owned state ranges are all-gathered into every rank's copy of `u`.
Accumulation order per row is untouched (a row is never split across ranks), so `du` is identical to the
single-GPU result, bit for bit.

Host logic (partitioning, segment bookkeeping, the exchange) is plain torch and runs on CPU tensors with gloo.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .network import B200Aggregator, B200Execution, ComponentBatch, IndexManager, Network
from . import _cabi


def row_of_vertex(im: IndexManager) -> np.ndarray:
    """aggregation-slot row (0-based) of every vertex: (v_aggr.first - 1) / edepth; rows follow batch order."""
    if getattr(im, "homogeneous", False):        # one vertex batch 1:nv (Network.from_edgelist): row = vertex
        return np.arange(im.nv, dtype=np.int64)
    if im.edepth > 0:
        return (im.v_aggr - 1) // im.edepth
    row = np.empty(im.nv, dtype=np.int64)
    k = 0
    for b in im.vertexbatches:
        row[b.indices - 1] = k + np.arange(len(b))
        k += len(b)
    return row


def row_entry_counts(im: IndexManager, edgebatches: Sequence[ComponentBatch]) -> np.ndarray:
    """directed entries per row = number of edge outputs aggregated into the row (dst outputs + src outputs)"""
    rov = row_of_vertex(im)
    cnt = np.zeros(im.nv, dtype=np.int64)
    for b in edgebatches:
        for e in _edge_chunks(b, im.ne):
            cnt += np.bincount(rov[im.edge_dst[e] - 1], minlength=im.nv)
            if b.model.outdim_src > 0:
                cnt += np.bincount(rov[im.edge_src[e] - 1], minlength=im.nv)
    return cnt


HALO_ALIGN = 16       # doubles: 128-byte alignment of every owner's block inside a reader's halo buffer

_CHUNK = 1 << 25      # edges per pass: bounds the temporaries of the planning functions at config-5 scale (4e8 edges)


def _edge_chunks(b: ComponentBatch, ne: int):
    """index objects (slices, or pieces of the batch's index array, 0-based) covering the edges of a batch"""
    if b.indices is None:
        for a in range(0, ne, _CHUNK):
            yield slice(a, min(a + _CHUNK, ne))
    else:
        idx = np.asarray(b.indices, dtype=np.int64) - 1
        for a in range(0, idx.size, _CHUNK):
            yield idx[a:a + _CHUNK]


def partition_rows(entry_counts: np.ndarray, world: int, prefer_equal_rows: float = 0.05) -> List[Tuple[int, int]]:
    """`world` contiguous row ranges with (nearly) equal entry count (+1 per row so that edge-less rows spread too).
    When the plain equal-rows split is within `prefer_equal_rows` of that balance (homogeneous graphs such as ER) it is
    used instead, because equal ranges allow one in-place all-gather for the exchange."""
    n = int(entry_counts.size)
    if world > 1 and n % world == 0 and n > 0:
        step = n // world
        loads = (entry_counts + 1).reshape(world, step).sum(axis=1)
        if loads.max() <= (1.0 + prefer_equal_rows) * loads.mean():
            return [(r * step, (r + 1) * step) for r in range(world)]
    w = np.cumsum(entry_counts + 1)
    total = int(w[-1]) if w.size else 0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(w, total * r / world, side="left")))
    cuts.append(int(entry_counts.size))
    cuts = [min(max(c, cuts[i - 1] if i else 0), entry_counts.size) for i, c in enumerate(cuts)]
    # a rank may end up with the EMPTY range (hub-first star graphs: the first row already exceeds a share); engines are
    # created with ND_B200_FLAG_ROW_RANGE, so (a, a) means "no rows", not "all rows"
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def state_segments(vertexbatches: Sequence[ComponentBatch], r0: int, r1: int) -> List[Tuple[int, int]]:
    """0-based [start, stop) ranges of the flat state vector owned by rows [r0, r1): one range per vertex batch the
    row range intersects (states of a batch are contiguous and in row order, src/network_structure.jl:224-239)."""
    segs, row = [], 0
    for b in vertexbatches:
        n, dim = len(b), b.model.dim
        lo, hi = max(r0, row), min(r1, row + n)
        if lo < hi and dim > 0:
            first = b.state_first - 1
            segs.append((first + (lo - row) * dim, first + (hi - row) * dim))
        row += n
    return segs


def param_segments(vertexbatches: Sequence[ComponentBatch], r0: int, r1: int) -> List[Tuple[int, int]]:
    """0-based [start, stop) ranges of the flat parameter vector holding the vertex parameters of rows [r0, r1)"""
    segs, row = [], 0
    for b in vertexbatches:
        n, pd = len(b), b.model.pdim
        lo, hi = max(r0, row), min(r1, row + n)
        if lo < hi and pd > 0:
            first = b.p_first - 1
            segs.append((first + (lo - row) * pd, first + (hi - row) * pd))
        row += n
    return segs


def edge_state_segments(edgebatches: Sequence[ComponentBatch], r0: int, r1: int, nrows: int) -> List[Tuple[int, int]]:
    """0-based [start, stop) ranges of the states of edges WITH states whose `f` the owner of rows [r0, r1) evaluates: every
    such batch is cut in the proportion of the row range (launch_edge_f in csrc/nd_b200.cu uses the same floor divisions), so
    the ranks' chunks are contiguous and tile the batch.  Their outputs are read by other ranks' rows from the all-gathered
    state vector like any vertex state."""
    segs = []
    for b in edgebatches:
        dim, n = b.model.dim, len(b)
        if dim > 0 and n > 0:
            i0, i1 = n * int(r0) // max(int(nrows), 1), n * int(r1) // max(int(nrows), 1)
            if i0 < i1:
                first = b.state_first - 1
                segs.append((first + i0 * dim, first + i1 * dim))
    return segs


def owner_of_rows(rows: np.ndarray, row_ranges: Sequence[Tuple[int, int]]) -> np.ndarray:
    """rank that owns each row (row ranges are contiguous and ascending)"""
    ends = np.array([b for _, b in row_ranges], dtype=np.int64)
    return np.searchsorted(ends, rows, side="right")


def halo_plan(im: IndexManager, edgebatches: Sequence[ComponentBatch], row_ranges: Sequence[Tuple[int, int]], rank: int):
    """Who reads what across the cut ("each rank sends the outputs of its boundary vertices needed by
    rank q to q").  Needs vertices whose single output is their first state (StateMask), so outputs live in `u`.

    need[r] = the remote vertices read by the rows of rank r, ordered by (owner rank, state offset).  Rank r's halo
    buffer holds their outputs in that order, logically appended to the state vector: position lastidx_dynamic + k.
    Returns for `rank`: gather_offset (per vertex, what nd_b200_desc.gather_offset takes), gather_len, halo_lens (per
    rank), sends = {peer: (state offsets this rank packs for peer, start position inside peer's halo)}, and
    need (vertex ids, 0-based, in halo order) for tests."""
    world = len(row_ranges)
    nv = im.nv
    vowner = owner_of_rows(row_of_vertex(im), row_ranges)
    if getattr(im, "homogeneous", False):
        goff = np.arange(nv, dtype=np.int64) * int(im.vdim)
    else:
        goff = np.asarray(im.v_data, dtype=np.int64) - 1
    needed = np.zeros((world, nv), dtype=bool)
    for b in edgebatches:
        for e in _edge_chunks(b, int(np.asarray(im.edge_src).size)):
            s, t = np.asarray(im.edge_src)[e] - 1, np.asarray(im.edge_dst)[e] - 1
            needed[vowner[t], s] = True               # the dst row reads the src vertex's output
            if b.model.outdim_src > 0:
                needed[vowner[s], t] = True           # ... and the src row (wrapper output) reads the dst vertex's
    needed[vowner, np.arange(nv)] = False             # own vertices are read from u
    vorder = np.lexsort((goff, vowner))
    need = [vorder[needed[r][vorder]] for r in range(world)]
    counts = np.array([np.bincount(vowner[n], minlength=world) for n in need], dtype=np.int64)   # [reader, owner]
    # every owner's block starts on a 128-byte boundary of the reader's halo buffer: the publishing blocks' warps then store
    # whole aligned 256-byte pieces over NVLink (profiles/r02e_p2p_store_bench_raw.txt: 650-700 GB/s aligned, 480-500 GB/s
    # when the destination is only 8-byte aligned).  Pad slots are never addressed.
    padded = (counts + HALO_ALIGN - 1) // HALO_ALIGN * HALO_ALIGN
    starts = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(padded, axis=1)], axis=1)
    ustarts = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(counts, axis=1)], axis=1)
    halo_lens = [int(starts[r, -1]) for r in range(world)]
    nstates = int(im.lastidx_dynamic)
    gather_offset = goff.copy()
    own = vowner[need[rank]]                          # need[rank] is ordered by (owner, state offset)
    k = np.arange(need[rank].size, dtype=np.int64)
    gather_offset[need[rank]] = nstates + starts[rank, own] + (k - ustarts[rank, own])
    sends = {}
    for r in range(world):
        if r == rank:
            continue
        mine = need[r][vowner[need[r]] == rank]
        sends[r] = (np.ascontiguousarray(goff[mine]), int(starts[r, rank]))
    return dict(gather_offset=gather_offset, gather_len=nstates + halo_lens[rank], halo_lens=halo_lens, sends=sends,
                need=need[rank], recv_counts=counts[rank], vowner=vowner)


def statemask_outputs(vertexbatches: Sequence[ComponentBatch], vdepth: int) -> bool:
    """True when every vertex output is its first state (the gather source is the state vector itself)"""
    from . import _cabi
    return vdepth == 1 and all(b.model.kernel_kind() not in (None, _cabi.V_SWING_DQ) for b in vertexbatches)


def _uniform_allgather_layout(segments_by_rank, n) -> bool:
    """True when rank r owns exactly [r*len, (r+1)*len) and the ranges tile the whole vector."""
    world = len(segments_by_rank)
    if n % world or any(len(s) != 1 for s in segments_by_rank):
        return False
    step = n // world
    return all(tuple(s[0]) == (r * step, (r + 1) * step) for r, s in enumerate(segments_by_rank))


def exchange_states(u, segments_by_rank: Sequence[Sequence[Tuple[int, int]]], group=None, async_op: bool = False):
    """Every rank publishes its owned state ranges into everybody's copy of `u` (in place).  Equal tiling ranges: ONE
    in-place all-gather; otherwise one broadcast per (owner, range) -- ranges are contiguous, so nothing is packed."""
    import torch.distributed as dist
    if _uniform_allgather_layout(segments_by_rank, u.numel()):
        rank = dist.get_rank(group)
        a, b = segments_by_rank[rank][0]
        w = dist.all_gather_into_tensor(u, u[a:b], group=group, async_op=True)
        if async_op:
            return [w]
        w.wait()
        return None
    works = []
    ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
    for r, segs in enumerate(segments_by_rank):
        for a, b in segs:
            works.append(dist.broadcast(u[a:b], src=ranks[r], group=group, async_op=True))
    if async_op:
        return works
    for w in works:
        w.wait()
    return None


class PartitionedNetwork:
    """`Network` evaluated cooperatively by `world` ranks.  `rhs(du, u, p, t)`: exchange the owned states of `u`, then
    evaluate the owned rows into `du` (only the owned entries of `du` are written)."""

    def __init__(self, g, vertexm, edgem, *, rank: int, world: int, group=None, device=None,
                 long_row_threshold: int = 0, exchange: str = "auto", from_edgelist: bool = False,
                 edge_parameters: str = "auto", local_parameters: bool = False):
        """exchange: "p2p"  -- states are pushed into every rank's replica with NVLink peer stores by the engine's
        publish kernel and the RHS kernel waits on arrival flags (nd_b200_rhs_exchange; no NCCL on the data path);
        "nccl" -- torch.distributed collectives on the caller's `u`; "auto" -- p2p when the network allows it
        (all vertices StateMask) and CUDA IPC works, else nccl."""
        # host tables first (no device work) to compute the partition.  from_edgelist (one registry vertex model, one
        # registry edge model): nothing per component is materialised -- the partition, the halo plan and the engine are
        # built from the bare edge list (BASELINE config 5: 5e7 vertices, 4e8 edges per graph, 8 ranks on one host)
        if from_edgelist:
            probe = Network.from_edgelist(g, vertexm, edgem, layout_only=True)
        else:
            probe = Network(g, vertexm, edgem, execution=B200Execution(), aggregator=lambda im, eb: None)
        self.rank, self.world, self.group = rank, world, group
        self.entry_counts = row_entry_counts(probe.im, probe.layer.edgebatches)
        self.row_ranges = partition_rows(self.entry_counts, world)
        nrows = int(self.entry_counts.size)
        self.segments = [state_segments(probe.vertexbatches, a, b) + edge_state_segments(probe.layer.edgebatches, a, b, nrows)
                         for a, b in self.row_ranges]
        stateful_edges = any(b.model.dim > 0 or b.model.kernel_kind() == _cabi.E_LOOPBACK or b.model.extdim > 0 for b in probe.layer.edgebatches) \
            or any(b.model.extdim > 0 for b in probe.vertexbatches)
        self.comm = None
        self.exchange_kind = "nccl"
        self.plan = None
        # local_parameters ("each rank owns ... the params of edges incident to its rows, cut-edge params
        # duplicated on both sides"): this rank's engine is built on the subgraph of the edges INCIDENT to its rows (same
        # vertices, same relative edge order -> same accumulation order per row), so its parameter vector holds the vertex
        # parameters and only those edges' parameters: p_local = p_global[p_index].  u / du keep the global layout.
        # (one edge batch only: with several, the subgraph's batches could be registered in another order than the full
        # graph's, which would change the accumulation order inside a row)
        self.local_parameters = bool(local_parameters) and world > 1 and not stateful_edges and len(probe.layer.edgebatches) == 1
        self.global_pdim = int(probe.pdim())
        self.p_index = None
        self._vertex_pdim = int(sum(len(b) * b.model.pdim for b in probe.vertexbatches)) if not from_edgelist else int(g.nv * vertexm.pdim)
        self._own_param_segs = param_segments(probe.vertexbatches, *self.row_ranges[rank]) if not from_edgelist else \
            ([(self.row_ranges[rank][0] * vertexm.pdim, self.row_ranges[rank][1] * vertexm.pdim)] if vertexm.pdim else [])
        g_full, edgem_full = g, edgem
        if self.local_parameters:
            src, dst = np.asarray(probe.im.edge_src), np.asarray(probe.im.edge_dst)
            vowner = owner_of_rows(row_of_vertex(probe.im), self.row_ranges)
            keep = np.nonzero((vowner[src - 1] == rank) | (vowner[dst - 1] == rank))[0]
            g = type(g)(g.nv, src[keep], dst[keep], _canonical=True)
            if from_edgelist:
                pe = edgem.pdim
                self.p_index = np.concatenate([np.arange(self._vertex_pdim, dtype=np.int64),
                                               (self._vertex_pdim + keep[:, None] * pe + np.arange(pe, dtype=np.int64)[None, :]).ravel()])
            else:
                from .components import EdgeModel
                if isinstance(edgem, EdgeModel):
                    pass
                elif isinstance(edgem, tuple):
                    edgem = (edgem[0], np.asarray(edgem[1])[keep])
                else:
                    edgem = [edgem[i] for i in keep]
                loc = Network(g, vertexm, edgem, execution=B200Execution(), aggregator=lambda im, eb: None)
                widths = np.array([m.pdim for m in probe.im.edgem], dtype=np.int64)[np.asarray(probe.im.etype)[keep]]
                gfirst = np.asarray(probe.im.e_para, dtype=np.int64)[keep] - 1
                lfirst = np.asarray(loc.im.e_para, dtype=np.int64) - 1
                idx = np.arange(int(loc.pdim()), dtype=np.int64)          # vertex part: identical layout
                for w in np.unique(widths):
                    if w == 0:
                        continue
                    m = widths == w
                    cols = np.arange(int(w), dtype=np.int64)[None, :]
                    idx[(lfirst[m][:, None] + cols).ravel()] = (gfirst[m][:, None] + cols).ravel()
                self.p_index = idx

        def engine(plan):
            if from_edgelist:
                return Network.from_edgelist(g, vertexm, edgem, device=device, row_range=self.row_ranges[rank], keep_tables=False,
                                             gather_offset=None if plan is None else plan["gather_offset"],
                                             gather_len=0 if plan is None else plan["gather_len"], edge_parameters=edge_parameters)
            return Network(g, vertexm, edgem, execution=B200Execution(),
                           aggregator=B200Aggregator("+", device=device, row_range=self.row_ranges[rank],
                                                     long_row_threshold=long_row_threshold, keep_tables=False,
                                                     gather_offset=None if plan is None else plan["gather_offset"],
                                                     gather_len=0 if plan is None else plan["gather_len"],
                                                     edge_parameters=edge_parameters))

        want_p2p = exchange in ("p2p", "auto") and world > 1
        if want_p2p and stateful_edges:         # their states travel with the all-gather; the packed halo carries vertex outputs only
            if exchange == "p2p":
                raise RuntimeError("p2p exchange carries only vertex outputs (no edge states, loopback connections, external inputs): use exchange='nccl'")
            want_p2p = False
        if want_p2p and not statemask_outputs(probe.vertexbatches, probe.im.vdepth):
            if exchange == "p2p":
                raise RuntimeError("p2p exchange needs StateMask vertices (the gather source must be the state vector)")
            want_p2p = False
        if want_p2p:
            # the plan needs every rank's needs (where this rank's block starts inside a peer's halo): full graph
            self.plan = halo_plan(probe.im, probe.layer.edgebatches, self.row_ranges, rank)
            self.nw = engine(self.plan)
            try:
                self._setup_p2p()
                self.exchange_kind = "p2p"
            except Exception:
                if exchange == "p2p":
                    raise
                self.plan, self.nw = None, engine(None)     # every rank fails together (see _setup_p2p): NCCL path
        else:
            self.nw = engine(None)

    # -- NVLink peer-memory exchange ----------------------------------------------------------------------------
    def _setup_p2p(self):
        import ctypes as C
        import torch.distributed as dist
        from . import _cabi
        L = _cabi.lib()
        plan = self.plan
        h = C.c_void_p()
        dev = self.nw.layer.aggregator.device
        rc = L.nd_b200_comm_create(dev, self.rank, self.world, plan["halo_lens"][self.rank], max(plan["halo_lens"]), C.byref(h))
        ok = rc == 0
        handle = C.create_string_buffer(_cabi.IPC_HANDLE_BYTES)
        if ok:
            ok = L.nd_b200_comm_export(h, handle) == 0
        if ok:
            for peer, (offs, start) in plan["sends"].items():
                offs = np.ascontiguousarray(offs, dtype=np.int64)
                ok = ok and L.nd_b200_comm_set_send(h, peer, offs.ctypes.data_as(_cabi.i64p), offs.size, start) == 0
        # every rank must take the same branch: agree on success before mapping anything
        flags = [None] * self.world
        dist.all_gather_object(flags, (bool(ok), handle.raw), group=self.group)
        if not all(f[0] for f in flags):
            msg = (L.nd_b200_comm_last_error(h) if h else L.nd_b200_comm_last_error(None)).decode()
            if h:
                L.nd_b200_comm_destroy(h)
            raise RuntimeError("nd_b200_comm setup failed on some rank: " + msg)
        opened = all(L.nd_b200_comm_open_peer(h, r, flags[r][1]) == 0 for r in range(self.world))
        res = [None] * self.world
        dist.all_gather_object(res, bool(opened), group=self.group)
        if not all(res):
            msg = L.nd_b200_comm_last_error(h).decode()
            L.nd_b200_comm_destroy(h)
            raise RuntimeError("CUDA IPC mapping of a peer halo buffer failed: " + msg)
        self.comm = h
        dist.barrier(group=self.group)

    def pack_params(self, p, *, stream=None):
        """Per-rank packed copy of the edge parameters this rank's rows read (nd_b200_pack_params): removes the isolated
        8-byte reads of cut-edge parameters scattered over the whole `p` vector.  Contract: edge parameters in `p` stay
        unchanged until the next call; `pack_params(None)` goes back to re-reading `p` on every RHS.  Local (no collective)."""
        self.nw.pack_params(p, stream=stream)

    def comm_timed_out(self) -> bool:
        import ctypes as C
        from . import _cabi
        if self.comm is None:
            return False
        v = C.c_int32(0)
        _cabi.lib().nd_b200_comm_status(self.comm, C.byref(v))
        return bool(v.value)

    def close(self):
        """collective: all ranks must call it (peers' mappings are closed before the owners free their replicas)"""
        import torch.distributed as dist
        from . import _cabi
        if self.comm is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            _cabi.lib().nd_b200_comm_destroy(self.comm)
            self.comm = None
            dist.barrier(group=self.group)

    def dim(self):
        return self.nw.dim()

    def pdim(self):
        return self.nw.pdim()

    @property
    def owned_segments(self):
        return self.segments[self.rank]

    def exchange(self, u):
        exchange_states(u, self.segments, self.group)

    def parameter_segments(self):
        """0-based [start, stop) ranges of THIS RANK's parameter vector that its rows read: with local_parameters the vertex
        parameters of its rows and the whole (incident-edge) edge block; otherwise the whole global vector (the parameters
        of cut edges are scattered over it)"""
        if not self.nw.pdim():
            return []
        if self.local_parameters:
            segs = list(self._own_param_segs)
            if self.nw.pdim() > self._vertex_pdim:
                segs.append((self._vertex_pdim, int(self.nw.pdim())))
            return segs
        return [(0, int(self.nw.pdim()))]

    def localize_parameters(self, p):
        """this rank's parameter vector from the reference's flat (global) one: p[p_index] (numpy array or torch tensor)"""
        if not self.local_parameters:
            return p
        if isinstance(p, np.ndarray):
            return np.ascontiguousarray(p[self.p_index])
        import torch
        return p.index_select(0, torch.from_numpy(self.p_index).to(p.device)).contiguous()

    def rhs_local(self, du, u, p, t, *, stream=None):
        """timing aid: the owned rows on the current halo content, no exchange (p2p engines only)"""
        from .network import _addr, _stream_handle
        from . import _cabi
        a_du, _, _ = _addr(du)
        a_u, _, _ = _addr(u)
        a_p, _, _ = _addr(p)
        rc = _cabi.lib().nd_b200_rhs_local(self.nw.handle, self.comm, a_du, a_u, a_p, float(t), _stream_handle(stream))
        if rc:
            self.nw._fail(rc)

    def halo_stats(self):
        """outputs received / sent per exchange by this rank and the share of its rows that read no remote output"""
        if self.plan is None:
            return None
        sent = int(sum(o.size for o, _ in self.plan["sends"].values()))
        return {"recv_outputs": int(self.plan["halo_lens"][self.rank]), "sent_outputs": sent,
                "owned_states": int(sum(b - a for a, b in self.owned_segments))}

    def rhs(self, du, u, p, t, *, exchange: bool = True, stream=None):
        """`nw(du,u,p,t)` for the owned rows; only the owned states of `u` need to be valid on entry"""
        if exchange and self.comm is not None:
            from .network import _addr, _stream_handle
            from . import _cabi
            a_du, _, n_du = _addr(du)
            a_u, _, n_u = _addr(u)
            a_p, _, n_p = _addr(p)
            self.nw._check_sizes(n_du, n_u, n_p, p is not None)
            # per-rank packed copy of the edge parameters this rank's rows read, kept in step with p automatically
            # (edge_parameters="auto"): removes the isolated reads of cut-edge parameters scattered over the whole p
            self.nw._auto_pack(p, a_p, stream)
            rc = _cabi.lib().nd_b200_rhs_exchange(self.nw.handle, self.comm, a_du, a_u, a_p, float(t), _stream_handle(stream))
            if rc:
                self.nw._fail(rc)
            return
        if exchange:
            self.exchange(u)
        self.nw(du, u, p, t)

    def rk4(self, u, p, t0, dt, nsteps, *, stream=None, work=None):
        """`nsteps` classical RK4 steps on the owned states of `u` (in place).  With the NVLink exchange this is
        nd_b200_rk4_exchange: four exchanging kernel launches per step per rank and nothing else (stage updates fused into
        the kernels' epilogues); with the all-gather exchange the stages are driven from the host (`rk4_step`)."""
        if self.comm is not None:
            from .network import _addr, _stream_handle
            from . import _cabi
            a_u, _, n_u = _addr(u)
            a_p, _, n_p = _addr(p)
            self.nw._check_sizes(n_u, n_u, n_p, p is not None)
            self.nw._auto_pack(p, a_p, stream)
            rc = _cabi.lib().nd_b200_rk4_exchange(self.nw.handle, self.comm, a_u, a_p, float(t0), float(dt), int(nsteps), _stream_handle(stream))
            if rc:
                self.nw._fail(rc)
            return
        work = {} if work is None else work
        for k in range(int(nsteps)):
            self.rk4_step(u, p, t0 + k * dt, dt, work)

    def rk4_step(self, u, p, t, dt, work):
        """one classical RK4 step on the owned states (same operation order as the single-GPU engine); `work` is a
        dict of scratch tensors reused across steps"""
        import torch
        k = work.setdefault("k", [torch.empty_like(u) for _ in range(4)])
        tmp = work.setdefault("tmp", torch.empty_like(u))
        h2 = 0.5 * dt
        self.rhs(k[0], u, p, t)
        for a, b in self.owned_segments:
            tmp[a:b] = u[a:b] + h2 * k[0][a:b]
        self.rhs(k[1], tmp, p, t + h2)
        for a, b in self.owned_segments:
            tmp[a:b] = u[a:b] + h2 * k[1][a:b]
        self.rhs(k[2], tmp, p, t + h2)
        for a, b in self.owned_segments:
            tmp[a:b] = u[a:b] + dt * k[2][a:b]
        self.rhs(k[3], tmp, p, t + dt)
        for a, b in self.owned_segments:
            u[a:b] = u[a:b] + (dt / 6.0) * (((k[0][a:b] + 2.0 * k[1][a:b]) + 2.0 * k[2][a:b]) + k[3][a:b])
