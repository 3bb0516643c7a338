"""`Network(g, vertexm, edgem; execution=B200Execution(), aggregator=B200Aggregator(+))` -- the reference's
constructor and call interface for the RHS path, with the B200 engine plugged into its two extension points.

Mirrors (reference file:line):
  * `Network(g, vertexm, edgem; execution, aggregator, ...)`      src/construction.jl:31-236
  * `IndexManager`, `register_vertices!`, `register_edges!`        src/network_structure.jl:1-55,224-289
  * batching by `_component_hash` / `find_identical`               src/construction.jl:238-256, src/utils.jl:197-217
  * `ExecutionStyle{buffered}` singletons                          src/executionstyles.jl:13-47
  * aggregator constructor convention `(im, batches) -> Aggregator`  src/aggregators.jl:1-16,137
  * `(nw::Network)(du, u, p, t)`                                    src/coreloop.jl:1-102
  * `get_buffers`                                                   src/coreloop.jl:103-109
The host side only builds index tables (numpy, vectorised); all arithmetic runs in libnd_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import os
import weakref

import numpy as np

from . import _cabi
from .components import ArgumentError, EdgeModel, VertexModel


# ---------------------------------------------------------------------------------------------------
# execution style / aggregator tags
# ---------------------------------------------------------------------------------------------------
class ExecutionStyle:
    buffered = False


class B200Execution(ExecutionStyle):
    """Field-less tag like the reference's execution styles (only its type is stored in `Network{EX,...}`,
    src/network_structure.jl:83,102-116).  `buffered=False`: the engine never materialises a gather buffer, so
    the Lazy provider is selected (src/construction.jl:210-214)."""
    buffered = False

    def __repr__(self):
        return "B200Execution{false}()"


def usebuffer(ex) -> bool:
    return bool(getattr(ex, "buffered", False))


def iscudacompatible(x) -> bool:
    return isinstance(x, (B200Execution, B200Aggregator)) or x in (B200Execution, B200Aggregator)


@dataclass
class ComponentBatch:
    """src/network_structure.jl:176-222 (strides flattened to first + width)."""
    kind: str                 # "vertex" | "edge"
    model: object
    indices: np.ndarray       # 1-based component ids, ascending
    state_first: int
    p_first: int
    in_first: int             # aggbuf (vertices) / gbuf (edges)
    out_first: int

    count: int = -1           # batches described without an index array (indices=None: components 1:count)

    def __len__(self):
        return int(self.indices.size) if self.indices is not None else int(self.count)


class IndexManager:
    """All flat index ranges (1-based firsts; widths come from the models)."""

    def __init__(self, g, vertexm: Sequence[VertexModel], vtype: np.ndarray, edgem: Sequence[EdgeModel],
                 etype: np.ndarray):
        self.g = g
        self.nv, self.ne = g.nv, g.ne
        self.edge_src, self.edge_dst = g.src, g.dst            # edgevec, src/network_structure.jl:37
        self.vertexm, self.edgem = list(vertexm), list(edgem)
        self.vtype, self.etype = vtype, etype
        vd = {m.outdim for m in self.vertexm}
        if len(vd) != 1:                                        # src/construction.jl:99-102
            raise ArgumentError("All vertices must have the same output dimension")
        ed = {m.outdim_dst for m in self.edgem}
        if len(ed) > 1:                                         # src/construction.jl:103-106
            raise ArgumentError("All edges must have the same output dimension")
        self.vdepth = vd.pop()
        self.edepth = ed.pop() if ed else 0
        self.lastidx_dynamic = self.lastidx_out = self.lastidx_p = 0
        self.lastidx_aggr = self.lastidx_gbuf = self.lastidx_extbuf = 0
        z = lambda n: np.zeros(n, dtype=np.int64)
        self.v_data, self.v_out, self.v_para, self.v_aggr = z(self.nv), z(self.nv), z(self.nv), z(self.nv)
        self.e_data, self.e_out_src, self.e_out_dst = z(self.ne), z(self.ne), z(self.ne)
        self.e_para, self.e_gbuf_src, self.e_gbuf_dst = z(self.ne), z(self.ne), z(self.ne)
        self.v_ext, self.e_ext = z(self.nv), z(self.ne)

    def _next(self, which: str, count: int, width: int) -> np.ndarray:
        last = getattr(self, which)
        setattr(self, which, last + count * width)
        return last + 1 + np.arange(count, dtype=np.int64) * width

    def register_vertices(self, idxs: np.ndarray, m: VertexModel) -> ComponentBatch:
        """src/network_structure.jl:224-239"""
        n = idxs.size
        i0 = idxs - 1
        self.v_data[i0] = self._next("lastidx_dynamic", n, m.dim)
        self.v_out[i0] = self._next("lastidx_out", n, m.outdim)
        self.v_para[i0] = self._next("lastidx_p", n, m.pdim)
        self.v_aggr[i0] = self._next("lastidx_aggr", n, self.edepth)
        self.v_ext[i0] = self._next("lastidx_extbuf", n, m.extdim)
        f = i0[0]
        return ComponentBatch("vertex", m, idxs, int(self.v_data[f]), int(self.v_para[f]), int(self.v_aggr[f]),
                              int(self.v_out[f]))

    def register_edges(self, idxs: np.ndarray, m: EdgeModel) -> ComponentBatch:
        """src/network_structure.jl:240-258: src output range first, dst range immediately after"""
        n = idxs.size
        i0 = idxs - 1
        self.e_data[i0] = self._next("lastidx_dynamic", n, m.dim)
        w = m.outdim_src + m.outdim_dst
        first = self._next("lastidx_out", n, w)
        self.e_out_src[i0] = first
        self.e_out_dst[i0] = first + m.outdim_src
        self.e_para[i0] = self._next("lastidx_p", n, m.pdim)
        gf = self._next("lastidx_gbuf", n, 2 * self.vdepth)
        self.e_gbuf_src[i0] = gf
        self.e_gbuf_dst[i0] = gf + self.vdepth
        self.e_ext[i0] = self._next("lastidx_extbuf", n, m.extdim)
        f = i0[0]
        return ComponentBatch("edge", m, idxs, int(self.e_data[f]), int(self.e_para[f]), int(self.e_gbuf_src[f]),
                              int(self.e_out_src[f]))


def resolve_extin(im: "IndexManager", ref) -> int:
    """One external-input reference -> the ExtMap entry the C ABI takes (src/external_inputs.jl:36-50): > 0 = 1-based
    index into u (a state), < 0 = -(1-based index into o) (an output)."""
    from .components import EIndex, VIndex
    if isinstance(ref, VIndex):
        m, data, out = im.vertexm[im.vtype[ref.comp - 1]], im.v_data[ref.comp - 1], im.v_out[ref.comp - 1]
        nout = m.outdim
    elif isinstance(ref, EIndex):
        m, data, out = im.edgem[im.etype[ref.comp - 1]], im.e_data[ref.comp - 1], im.e_out_src[ref.comp - 1]
        nout = m.outdim_src + m.outdim_dst
    else:
        raise ArgumentError(f"external input {ref!r} is neither a VIndex nor an EIndex")
    sub = ref.sub
    if isinstance(sub, str):
        if sub not in m.sym:
            raise ArgumentError(f"Cannot resolve external input {ref!r}: {sub!r} is not a state symbol of {m.name}")
        sub = m.sym.index(sub) + 1
    if isinstance(sub, (int, np.integer)):
        if not 1 <= sub <= m.dim:
            raise ArgumentError(f"Cannot resolve external input {ref!r}: state index outside 1..{m.dim}")
        return int(data + sub - 1)
    if isinstance(sub, tuple) and len(sub) == 2 and sub[0] == "out" and 1 <= sub[1] <= nout:
        return -int(out + sub[1] - 1)
    raise ArgumentError(f"Cannot resolve external input {ref!r}")


def find_identical(keys: np.ndarray) -> List[np.ndarray]:
    """src/utils.jl:197-217: groups in first-occurrence order, members ascending (1-based)."""
    uniq, first = np.unique(keys, return_index=True)
    order = np.argsort(first, kind="stable")
    return [np.nonzero(keys == uniq[k])[0].astype(np.int64) + 1 for k in order]


def _expand_models(models, n, what):
    """Accept one model (broadcast, src/construction.jl:42-46), a list of n models, or a
    (unique_models, type_index_array) pair for very large networks."""
    if isinstance(models, (VertexModel, EdgeModel)):
        return [models], np.zeros(n, dtype=np.int64), True
    if isinstance(models, tuple) and len(models) == 2 and isinstance(models[1], np.ndarray):
        uniq, types = list(models[0]), np.asarray(models[1], dtype=np.int64)
        if types.size != n:
            raise ArgumentError(f"Number of {what} models does not match the graph")
        # merge models with equal component hash (object identity is not what the reference batches on)
        hashes, remap = {}, np.zeros(len(uniq), dtype=np.int64)
        merged = []
        for i, m in enumerate(uniq):
            h = (m.component_hash(), m.extin)      # external-input references are per component, not part of the batch hash
            if h not in hashes:
                hashes[h] = len(merged)
                merged.append(m)
            remap[i] = hashes[h]
        return merged, remap[types], False
    models = list(models)
    if len(models) != n:                                        # src/construction.jl:48-51
        raise ArgumentError(f"Number of {what} models does not match the graph")
    hashes, uniq, types = {}, [], np.empty(n, dtype=np.int64)
    for i, m in enumerate(models):
        h = (m.component_hash(), m.extin)
        k = hashes.get(h)
        if k is None:
            k = hashes[h] = len(uniq)
            uniq.append(m)
        types[i] = k
    return uniq, types, False


# ---------------------------------------------------------------------------------------------------
# the aggregator that owns the engine
# ---------------------------------------------------------------------------------------------------
class HaloTimeoutError(RuntimeError):
    """multi-GPU: a rank gave up waiting for a peer's boundary outputs (ND_B200_ETIMEOUT); results since then are NaN"""


class B200Aggregator:
    """`B200Aggregator(+)` is, like every reference aggregator, a constructor closure: calling it with
    `(im, edgebatches)` after all batches are registered (src/construction.jl:198) builds the aggregator -- here
    the whole device engine, because the execution-style tag cannot carry state.  Keeps an `f` field
    (read as `nw.layer.aggregator.f`, test/testutils.jl:50)."""

    def __init__(self, f="+", *, device: Optional[int] = None, row_range=None, long_row_threshold: int = 0,
                 keep_tables: bool = True, host_only: bool = False, gather_offset=None, gather_len: int = 0,
                 edge_parameters: str = "auto"):
        if f not in ("+", sum, np.add) and getattr(f, "__name__", "") != "add":
            raise ArgumentError("B200Aggregator only supports + as the reducer (no CPU fallback for others)")
        if edge_parameters not in ("auto", "live"):
            raise ArgumentError("edge_parameters must be 'auto' or 'live'")
        self.f = "+"
        # edge_parameters: "live" = every call re-reads the edge parameters from the caller's p through per-entry offsets;
        # "auto" (default) = the same results, but when p is a device tensor that carries a modification counter (torch's
        # `_version`) and two consecutive calls see the same unmodified p, the engine's packed per-entry copy is refreshed
        # (nd_b200_pack_params) and used until p changes -- the explicit refresh of SURVEY.md section 7, made automatic.
        self.edge_parameters = edge_parameters
        # host_only: build the engine's tables without touching a device (layout tests); such a network cannot be called
        # gather_offset / gather_len: multi-GPU packed halo layout (nd_b200_desc.gather_offset), see distributed.py
        self._opts = dict(device=device, row_range=row_range, long_row_threshold=long_row_threshold,
                          keep_tables=keep_tables, host_only=host_only, gather_offset=gather_offset,
                          gather_len=gather_len)
        self.handle = None
        self._keep = None

    def __repr__(self):
        return "B200Aggregator(+)"

    # -- constructor-closure protocol ------------------------------------------------------------
    def __call__(self, im: IndexManager, edgebatches: List[ComponentBatch]):
        """`aggregator(im, edgebatches)`; the vertex batches are read from the fully populated IndexManager
        (in Julia: rebuilt from im.v_* and im.vertexm, which are complete at src/construction.jl:198)."""
        agg = B200Aggregator(self.f, edge_parameters=self.edge_parameters, **self._opts)
        agg._build(im, im.vertexbatches, edgebatches)
        return agg

    def _build(self, im, vertexbatches, edgebatches):
        L = self._L = _cabi.lib()           # the library that owns the handle (also the one that destroys it)
        vb = (_cabi.VBatch * len(vertexbatches))()
        keep = []
        customs = {}     # custom_spec -> kind id: user-supplied CUDA component functions (compiled by the engine, NVRTC)

        def custom_kind(model):
            spec = model.custom_spec()
            if spec is None:
                return None
            if spec not in customs:
                customs[spec] = _cabi.CUSTOM_KIND_BASE + len(customs)
            return customs[spec]

        def ext_table(cb, b, models, types):
            """ExtMap entries of a batch: the components of a batch share f (and extdim) but each has its own references"""
            if b.model.extdim == 0:
                return
            tab = np.empty((len(b), b.model.extdim), dtype=np.int64)
            for r, comp in enumerate(b.indices):
                refs = models[types[comp - 1]].extin
                tab[r] = [resolve_extin(im, ref) for ref in refs]
            keep.append(tab)
            cb.extdim = b.model.extdim
            cb.ext_src = tab.ctypes.data_as(_cabi.i64p)

        for k, b in enumerate(vertexbatches):
            kind = b.model.kernel_kind()
            if kind is None:
                kind = custom_kind(b.model)
            if kind is None:
                raise ArgumentError(f"vertex model {b.model.name!r} has no kernel in the B200 registry "
                                    "(unsupported components raise instead of falling back to the CPU)")
            idx = np.ascontiguousarray(b.indices, dtype=np.int64)
            keep.append(idx)
            vb[k] = _cabi.VBatch(kind, b.model.dim, b.model.pdim, b.model.outdim, idx.size,
                                 idx.ctypes.data_as(_cabi.i64p), b.state_first, b.p_first, b.out_first, b.in_first)
            ext_table(vb[k], b, im.vertexm, im.vtype)
        eb = (_cabi.EBatch * max(1, len(edgebatches)))()
        for k, b in enumerate(edgebatches):
            kind = b.model.kernel_kind()
            if kind is None:
                kind = custom_kind(b.model)
            if kind is None:
                raise ArgumentError(f"edge model {b.model.name!r} has no kernel in the B200 registry "
                                    "(unsupported components raise instead of falling back to the CPU)")
            idx = np.ascontiguousarray(b.indices, dtype=np.int64)
            keep.append(idx)
            masks = b.model.state_masks() or (0, 0)
            eb[k] = _cabi.EBatch(kind, b.model.coupling, b.model.dim, b.model.pdim, b.model.outdim_src,
                                 b.model.outdim_dst, idx.size, idx.ctypes.data_as(_cabi.i64p), b.state_first,
                                 b.p_first, b.out_first, b.in_first, masks[0], masks[1])
            ext_table(eb[k], b, im.edgem, im.etype)
        dev = self._opts["device"]
        if dev is None:
            dev = _current_device()
        rr = self._opts["row_range"] or (0, 0)
        src = np.ascontiguousarray(im.edge_src, dtype=np.int64)
        dst = np.ascontiguousarray(im.edge_dst, dtype=np.int64)
        desc = _cabi.Desc(_cabi.ABI_VERSION, int(dev), im.nv, im.ne, src.ctypes.data_as(_cabi.i64p),
                          dst.ctypes.data_as(_cabi.i64p), im.vdepth, im.edepth, len(vertexbatches),
                          len(edgebatches), vb, eb, im.lastidx_dynamic, im.lastidx_p, im.lastidx_out,
                          im.lastidx_aggr, int(rr[0]), int(rr[1]), int(self._opts["long_row_threshold"]),
                          (0 if self._opts["keep_tables"] else _cabi.FLAG_NO_EXPORT)
                          | (_cabi.FLAG_HOST_ONLY if self._opts["host_only"] else 0)
                          | (_cabi.FLAG_ROW_RANGE if self._opts["row_range"] is not None else 0),
                          None, 0)
        if customs:
            ck = (_cabi.CustomKind * len(customs))()
            for i, (spec, kid) in enumerate(customs.items()):
                role, dim, pdim, outdim, two_sided, f_body, g_body, extdim, g_ff = spec
                fb, gb = f_body.encode(), (g_body.encode() if g_body is not None else None)
                keep += [fb, gb]
                ck[i] = _cabi.CustomKind(kid, role, dim, pdim, outdim, two_sided, fb, gb, extdim, g_ff)
            keep.append(ck)
            desc.n_custom = len(customs)
            desc.custom = ck
        if self._opts["gather_offset"] is not None:
            go = np.ascontiguousarray(self._opts["gather_offset"], dtype=np.int64)
            if go.size != im.nv:
                raise ArgumentError("gather_offset needs one entry per vertex")
            keep.append(go)
            desc.gather_offset = go.ctypes.data_as(_cabi.i64p)
            desc.gather_len = int(self._opts["gather_len"])
        keep += [vb, eb, src, dst]
        self._desc, self._keep = desc, keep          # kept alive with the engine (tests mutate copies of it)
        h = C.c_void_p()
        rc = L.nd_b200_create(C.byref(desc), C.byref(h))
        if rc != _cabi.OK:
            msg = L.nd_b200_last_error(None).decode()
            raise (ArgumentError if rc in (_cabi.EINVAL, _cabi.EUNSUPPORTED) else RuntimeError)(msg)
        self.handle = h
        self.device = int(dev)
        self._sizes = (im.lastidx_out, im.lastidx_aggr)

    def aggregate(self, aggbuf, o, *, stream=None):
        """`aggregate!(aggregator, aggbuf, o)` (src/aggregators.jl:140-151): add the edge-output block of the device
        vector `o` into the device vector `aggbuf`, per slot in ascending `o` order, on top of its present content."""
        a_a, dev_a, n_a = _addr(aggbuf)
        a_o, dev_o, n_o = _addr(o)
        if not (dev_a and dev_o):
            raise ArgumentError("aggregate needs device-resident aggbuf and o")
        if (n_a, n_o) != (self._sizes[1], self._sizes[0]):
            raise ArgumentError(f"aggregate: aggbuf / o have {n_a} / {n_o} entries, expected {self._sizes[1]} / {self._sizes[0]}")
        rc = self._L.nd_b200_aggregate(self.handle, a_a, a_o, _stream_handle(stream))
        if rc:
            msg = self._L.nd_b200_last_error(self.handle).decode()
            raise (ArgumentError if rc in (_cabi.EINVAL, _cabi.EUNSUPPORTED) else RuntimeError)(msg)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                (getattr(self, "_L", None) or _cabi.lib()).nd_b200_destroy(h)
            except Exception:
                pass


def get_aggr_constructor(agg: B200Aggregator):
    """src/aggregators.jl:287-292: recover the constructor closure from a built aggregator."""
    return B200Aggregator(agg.f, edge_parameters=agg.edge_parameters, **agg._opts)


def _current_device() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def _addr(x):
    """raw address + (is_device, nbytes/8) of a flat float64 buffer: torch tensor, numpy array, CUDA array
    interface object, or None."""
    if x is None:
        return None, None, 0
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64 or not x.flags.c_contiguous:
            raise ArgumentError("expected a contiguous Float64 vector")
        return x.ctypes.data, False, x.size
    if hasattr(x, "data_ptr"):  # torch
        import torch
        if x.dtype != torch.float64 or not x.is_contiguous():
            raise ArgumentError("expected a contiguous Float64 vector")
        return x.data_ptr(), bool(x.is_cuda), x.numel()
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is not None:
        if cai["typestr"] not in ("<f8", "=f8"):
            raise ArgumentError("expected a Float64 device vector")
        return cai["data"][0], True, int(np.prod(cai["shape"]))
    raise ArgumentError(f"unsupported array type {type(x)}")


def _stream_handle(stream) -> Optional[int]:
    if stream is None:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.cuda.current_stream().cuda_stream or None
        except Exception:
            return None
        return None
    return getattr(stream, "cuda_stream", stream) or None


class NetworkLayer:
    """src/network_structure.jl:57-75 (only what the RHS path reads)"""

    def __init__(self, g, edgebatches, aggregator, edepth, vdepth):
        self.g, self.edgebatches, self.aggregator = g, edgebatches, aggregator
        self.edepth, self.vdepth = edepth, vdepth


class Network:
    """`Network(g, vertexm, edgem; execution=B200Execution(), aggregator=B200Aggregator("+"))`."""

    def __init__(self, g, vertexm, edgem, *, execution=None, aggregator=None, verbose=False):
        execution = B200Execution() if execution is None else execution
        aggregator = B200Aggregator("+") if aggregator is None else aggregator
        if not isinstance(execution, ExecutionStyle):                       # src/construction.jl:96
            raise ArgumentError("execution must be an ExecutionStyle")
        vm, vtype, _ = _expand_models(vertexm, g.nv, "vertex")
        em, etype, _ = _expand_models(edgem, g.ne, "edge")
        for m in vm:
            if not isinstance(m, VertexModel):
                raise ArgumentError("vertexm must be VertexModel(s)")
        for m in em:
            if not isinstance(m, EdgeModel):
                raise ArgumentError("edgem must be EdgeModel(s)")
        im = IndexManager(g, vm, vtype, em, etype)
        self.im = im
        # batches: all vertex batches first, then all edge batches (src/construction.jl:171-195)
        # batches are formed on the component hash (src/construction.jl:245-256); models that differ only in WHAT their
        # external inputs refer to share a batch
        def batch_keys(models, types):
            ids = {}
            per_model = np.array([ids.setdefault(m.component_hash(), len(ids)) for m in models], dtype=np.int64)
            return per_model[types]
        self.vertexbatches = [im.register_vertices(idxs, vm[vtype[idxs[0] - 1]]) for idxs in find_identical(batch_keys(vm, vtype))]
        edgebatches = [im.register_edges(idxs, em[etype[idxs[0] - 1]]) for idxs in find_identical(batch_keys(em, etype))] if g.ne else []
        if verbose:
            for b in self.vertexbatches + edgebatches:
                print(f" - {b.kind} batch {b.model.name}: {len(b)} components")
        im.vertexbatches = self.vertexbatches
        agg = aggregator(im, edgebatches)                                   # src/construction.jl:198
        self.layer = NetworkLayer(g, edgebatches, agg, im.edepth, im.vdepth)
        self.execution = execution
        self._L = _cabi.lib()

    @classmethod
    def from_edgelist(cls, g, vertexm: VertexModel, edgem: EdgeModel, *, device: Optional[int] = None, row_range=None,
                      keep_tables: bool = False, host_only: bool = False, gather_offset=None, gather_len: int = 0,
                      layout_only: bool = False, edge_parameters: str = "auto"):
        """Homogeneous network straight from the edge list (`nd_b200_create_from_edgelist`): one registry
        vertex model, one registry edge model.  Same flat `u` / `p` layout and same engine as `Network(g, vertexm, edgem)`,
        without the per-component host tables (BASELINE config 5 has 4e8 edges: six Int64 tables would be 19 GB)."""
        from types import SimpleNamespace
        vk, ek = vertexm.kernel_kind(), edgem.kernel_kind()
        if vk is None or ek is None:
            raise ArgumentError("from_edgelist needs registry models (no CPU fallback)")
        self = cls.__new__(cls)
        self._L = L = None if layout_only else _cabi.lib()
        nv, ne = g.nv, g.ne
        src = np.ascontiguousarray(g.src, dtype=np.int64)
        dst = np.ascontiguousarray(g.dst, dtype=np.int64)
        osrc = edgem.outdim_src if ne else 0
        edepth = edgem.outdim_dst if ne else vertexm.outdim
        self.im = SimpleNamespace(nv=nv, ne=ne, vdepth=vertexm.outdim, edepth=edepth, homogeneous=True, edge_src=src, edge_dst=dst,
                                  vdim=vertexm.dim,
                                  lastidx_dynamic=nv * vertexm.dim, lastidx_p=nv * vertexm.pdim + ne * edgem.pdim,
                                  lastidx_out=nv * vertexm.outdim + ne * (osrc + edgem.outdim_dst), lastidx_aggr=nv * edepth)
        dev = _current_device() if device is None else device
        rr = row_range or (0, 0)
        flags = (0 if keep_tables else _cabi.FLAG_NO_EXPORT) | (_cabi.FLAG_HOST_ONLY if host_only else 0) \
            | (_cabi.FLAG_ROW_RANGE if row_range is not None else 0)
        h = C.c_void_p()
        go = None
        if gather_offset is not None:           # row-partitioned engine with a packed halo (distributed.py)
            go = np.ascontiguousarray(gather_offset, dtype=np.int64)
            if go.size != nv:
                raise ArgumentError("gather_offset needs one entry per vertex")
        agg = None
        if not layout_only:     # layout_only: sizes and batches only, no engine (what distributed.py plans the partition on)
            rc = L.nd_b200_create_from_edgelist(int(dev), nv, ne, src.ctypes.data_as(_cabi.i64p), dst.ctypes.data_as(_cabi.i64p),
                                                vk, ek, edgem.coupling, int(rr[0]), int(rr[1]), flags,
                                                go.ctypes.data_as(_cabi.i64p) if go is not None else None, int(gather_len), C.byref(h))
            if rc != _cabi.OK:
                msg = L.nd_b200_last_error(None).decode()
                raise (ArgumentError if rc in (_cabi.EINVAL, _cabi.EUNSUPPORTED) else RuntimeError)(msg)
            agg = B200Aggregator("+", device=dev, row_range=row_range, keep_tables=keep_tables, host_only=host_only,
                                 edge_parameters=edge_parameters)
            agg.handle, agg.device, agg._L = h, int(dev), L
            agg._sizes = (self.im.lastidx_out, self.im.lastidx_aggr)
        # the one vertex batch / edge batch of the layout (what register_vertices! / register_edges! would return), without
        # per-component index arrays: indices=None stands for 1:n
        self.vertexbatches = [ComponentBatch("vertex", vertexm, None, 1, 1, 1, 1)]
        ebatches = [ComponentBatch("edge", edgem, None, nv * vertexm.dim + 1, nv * vertexm.pdim + 1, 1, nv * vertexm.outdim + 1)] if ne else []
        self.vertexbatches[0].count = nv
        for b in ebatches:
            b.count = ne
        self.layer = NetworkLayer(g, ebatches, agg, edepth, vertexm.outdim)
        self.execution = B200Execution()
        return self

    # -- sizes (src/network_structure.jl:118-131) ---------------------------------------------------
    def dim(self):
        return self.im.lastidx_dynamic

    def pdim(self):
        return self.im.lastidx_p

    @property
    def handle(self):
        agg = self.layer.aggregator
        return None if agg is None else agg.handle

    def _fail(self, rc):
        msg = self._L.nd_b200_last_error(self.handle).decode()
        if rc == _cabi.ETIMEOUT:
            raise HaloTimeoutError(msg)
        raise (ArgumentError if rc in (_cabi.EINVAL, _cabi.EUNSUPPORTED) else RuntimeError)(msg)

    def _check_sizes(self, n_du, n_u, n_p, has_p):
        if n_du != self.im.lastidx_dynamic or n_u != self.im.lastidx_dynamic:   # src/coreloop.jl:2-4
            raise ArgumentError(f"du or u does not have expected size {self.im.lastidx_dynamic}")
        if self.im.lastidx_p > 0 and (not has_p or n_p != self.im.lastidx_p):    # src/coreloop.jl:5-7
            raise ArgumentError(f"p does not has expecte size {self.im.lastidx_p}")

    # -- the RHS -----------------------------------------------------------------------------------
    def __call__(self, du, u, p, t, *, stream=None, perturb=None, RET="du"):
        """In-place `nw(du, u, p, t)`; returns None (src/coreloop.jl:101).  Device vectors run asynchronously on
        the current stream; numpy (host) vectors take the end-to-end host path (H2D, RHS, D2H, synchronise)."""
        if perturb is not None or RET != "du":
            raise ArgumentError("B200Execution supports neither `perturb` nor RET != :du; keep a CPU twin network "
                                "for initialisation / linear analysis")
        a_du, dev_du, n_du = _addr(du)
        a_u, dev_u, n_u = _addr(u)
        a_p, dev_p, n_p = _addr(p)
        self._check_sizes(n_du, n_u, n_p, p is not None)
        kinds = {d for d in (dev_du, dev_u, dev_p) if d is not None}
        if len(kinds) != 1:
            raise ArgumentError("du, u and p must all live on the device or all on the host")
        if kinds.pop():
            self._auto_pack(p, a_p, stream)
            rc = self._L.nd_b200_rhs(self.handle, a_du, a_u, a_p, float(t), _stream_handle(stream))
        else:
            rc = self._L.nd_b200_rhs_host(self.handle, a_du, a_u, a_p, float(t))
        if rc:
            self._fail(rc)
        return None

    def _auto_pack(self, p, a_p, stream):
        """edge_parameters="auto": keep the engine's packed copy of the edge parameters in step with p.  Reference
        semantics are preserved -- p may be changed between calls (callbacks, docs/examples/cascading_failure.jl:107-110);
        every in-place change of a torch tensor bumps its `_version`, which invalidates the packed copy.  The copy is only
        (re)built when the SAME unmodified p is seen on two consecutive calls (an integrator's stages), so a parameter
        vector that changes on every call never pays for packing."""
        st = self.__dict__.setdefault("_pack_state", {"manual": False, "packable": True, "packed": None, "last": None})
        if st["manual"] or not st["packable"] or p is None:
            return
        if getattr(self.layer.aggregator, "edge_parameters", "live") != "auto":
            return
        ver = getattr(p, "_version", None)
        if ver is None:
            return
        # identity = the tensor OBJECT (held weakly: a new tensor that the caching allocator places at the same address
        # with the same counter is a different p), its address and its modification counter
        key = (a_p, int(ver))
        try:
            same_obj = st.get("ref") is not None and st["ref"]() is p
            if not same_obj:
                st["ref"] = weakref.ref(p)
        except TypeError:
            return
        if not same_obj:                # another parameter vector: forget the packed copy, start counting afresh
            if st["packed"] is not None:
                self._L.nd_b200_pack_params(self.handle, None, _stream_handle(stream))
                st["packed"] = None
            st["last"] = key
            return
        if st["packed"] == key:
            pass
        elif st["last"] == key:
            rc = self._L.nd_b200_pack_params(self.handle, a_p, _stream_handle(stream))
            if rc:                      # no packed kernels for this network (several edge batches, user-supplied kinds ...)
                st["packable"] = False
            else:
                st["packed"] = key
        elif st["packed"] is not None:
            self._L.nd_b200_pack_params(self.handle, None, _stream_handle(stream))
            st["packed"] = None
        st["last"] = key

    def get_buffers(self, o, aggbuf, u, p, t, *, stream=None):
        """`get_buffers(nw, u, p, t)` (src/coreloop.jl:103-109) into caller-provided device vectors."""
        a_o, dev_o, n_o = _addr(o)
        a_a, dev_a, n_a = _addr(aggbuf)
        a_u, dev_u, n_u = _addr(u)
        a_p, dev_p, n_p = _addr(p)
        if any(d is False for d in (dev_o, dev_a, dev_u, dev_p)):
            raise ArgumentError("get_buffers needs device-resident o, aggbuf, u and p")
        if n_o != self.im.lastidx_out or n_a != self.im.lastidx_aggr:
            raise ArgumentError("output / aggregation buffer has the wrong size")
        self._check_sizes(n_u, n_u, n_p, p is not None)
        self._auto_pack(p, a_p, stream)
        rc = self._L.nd_b200_get_buffers(self.handle, a_o, a_a, a_u, a_p, float(t), _stream_handle(stream))
        if rc:
            self._fail(rc)

    def rk4(self, u, p, t0, dt, nsteps, *, stream=None):
        """fixed-step classical RK4, u advanced in place on the device"""
        a_u, dev, n_u = _addr(u)
        a_p, _, n_p = _addr(p)
        if not dev:
            raise ArgumentError("rk4 needs device-resident u")
        self._check_sizes(n_u, n_u, n_p, p is not None)
        self._auto_pack(p, a_p, stream)
        rc = self._L.nd_b200_rk4(self.handle, a_u, a_p, float(t0), float(dt), int(nsteps), _stream_handle(stream))
        if rc:
            self._fail(rc)

    def pack_params(self, p, *, stream=None):
        """Optional contract (nd_b200_pack_params): copy the edge parameters of device vector `p` into the engine's
        per-entry array; until the next call the kernels read edge parameters from that copy (coalesced) and the caller
        must not change them in `p`.  `pack_params(None)` returns to re-reading `p` on every call (the default)."""
        a_p, dev, n_p = _addr(p)
        if p is not None and (not dev or n_p != self.im.lastidx_p):
            raise ArgumentError(f"pack_params needs a device vector of size {self.im.lastidx_p}")
        rc = self._L.nd_b200_pack_params(self.handle, a_p, _stream_handle(stream))
        if rc:
            self._fail(rc)
        # an explicit call takes the automatic refresh out of the picture: the caller owns the contract from here on
        st = self.__dict__.setdefault("_pack_state", {"manual": False, "packable": True, "packed": None, "last": None})
        st["manual"], st["packed"], st["last"] = True, None, None

    # -- introspection -----------------------------------------------------------------------------
    def engine_sizes(self):
        s = (C.c_int64 * 8)()
        rc = self._L.nd_b200_export_sizes(self.handle, s)
        if rc:
            self._fail(rc)
        keys = ["nrows", "nentries", "nblocks", "n_long_rows", "gather_from_u", "launches_per_rhs", "row_begin",
                "row_end"]
        return dict(zip(keys, [int(x) for x in s]))

    def export_tables(self):
        sz = self.engine_sizes()
        rowptr = np.zeros(sz["nrows"] + 1, dtype=np.int64)
        n = max(sz["nentries"], 1)
        nbr, eid, side = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int32)
        rc = self._L.nd_b200_export_tables(self.handle, rowptr.ctypes.data_as(_cabi.i64p),
                                           nbr.ctypes.data_as(_cabi.i64p), eid.ctypes.data_as(_cabi.i64p),
                                           side.ctypes.data_as(_cabi.i32p))
        if rc:
            self._fail(rc)
        k = sz["nentries"]
        return rowptr, nbr[:k], eid[:k], side[:k]

    def kernel_name(self) -> str:
        """name of the kernel family that evaluates this network (rhs_jag_kernel unless ND_B200_KERNEL selects a tile kernel)"""
        return self._L.nd_b200_kernel_name(self.handle).decode()

    def custom_source(self) -> Optional[str]:
        """the CUDA source the engine generated for user-supplied component kinds (None for registry-only networks)"""
        src = self._L.nd_b200_custom_source(self.handle)
        return src.decode() if src else None

    def export_jag(self):
        """the jagged device layout of the default kernel (host_only engines): slices[n,4], lanes[n,32], longs[m,4],
        order[entries] -- see include/nd_b200.h"""
        sz = np.zeros(6, dtype=np.int64)
        rc = self._L.nd_b200_export_jag_sizes(self.handle, sz.ctypes.data_as(_cabi.i64p))
        if rc:
            self._fail(rc)
        ns, nl, split, ne, wait_from, halo_len = (int(v) for v in sz)
        if ns < 0:
            raise ArgumentError("engine does not use the jagged layout")
        slices = np.zeros((max(ns, 1), 4), dtype=np.int32)
        lanes = np.zeros((max(ns, 1), 32), dtype=np.uint16)
        longs = np.zeros((max(nl, 1), 4), dtype=np.int32)
        order = np.zeros(max(ne, 1), dtype=np.int32)
        rc = self._L.nd_b200_export_jag(self.handle, slices.ctypes.data_as(_cabi.i32p),
                                        lanes.ctypes.data_as(C.POINTER(C.c_uint16)), longs.ctypes.data_as(_cabi.i32p),
                                        order.ctypes.data_as(_cabi.i32p))
        if rc:
            self._fail(rc)
        return dict(slices=slices[:ns], lanes=lanes[:ns], longs=longs[:nl], order=order[:ne], split=split,
                    wait_from=wait_from, halo_len=halo_len)

    def launch_count(self) -> int:
        return int(self._L.nd_b200_launch_count(self.handle))

    def set_timing(self, enabled: bool):
        self._L.nd_b200_set_timing(self.handle, int(enabled))

    def timings(self):
        f, pre, n = C.c_double(), C.c_double(), C.c_int64()
        rc = self._L.nd_b200_timings(self.handle, C.byref(f), C.byref(pre), C.byref(n))
        if rc:
            self._fail(rc)
        return dict(fused_ms=f.value, prepass_ms=pre.value, ncalls=int(n.value))


def dim(nw: Network) -> int:
    return nw.dim()


def pdim(nw: Network) -> int:
    return nw.pdim()


def pinned_empty(n: int) -> np.ndarray:
    """float64 host vector in page-locked memory (for the host-buffer RHS path)."""
    L = _cabi.lib()
    ptr = L.nd_b200_host_alloc(int(n) * 8)
    if not ptr:
        raise RuntimeError("cudaHostAlloc failed")
    buf = (C.c_double * int(n)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float64, count=int(n))
    _PINNED[arr.ctypes.data] = ptr
    return arr


_PINNED = {}
