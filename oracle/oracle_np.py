"""Slow pure-Python/numpy twin of the oracle (TEST INFRASTRUCTURE ONLY).

An independent second restatement of the same reference code, written with explicit
per-component loops and 1-based ranges, used to cross-check nd_oracle.c on small graphs.
Citations are relative to the NetworkDynamics.jl tree.
"""
from __future__ import annotations

import math

import numpy as np

from . import oracle as O


class Rng:
    """1-based inclusive UnitRange first:last"""
    __slots__ = ("first", "last")

    def __init__(self, first, last):
        self.first, self.last = first, last

    def idx(self):  # 0-based numpy indices
        return np.arange(self.first - 1, self.last)

    def __len__(self):
        return self.last - self.first + 1

    def __eq__(self, o):
        return (self.first, self.last) == (o.first, o.last)

    def __repr__(self):
        return f"{self.first}:{self.last}"


def find_identical(v):
    """src/utils.jl:197-217"""
    idxs_per_type, unique = [], []
    for i, x in enumerate(v):
        for j, y in enumerate(unique):
            if x == y:
                idxs_per_type[j].append(i + 1)
                break
        else:
            unique.append(x)
            idxs_per_type.append([i + 1])
    return idxs_per_type


class IndexManager:
    """src/network_structure.jl:1-55,224-289 and src/construction.jl:154-195"""

    def __init__(self, nv, esrc, edst, vspecs, vtype, especs, etype):
        self.nv, self.ne = nv, len(esrc)
        self.edgevec = list(zip([int(s) for s in esrc], [int(d) for d in edst]))
        self.vspecs, self.especs = vspecs, especs
        self.vdepth = vspecs[vtype[0]].outdim
        self.edepth = especs[etype[0]].outdim_dst if self.ne else 0
        self.last = dict(dynamic=0, out=0, p=0, aggr=0, gbuf=0, ext=0)
        self.v_data, self.v_out, self.v_para, self.v_aggr, self.v_ext = ({} for _ in range(5))
        self.e_data, self.e_out, self.e_para, self.e_gbufr, self.e_ext = ({} for _ in range(5))
        self.vbatches = [(vtype[idxs[0] - 1], idxs) for idxs in find_identical(list(vtype))]
        self.ebatches = [(etype[idxs[0] - 1], idxs) for idxs in find_identical(list(etype))]
        for spec, idxs in self.vbatches:
            s = vspecs[spec]
            for i in idxs:
                self.v_data[i] = self._next("dynamic", s.dim)
                self.v_out[i] = self._next("out", s.outdim)
                self.v_para[i] = self._next("p", s.pdim)
                self.v_aggr[i] = self._next("aggr", self.edepth)
                self.v_ext[i] = self._next("ext", getattr(s, "extdim", 0))       # src/network_structure.jl:230
        for spec, idxs in self.ebatches:
            s = especs[spec]
            for i in idxs:
                self.e_data[i] = self._next("dynamic", s.dim)
                self.e_out[i] = (self._next("out", s.outdim_src), self._next("out", s.outdim_dst))
                self.e_para[i] = self._next("p", s.pdim)
                self.e_gbufr[i] = (self._next("gbuf", self.vdepth), self._next("gbuf", self.vdepth))
                self.e_ext[i] = self._next("ext", getattr(s, "extdim", 0))       # src/network_structure.jl:249

    def _next(self, which, n):
        last = self.last[which]
        self.last[which] = last + n
        return Rng(last + 1, last + n)

    def gbuf_map(self):
        """src/gbufs.jl:13-18"""
        m = np.zeros(self.last["gbuf"], dtype=np.int64)
        for i, (s, d) in enumerate(self.edgevec, start=1):
            m[self.e_gbufr[i][0].idx()] = self.v_out[s].idx() + 1
            m[self.e_gbufr[i][1].idx()] = self.v_out[d].idx() + 1
        return m

    def aggregation_map(self):
        """src/aggregators.jl:58-86 -> (first (1-based), map)"""
        m = np.zeros(self.last["out"], dtype=np.int64)
        for _, idxs in self.ebatches:
            for eidx in idxs:
                s, d = self.edgevec[eidx - 1]
                m[self.e_out[eidx][1].idx()] = self.v_aggr[d].idx() + 1
                if len(self.e_out[eidx][0]) > 0:
                    m[self.e_out[eidx][0].idx()] = self.v_aggr[s].idx() + 1
        nz = np.nonzero(m)[0]
        if nz.size == 0:
            return 1, m[:0]
        return int(nz[0]) + 1, m[nz[0]:nz[-1] + 1]


class PyKind:
    """A component kind given as host callables (tests of user-supplied CUDA kinds): for vertices f(v, esum, p, t) -> dv
    and optionally g(v, p, t) -> out (None = StateMask(1:outdim)); for edges g(v_src, v_dst, p, t) -> e_dst, or with the
    Fiducial wrapper -> (e_src, e_dst) (src/component_functions.jl:189-203); for edges with states
    f(e, v_src, v_dst, p, t) -> de (their outputs are StateMasks, ESpec.mask_src / mask_dst).  Components with external
    inputs (spec.extdim > 0) take them right after the inputs: f(v, esum, ext, p, t) / f(e, v_src, v_dst, ext, p, t)."""

    def __init__(self, f=None, g=None):
        self.f, self.g = f, g


def _vertex_g(kind, u, p, t=0.0, ins=None):
    if isinstance(kind, PyKind):
        if kind.g is None:
            return None
        return [float(x) for x in (kind.g(u, ins, p, t) if ins is not None else kind.g(u, p, t))]
    if kind == O.V_SWING_DQ:
        return [p[3] * math.cos(u[0]), p[3] * math.sin(u[0])]
    return None  # StateMask handled by caller


def _edge_g(kind, vs, vd, p, t=0.0):
    if isinstance(kind, PyKind):
        return kind.g(vs, vd, p, t)
    if kind == O.E_DIFFUSION:
        return [p[0] * (vs[0] - vd[0])]
    if kind == O.E_DIFFUSION_NOP:
        return [vs[0] - vd[0]]
    if kind == O.E_KURAMOTO:
        return [p[0] * math.sin(vs[0] - vd[0])]
    if kind == O.E_LOOPBACK:                               # src/post_utils.jl:105-108
        return [-1.0 * x for x in vs]
    if kind == O.E_DIFFUSION_FID:                          # two-sided: (e_src, e_dst)
        ed = p[0] * (vs[0] - vd[0])
        return [-ed], [ed]
    if kind == O.E_LINE_DQ:
        R, X, active = p
        dr, di = vs[0] - vd[0], vs[1] - vd[1]
        den = R * R + X * X
        return [active * ((R * dr + X * di) / den), active * ((R * di - X * dr) / den)]
    raise ValueError(kind)


def _edge_f(kind, e, vs, vd, p, t=0.0, ext=None):
    """f of edges with states: test/ComponentLibrary.jl:30-34, test/diffusion_test.jl:96-100"""
    if isinstance(kind, PyKind):
        return [float(x) for x in (kind.f(e, vs, vd, ext, p, t) if ext is not None else kind.f(e, vs, vd, p, t))]
    if kind == O.E_DIFFUSION_ODE:
        tau = p[0]
        return [1.0 / tau * (math.sin(vs[0] - vd[0]) - e[0]), 1.0 / tau * (math.sin(vd[0] - vs[0]) - e[1])]
    if kind == O.E_RELAX_ODE:
        return [vs[0] - vd[0] - e[0], vd[0] - vs[0] - e[1]]
    raise ValueError(kind)


def _vertex_f(kind, v, acc, p, t=0.0, ext=None):
    if isinstance(kind, PyKind):
        return [float(x) for x in (kind.f(v, acc, ext, p, t) if ext is not None else kind.f(v, acc, p, t))]
    if kind == O.V_DIFFUSION:
        return [acc[0]]
    if kind == O.V_KURAMOTO_FIRST:
        return [p[0] + acc[0]]
    if kind == O.V_KURAMOTO_SECOND:
        M, D, Pm = p
        return [v[1], 1.0 / M * (Pm - D * v[1] + acc[0])]
    if kind == O.V_KURAMOTO_SECOND_BENCH:
        x = p[0] - 1.0 * v[1]
        x += acc[0]
        return [v[1], x]
    if kind == O.V_SWING_DQ:
        M, D, Pmech, V = p
        ur, ui = V * math.cos(v[0]), V * math.sin(v[0])
        Pel = ur * acc[0] + ui * acc[1]
        return [v[1], 1.0 / M * (Pmech + (-D * v[1]) + Pel)]
    raise ValueError(kind)


def rhs(im: IndexManager, u, p, t=0.0, extmap=None):
    """src/coreloop.jl:1-102 with SequentialExecution{true} + SequentialAggregator(+).  `extmap`: the ExtMap
    (src/external_inputs.jl:1-34) as one signed index per slot of the external-input buffer: > 0 = StateBufIdx (1-based
    index into u), < 0 = -(OutBufIdx) (1-based index into o)."""
    u = [float(x) for x in u]
    p = [float(x) for x in p] if p is not None else []
    du = [0.0] * im.last["dynamic"]
    o = [float("nan")] * im.last["out"]
    aggbuf = [0.0] * im.last["aggr"]
    sl = lambda a, r: a[r.first - 1:r.last]
    for spec, idxs in im.vbatches:  # PASS 1: g of vertices without feed forward
        s = im.vspecs[spec]
        if getattr(s, "ff", False):
            continue
        for i in idxs:
            out = _vertex_g(s.kind, sl(u, im.v_data[i]), sl(p, im.v_para[i]), t)
            if out is None:
                out = sl(u, im.v_data[i])[:s.outdim]
            o[im.v_out[i].first - 1:im.v_out[i].last] = out
    for spec, idxs in im.ebatches:  # PASS 2: g of edges without ff = StateMasks of the edge's own states
        s = im.especs[spec]
        if s.dim == 0:
            continue
        for i in idxs:
            ue = sl(u, im.e_data[i])
            odst = ue[s.mask_dst - 1:s.mask_dst - 1 + s.outdim_dst]
            o[im.e_out[i][1].first - 1:im.e_out[i][1].last] = odst
            if s.coupling == O.ANTISYMMETRIC:
                o[im.e_out[i][0].first - 1:im.e_out[i][0].last] = [-x for x in odst]
            elif s.coupling == O.SYMMETRIC:
                o[im.e_out[i][0].first - 1:im.e_out[i][0].last] = list(odst)
            elif s.coupling == O.FIDUCIAL:
                o[im.e_out[i][0].first - 1:im.e_out[i][0].last] = ue[s.mask_src - 1:s.mask_src - 1 + s.outdim_src]
    for spec, idxs in im.ebatches:  # apply_loopback!, src/coreloop.jl:47 + src/post_utils.jl:213-234
        if im.especs[spec].kind != O.E_LOOPBACK:
            continue
        for i in idxs:
            src, dst = im.edgevec[i - 1]
            aggbuf[im.v_aggr[src].first - 1:im.v_aggr[src].last] = sl(o, im.v_out[dst])
    for spec, idxs in im.vbatches:  # PASS 3: g of feed-forward vertices (injectors), src/coreloop.jl:55
        s = im.vspecs[spec]
        if not getattr(s, "ff", False):
            continue
        for i in idxs:
            o[im.v_out[i].first - 1:im.v_out[i].last] = _vertex_g(s.kind, sl(u, im.v_data[i]), sl(p, im.v_para[i]), t, sl(aggbuf, im.v_aggr[i]))
    extbuf = [float("nan")] * im.last["ext"]        # collect_externals!, src/coreloop.jl:61 + src/external_inputs.jl:52-66
    if im.last["ext"]:
        assert extmap is not None and len(extmap) == im.last["ext"]
        for k, src in enumerate(extmap):
            extbuf[k] = u[src - 1] if src > 0 else o[-src - 1]
    gmap = im.gbuf_map()
    gbuf = [o[k - 1] for k in gmap]  # gather!
    for spec, idxs in im.ebatches:  # PASS 4: f of edges without ff
        s = im.especs[spec]
        if s.dim == 0:
            continue
        for i in idxs:
            vs, vd = sl(gbuf, im.e_gbufr[i][0]), sl(gbuf, im.e_gbufr[i][1])
            ext = sl(extbuf, im.e_ext[i]) if getattr(s, "extdim", 0) else None
            du[im.e_data[i].first - 1:im.e_data[i].last] = _edge_f(s.kind, sl(u, im.e_data[i]), vs, vd, sl(p, im.e_para[i]), t, ext)
    for spec, idxs in im.ebatches:  # PASS 5
        s = im.especs[spec]
        if s.dim != 0:
            continue
        for i in idxs:
            vs, vd = sl(gbuf, im.e_gbufr[i][0]), sl(gbuf, im.e_gbufr[i][1])
            odst = _edge_g(s.kind, vs, vd, sl(p, im.e_para[i]), t)
            if s.coupling == O.FIDUCIAL:                       # two-sided g writes both outputs itself
                osrc, odst = odst
                o[im.e_out[i][0].first - 1:im.e_out[i][0].last] = [float(x) for x in osrc]
            odst = [float(x) for x in odst]
            o[im.e_out[i][1].first - 1:im.e_out[i][1].last] = odst
            if s.coupling == O.ANTISYMMETRIC:
                o[im.e_out[i][0].first - 1:im.e_out[i][0].last] = [-x for x in odst]
            elif s.coupling == O.SYMMETRIC:
                o[im.e_out[i][0].first - 1:im.e_out[i][0].last] = list(odst)
    first, amap = im.aggregation_map()  # aggregate!
    for k, dst in enumerate(amap):
        if dst != 0:
            aggbuf[dst - 1] = aggbuf[dst - 1] + o[first - 1 + k]
    for spec, idxs in im.vbatches:  # PASS 6
        s = im.vspecs[spec]
        for i in idxs:
            ext = sl(extbuf, im.v_ext[i]) if getattr(s, "extdim", 0) else None
            dv = _vertex_f(s.kind, sl(u, im.v_data[i]), sl(aggbuf, im.v_aggr[i]), sl(p, im.v_para[i]), t, ext)
            du[im.v_data[i].first - 1:im.v_data[i].last] = dv
    return np.array(du), np.array(o), np.array(aggbuf)
