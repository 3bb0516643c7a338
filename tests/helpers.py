"""Shared test helpers: build the oracle network that corresponds to a product-side model assignment, the parity
metric of SURVEY.md section 8d, and the benchmark-style random inputs."""
import numpy as np

from oracle import oracle as O


def vspec_of(m):
    kind = m.kernel_kind()
    return O.VSpec(O.V_OPAQUE if kind is None else kind, m.dim, m.pdim, m.outdim)


def espec_of(m):
    kind = m.kernel_kind()
    coupling = m.coupling if m.coupling is not None else O.FIDUCIAL
    masks = m.state_masks() or (0, 0)
    return O.ESpec(O.E_OPAQUE if kind is None else kind, coupling, m.dim, m.pdim, m.outdim_src, m.outdim_dst, masks[0], masks[1])


def model_types(models, n):
    """(unique models, type index per component) with the reference's batching equality (component hash)."""
    from ndb200 import EdgeModel, VertexModel
    if isinstance(models, (VertexModel, EdgeModel)):
        return [models], np.zeros(n, dtype=np.int32)
    if isinstance(models, tuple):
        uniq, types = models
        return list(uniq), np.asarray(types, dtype=np.int32)
    hashes, uniq, types = {}, [], np.empty(n, dtype=np.int32)
    for i, m in enumerate(models):
        h = m.component_hash()
        if h not in hashes:
            hashes[h] = len(uniq)
            uniq.append(m)
        types[i] = hashes[h]
    return uniq, types


def oracle_network(g, vertexm, edgem):
    vm, vt = model_types(vertexm, g.nv)
    em, et = model_types(edgem, g.ne)
    return O.OracleNetwork(g.nv, g.src, g.dst, [vspec_of(m) for m in vm], vt, [espec_of(m) for m in em], et)


def floored_rel_err(x, ref):
    """max_i |x_i - ref_i| / max(|ref_i|, 1e-3*||ref||_inf)   (SURVEY.md section 8d parity metric)"""
    x, ref = np.asarray(x), np.asarray(ref)
    floor = 1e-3 * np.max(np.abs(ref)) if ref.size else 0.0
    den = np.maximum(np.abs(ref), max(floor, 1e-300))
    return float(np.max(np.abs(x - ref) / den)) if ref.size else 0.0


def rand_inputs(nw_dim, nw_pdim, seed=1, layout=None):
    """u ~ U[0,1), p ~ U[0,1) from seeded streams (benchmark/benchmark_compat.jl:6-14).  `layout` is an optional
    callable p -> p that conditions parameters which must stay away from 0 (M, R, X ...)."""
    u = np.random.default_rng(seed).random(nw_dim)
    p = np.random.default_rng(seed + 1000).random(nw_pdim)
    if layout is not None:
        p = layout(p)
    return u, p


def condition_params(nw, p):
    """Keep divisors benign: kuramoto_second M in [0.5,1.5); swing_dq M in [0.5,1.5), D in
    [0.05,0.15), V in [0.9,1.1); line_dq R in [0.01,0.1), X in [0.1,1), active = 1."""
    p = p.copy()
    for b in nw.vertexbatches:
        f, n, w = b.p_first - 1, len(b), b.model.pdim
        blk = p[f:f + n * w].reshape(n, w) if w else None
        if b.model.name == "kuramoto_second":
            blk[:, 0] += 0.5
        elif b.model.name == "swing_dq":
            blk[:, 0] += 0.5
            blk[:, 1] = 0.05 + 0.1 * blk[:, 1]
            blk[:, 3] = 0.9 + 0.2 * blk[:, 3]
    for b in nw.layer.edgebatches:
        f, n, w = b.p_first - 1, len(b), b.model.pdim
        if b.model.name == "line_dq":
            blk = p[f:f + n * w].reshape(n, w)
            blk[:, 0] = 0.01 + 0.09 * blk[:, 0]
            blk[:, 1] = 0.1 + 0.9 * blk[:, 1]
            blk[:, 2] = 1.0
    return p


def null_aggregator(im, edgebatches):
    """an aggregator closure that builds nothing: lets the host-side table construction run without a GPU"""
    return None
