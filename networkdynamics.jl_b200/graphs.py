"""Graphs in the edge order the reference uses.

`IndexManager` numbers edges by `collect(edges(g))` (src/network_structure.jl:37): for a Graphs.jl
`SimpleGraph` that is every pair once with src < dst, sorted by (src, dst); for a `SimpleDiGraph`
every arc sorted by (src, dst) (docs/src/mathematical_model.md:108-115).  The generators below are
our own seeded generators (Graphs.jl is not available here); only the canonical edge order matters
for parity.  Vertex ids are 1-based like the reference's.
"""
from __future__ import annotations

import numpy as np


def _sorted_unique(keys: np.ndarray) -> np.ndarray:
    """np.unique(keys) for int64 keys via sort + neighbour compare (numpy 2.3's hash-based unique takes a minute on the
    3e7 keys of an 8-GPU benchmark graph, the vectorised sort under a second)"""
    k = np.sort(np.asarray(keys, dtype=np.int64), kind="stable")
    if k.size == 0:
        return k
    keep = np.empty(k.size, dtype=bool)
    keep[0] = True
    np.not_equal(k[1:], k[:-1], out=keep[1:])
    return k[keep]


class SimpleGraph:
    """Undirected simple graph with edges in Graphs.jl `edges(g)` order."""
    directed = False

    def __init__(self, nv: int, src, dst, _canonical: bool = False):
        self.nv = int(nv)
        s = np.asarray(src, dtype=np.int64).ravel()
        d = np.asarray(dst, dtype=np.int64).ravel()
        if not _canonical:
            if s.size and (min(s.min(), d.min()) < 1 or max(s.max(), d.max()) > self.nv):
                raise ValueError("edge endpoint outside 1:nv")
            lo, hi = np.minimum(s, d), np.maximum(s, d)
            keep = lo != hi  # SimpleGraph has no self loops
            key = _sorted_unique(lo[keep] * (self.nv + 1) + hi[keep])  # sorted by (src,dst), multi-edges collapse
            s, d = key // (self.nv + 1), key % (self.nv + 1)
        self.src, self.dst = np.ascontiguousarray(s), np.ascontiguousarray(d)

    @property
    def ne(self) -> int:
        return int(self.src.size)

    def laplacian(self):
        L = np.zeros((self.nv, self.nv))
        for s, d in zip(self.src - 1, self.dst - 1):
            L[s, s] += 1; L[d, d] += 1; L[s, d] -= 1; L[d, s] -= 1
        return L


class SimpleDiGraph(SimpleGraph):
    """Directed simple graph; arcs sorted by (src, dst)."""
    directed = True

    def __init__(self, nv: int, src, dst):
        self.nv = int(nv)
        s = np.asarray(src, dtype=np.int64).ravel()
        d = np.asarray(dst, dtype=np.int64).ravel()
        keep = s != d
        key = _sorted_unique(s[keep] * (self.nv + 1) + d[keep])
        self.src = np.ascontiguousarray(key // (self.nv + 1))
        self.dst = np.ascontiguousarray(key % (self.nv + 1))


def nv(g):
    return g.nv


def ne(g):
    return g.ne


# ---------------------------------------------------------------------------------------------------
# generators (seeded, vectorised)
# ---------------------------------------------------------------------------------------------------
def complete_graph(n: int) -> SimpleGraph:
    i, j = np.triu_indices(n, k=1)
    return SimpleGraph(n, i + 1, j + 1)


def path_graph(n: int) -> SimpleGraph:
    a = np.arange(1, n, dtype=np.int64)
    return SimpleGraph(n, a, a + 1)


def grid_graph(nx: int, ny: int) -> SimpleGraph:
    """nx x ny 4-neighbour lattice; vertex (x, y) -> 1 + x + nx*y."""
    idx = np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    s = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()]) + 1
    d = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()]) + 1
    return SimpleGraph(nx * ny, s, d)


def watts_strogatz(n: int, k: int, beta: float, seed: int = 1, directed: bool = False):
    """Ring of n vertices, each joined to its k//2 successors, far end rewired with probability beta
    (self loops / duplicates produced by rewiring are dropped)."""
    rng = np.random.default_rng(seed)
    base = np.arange(n, dtype=np.int64)
    s = np.repeat(base, k // 2)
    off = np.tile(np.arange(1, k // 2 + 1, dtype=np.int64), n)
    d = (s + off) % n
    rew = rng.random(s.size) < beta
    d = np.where(rew, rng.integers(0, n, size=s.size), d)
    if directed:
        flip = rng.random(s.size) < 0.5
        s, d = np.where(flip, d, s), np.where(flip, s, d)
        return SimpleDiGraph(n, s + 1, d + 1)
    return SimpleGraph(n, s + 1, d + 1)


def erdos_renyi(n: int, m: int, seed: int = 1) -> SimpleGraph:
    """G(n, M): M distinct unordered pairs, uniformly."""
    rng = np.random.default_rng(seed)
    keys = np.empty(0, dtype=np.int64)
    while keys.size < m:
        need = int((m - keys.size) * 1.05) + 16
        a = rng.integers(1, n + 1, size=need, dtype=np.int64)
        b = rng.integers(1, n + 1, size=need, dtype=np.int64)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        k = (lo * (n + 1) + hi)[lo != hi]
        keys = _sorted_unique(np.concatenate([keys, k]))
    if keys.size > m:
        keys = np.sort(rng.permutation(keys)[:m])
    return SimpleGraph(n, keys // (n + 1), keys % (n + 1), _canonical=True)


def barabasi_albert(n: int, m: int, seed: int = 1) -> SimpleGraph:
    """Preferential attachment (Batagelj-Brandes slot sampling, resolved by pointer jumping so that it
    vectorises): vertex v >= 1 draws m targets proportionally to degree among earlier slots.  Self loops and
    duplicate edges are dropped, so E is slightly below m*(n-1)."""
    rng = np.random.default_rng(seed)
    ne_ = m * (n - 1)
    i = np.arange(ne_, dtype=np.int64)
    # slot 0 = vertex 0; slot 1+2i = source of edge i; slot 2+2i = target of edge i
    r = (rng.random(ne_) * (1 + 2 * i)).astype(np.int64)
    ptr = r.copy()
    while True:
        is_target = (ptr > 0) & (ptr % 2 == 0)
        if not is_target.any():
            break
        ptr[is_target] = r[(ptr[is_target] - 2) // 2]
    src = 1 + i // m
    dst = np.where(ptr == 0, 0, 1 + ((ptr - 1) // 2) // m)
    return SimpleGraph(n, src + 1, dst + 1)


# ---------------------------------------------------------------------------------------------------
# locality ordering for the multi-GPU vertex partition ("contiguous ranges of a locality-ordered graph")
# ---------------------------------------------------------------------------------------------------
def locality_order(g, method: str = "rcm") -> np.ndarray:
    """A vertex ordering under which neighbours get nearby ids, so that contiguous id ranges (what PartitionedNetwork
    gives every rank) cut few edges: reverse Cuthill-McKee on the symmetrised adjacency structure.
    Returns `order` (0-based): order[k] = the OLD vertex (0-based) that becomes new vertex k."""
    if method != "rcm":
        raise ValueError("locality_order: only 'rcm' is implemented")
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    n = g.nv
    s, d = np.asarray(g.src) - 1, np.asarray(g.dst) - 1
    a = coo_matrix((np.ones(2 * s.size, dtype=np.int8), (np.concatenate([s, d]), np.concatenate([d, s]))), shape=(n, n)).tocsr()
    return np.asarray(reverse_cuthill_mckee(a, symmetric_mode=True), dtype=np.int64)


def permute_graph(g, order: np.ndarray):
    """Relabel the vertices of `g` (new vertex k = old vertex order[k]) and return (g2, edge_order): g2 in the reference's
    canonical edge order, edge_order[j] = the OLD edge index (0-based) that is edge j of g2 -- what a caller needs to carry
    per-vertex data (`x_new = x_old[order]`) and per-edge data (`y_new = y_old[edge_order]`) over to the relabelled network.
    The sequential accumulation order is then defined on the new ids."""
    order = np.asarray(order, dtype=np.int64)
    n = g.nv
    if order.size != n or not np.array_equal(np.sort(order), np.arange(n)):
        raise ValueError("order must be a permutation of 0..nv-1")
    new_of_old = np.empty(n, dtype=np.int64)
    new_of_old[order] = np.arange(n)
    s, d = new_of_old[np.asarray(g.src) - 1], new_of_old[np.asarray(g.dst) - 1]
    if not g.directed:
        s, d = np.minimum(s, d), np.maximum(s, d)
    edge_order = np.argsort(s * n + d, kind="stable")
    cls = SimpleDiGraph if g.directed else SimpleGraph
    g2 = cls(n, s[edge_order] + 1, d[edge_order] + 1)
    assert g2.ne == g.ne
    return g2, edge_order
