"""networkdynamics.jl_b200 -- B200-native engine for ONE path of NetworkDynamics.jl: the network right-hand side
`nw(du, u, p, t)` (src/coreloop.jl), behind the reference's `ExecutionStyle` + `Aggregator` plug-in seam.

Layout of this package (only what the path needs):
  csrc/            hand-written sm_100a kernels + the C ABI (include/nd_b200.h)
  _cabi.py         ctypes view of the C ABI, in-tree nvcc build
  components.py    VertexModel / EdgeModel / StateMask / wrappers + the `Lib` model zoo with registered kernels
  graphs.py        graphs in Graphs.jl `edges(g)` order + seeded generators
  network.py       IndexManager, Network, B200Execution, B200Aggregator
  distributed.py   vertex-partitioned multi-GPU RHS (one process per GPU)
"""
from . import _cabi
from ._cabi import build
from .components import (AntiSymmetric, ArgumentError, CudaFunction, Directed, EdgeModel, EIndex, Fiducial, Lib,
                         RegisteredFunction, StateMask, Symmetric, VertexModel, VIndex)
from .graphs import (SimpleDiGraph, SimpleGraph, barabasi_albert, complete_graph, erdos_renyi, grid_graph, locality_order, ne,
                     nv, path_graph, permute_graph, watts_strogatz)
from .network import (B200Aggregator, B200Execution, ComponentBatch, ExecutionStyle, HaloTimeoutError, IndexManager, Network, dim,
                      find_identical, get_aggr_constructor, iscudacompatible, pdim, pinned_empty, usebuffer)

__all__ = [n for n in dir() if not n.startswith("_")]
