"""Host-side mirror of the reference's component model types, for the RHS path only.

Mirrors `VertexModel` / `EdgeModel` (src/component_functions.jl:251-329,532-567), `StateMask` (:81-99) and the
edge output wrappers `AntiSymmetric` / `Symmetric` / `Directed` (:117-175).  In the reference `f` and `g` are
arbitrary Julia functions; the B200 engine only runs hand-written kernels, so here `f`/`g` are tokens of a
kernel registry (`RegisteredFunction`).  A model whose functions are not registered can still be described
(layout queries work) but `Network(...; aggregator=B200Aggregator(+))` raises `ArgumentError` for it -- there is
no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

from . import _cabi


class ArgumentError(ValueError):
    """Julia's ArgumentError (what the reference throws on size / type mismatches, src/coreloop.jl:3,6)."""


@dataclass(frozen=True)
class RegisteredFunction:
    """A component function with a hand-written sm_100a kernel.  `ref` cites the Julia source it restates."""
    name: str
    kind: int
    role: str  # "vertex_f" | "vertex_g" | "edge_g" | "edge_g2" (two-sided static g) | "edge_f" (f of an edge with states)
    ref: str = ""

    def __call__(self, *a, **k):  # pragma: no cover - these are device kernels, not host callables
        raise RuntimeError(f"{self.name} is a device kernel token; evaluate it through Network(...)")


@dataclass(frozen=True)
class CudaFunction:
    """A user-supplied component function stated as the BODY of a CUDA C++ device function; it is spliced into the fused
    kernels and compiled for sm_100a (NVRTC) when the network is built -- the engine-side counterpart of the reference
    accepting arbitrary Julia functions (and of MTK models printed with Symbolics' C target,
    ext/NetworkDynamicsMTKExt.jl:497-518).  Arguments visible to the body, by role (include/nd_b200.h):
      vertex_f : double* dv, const double* v, const double* esum, const double* p, double t
      vertex_g : double* out, const double* v, const double* p, double t
      vertex_gff : double* out, const double* v, const double* ins, const double* p, double t
                 (feed-forward g of an "injector" leaf behind a LoopbackConnection: `ins` = the hub's output; dim may be 0)
      edge_g   : double* e_dst, const double* v_src, const double* v_dst, const double* p, double t
      edge_g2  : double* e_src, double* e_dst, const double* v_src, const double* v_dst, const double* p, double t
                 (the two-sided form; use it unwrapped or as Fiducial(...))
      edge_f   : double* de, const double* e, const double* v_src, const double* v_dst, const double* p, double t
                 (f of an edge with states; its g must be StateMasks: AntiSymmetric(1), Fiducial(dst=1, src=2), ...)
    `py` is an optional host restatement (same arguments as numpy arrays, returning the outputs) used by tests only."""
    name: str
    role: str
    body: str
    py: Optional[object] = field(default=None, compare=False, hash=False)

    def __call__(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{self.name} is device code; evaluate it through Network(...)")


class Fiducial:
    """`Fiducial(src, dst)` / `Fiducial(src=..., dst=...)` (src/component_functions.jl:177-203): two single-sided output
    functions.  For edges with states both are StateMasks (ints / index tuples are wrapped like in the reference,
    :191-194).  For static edges the engine takes the pair as ONE two-sided function `Fiducial(g)` with
    g(osrc, odst, vsrc, vdst, p, t) -- the form the reference itself evaluates (src/coreloop.jl:208)."""
    coupling = _cabi.FIDUCIAL

    def __init__(self, *args, src=None, dst=None, g=None):
        if len(args) == 1 and src is None and dst is None and g is None:
            g = args[0]
        elif len(args) == 2 and src is None and dst is None:
            src, dst = args
        elif args:
            raise ArgumentError("Fiducial(g), Fiducial(src, dst) or Fiducial(src=..., dst=...)")
        wrap = lambda m: m if (m is None or isinstance(m, StateMask) or callable(m)) else StateMask(m)
        self.g, self.src, self.dst = g, wrap(src), wrap(dst)
        if (self.g is None) == (self.src is None or self.dst is None):
            raise ArgumentError("Fiducial needs either one two-sided function or both src and dst")

    def _key(self):
        return (self.g, self.src, self.dst)

    def __eq__(self, o):
        return isinstance(o, Fiducial) and self._key() == o._key()

    def __hash__(self):
        return hash(("Fiducial",) + self._key())

    def __repr__(self):
        return f"Fiducial({self.g!r})" if self.g is not None else f"Fiducial(src={self.src!r}, dst={self.dst!r})"


@dataclass(frozen=True)
class StateMask:
    """`StateMask(idxs)`: output k is state idxs[k] (1-based), src/component_functions.jl:81-99."""
    idxs: Tuple[int, ...]

    def __init__(self, idxs):
        if isinstance(idxs, int):
            idxs = (idxs,)
        object.__setattr__(self, "idxs", tuple(int(i) for i in idxs))

    def contiguous_first(self):
        """first index when the mask is a contiguous ascending range (what the engine's StateMask reads need), else None"""
        i = self.idxs
        return i[0] if i and i == tuple(range(i[0], i[0] + len(i))) else None


@dataclass(frozen=True)
class AntiSymmetric:
    """osrc = -odst, src/component_functions.jl:117-127; a number / index tuple wraps the StateMask like in the reference"""
    g: object
    coupling = _cabi.ANTISYMMETRIC

    def __post_init__(self):
        if isinstance(self.g, (int, tuple, list, range)):
            object.__setattr__(self, "g", StateMask(self.g))


@dataclass(frozen=True)
class Symmetric:
    """osrc = odst, src/component_functions.jl:142-152; a number / index tuple wraps the StateMask like in the reference"""
    g: object
    coupling = _cabi.SYMMETRIC

    def __post_init__(self):
        if isinstance(self.g, (int, tuple, list, range)):
            object.__setattr__(self, "g", StateMask(self.g))


@dataclass(frozen=True)
class Directed:
    """no src output (outdim.src = 0), src/component_functions.jl:167-175; a number / index tuple wraps the StateMask like in the reference"""
    g: object
    coupling = _cabi.DIRECTED

    def __post_init__(self):
        if isinstance(self.g, (int, tuple, list, range)):
            object.__setattr__(self, "g", StateMask(self.g))


@dataclass(frozen=True)
class VIndex:
    """`VIndex(i, sub)`: state / output `sub` of vertex i (1-based) as an external input (src/external_inputs.jl:36-50):
    `sub` = a state symbol of the model (`sym`), a 1-based state index, or ("out", k) for output k."""
    comp: int
    sub: object


@dataclass(frozen=True)
class EIndex:
    """`EIndex(i, sub)`: the same for edge i; ("out", k) counts over the flat outputs (src outputs, then dst outputs)."""
    comp: int
    sub: object


@dataclass(frozen=True)
class VertexModel:
    f: Optional[object]
    g: object                      # StateMask or a RegisteredFunction (NoFeedForward g)
    dim: int
    pdim: int = 0
    outdim: Optional[int] = None
    sym: Tuple[str, ...] = ()
    psym: Tuple[str, ...] = ()
    name: str = "VertexModel"
    extin: Tuple[object, ...] = ()   # external inputs: VIndex / EIndex refs; f then takes (dv, v, esum, ext, p, t)

    @property
    def extdim(self) -> int:
        return len(self.extin)

    def __post_init__(self):
        if self.outdim is None:
            if isinstance(self.g, StateMask):
                object.__setattr__(self, "outdim", len(self.g.idxs))
            else:
                raise ArgumentError("outdim required when g is not a StateMask")

    def component_hash(self):
        """What batches are formed on: src/construction.jl:245-256 (name/metadata are not part of it)."""
        return ("V", self.f, self.g, self.dim, self.outdim, self.pdim, self.extdim)

    def custom_spec(self):
        """(role, dim, pdim, outdim, two_sided, f_body, g_body, extdim) when f is user-supplied CUDA code, else None"""
        if not isinstance(self.f, CudaFunction) or self.f.role != "vertex_f":
            return None
        if isinstance(self.g, CudaFunction) and self.g.role in ("vertex_g", "vertex_gff"):
            g_body = self.g.body
        elif isinstance(self.g, StateMask) and self.g.idxs == tuple(range(1, self.outdim + 1)):
            g_body = None
        else:
            return None
        return (0, self.dim, self.pdim, self.outdim, 0, self.f.body, g_body, self.extdim, int(self.hasff))

    @property
    def hasff(self) -> bool:
        """feed-forward output function (src/component_functions.jl:63): allowed for injector leaves only"""
        return isinstance(self.g, CudaFunction) and self.g.role == "vertex_gff"

    def kernel_kind(self) -> Optional[int]:
        f = self.f
        if not isinstance(f, RegisteredFunction) or f.role != "vertex_f" or self.extdim:
            return None
        if f.kind == _cabi.V_SWING_DQ:
            return f.kind if isinstance(self.g, RegisteredFunction) and self.g.kind == f.kind else None
        # the StateMask kernels read output k from state k: mask must be 1:outdim
        if isinstance(self.g, StateMask) and self.g.idxs == tuple(range(1, self.outdim + 1)):
            return f.kind
        return None


@dataclass(frozen=True)
class EdgeModel:
    g: object                      # AntiSymmetric / Symmetric / Directed / Fiducial wrapping a function or, for edges with
                                   # states (f given, dim > 0), StateMasks
    outdim: int = 1
    pdim: int = 0
    dim: int = 0
    f: Optional[object] = None
    psym: Tuple[str, ...] = ()
    name: str = "EdgeModel"
    sym: Tuple[str, ...] = ()
    extin: Tuple[object, ...] = ()   # external inputs (edges with states only): f takes (de, e, v_src, v_dst, ext, p, t)

    @property
    def extdim(self) -> int:
        return len(self.extin)

    @property
    def coupling(self) -> Optional[int]:
        if isinstance(self.g, CudaFunction) and self.g.role == "edge_g2":
            return _cabi.FIDUCIAL
        return getattr(self.g, "coupling", None)

    @property
    def outdim_src(self) -> int:
        return 0 if isinstance(self.g, Directed) else self.outdim

    @property
    def outdim_dst(self) -> int:
        return self.outdim

    def component_hash(self):
        return ("E", self.f, self.g, self.dim, self.outdim_src, self.outdim_dst, self.pdim, self.extdim)

    def state_masks(self):
        """(mask_src_first, mask_dst_first), 1-based, for an edge with states whose outputs are contiguous StateMasks of
        the right width; None otherwise (src/component_functions.jl:81-99,117-203)"""
        if self.dim <= 0 or self.f is None:
            return None
        if isinstance(self.g, Fiducial):
            ms, md = self.g.src, self.g.dst
            if not (isinstance(ms, StateMask) and isinstance(md, StateMask)) or len(ms.idxs) != self.outdim or len(md.idxs) != self.outdim:
                return None
            a, b = ms.contiguous_first(), md.contiguous_first()
            return None if a is None or b is None or max(a, b) + self.outdim - 1 > self.dim else (a, b)
        md = getattr(self.g, "g", None)
        if not isinstance(md, StateMask) or len(md.idxs) != self.outdim:
            return None
        b = md.contiguous_first()
        return None if b is None or b + self.outdim - 1 > self.dim else (0, b)

    def custom_spec(self):
        if self.dim > 0:              # edge with states: user-supplied f, StateMask outputs
            if isinstance(self.f, CudaFunction) and self.f.role == "edge_f" and self.state_masks() is not None:
                return (1, self.dim, self.pdim, self.outdim, 0, self.f.body, None, self.extdim, 0)
            return None
        if self.f is not None or self.extdim:
            return None
        if isinstance(self.g, CudaFunction) and self.g.role == "edge_g2":       # unwrapped two-sided g
            return (1, 0, self.pdim, self.outdim, 1, self.g.body, None, 0, 0)
        inner = getattr(self.g, "g", None)
        if isinstance(inner, CudaFunction):
            if isinstance(self.g, Fiducial):
                return (1, 0, self.pdim, self.outdim, 1, inner.body, None, 0, 0) if inner.role == "edge_g2" else None
            return (1, 0, self.pdim, self.outdim, 0, inner.body, None, 0, 0) if inner.role == "edge_g" else None
        return None

    def kernel_kind(self) -> Optional[int]:
        inner = getattr(self.g, "g", None)
        if self.coupling is None or self.extdim:
            return None
        if self.dim > 0:              # edge with states: registered f, StateMask outputs
            if isinstance(self.f, RegisteredFunction) and self.f.role == "edge_f" and self.state_masks() is not None:
                return self.f.kind
            return None
        if self.f is not None:
            return None
        if isinstance(inner, RegisteredFunction) and inner.role == ("edge_g2" if isinstance(self.g, Fiducial) else "edge_g"):
            return inner.kind
        return None


class Lib:
    """The model zoo of test/ComponentLibrary.jl (module `Lib`) and benchmark/benchmark_models.jl, restricted
    to the models with registered kernels."""
    diffusionedge = RegisteredFunction("diffusionedge!", _cabi.E_DIFFUSION, "edge_g", "test/ComponentLibrary.jl:8-10")
    diffusionedge_nop = RegisteredFunction("diffusionedge!", _cabi.E_DIFFUSION_NOP, "edge_g", "benchmark/benchmark_models.jl:5-8")
    kuramoto_edge_f = RegisteredFunction("kuramoto_edge!", _cabi.E_KURAMOTO, "edge_g", "test/ComponentLibrary.jl:51-53")
    line_dq_f = RegisteredFunction("StaticPowerLineDQ", _cabi.E_LINE_DQ, "edge_g", "test/ComponentLibrary.jl:212-245")
    diffusionedge_fid = RegisteredFunction("diffusionedge_fid!", _cabi.E_DIFFUSION_FID, "edge_g2", "test/ComponentLibrary.jl:22-25")
    diffusion_dedge = RegisteredFunction("diffusion_dedge!", _cabi.E_DIFFUSION_ODE, "edge_f", "test/ComponentLibrary.jl:30-34")
    real_ode_edge = RegisteredFunction("real_ode_edge!", _cabi.E_RELAX_ODE, "edge_f", "test/diffusion_test.jl:96-100")
    loopback_g = RegisteredFunction("LOOPBACK_G", _cabi.E_LOOPBACK, "edge_g", "src/post_utils.jl:105-108")
    diffusionvertex = RegisteredFunction("diffusionvertex!", _cabi.V_DIFFUSION, "vertex_f", "test/ComponentLibrary.jl:42-45")
    kuramoto_vertex = RegisteredFunction("kuramoto_vertex!", _cabi.V_KURAMOTO_FIRST, "vertex_f", "test/ComponentLibrary.jl:69-71")
    kuramoto_inertia = RegisteredFunction("kuramoto_inertia!", _cabi.V_KURAMOTO_SECOND, "vertex_f", "test/ComponentLibrary.jl:59-63")
    kuramoto_inertia_bench = RegisteredFunction("kuramoto_inertia!", _cabi.V_KURAMOTO_SECOND_BENCH, "vertex_f", "benchmark/benchmark_models.jl:37-41")
    swing_dq_f = RegisteredFunction("SwingDQ.f", _cabi.V_SWING_DQ, "vertex_f", "test/ComponentLibrary.jl:139-161")
    swing_dq_g = RegisteredFunction("SwingDQ.g", _cabi.V_SWING_DQ, "vertex_g", "test/ComponentLibrary.jl:158-159")

    @staticmethod
    def diffusion_edge():
        return EdgeModel(g=AntiSymmetric(Lib.diffusionedge), outdim=1, pdim=1, name="diff_edge")

    @staticmethod
    def diffusion_edge_nop():
        """benchmark variant without a parameter (benchmark/benchmark_models.jl:9)"""
        return EdgeModel(g=AntiSymmetric(Lib.diffusionedge_nop), outdim=1, pdim=0, name="diff_edge")

    @staticmethod
    def diffusion_edge_fid():
        """test/ComponentLibrary.jl:26-28: two-sided static g, no wrapper"""
        return EdgeModel(g=Fiducial(Lib.diffusionedge_fid), outdim=1, pdim=1, name="diff_edge_fid")

    @staticmethod
    def loopback(outdim: int = 1):
        """`LoopbackConnection(; potential, flow)` (src/post_utils.jl:110-185): Directed(LOOPBACK_G) from an injector leaf to
        its hub -- the hub receives minus the injector's output, the injector's input is the hub's output"""
        return EdgeModel(g=Directed(Lib.loopback_g), outdim=outdim, pdim=0, name="loopback")

    @staticmethod
    def diffusion_odeedge():
        """test/ComponentLibrary.jl:35-40: edge with two states, outputs Fiducial(dst=1:1, src=2:2)"""
        return EdgeModel(f=Lib.diffusion_dedge, dim=2, pdim=1, psym=("τ",), g=Fiducial(dst=(1,), src=(2,)), outdim=1,
                         name="diff_edge_ode")

    @staticmethod
    def relax_odeedge():
        """test/diffusion_test.jl:101: EdgeModel(; f=real_ode_edge!, dim=2, g=Fiducial(2,1))"""
        return EdgeModel(f=Lib.real_ode_edge, dim=2, pdim=0, g=Fiducial(2, 1), outdim=1, name="real_ode_edge")

    @staticmethod
    def diffusion_vertex():
        return VertexModel(f=Lib.diffusionvertex, dim=1, pdim=0, g=StateMask((1,)), name="diffusion_vertex")

    @staticmethod
    def kuramoto_edge():
        return EdgeModel(g=AntiSymmetric(Lib.kuramoto_edge_f), outdim=1, pdim=1, psym=("K",), name="kuramoto_edge")

    @staticmethod
    def kuramoto_first():
        return VertexModel(f=Lib.kuramoto_vertex, dim=1, pdim=1, g=StateMask(1), sym=("θ",), psym=("ω",),
                           name="kuramoto_first")

    @staticmethod
    def kuramoto_second():
        return VertexModel(f=Lib.kuramoto_inertia, dim=2, pdim=3, g=StateMask(1), sym=("δ", "ω"),
                           psym=("M", "D", "Pm"), name="kuramoto_second")

    @staticmethod
    def kuramoto_second_bench():
        return VertexModel(f=Lib.kuramoto_inertia_bench, dim=2, pdim=1, g=StateMask((1,)), sym=("θ", "ω"),
                           name="kuramoto_vertex_2d")

    @staticmethod
    def swing_dq():
        return VertexModel(f=Lib.swing_dq_f, g=Lib.swing_dq_g, dim=2, pdim=4, outdim=2, sym=("θ", "ω"),
                           psym=("M", "D", "Pmech", "V"), name="swing_dq")

    @staticmethod
    def line_dq():
        return EdgeModel(g=AntiSymmetric(Lib.line_dq_f), outdim=2, pdim=3, psym=("R", "X", "active"), name="line_dq")
