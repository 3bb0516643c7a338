# NetworkDynamicsB200.jl -- reference-side binding of libnd_b200.so (include/nd_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.  It is the glue a maintainer
# would add (as a package extension next to ext/NetworkDynamicsCUDAExt.jl) so that
#
#     nw = Network(g, vertexm, edgem; execution=B200Execution(), aggregator=B200Aggregator(+))
#     nw(du, u, p, t)          # du, u, p :: CuArray{Float64}
#
# runs the hand-written sm_100a engine instead of the KernelAbstractions path.  The Python mirror in
# networkdynamics.jl_b200/network.py implements the same flattening and is what the tests exercise.
module NetworkDynamicsB200

using NetworkDynamics
using NetworkDynamics: ExecutionStyle, Aggregator, IndexManager, ComponentBatch, Network,
                       AntiSymmetric, Symmetric, Directed, Fiducial, StateMask, _find_identical_components,
                       compf, compg, dim, pdim, outdim
import CUDA
using CUDA: CuArray, CuVector, CuPtr, stream
import NetworkDynamics: aggregate!, get_aggr_constructor, iscudacompatible, aggfun

const libnd_b200 = get(ENV, "ND_B200_LIB", "libnd_b200.so")
const ABI_VERSION = Cint(5)

# ---- tags ---------------------------------------------------------------------------------------------------------
"`ExecutionStyle` tag (field-less: only its type is stored in `Network{EX,...}`, src/network_structure.jl:83,116)."
struct B200Execution{buffered} <: ExecutionStyle{buffered} end
B200Execution() = B200Execution{false}()      # Lazy gather provider: nothing is materialised (src/construction.jl:210-214)
iscudacompatible(::Type{<:B200Execution}) = true

# ---- kernel registry ------------------------------------------------------------------------------------------------
# Users opt a component function into a hand-written kernel by adding a method; anything else raises ArgumentError.
vertex_kernel(f, g) = nothing
edge_kernel(g) = nothing
edge_f_kernel(f) = nothing             # f of an edge WITH states ("ODE edge"); its g must be StateMasks
const V_DIFFUSION, V_KURAMOTO_FIRST, V_KURAMOTO_SECOND, V_KURAMOTO_SECOND_BENCH, V_SWING_DQ = Cint.(0:4)
const E_DIFFUSION, E_DIFFUSION_NOP, E_KURAMOTO, E_LINE_DQ, E_DIFFUSION_ODE, E_RELAX_ODE, E_DIFFUSION_FID, E_LOOPBACK = Cint.(0:7)
edge_kernel(::typeof(NetworkDynamics.LOOPBACK_G)) = E_LOOPBACK      # LoopbackConnection (src/post_utils.jl:105-185)
# e.g. (test/ComponentLibrary.jl):
#   NetworkDynamicsB200.edge_kernel(::typeof(Lib.kuramoto_edge!))       = NetworkDynamicsB200.E_KURAMOTO
#   NetworkDynamicsB200.vertex_kernel(::typeof(Lib.kuramoto_vertex!), ::StateMask) = NetworkDynamicsB200.V_KURAMOTO_FIRST
#   NetworkDynamicsB200.edge_kernel(::typeof(Lib.diffusionedge_fid!))    = NetworkDynamicsB200.E_DIFFUSION_FID   # two-sided, unwrapped
#   NetworkDynamicsB200.edge_f_kernel(::typeof(Lib.diffusion_dedge!))    = NetworkDynamicsB200.E_DIFFUSION_ODE   # g = Fiducial(dst=1:1, src=2:2)

# contiguous StateMask -> its first index (1-based); anything else is not an engine-readable output of a stateful edge
_mask_first(m::StateMask) = (ix = collect(m.idxs); ix == collect(first(ix):first(ix)+length(ix)-1) ? Cint(first(ix)) : nothing)
_mask_first(_) = nothing

coupling_of(::AntiSymmetric) = Cint(0)
coupling_of(::Symmetric) = Cint(1)
coupling_of(::Directed) = Cint(2)
coupling_of(::Fiducial) = Cint(3)          # static: two-sided kinds; with states: Fiducial(src=mask, dst=mask)
coupling_of(x) = throw(ArgumentError("B200 engine: edge output wrapper $(typeof(x)) is not supported (no CPU fallback)"))

# ---- user-supplied kinds (run-time compiled by the engine, NVRTC; include/nd_b200.h: nd_b200_custom_kind) -----------
# A component function that has no hand-written kernel can still run on the engine if it can be stated as the BODY of a
# CUDA C++ device function with the reference's own argument list (dv, v, esum, p, t / out, v, p, t /
# e_dst, v_src, v_dst, p, t / e_src, e_dst, v_src, v_dst, p, t), e.g.
#   NetworkDynamicsB200.cuda_source(::typeof(myedge!)) = "e_dst[0] = p[0] * sin(v_src[0] - v_dst[0]);"
# For ModelingToolkit components the body is what Symbolics prints for the generated function:
#   cuda_source(f::RuntimeGeneratedFunction) = Symbolics.build_function(rhs_exprs, args...; target = Symbolics.CTarget())
# (ext/NetworkDynamicsMTKExt.jl:497-518 builds the same expressions into a RuntimeGeneratedFunction).
cuda_source(f) = nothing
const CUSTOM_KIND_BASE = Cint(1000)
struct CCustomKind
    kind::Cint; role::Cint; dim::Cint; pdim::Cint; outdim::Cint; two_sided::Cint
    f_body::Cstring; g_body::Cstring
    extdim::Cint; g_ff::Cint          # g_ff: vertex g is feed forward, g(out, v, ins, p, t) (injector leaves only)
end

# ---- C structs (mirror include/nd_b200.h) ---------------------------------------------------------------------------
struct CVBatch
    kind::Cint; dim::Cint; pdim::Cint; outdim::Cint
    count::Int64; indices::Ptr{Int64}
    state_first::Int64; p_first::Int64; out_first::Int64; aggr_first::Int64
    extdim::Cint; reserved::Cint; ext_src::Ptr{Int64}     # external inputs: ExtMap entries of the batch (see _ext_table)
end
struct CEBatch
    kind::Cint; coupling::Cint; dim::Cint; pdim::Cint; outdim_src::Cint; outdim_dst::Cint
    count::Int64; indices::Ptr{Int64}
    state_first::Int64; p_first::Int64; out_first::Int64; gbuf_first::Int64
    mask_src_first::Cint; mask_dst_first::Cint      # edges with states: StateMask outputs (0 for static edges)
    extdim::Cint; reserved::Cint; ext_src::Ptr{Int64}
end
struct CDesc
    abi_version::Cint; device::Cint
    nv::Int64; ne::Int64
    edge_src::Ptr{Int64}; edge_dst::Ptr{Int64}
    vdepth::Cint; edepth::Cint; n_vbatches::Cint; n_ebatches::Cint
    vbatches::Ptr{CVBatch}; ebatches::Ptr{CEBatch}
    lastidx_dynamic::Int64; lastidx_p::Int64; lastidx_out::Int64; lastidx_aggr::Int64
    row_begin::Int64; row_end::Int64
    long_row_threshold::Cint; flags::Cint
    gather_offset::Ptr{Int64}; gather_len::Int64        # multi-GPU packed halo (C_NULL, 0: single GPU)
    n_custom::Cint; reserved::Cint; custom::Ptr{CCustomKind}
end

# ---- the aggregator owns the engine ---------------------------------------------------------------------------------
mutable struct B200Aggregator{F} <: Aggregator
    f::F                      # read as nw.layer.aggregator.f (test/testutils.jl:50) and by aggfun
    handle::Ptr{Cvoid}
end
"Constructor-closure convention of every reference aggregator (src/aggregators.jl:1-16,137)."
function B200Aggregator(f)
    f === (+) || throw(ArgumentError("B200Aggregator supports only + (no CPU fallback for other reducers)"))
    (im, batches) -> B200Aggregator(im, batches, f)
end
get_aggr_constructor(a::B200Aggregator) = B200Aggregator(a.f)

"""
`aggregate!(a, aggbuf, o)` (src/aggregators.jl:140-151) on device vectors: adds the edge-output block of `o` into `aggbuf`,
per slot in ascending `o` order, on top of its content (test/aggregators_test.jl:69-79).  The RHS method below does not
use it -- the engine's row kernel keeps the sums in registers.
"""
function aggregate!(a::B200Aggregator, aggbuf::CuVector{Float64}, o::CuVector{Float64})
    _check(ccall((:nd_b200_aggregate, libnd_b200), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                 a.handle, pointer(aggbuf), pointer(o), stream().handle), a.handle)
    nothing
end

function _check(rc, handle)
    rc == 0 && return
    msg = unsafe_string(ccall((:nd_b200_last_error, libnd_b200), Cstring, (Ptr{Cvoid},), handle))
    rc in (1, 2) ? throw(ArgumentError(msg)) : error(msg)
end

function B200Aggregator(im::IndexManager, edgebatches, f)
    # vertex batches: same grouping the constructor used (src/construction.jl:156-168,238-243); `im` is fully
    # populated when the aggregator closure runs (src/construction.jl:198)
    vidxs = _find_identical_components(im.vertexm)
    keep = Any[]
    customs = CCustomKind[]                  # user-supplied kinds referenced by the batches
    # external inputs (src/external_inputs.jl): the reference's own ExtMap, one signed index per slot of the external-input
    # buffer -- StateBufIdx(i) -> +i (into u), OutBufIdx(i) -> -i (into o) -- laid out per batch as count x extdim
    extmap = NetworkDynamics.has_external_input(im) ? NetworkDynamics.ExtMap(im).map : nothing
    function _ext_table(ranges, idxs, xd)
        xd == 0 && return Ptr{Int64}(C_NULL)
        tab = Int64[(m = extmap[r[k]]; m isa NetworkDynamics.StateBufIdx ? m.idx : -m.idx) for i in idxs for r in (ranges[i],) for k in 1:xd]
        push!(keep, tab)
        pointer(tab)
    end
    function custom_kind!(role, d, pd, od, two_sided, fsrc, gsrc, xd=0, gff=0)
        fb = Base.unsafe_convert(Cstring, Base.cconvert(Cstring, fsrc)); push!(keep, fsrc)
        gb = isnothing(gsrc) ? Cstring(C_NULL) : (push!(keep, gsrc); Base.unsafe_convert(Cstring, Base.cconvert(Cstring, gsrc)))
        push!(customs, CCustomKind(CUSTOM_KIND_BASE + length(customs), role, d, pd, od, two_sided, fb, gb, xd, gff))
        customs[end].kind
    end
    vb = map(vidxs) do idxs
        m = im.vertexm[first(idxs)]
        kind = vertex_kernel(compf(m), compg(m))
        if isnothing(kind) && (!isnothing(cuda_source(compf(m))) || (isnothing(compf(m)) && !isnothing(cuda_source(compg(m)))))   # user-supplied kind
            gs = compg(m) isa StateMask ? nothing : cuda_source(compg(m))
            # feed-forward g (injector leaves behind a LoopbackConnection): the body takes `ins` after `v`; a vertex without
            # states (PureFeedForward) has f === nothing: its body is empty
            gff = NetworkDynamics.hasff(m) ? 1 : 0
            fsrc = isnothing(compf(m)) ? "" : cuda_source(compf(m))
            kind = custom_kind!(0, dim(m), pdim(m), outdim(m), 0, fsrc, gs, NetworkDynamics.extdim(m), gff)
        end
        isnothing(kind) && throw(ArgumentError("vertex model $(m.name) has neither a registry kernel nor a cuda_source (no CPU fallback)"))
        xd = NetworkDynamics.extdim(m)
        (xd > 0 && kind < CUSTOM_KIND_BASE) && throw(ArgumentError("B200 engine: external inputs need a cuda_source vertex function (no CPU fallback)"))
        ix = Vector{Int64}(idxs); push!(keep, ix)
        i1 = first(idxs)
        CVBatch(kind, dim(m), pdim(m), outdim(m), length(ix), pointer(ix),
                first(im.v_data[i1]), first(im.v_para[i1]), first(im.v_out[i1]), first(im.v_aggr[i1]),
                xd, 0, _ext_table(im.v_ext, idxs, xd))
    end
    eb = map(collect(edgebatches)) do b
        g = compg(b)
        od = outdim(b)
        msrc, mdst = Cint(0), Cint(0)
        if dim(b) > 0
            # edge with states (src/coreloop.jl:41,76): f is a registry kernel or a cuda_source body, the outputs must be
            # contiguous StateMasks -- AntiSymmetric(1), Symmetric(1:2), Directed(1), Fiducial(dst=1:1, src=2:2), ...
            g isa NetworkDynamics.SingleSidedOutputWrapper || throw(ArgumentError("B200 engine: an edge with states needs StateMask outputs (no CPU fallback)"))
            md = _mask_first(g isa Fiducial ? g.dst : g.g)
            ms = g isa Fiducial ? _mask_first(g.src) : Cint(0)
            (isnothing(md) || isnothing(ms)) && throw(ArgumentError("B200 engine: edge outputs must be contiguous StateMasks of the edge's states"))
            msrc, mdst = ms, md
            kind = edge_f_kernel(compf(b))
            if isnothing(kind) && !isnothing(cuda_source(compf(b)))
                kind = custom_kind!(1, dim(b), pdim(b), od.dst, 0, cuda_source(compf(b)), nothing, NetworkDynamics.extdim(b))
            end
        else
            inner = g isa NetworkDynamics.SingleSidedOutputWrapper && !(g isa Fiducial) ? g.g : g
            kind = edge_kernel(inner)
            if isnothing(kind) && !isnothing(cuda_source(inner))     # user-supplied kind (two-sided body: Fiducial / unwrapped)
                two = (g isa Fiducial || !(g isa NetworkDynamics.SingleSidedOutputWrapper)) ? 1 : 0
                kind = custom_kind!(1, 0, pdim(b), od.dst, two, cuda_source(inner), nothing)
            end
        end
        isnothing(kind) &&
            throw(ArgumentError("edge batch $(typeof(g)) has neither a registry kernel nor a cuda_source (no CPU fallback)"))
        ix = Vector{Int64}(b.indices); push!(keep, ix)
        coupling = g isa NetworkDynamics.SingleSidedOutputWrapper ? coupling_of(g) : Cint(3)   # unwrapped two-sided g
        xd = NetworkDynamics.extdim(b)
        (xd > 0 && (dim(b) == 0 || kind < CUSTOM_KIND_BASE)) &&
            throw(ArgumentError("B200 engine: external inputs of an edge need an edge with states and a cuda_source f (no CPU fallback)"))
        CEBatch(kind, coupling, dim(b), pdim(b), od.src, od.dst, length(ix), pointer(ix),
                b.statestride.first, b.pstride.first, b.outbufstride.first, b.inbufstride.first, msrc, mdst,
                xd, 0, _ext_table(im.e_ext, b.indices, xd))
    end
    esrc = Int64[e.src for e in im.edgevec]; edst = Int64[e.dst for e in im.edgevec]
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep vb eb esrc edst customs begin
        desc = CDesc(ABI_VERSION, Cint(CUDA.deviceid()), length(im.vertexm), length(im.edgevec),
                     pointer(esrc), pointer(edst), im.vdepth, im.edepth, length(vb), length(eb),
                     pointer(vb), pointer(eb), im.lastidx_dynamic, im.lastidx_p, im.lastidx_out, im.lastidx_aggr,
                     0, 0, Cint(0), Cint(0), Ptr{Int64}(C_NULL), 0,
                     Cint(length(customs)), Cint(0), isempty(customs) ? Ptr{CCustomKind}(C_NULL) : pointer(customs))
        rc = ccall((:nd_b200_create, libnd_b200), Cint, (Ref{CDesc}, Ref{Ptr{Cvoid}}), desc, handle)
        _check(rc, C_NULL)
    end
    a = B200Aggregator{typeof(f)}(f, handle[])
    finalizer(x -> ccall((:nd_b200_destroy, libnd_b200), Cvoid, (Ptr{Cvoid},), x.handle), a)
    a
end

# ---- the RHS ---------------------------------------------------------------------------------------------------------
# Replaces (nw::Network)(du,u,p,t) (src/coreloop.jl:1-102) for B200Execution networks.
function (nw::Network{<:B200Execution})(du, u, p, t; perturb=nothing, perturb_maps=nothing, RET=Val(:du))
    (isnothing(perturb) && RET == Val(:du)) ||
        throw(ArgumentError("B200Execution supports neither `perturb` nor RET != :du; keep a CPU twin " *
                            "Network(nw; execution=SequentialExecution{true}(), aggregator=SequentialAggregator(+))"))
    (du isa CuArray{Float64} && u isa CuArray{Float64}) || throw(ArgumentError("B200Execution needs CuArray{Float64} du/u"))
    (length(du) == length(u) == nw.im.lastidx_dynamic) ||
        throw(ArgumentError("du or u does not have expected size $(nw.im.lastidx_dynamic)"))
    pp = NetworkDynamics.pdim(nw) > 0 ? pointer(p::CuArray{Float64}) : CuPtr{Float64}(0)
    NetworkDynamics.pdim(nw) > 0 && length(p) != nw.im.lastidx_p &&
        throw(ArgumentError("p does not has expecte size $(nw.im.lastidx_p)"))
    h = nw.layer.aggregator.handle
    rc = ccall((:nd_b200_rhs, libnd_b200), Cint,
               (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Ptr{Cvoid}),
               h, pointer(du), pointer(u), pp, Float64(t), stream().handle)   # CUDA.jl task-local stream: no host sync
    _check(rc, h)
    nothing
end

"Fixed-step classical RK4 with stage updates fused into the RHS kernels (device-resident `u`, advanced in place)."
function rk4!(nw::Network{<:B200Execution}, u::CuArray{Float64}, p, t0, dt, nsteps)
    h = nw.layer.aggregator.handle
    pp = NetworkDynamics.pdim(nw) > 0 ? pointer(p::CuArray{Float64}) : CuPtr{Float64}(0)
    rc = ccall((:nd_b200_rk4, libnd_b200), Cint,
               (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Cdouble, Int64, Ptr{Cvoid}),
               h, pointer(u), pp, Float64(t0), Float64(dt), Int64(nsteps), stream().handle)
    _check(rc, h)
    u
end

"""
    pack_params!(nw, p)          # p::CuArray{Float64}, or `nothing` to undo

Optional contract (no reference counterpart; `nd_b200_pack_params`): the engine copies the edge parameters into its own
per-entry array and reads them coalesced until the next call.  The caller promises not to change edge parameters in `p`
in between (e.g. call it again from the `affect!` of a callback that mutates `p`).  Default: `p` is re-read on every call.
"""
function pack_params!(nw::Network{<:B200Execution}, p)
    h = nw.layer.aggregator.handle
    pp = isnothing(p) ? CuPtr{Float64}(0) : pointer(p::CuArray{Float64})
    _check(ccall((:nd_b200_pack_params, libnd_b200), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Ptr{Cvoid}), h, pp, stream().handle), h)
    nothing
end

# ---- printing (src/show.jl:46-49 prints every reference aggregator as Name(repr(f))) ----------------------------------
Base.show(io::IO, s::B200Aggregator) = print(io, "B200Aggregator($(repr(s.f)))")
Base.show(io::IO, ::B200Execution{buffered}) where {buffered} = print(io, "B200Execution{$buffered}()")

# ---- get_buffers (src/coreloop.jl:103-109: `nw(nothing, u, p, t; RET=Val(:buf_init))` returns (o, aggbuf, extbuf)) ------
# The engine never materialises `o` / `aggbuf` during the RHS; nd_b200_get_buffers fills caller-owned device vectors in the
# reference's layouts (vertex outputs, then per edge the src range and the dst range; aggregation slots).  External inputs
# have no buffer in the engine (they are gathered inside the vertex phase), so the third element is `nothing` unless the
# network has them, in which case the call is refused like every unsupported path (no CPU fallback).
function NetworkDynamics.get_buffers(nw::Network{<:B200Execution}, u::CuArray{Float64}, p, t; initbufs=true, kwargs...)
    isempty(kwargs) || throw(ArgumentError("B200Execution: get_buffers takes no perturbation keywords"))
    NetworkDynamics.has_external_input(nw.im) &&
        throw(ArgumentError("B200Execution: get_buffers of a network with external inputs is not supported (no CPU fallback)"))
    length(u) == nw.im.lastidx_dynamic || throw(ArgumentError("u does not have expected size $(nw.im.lastidx_dynamic)"))
    h = nw.layer.aggregator.handle
    o = CUDA.fill(NaN, nw.im.lastidx_out)            # fill!(o, NaN), src/coreloop.jl:26 (entries no component writes stay NaN)
    aggbuf = CUDA.zeros(Float64, nw.im.lastidx_aggr)
    pp = NetworkDynamics.pdim(nw) > 0 ? pointer(p::CuArray{Float64}) : CuPtr{Float64}(0)
    rc = ccall((:nd_b200_get_buffers, libnd_b200), Cint,
               (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, Ptr{Cvoid}),
               h, pointer(o), pointer(aggbuf), pointer(u), pp, Float64(t), stream().handle)
    _check(rc, h)
    (o, aggbuf, nothing)
end

export B200Execution, B200Aggregator, rk4!, pack_params!
end # module
