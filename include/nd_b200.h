/*
 * nd_b200.h -- C ABI of the B200-native network right-hand-side engine.
 *
 * This is the drop-in boundary for ONE path of JuliaDynamics/NetworkDynamics.jl: the network RHS
 * `nw(du, u, p, t)` (src/coreloop.jl:1-102).  The entry points are what a Julia `ccall` binding
 * of `B200Execution` / `B200Aggregator` needs (see INTEGRATION.md and julia/NetworkDynamicsB200.jl):
 * plain pointers, sizes and status codes; no C++ types, no torch types, no exceptions.
 *
 * Index conventions: everything the CALLER passes in a descriptor is in the reference's own
 * convention -- 1-based Int64, exactly the numbers held by `IndexManager`
 * (src/network_structure.jl:1-55) and `ComponentBatch` (src/network_structure.jl:176-222).
 * `du`, `u`, `p` are raw DEVICE pointers in the reference's flat layout
 * (u = [vertex batches | edge batches], src/network_structure.jl:224-258).
 */
#ifndef ND_B200_H
#define ND_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ND_B200_ABI_VERSION 5

/* status codes (Julia glue rethrows: EINVAL -> ArgumentError, others -> ErrorException) */
enum {
  ND_B200_OK = 0,
  ND_B200_EINVAL = 1,        /* size / layout mismatch (src/coreloop.jl:2-7 raise ArgumentError) */
  ND_B200_EUNSUPPORTED = 2,  /* component kind / feature outside the kernel registry: NO CPU fallback */
  ND_B200_ECUDA = 3,
  ND_B200_ENOMEM = 4,
  ND_B200_ETIMEOUT = 5       /* multi-GPU: a peer's boundary outputs did not arrive within the spin budget (sticky) */
};

/* ---- kernel registry: vertex models ------------------------------------------------------- */
enum {
  ND_B200_V_DIFFUSION = 0,             /* test/ComponentLibrary.jl:42-46, benchmark_models.jl:16-20 */
  ND_B200_V_KURAMOTO_FIRST = 1,        /* test/ComponentLibrary.jl:69-72, benchmark_models.jl:32-35 */
  ND_B200_V_KURAMOTO_SECOND = 2,       /* test/ComponentLibrary.jl:59-67, p=(M,D,Pm)               */
  ND_B200_V_KURAMOTO_SECOND_BENCH = 3, /* benchmark/benchmark_models.jl:37-42, p=(P,)              */
  ND_B200_V_SWING_DQ = 4               /* test/ComponentLibrary.jl:139-161, p=(M,D,Pmech,V)        */
};
/* ---- kernel registry: edge models --------------------------------------------------------- */
enum {
  ND_B200_E_DIFFUSION = 0,             /* test/ComponentLibrary.jl:8-13,  e = p*(vs-vd)            */
  ND_B200_E_DIFFUSION_NOP = 1,         /* benchmark/benchmark_models.jl:5-9, e = vs-vd             */
  ND_B200_E_KURAMOTO = 2,              /* test/ComponentLibrary.jl:51-57, e = K*sin(ths-thd)       */
  ND_B200_E_LINE_DQ = 3,               /* test/ComponentLibrary.jl:212-245, p=(R,X,active)         */
  /* edges WITH states (dim > 0): f(de,e,vsrc,vdst,p,t) registered here, outputs are StateMasks (ebatch.mask_*) */
  ND_B200_E_DIFFUSION_ODE = 4,         /* test/ComponentLibrary.jl:30-40, dim 2, p=(tau,)          */
  ND_B200_E_RELAX_ODE = 5,             /* test/diffusion_test.jl:96-101, dim 2, no parameters      */
  ND_B200_E_DIFFUSION_FID = 6,         /* test/ComponentLibrary.jl:22-28, static two-sided g (coupling = ND_B200_FIDUCIAL) */
  /* LoopbackConnection (src/post_utils.jl:105-190): Directed(LOOPBACK_G), out_dst = -in_src, from an "injector" leaf to
   * its hub; the injector's input is the hub's output (apply_loopback!, src/coreloop.jl:47).  vdepth == edepth. */
  ND_B200_E_LOOPBACK = 7
};
/* edge output wrappers, src/component_functions.jl:117-203 */
enum { ND_B200_ANTISYMMETRIC = 0, ND_B200_SYMMETRIC = 1, ND_B200_DIRECTED = 2,
       ND_B200_FIDUCIAL = 3 /* static edges: the edge's own two-sided g(osrc, odst, ...) (ND_B200_E_DIFFUSION_FID or a
                               user-supplied kind); edges with states: Fiducial(src=mask, dst=mask) */ };

/* One `ComponentBatch` of vertices (src/network_structure.jl:176-222 + register_vertices! :224-239).
 * `*_first` are the `first` fields of the batch's BatchStrides (1-based); widths are the strides. */
typedef struct nd_b200_vbatch {
  int32_t kind;            /* ND_B200_V_*                                    */
  int32_t dim, pdim, outdim;
  int64_t count;
  const int64_t* indices;  /* batch.indices: 1-based vertex ids              */
  int64_t state_first;     /* statestride.first                              */
  int64_t p_first;         /* pstride.first                                  */
  int64_t out_first;       /* outbufstride.first                             */
  int64_t aggr_first;      /* inbufstride.first (aggbuf)                     */
  /* external inputs (src/external_inputs.jl; user-supplied kinds only): extdim scalars per component, gathered before f
   * like collect_externals! (src/coreloop.jl:61).  ext_src[i*extdim + k] is entry extbuf_range(batch,i)[k] of the
   * reference's ExtMap: > 0 = StateBufIdx (1-based index into u), < 0 = -(OutBufIdx) (1-based index into o; must be an
   * output of a vertex or of an edge with states -- outputs of feed-forward components are refused like in the
   * reference, src/external_inputs.jl:42-44).  extdim = 0 / ext_src = NULL: none. */
  int32_t extdim, reserved;
  const int64_t* ext_src;
} nd_b200_vbatch;

/* One `ComponentBatch` of edges (register_edges!, src/network_structure.jl:240-258). */
typedef struct nd_b200_ebatch {
  int32_t kind;            /* ND_B200_E_*                                    */
  int32_t coupling;        /* ND_B200_ANTISYMMETRIC | SYMMETRIC | DIRECTED   */
  int32_t dim, pdim, outdim_src, outdim_dst;
  int64_t count;
  const int64_t* indices;  /* batch.indices: 1-based edge ids                */
  int64_t state_first, p_first, out_first, gbuf_first;
  /* edges with states (dim > 0; "ODE edges", src/coreloop.jl:41,76): the output function must be a StateMask
   * (src/component_functions.jl:81-99) over a contiguous index range -- dst output k = state mask_dst_first + k
   * (1-based), wrapped by AntiSymmetric / Symmetric / Directed, or Fiducial(src=..., dst=...) with its own
   * mask_src_first.  0 when dim == 0. */
  int32_t mask_src_first, mask_dst_first;
  int32_t extdim, reserved;      /* external inputs of edges with states, as in nd_b200_vbatch */
  const int64_t* ext_src;
} nd_b200_ebatch;

/* ---- user-supplied component kinds (runtime-compiled, NVRTC) ---------------------------------------------------------
 * Lifts the fixed registry: a component function the caller can state as CUDA C++ -- hand-written, or
 * printed from a symbolic model (ModelingToolkit / Symbolics `build_function(...; target=CTarget())`, the path
 * ext/NetworkDynamicsMTKExt.jl:497-518 takes to a RuntimeGeneratedFunction) -- is spliced into the same fused kernels
 * and compiled for sm_100a when the engine is created.  Bodies see exactly the arguments of the reference's
 * component functions (src/component_functions.jl:251-329,532-567), as plain double pointers:
 *   vertex f : void f(double* dv, const double* v, const double* esum, const double* p, double t)
 *   vertex g : void g(double* out, const double* v, const double* p, double t)        (NULL = StateMask(1:outdim))
 *   edge g   : void g(double* e_dst, const double* v_src, const double* v_dst, const double* p, double t)
 *              wrapped by AntiSymmetric / Symmetric / Directed (ebatch.coupling), or, with two_sided = 1 and
 *              coupling = ND_B200_FIDUCIAL, the reference's own two-sided form
 *              void g(double* e_src, double* e_dst, const double* v_src, const double* v_dst, const double* p, double t)
 *   edge f   : void f(double* de, const double* e, const double* v_src, const double* v_dst, const double* p, double t)
 *              for edges with states (custom_kind.dim > 0); their outputs are the StateMasks of ebatch.mask_*
 *   with external inputs (custom_kind.extdim > 0) f takes them like in the reference, right after the inputs:
 *              vertex f(dv, v, esum, ext, p, t);   edge f(de, e, v_src, v_dst, ext, p, t)
 * Compiled with --fmad=false like the registry kernels.  A body that does not compile fails nd_b200_create with
 * ND_B200_EINVAL and the NVRTC log in nd_b200_last_error. */
#define ND_B200_CUSTOM_KIND_BASE 1000
typedef struct nd_b200_custom_kind {
  int32_t kind;            /* >= ND_B200_CUSTOM_KIND_BASE; referenced by vbatch.kind / ebatch.kind */
  int32_t role;            /* 0 = vertex, 1 = edge                                                */
  int32_t dim, pdim, outdim;  /* vertex: states, parameters, outputs; edge: states (0 = static), parameters, outdim.dst */
  int32_t two_sided;       /* edge only: body has the Fiducial signature                         */
  const char* f_body;      /* vertex: body of f; static edge: body of g; edge with states: body of f */
  const char* g_body;      /* vertex: body of g, or NULL for StateMask(1:outdim); edge: NULL      */
  int32_t extdim;          /* number of external inputs f takes (0: none)                         */
  int32_t g_ff;            /* vertex only: g is feed forward -- void g(double* out, const double* v, const double* ins,
                              const double* p, double t); allowed for leaves behind a LoopbackConnection
                              (src/construction.jl:52-80); dim may be 0 (PureFeedForward)            */
} nd_b200_custom_kind;

typedef struct nd_b200_desc {
  int32_t abi_version;     /* ND_B200_ABI_VERSION                            */
  int32_t device;          /* CUDA device ordinal                            */
  int64_t nv, ne;
  const int64_t* edge_src; /* im.edgevec[i].src, 1-based, edges(g) order     */
  const int64_t* edge_dst;
  int32_t vdepth, edepth;  /* im.vdepth, im.edepth                           */
  int32_t n_vbatches, n_ebatches;
  const nd_b200_vbatch* vbatches;
  const nd_b200_ebatch* ebatches;
  int64_t lastidx_dynamic, lastidx_p, lastidx_out, lastidx_aggr;
  /* Multi-GPU vertex partition: this engine evaluates aggregation-slot rows
   * [row_begin, row_end) (0-based, slot = (v_aggr.first-1)/edepth) and writes only their
   * states in du.  row_end <= 0 means "all rows".  Edge batches WITH states: the engine also
   * evaluates f for edges [count*row_begin/nrows, count*row_end/nrows) of each such batch (0-based
   * positions in the batch, integer division) and writes their states -- contiguous chunks that tile
   * the batch when the ranks' row ranges tile the rows.  Needs the complete u (all-gather exchange);
   * not available together with gather_offset.                              */
  int64_t row_begin, row_end;
  /* rows with more than this many incoming entries are reduced by a whole thread block with a
   * fixed-shape tree instead of one sequential thread; <= 0 -> default (128); INT32_MAX ->
   * strictly sequential per-row accumulation everywhere (debug).            */
  int32_t long_row_threshold;
  int32_t flags;           /* ND_B200_FLAG_*                                 */
  /* Multi-GPU packed halo (optional, NULL = none; needs StateMask vertices with one output, i.e. the gather
   * source is the state vector).  gather_offset[v-1] = 0-based position of vertex v's output in the gather
   * source of THIS engine: positions < lastidx_dynamic address the caller's state vector u (vertices of owned
   * rows must map to their own state), positions in [lastidx_dynamic, gather_len) address the halo buffer that
   * the peers fill through nd_b200_rhs_exchange.  Vertices this engine never reads may hold any valid value. */
  const int64_t* gather_offset;
  int64_t gather_len;
  int32_t n_custom;        /* user-supplied component kinds referenced by the batches */
  int32_t reserved;
  const nd_b200_custom_kind* custom;
} nd_b200_desc;

#define ND_B200_FLAG_NO_EXPORT 1  /* do not keep host copies of the CSR for nd_b200_export_tables */
#define ND_B200_FLAG_ROW_RANGE 4  /* row_begin / row_end are a row partition even when it is EMPTY (row_begin == row_end);
                                   * without the flag row_end <= 0 means "all rows" */
#define ND_B200_FLAG_HOST_ONLY 2  /* build every table on the host and stop: no CUDA call is made, the engine can only
                                     export its tables (layout tests on machines without a GPU) */

typedef struct nd_b200_engine nd_b200_engine;

/* Replaces: aggregator construction `aggregator(im, edgebatches)` (src/construction.jl:198) plus the
 * device adaptation of ext/NetworkDynamicsCUDAExt.jl:16-45.  Builds the destination-sorted CSR. */
int nd_b200_create(const nd_b200_desc* desc, nd_b200_engine** out);
/* The same for a HOMOGENEOUS network given as a bare edge list (BASELINE config 5: 5e7 vertices, 4e8 edges): one registry
 * vertex kind, one registry edge kind + wrapper; the flat u / p layout is the one the reference's constructor produces
 * for such a network (vertex states, then nothing; vertex parameters, then edge parameters in edges(g) order).
 * edge_src / edge_dst: 1-based, edges(g) order.  Saves the caller the per-component tables of nd_b200_desc.
 * gather_offset / gather_len: as in nd_b200_desc (row-partitioned engine with a packed halo). */
int nd_b200_create_from_edgelist(int32_t device, int64_t nv, int64_t ne, const int64_t* edge_src, const int64_t* edge_dst,
                                 int32_t vertex_kind, int32_t edge_kind, int32_t coupling, int64_t row_begin,
                                 int64_t row_end, int32_t flags, const int64_t* gather_offset /* NULL: no halo layout */,
                                 int64_t gather_len, nd_b200_engine** out);
void nd_b200_destroy(nd_b200_engine*);
/* message of the last failing call on this engine (engine == NULL: last failing create on this thread) */
const char* nd_b200_last_error(const nd_b200_engine*);
int nd_b200_abi_version(void);

/* Replaces `(nw::Network)(du,u,p,t)` (src/coreloop.jl:1-102).  Device pointers; asynchronous on
 * `stream` (a cudaStream_t; NULL = legacy default stream); no host synchronisation; p may be NULL
 * when lastidx_p == 0.  Fully defines du for the rows this engine owns. */
int nd_b200_rhs(nd_b200_engine*, double* du, const double* u, const double* p, double t, void* stream);

/* Same call with HOST buffers (pageable or pinned): H2D copies of u and p, the RHS, D2H copy of du,
 * then a stream synchronise.  This is the end-to-end form a CPU-resident caller would use. */
int nd_b200_rhs_host(nd_b200_engine*, double* du_host, const double* u_host, const double* p_host, double t);

/* Optional contract with the caller (no reference counterpart): copy the edge parameters into the engine's own per-entry
 * array, in the entry order of its device layout, and evaluate every following RHS / RK4 / get_buffers call from that
 * copy -- the kernels then read parameters coalesced with the index stream instead of through one scattered 8-byte read
 * per entry.  The caller guarantees that the EDGE parameters in p do not change until the next nd_b200_pack_params;
 * vertex parameters are still read from p on every call.  p == NULL returns to reading p on every call (the default, the
 * reference's semantics: callbacks may mutate p between calls).  Available for networks with ONE registry edge batch
 * that has parameters, default launch shape; ND_B200_EUNSUPPORTED otherwise.  nd_b200_rk4 does this by itself for the
 * duration of a call (p is constant there) when p exceeds 256 MB and the call takes >= 4 steps; environment
 * ND_B200_RK4_PACK=1 / 0 forces / forbids it. */
int nd_b200_pack_params(nd_b200_engine*, const double* p, void* stream);

/* Replaces `get_buffers(nw,u,p,t)` = RET=Val(:buf_init) (src/coreloop.jl:33-36,92-94,103-109):
 * materialises the output buffer o (lastidx_out) and the aggregation buffer (lastidx_aggr). */
int nd_b200_get_buffers(nd_b200_engine*, double* o, double* aggbuf, const double* u, const double* p,
                        double t, void* stream);

/* Replaces `aggregate!(aggregator, aggbuf, o)` (src/aggregators.jl:140-151, the SequentialAggregator sweep; call site
 * src/coreloop.jl:90): adds the edge-output block of a materialised output buffer o (lastidx_out, device) into aggbuf
 * (lastidx_aggr, device) -- per slot in ascending `o` order, ON TOP of aggbuf's present content, which is what
 * test/aggregators_test.jl:69-79 requires of every Aggregator.  The RHS itself never materialises o (the sums live in
 * registers); this entry point exists so that the Aggregator plug-in is complete on its own.  Full (unpartitioned)
 * engines created without ND_B200_FLAG_NO_EXPORT. */
int nd_b200_aggregate(nd_b200_engine*, double* aggbuf, const double* o, void* stream);

/* Classical fixed-step RK4 on device-resident u (in place), stage updates fused into the RHS
 * kernels, step captured in a CUDA graph.  Replaces `solve(ODEProblem(nw,...), RK4(); dt, adaptive=false)`
 * around src/post_utils.jl:43-100 for the registry models. */
int nd_b200_rk4(nd_b200_engine*, double* u, const double* p, double t0, double dt, int64_t nsteps, void* stream);

/* sizes[0]=nrows (owned) [1]=nentries [2]=nblocks [3]=n_long_rows [4]=gather_from_u (0/1)
 * [5]=kernel launches per RHS [6]=row_begin [7]=row_end */
int nd_b200_export_sizes(const nd_b200_engine*, int64_t sizes[8]);
/* The CSR the engine built, for the bit-exact index check: rowptr[nrows+1] (0-based offsets),
 * and per entry the 1-based neighbour vertex id, 1-based edge id, and side (0: this row is the
 * edge's dst, 1: this row is the edge's src).  Entries of a row are in accumulation order. */
int nd_b200_export_tables(const nd_b200_engine*, int64_t* rowptr, int64_t* nbr_vertex, int64_t* edge_id,
                          int32_t* side);

/* The jagged device layout of the default kernel (rhs_jag_kernel), ND_B200_FLAG_HOST_ONLY engines only.
 * sizes[0]=slices (-1: engine uses a tile kernel) [1]=rows reduced by a whole block [2]=row split width [3]=entries
 * [4]=first slice (tile, for the tile kernel) that reads the halo [5]=outputs in the halo (-1: no halo layout).
 * slices[4*s..] = {entry base, first row, vertex batch, max parts}; lanes[32*s+l] = len | rowrel<<6 (7 bits) | head<<13 |
 * valid<<14; longs[4*k..] = {entry base, row, entries, vertex batch}; order[k] = CSR entry (as exported by
 * nd_b200_export_tables) stored at jagged position k. */
int nd_b200_export_jag_sizes(const nd_b200_engine*, int64_t sizes[6]);
int nd_b200_export_jag(const nd_b200_engine*, int32_t* slices, uint16_t* lanes, int32_t* longs, int32_t* order);

/* the CUDA source generated for an engine with user-supplied kinds (NULL otherwise); owned by the engine */
const char* nd_b200_custom_source(const nd_b200_engine*);

/* name of the kernel family that evaluates this engine's RHS: "rhs_fused_kernel" (tile kernel), "rhs_jag_kernel" (jagged
 * warp slices), "rhs_js_kernel" (jagged slices streamed through shared memory by TMA bulk copies, persistent warps),
 * "edge_pass_kernel+row_pass_kernel" (split mode); static string */
const char* nd_b200_kernel_name(const nd_b200_engine*);

/* kernel launches issued by this engine since creation (bench.py's gpu_launches) */
int64_t nd_b200_launch_count(const nd_b200_engine*);
/* average device time [ms] of the fused RHS kernel over the last `nd_b200_rhs` calls made while
 * timing was enabled (CUDA events on the caller's stream). */
int nd_b200_set_timing(nd_b200_engine*, int enabled);
int nd_b200_timings(nd_b200_engine*, double* fused_ms_avg, double* prepass_ms_avg, int64_t* ncalls);

/* ---- multi-GPU: packed halo exchange over NVLink peer memory (one process per GPU) ---------------------------------
 * The reference has no multi-device path.  A `comm` owns this rank's double-buffered HALO buffer -- the
 * outputs of exactly the remote vertices its rows read, grouped by owner rank and sorted by state offset -- plus an
 * arrival-flag array, all in ONE device allocation that the other ranks map through CUDA IPC.  The engine is created
 * with nd_b200_desc.gather_offset pointing into [u | halo].  nd_b200_rhs_exchange = "pack the outputs every peer needs
 * from my state vector straight into that peer's halo buffer with NVLink stores, raise my flag everywhere" + "RHS kernel
 * whose interior tiles start at once and whose boundary tiles wait for the flags".  No NCCL, no host synchronisation on
 * the data path.  Collective: every rank must make the same sequence of calls, all on one stream per rank. */
typedef struct nd_b200_comm nd_b200_comm;
#define ND_B200_IPC_HANDLE_BYTES 64
/* halo_len: outputs in THIS rank's halo; max_halo_len: the largest halo_len among the ranks (one layout everywhere) */
int nd_b200_comm_create(int32_t device, int32_t rank, int32_t world, int64_t halo_len, int64_t max_halo_len,
                        nd_b200_comm** out);
/* opaque handle (ND_B200_IPC_HANDLE_BYTES bytes) to ship to the peers (e.g. torch.distributed all_gather_object) */
int nd_b200_comm_export(nd_b200_comm*, void* handle_out);
int nd_b200_comm_open_peer(nd_b200_comm*, int32_t peer, const void* handle);
/* what this rank sends to `peer` on every exchange: the 0-based state offsets (ascending) of the outputs that peer's
 * rows read from this rank, stored contiguously from position dst_offset of the peer's halo buffer */
int nd_b200_comm_set_send(nd_b200_comm*, int32_t peer, const int64_t* state_offsets, int64_t n, int64_t dst_offset);
/* replaces `(nw::Network)(du,u,p,t)` for a row-partitioned engine: only the owned states of u need to be valid */
int nd_b200_rhs_exchange(nd_b200_engine*, nd_b200_comm*, double* du, const double* u, const double* p, double t,
                         void* stream);
/* replaces `solve(ODEProblem(nw,...), RK4(); dt, adaptive=false)` for a row-partitioned engine: classical RK4 with four
 * exchanging launches per step (every stage packs the boundary outputs of its own input for the peers and applies the fused
 * stage update to the owned states).  Collective like nd_b200_rhs_exchange; only the owned states of u are read / advanced. */
int nd_b200_rk4_exchange(nd_b200_engine*, nd_b200_comm*, double* u, const double* p, double t0, double dt, int64_t nsteps,
                         void* stream);
/* timing aid (exposed halo time = exchange call - this): the owned rows evaluated on whatever the halo buffer
 * currently holds; no publish, no wait */
int nd_b200_rhs_local(nd_b200_engine*, nd_b200_comm*, double* du, const double* u, const double* p, double t,
                      void* stream);
/* *timed_out = 1 if some RHS kernel gave up waiting for a peer (~2 s spin budget; results are then invalid) */
int nd_b200_comm_status(nd_b200_comm*, int32_t* timed_out);
const char* nd_b200_comm_last_error(const nd_b200_comm*);
void nd_b200_comm_destroy(nd_b200_comm*);

/* pinned host memory for nd_b200_rhs_host callers */
void* nd_b200_host_alloc(int64_t bytes);
void nd_b200_host_free(void*);

#ifdef __cplusplus
}
#endif
#endif
