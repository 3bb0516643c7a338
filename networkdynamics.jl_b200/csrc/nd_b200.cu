// nd_b200.cu -- engine object + C ABI (include/nd_b200.h) of the B200-native network RHS.
//
// Construction (host, C++): from the reference's own tables (IndexManager ranges + ComponentBatches,
// src/network_structure.jl:1-55,176-258) build a destination-sorted CSR over aggregation slots whose
// per-row entry order is the accumulation order of SequentialAggregator (src/aggregators.jl:140-151):
// ascending position in the output buffer `o` = (edge batch, position in batch, src-out before dst-out).
// Evaluation: see nd_b200_kernels.cuh.
#include "nd_b200_kernels.cuh"

#include <nvrtc.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace ndb;

// Kernel launch.  The kernel name comes last so that template argument lists (they contain commas) survive the macro.
// ND_CUSIM: tests/cusim/ compiles this very file with g++ against an emulation of the CUDA runtime that executes the kernel
// sources thread by thread on the CPU (fibers, warp collectives, block barriers) so that the CPU test suite exercises the
// real kernels and launch logic.  That build is a test double: it is never part of libnd_b200.so and the package never loads it.
thread_local bool g_pdl_launch = false;   // see nd_launch_kernel below
#ifdef ND_CUSIM
#define ND_LAUNCH(grid, block, stream, args, ...) \
  cusim::launch(dim3((unsigned)(grid)), dim3((unsigned)(block)), (stream), [=]() { __VA_ARGS__ args; })
#else
// nd_launch_kernel: plain <<<>>> launch, or -- while g_pdl_launch is set (the stage kernels captured by nd_b200_rk4) -- a
// launch with cudaLaunchAttributeProgrammaticStreamSerialization (programmatic dependent launch, see pdl_wait() in the kernels)
#define ND_UNPACK(...) __VA_ARGS__
#define ND_LAUNCH(grid, block, stream, args, ...) nd_launch_kernel(__VA_ARGS__, dim3((unsigned)(grid)), dim3((unsigned)(block)), (stream), ND_UNPACK args)
template <class... KA, class... A>
inline void nd_launch_kernel(void (*kernel)(KA...), dim3 grid, dim3 block, cudaStream_t st, A&&... a) {
  if (g_pdl_launch) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(a)...);
  } else {
    kernel<<<grid, block, 0, st>>>(static_cast<KA>(a)...);
  }
}
#endif

namespace {

thread_local std::string g_create_error;

constexpr int MAX_VB = 64;
constexpr int MAX_EB = 255;   // edge batch id is stored per entry as uint8

struct HostVB { int kind, dim, pdim, outdim; long long count, state0, p0, out0, row0; int ff; };
struct HostEB { int kind, coupling, dim, pdim, osrc, odst; long long count, p0, out0; long long state0; int mask_src, mask_dst; };   // masks 0-based

}  // namespace

struct nd_b200_engine {
  int device = 0;
  std::string err;
  long long nv = 0, ne = 0;
  int vdepth = 1, edepth = 1;
  long long lastidx_dynamic = 0, lastidx_p = 0, lastidx_out = 0, lastidx_aggr = 0;
  long long nrows_total = 0, row_begin = 0, row_end = 0;
  long long nentries = 0;
  int nblocks = 0, n_long = 0;
  int long_thr = 128;
  int gather_from_u = 1;
  int ek = EK_GENERIC;        // edge kind of the single edge batch, or EK_GENERIC
  bool compact = false;       // tile layout with compact entry words (offset | local row << 23 | side << 30), kernels <..., CR = true>
  int block = 256, ept = 8;   // launch shape of the fused kernel
  std::vector<HostVB> hvb;
  std::vector<HostEB> heb;
  // device tables
  int *d_rowptr = nullptr, *d_nbr = nullptr, *d_epar = nullptr, *d_blk_row = nullptr;
  uint8_t* d_ebid = nullptr;
  VBDev* d_vb = nullptr;
  EBDev* d_eb = nullptr;
  double *d_vout[2] = {nullptr, nullptr};
  std::vector<std::pair<long long, long long>> own_segs;   // owned state ranges (0-based start, length)
  int4* d_tiles = nullptr;   // one descriptor per thread block
  int ntiles = 0;
  // evaluation mode: 0 = single fused kernel (default); 1 = edge pass + row pass around the edge-output buffer
  // ("edge once", ND_B200_KERNEL=split; not available for row-partitioned engines)
  int split = 0;
  long long ne_all = 0;       // edges in `o` order
  long long oedge_len = 0;    // scalars in the edge part of `o`
  long long oedge_base = 0;   // = nv*vdepth: position of the edge part inside `o`
  int *d_es = nullptr, *d_et = nullptr, *d_eepar = nullptr, *d_eooff = nullptr, *d_oidx = nullptr;
  uint8_t* d_eebid = nullptr;
  double* d_oedge = nullptr;
  // jagged layout (ND_B200_KERNEL=jag): warp slices, see rhs_jag_kernel
  int jag = 0, jag_u = 2, jag_wps = 0, jsplit = 32;
  // column-blocked evaluation (ND_B200_L2_BLOCKS, see nd_b200_create): `blocks` are complete engines over the SAME rows, each
  // holding only the entries whose neighbour output lies in [col_lo, col_hi) of the gather source; nd_b200_rhs runs them in
  // ascending column order, the row sums travel through d_blockacc
  std::vector<nd_b200_engine*> blocks;
  double* d_blockacc = nullptr;
  bool col_filter = false;                    // this engine IS one of those blocks
  long long col_lo = 0, col_hi = 0;
  bool cols_sorted = false;                   // every row's entries are in ascending neighbour-offset order (set by build_csr)
  int rk4_pdl = 0;                            // programmatic dependent launch between the stage kernels of nd_b200_rk4's graph (ND_B200_RK4_PDL)
  int jag_win = 0;                            // window mode of rhs_jag_kernel: every block = one 128-row window (slice table padded)
  int jag_persist = 0, jag_block = 128;   // persistent warps (rhs_jag_persist_kernel) / 64-thread blocks, ND_B200_JAG_PERSIST, ND_B200_JAG_BLOCK
  int num_sms = 148;
  unsigned long long* d_coop_bar = nullptr;   // arrival counter of the persistent RK4 kernel's grid barrier
  int jaga = 0, jaga_ch = 8, jaga_wps = 48;   // asynchronous-gather variant of the jagged kernel (rhs_jaga_kernel), columns per chunk
  int pf_dist = 0;                            // L2 prefetch distance of the jagged kernels' entry streams (entries), ND_B200_PF_DIST
  int jagb = 0;                               // batched variant (rhs_jagb_kernel); shares ND_B200_JAGA_CH / ND_B200_JAGA_WPS
  int nslices = 0, n_jag_blocks = 0, n_jlong = 0;
  int4 *d_jslices = nullptr, *d_jlong = nullptr;
  uint16_t* d_jlanes = nullptr;
  int* d_jnbr = nullptr;
  int2* d_jent = nullptr;
  uint8_t* d_jebid = nullptr;
  // streamed jagged kernel (rhs_js_kernel, ND_B200_KERNEL=js): persistent warps, each owning the contiguous slices
  // [jwarp.x, jwarp.y); jag_len = length of the (padded) entry stream; launch parameters
  int jstream = 0, js_u = 8, js_nst = 8, js_wps = 16;
  int n_jwarps = 0;
  long long jag_len = 0;
  int2* d_jwarp = nullptr;
  std::vector<int2> h_jwarp;
  // host-buffer pipeline (nd_b200_rhs_host): per thread block of the launch grid, the end of the parameter range it
  // reads and the rows it writes; launch_nblk >= 0 restricts a launch to blocks [P.blk_off, P.blk_off + launch_nblk)
  std::vector<int> blk_pmax, blk_rmin, blk_rmax;
  bool blk_rows_monotone = false;
  int launch_nblk = -1;
  cudaStream_t s_copy = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_pipe;
  int halo_base = INT_MAX;    // gather offsets >= halo_base address the halo buffer (multi-GPU packed halo)
  long long gather_len = 0;
  int wait_from = 0;          // first tile / slice that reads the halo
  // user-supplied component kinds: the kernels are compiled at creation time (NVRTC) from the same header
  struct CustomKind { int kind, role, dim, pdim, outdim, two_sided; std::string f_body, g_body; int extdim, g_ff; };
  std::vector<CustomKind> customs;
  bool custom = false;
  int c_pe = 0, c_maxdim = 1;          // template PE (largest edge pdim) and ND_MAX_VDIM of the generated kernels
  std::string custom_src;
  cudaLibrary_t c_lib = nullptr;
  cudaKernel_t c_fused = nullptr, c_jag = nullptr, c_vout = nullptr, c_eout = nullptr, c_ef = nullptr;
  // edge batches with states ("ODE edges"): their f runs in edge_f_kernel after the row kernel; per edge of the batch the
  // gather offsets of its two vertex outputs
  struct OdeBatch { int b; int* d_es; int* d_et; int* d_ext; int extdim; };
  std::vector<int*> d_vext;            // per vertex batch: external-input codes (nullptr: none)
  std::vector<int*> d_ffin;            // per vertex batch: hub output offsets of feed-forward (injector) vertices
  bool any_ff = false;
  int c_maxext = 0;                    // ND_MAX_EXT of the generated kernels
  std::vector<OdeBatch> ode;
  int c_maxedim = 1;                   // ND_MAX_EDIM of the generated kernels
  // packed edge parameters (nd_b200_pack_params): per entry, in the entry order of the layout in use
  double* d_ppack = nullptr;
  bool pack_on = false;
  int pack_pe = 0;                     // edge pdim of the single edge batch (0: packing unavailable)
  bool graph_packed = false;           // the captured RK4 graph was built with pack_on
  bool host_only = false;     // ND_B200_FLAG_HOST_ONLY: tables built, nothing uploaded (layout tests without a GPU)
  std::vector<int4> h_jslices, h_jlong;
  std::vector<uint16_t> h_jlanes;
  std::vector<int> h_jorder;  // jagged position -> CSR entry
  // get_buffers support (lazy)
  std::vector<std::vector<int>> h_esrc_off, h_edst_off;   // per edge batch, gather offsets
  std::vector<int*> d_esrc_off, d_edst_off;
  // nd_b200_aggregate support (lazy): per CSR entry the 0-based index of its output block in `o` (-1: the entry has no
  // slot in `o` -- the hub -> injector entry of a loopback edge)
  std::vector<long long> h_aggidx;
  long long* d_aggidx = nullptr;
  int* d_aggrow = nullptr;
  // host copies for export
  std::vector<long long> h_rowptr;
  std::vector<int> h_nbr_vid, h_eid;
  std::vector<uint8_t> h_side;
  // rk4
  double *d_tmpA = nullptr, *d_tmpB = nullptr, *d_ksum = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int graph_steps = 0;
  const double *graph_u = nullptr, *graph_p = nullptr;
  double graph_dt = 0.0;
  cudaStream_t cap_stream = nullptr;
  // host-buffer path
  cudaStream_t own_stream = nullptr;
  double *d_hu = nullptr, *d_hp = nullptr, *d_hdu = nullptr;
  // stats
  long long launches = 0;
  bool timing = false;
  std::vector<cudaEvent_t> ev;   // pairs: [2i] start, [2i+1] stop of the fused kernel; prepass pairs in ev_pre
  std::vector<cudaEvent_t> ev_pre;
  size_t ev_used = 0, ev_pre_used = 0;
};

namespace {

int fail(nd_b200_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CUDA_TRY(e, call)                                                                      \
  do {                                                                                         \
    cudaError_t _c = (call);                                                                   \
    if (_c != cudaSuccess) return fail(e, ND_B200_ECUDA, "%s: %s", #call, cudaGetErrorString(_c)); \
  } while (0)

template <typename T>
int upload(nd_b200_engine* e, T** dst, const std::vector<T>& src) {
  size_t bytes = sizeof(T) * std::max<size_t>(src.size(), 1);
  CUDA_TRY(e, cudaMalloc((void**)dst, bytes));
  if (!src.empty()) CUDA_TRY(e, cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
  return 0;
}

// registry: the (dim, pdim, outdim) each kernel was written for
bool vertex_kind_ok(const nd_b200_vbatch& b, std::string& why);
bool edge_kind_ok(const nd_b200_ebatch& b, int vdepth, std::string& why);
const nd_b200_engine::CustomKind* find_custom(const nd_b200_engine* e, int kind, int role) {
  for (const auto& c : e->customs)
    if (c.kind == kind && c.role == role) return &c;
  return nullptr;
}

bool vertex_kind_ok(const nd_b200_engine* e, const nd_b200_vbatch& b, std::string& why) {
  if (b.kind >= ND_B200_CUSTOM_KIND_BASE) {
    const auto* c = find_custom(e, b.kind, 0);
    if (!c) { why = "vertex kind " + std::to_string(b.kind) + " is not among the descriptor's custom kinds"; return false; }
    if (c->dim != b.dim || c->pdim != b.pdim || c->outdim != b.outdim || c->extdim != b.extdim) { why = "custom vertex kind " + std::to_string(b.kind) + " declared with other (dim,pdim,outdim,extdim)"; return false; }
    if (c->g_body.empty() && b.outdim > b.dim) { why = "custom vertex kind " + std::to_string(b.kind) + ": StateMask(1:outdim) needs outdim <= dim"; return false; }
    return true;
  }
  return vertex_kind_ok(b, why);
}

bool vertex_kind_ok(const nd_b200_vbatch& b, std::string& why) {
  if (b.extdim != 0) { why = "external inputs need a user-supplied vertex kind"; return false; }
  struct R { int kind, dim, pdim, outdim; };
  static const R reg[] = {{ND_B200_V_DIFFUSION, 1, 0, 1}, {ND_B200_V_KURAMOTO_FIRST, 1, 1, 1},
                          {ND_B200_V_KURAMOTO_SECOND, 2, 3, 1}, {ND_B200_V_KURAMOTO_SECOND_BENCH, 2, 1, 1},
                          {ND_B200_V_SWING_DQ, 2, 4, 2}};
  for (const R& r : reg)
    if (r.kind == b.kind) {
      if (r.dim != b.dim || r.pdim != b.pdim || r.outdim != b.outdim) {
        why = "vertex kind " + std::to_string(b.kind) + " registered with (dim,pdim,outdim)=(" + std::to_string(r.dim) + "," +
              std::to_string(r.pdim) + "," + std::to_string(r.outdim) + ")";
        return false;
      }
      return true;
    }
  why = "vertex kind " + std::to_string(b.kind) + " is not in the B200 kernel registry";
  return false;
}
bool edge_kind_ok(const nd_b200_engine* e, const nd_b200_ebatch& b, int vdepth, std::string& why) {
  if (b.kind >= ND_B200_CUSTOM_KIND_BASE) {
    const auto* c = find_custom(e, b.kind, 1);
    if (!c) { why = "edge kind " + std::to_string(b.kind) + " is not among the descriptor's custom kinds"; return false; }
    if (c->pdim != b.pdim || c->outdim != b.outdim_dst || c->dim != b.dim || c->extdim != b.extdim) { why = "custom edge kind " + std::to_string(b.kind) + " declared with other (dim,pdim,outdim,extdim)"; return false; }
    if (b.extdim > 0 && b.dim == 0) { why = "custom edge kind " + std::to_string(b.kind) + ": external inputs need an edge with states (static edges are feed-forward)"; return false; }
    if (b.dim > 0) {
      if (c->two_sided) { why = "custom edge kind " + std::to_string(b.kind) + ": an edge with states supplies f; its outputs are StateMasks"; return false; }
      return true;
    }
    if ((b.coupling == ND_B200_FIDUCIAL) != (c->two_sided != 0)) { why = "custom edge kind " + std::to_string(b.kind) + ": the Fiducial wrapper and a two-sided body go together"; return false; }
    return true;
  }
  if (b.dim == 0 && (b.coupling == ND_B200_FIDUCIAL) != (b.kind == ND_B200_E_DIFFUSION_FID)) { why = "a static Fiducial edge needs a two-sided kind (ND_B200_E_DIFFUSION_FID or a user-supplied one)"; return false; }
  return edge_kind_ok(b, vdepth, why);
}

bool edge_kind_ok(const nd_b200_ebatch& b, int vdepth, std::string& why) {
  if (b.extdim != 0) { why = "external inputs need a user-supplied edge kind"; return false; }
  if (b.kind == ND_B200_E_LOOPBACK) {     // LoopbackConnection: Directed(LOOPBACK_G), any depth with vdepth == edepth
    if (b.coupling != ND_B200_DIRECTED || b.pdim != 0 || b.dim != 0 || b.outdim_dst != vdepth) { why = "a loopback edge is Directed, without parameters or states, and forwards vdepth values"; return false; }
    return true;
  }
  struct R { int kind, pdim, odst, vdepth, dim; };
  static const R reg[] = {{ND_B200_E_DIFFUSION, 1, 1, 1, 0}, {ND_B200_E_DIFFUSION_NOP, 0, 1, 1, 0},
                          {ND_B200_E_KURAMOTO, 1, 1, 1, 0}, {ND_B200_E_LINE_DQ, 3, 2, 2, 0},
                          {ND_B200_E_DIFFUSION_ODE, 1, 1, 1, 2}, {ND_B200_E_RELAX_ODE, 0, 1, 1, 2}, {ND_B200_E_DIFFUSION_FID, 1, 1, 1, 0}};
  for (const R& r : reg)
    if (r.kind == b.kind) {
      if (r.pdim != b.pdim || r.odst != b.outdim_dst || r.vdepth != vdepth || r.dim != b.dim) {
        why = "edge kind " + std::to_string(b.kind) + " registered with (dim,pdim,outdim,vdepth)=(" + std::to_string(r.dim) + "," + std::to_string(r.pdim) + "," +
              std::to_string(r.odst) + "," + std::to_string(r.vdepth) + ")";
        return false;
      }
      return true;
    }
  why = "edge kind " + std::to_string(b.kind) + " is not in the B200 kernel registry";
  return false;
}

void fill_params(const nd_b200_engine* e, KParams& P) {
  memset(&P, 0, sizeof P);
  P.rowptr = e->d_rowptr; P.nbr = e->d_nbr; P.epar = e->d_epar; P.ebid = e->d_ebid; P.blk_row = e->d_blk_row;
  P.vb = e->d_vb; P.eb = e->d_eb; P.n_vb = (int)e->hvb.size(); P.n_eb = (int)e->heb.size();
  P.row_base = (int)e->row_begin; P.long_thr = e->long_thr; P.gather_from_u = e->gather_from_u;
  P.state_edges = e->ode.empty() ? 0 : 1;
  P.tiles = e->d_tiles; P.ntiles = e->ntiles; P.oidx = e->d_oidx; P.oedge = e->d_oedge;
  P.jslices = e->d_jslices; P.jlanes = e->d_jlanes; P.jnbr = e->d_jnbr; P.jent = e->d_jent; P.jebid = e->d_jebid;
  P.jlong = e->d_jlong; P.nslices = e->nslices; P.n_jag_blocks = e->n_jag_blocks;
  P.jwarp = e->d_jwarp; P.n_jwarps = e->n_jwarps; P.n_jlong = e->n_jlong;
  P.pf_dist = e->pf_dist; P.jag_len = e->jag_len;
  P.halo = nullptr; P.halo_base = e->halo_base; P.wait_from = e->wait_from;
  P.ppack = e->pack_on ? e->d_ppack : nullptr;
}

#include "nd_b200_launch.inc"
#include "nd_b200_custom.inc"
#include "nd_b200_build.inc"

int build_engine(nd_b200_engine* e, const nd_b200_desc* d) { return EngineBuilder(e, d).run(); }

int check_call(nd_b200_engine* e, const void* du, const void* u, const void* p) {
  if (!e) return ND_B200_EINVAL;
  if (e->host_only) return fail(e, ND_B200_EUNSUPPORTED, "engine was created with ND_B200_FLAG_HOST_ONLY (tables only, no device)");
  if (!du || !u) return fail(e, ND_B200_EINVAL, "du or u is NULL (expected size %lld)", e->lastidx_dynamic);
  if (e->lastidx_p > 0 && !p) return fail(e, ND_B200_EINVAL, "p is NULL but the network has %lld parameters", e->lastidx_p);
  return 0;
}

struct WaitSpec { const double* halo; const unsigned long long* flags; unsigned long long seq; int world; int* timeout; const HaloParams* pub; int n_pub; int* timeout_host; long long budget; };

int rhs_impl(nd_b200_engine* e, double* du, const double* u, const double* p, double t, cudaStream_t st, int mode,
             double* aggbuf, const WaitSpec* w = nullptr) {
  if (e->halo_base != INT_MAX && !w) return fail(e, ND_B200_EINVAL, "this engine reads a halo buffer (gather_offset): call it through nd_b200_rhs_exchange");
  if (!e->blocks.empty() && mode == MODE_DU && !w && !e->timing) {
    // column-blocked evaluation: block b adds its entries on top of the sums of blocks 0 .. b-1 (same entry order as one pass,
    // because every row's entries are in ascending column order); the last block applies the vertex model
    const size_t nb = e->blocks.size();
    for (size_t b = 0; b < nb; ++b) {
      nd_b200_engine* sub = e->blocks[b];
      KParams Q;
      fill_params(sub, Q);
      Q.u = u; Q.gsrc = u; Q.p = p; Q.t = t;
      Q.acc_in = b > 0 ? e->d_blockacc : nullptr;
      if (b + 1 < nb) { Q.mode = MODE_AGG; Q.aggbuf = e->d_blockacc; }
      else { Q.mode = MODE_DU; Q.du = du; }
      CUDA_TRY(e, launch_fused(sub, Q, st));
    }
    e->launches += (long long)nb;
    return ND_B200_OK;
  }
  KParams P;
  fill_params(e, P);
  P.u = u; P.p = p; P.du = du; P.mode = mode; P.aggbuf = aggbuf; P.t = t;
  P.gsrc = u;
  if (w) {
    P.halo = w->halo; P.wait_flags = w->flags; P.wait_seq = w->seq; P.wait_world = w->world; P.wait_timeout = w->timeout;
    P.wait_timeout_host = w->timeout_host; P.wait_budget = w->budget;
    if (w->pub) {
      P.H = *w->pub; P.n_pub = w->n_pub;
      // no block of this launch would wait for the peers (no row reads the halo, or no rows at all): add the fence block
      const bool waits = e->jag ? (e->wait_from < e->nslices || e->n_jlong > 0) : (e->wait_from < e->nblocks);
      P.fence = waits ? 0 : 1;
    }
  }
  if (e->timing) {
    if (ensure_events(e, e->ev, e->ev_used + 2) || ensure_events(e, e->ev_pre, e->ev_pre_used + 2)) return ND_B200_ECUDA;
  }
  if (!e->gather_from_u) {
    if (e->timing) CUDA_TRY(e, cudaEventRecord(e->ev_pre[e->ev_pre_used], st));
    CUDA_TRY(e, launch_vout(e, u, p, e->d_vout[0], st, t));
    if (e->timing) { CUDA_TRY(e, cudaEventRecord(e->ev_pre[e->ev_pre_used + 1], st)); e->ev_pre_used += 2; }
    P.gsrc = e->d_vout[0];
  }
  if (e->timing) CUDA_TRY(e, cudaEventRecord(e->ev[e->ev_used], st));
  CUDA_TRY(e, launch_fused(e, P, st));
  if (e->timing) { CUDA_TRY(e, cudaEventRecord(e->ev[e->ev_used + 1], st)); e->ev_used += 2; }
  if (!e->ode.empty() && mode == MODE_DU) CUDA_TRY(e, launch_edge_f(e, P, st));
  return ND_B200_OK;
}

void destroy_graph(nd_b200_engine* e) {
  if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
  if (e->graph) cudaGraphDestroy(e->graph);
  e->graph_exec = nullptr; e->graph = nullptr; e->graph_steps = 0;
}

// enqueue one fused RK4 step (4 launches) on st.  u is updated in place.
int rk4_step_enqueue(nd_b200_engine* e, double* u, const double* p, double t, double dt, cudaStream_t st) {
  KParams P;
  fill_params(e, P);
  P.p = p; P.mode = MODE_RK; P.u0 = u; P.ksum = e->d_ksum; P.h6 = dt / 6.0;
  const double h2 = 0.5 * dt;
  const double* in[4] = {u, e->d_tmpA, e->d_tmpB, e->d_tmpA};
  double* out[4] = {e->d_tmpA, e->d_tmpB, e->d_tmpA, u};
  const double hs[4] = {h2, h2, dt, 0.0};
  const double ts[4] = {t, t + h2, t + h2, t + dt};
  for (int s = 0; s < 4; ++s) {
    P.stage = s + 1; P.u = in[s]; P.unext = out[s]; P.hs = hs[s]; P.t = ts[s];
    if (e->gather_from_u) { P.gsrc = in[s]; P.vout_next = nullptr; }
    else if (e->custom) {
      // user-supplied output functions may read t or, for feed-forward vertices, another vertex's output: every stage runs
      // the output pre-pass on its own input at its own time instead of taking the outputs from the previous epilogue
      CUDA_TRY(e, launch_vout(e, in[s], p, e->d_vout[0], st, ts[s]));
      P.gsrc = e->d_vout[0]; P.vout_next = nullptr;
    }
    else { P.gsrc = e->d_vout[s & 1]; P.vout_next = e->d_vout[(s + 1) & 1]; }   // stage 4 leaves outputs of the new u in d_vout[0]
    CUDA_TRY(e, launch_fused(e, P, st));
    if (!e->ode.empty()) CUDA_TRY(e, launch_edge_f(e, P, st));
  }
  return ND_B200_OK;
}

}  // namespace

extern "C" {

int nd_b200_abi_version(void) { return ND_B200_ABI_VERSION; }

int nd_b200_create(const nd_b200_desc* desc, nd_b200_engine** out) {
  if (!desc || !out) return fail(nullptr, ND_B200_EINVAL, "null descriptor or output pointer");
  *out = nullptr;
  nd_b200_engine* e = new (std::nothrow) nd_b200_engine();
  if (!e) return fail(nullptr, ND_B200_ENOMEM, "out of host memory");
  int rc;
  try {
    rc = build_engine(e, desc);
  } catch (const std::bad_alloc&) {
    rc = fail(e, ND_B200_ENOMEM, "out of host memory while building the CSR");
  } catch (...) {
    rc = fail(e, ND_B200_EINVAL, "unexpected failure while building the CSR");
  }
  if (rc != ND_B200_OK) {
    g_create_error = e->err;
    nd_b200_destroy(e);
    return rc;
  }
  // ---- column-blocked evaluation (opt-in, ND_B200_L2_BLOCKS=<k>|auto) --------------------------------------------------------
  // When the vertex outputs do not fit the L2, every gather miss fills a 128-byte line from DRAM (config 5 at full size: 103.7 GB
  // per RHS).  With k blocks the RHS becomes k launches, block b gathering only from the b-th k-th of the outputs
  // (L2 resident for its whole launch).  Single GPU, jagged layout, one vertex output, registry kinds, rows in ascending neighbour
  // order (sorted edge lists) -- otherwise the switch is ignored.
  if (const char* sblk = getenv("ND_B200_L2_BLOCKS")) {
    const long long table_bytes = 8LL * e->lastidx_dynamic;
    long long k = !strcmp(sblk, "auto") ? (table_bytes > 96LL << 20 ? (table_bytes + (48LL << 20) - 1) / (48LL << 20) : 0) : atoll(sblk);
    k = std::min<long long>(k, 64);
    const bool ok = k >= 2 && e->jag && !e->jaga && !e->jagb && !e->jstream && !e->custom && !e->split && e->ode.empty() && e->vdepth == 1 &&
                    e->edepth == 1 && e->gather_from_u && e->halo_base == INT_MAX && e->row_end - e->row_begin == e->nrows_total &&
                    e->cols_sorted && !e->host_only && e->ek != EK_GENERIC && !e->jag_win && !e->jag_persist && e->jag_block != 64;
    if (ok) {
      const long long width = (e->lastidx_dynamic + k - 1) / k;
      for (long long b = 0; b < k && rc == ND_B200_OK; ++b) {
        nd_b200_engine* sub = new (std::nothrow) nd_b200_engine();
        if (!sub) { rc = fail(e, ND_B200_ENOMEM, "out of host memory"); break; }
        sub->col_filter = true; sub->col_lo = b * width; sub->col_hi = std::min<long long>((b + 1) * width, e->lastidx_dynamic);
        nd_b200_desc dsub = *desc;
        dsub.flags |= ND_B200_FLAG_NO_EXPORT;
        try { rc = build_engine(sub, &dsub); } catch (...) { rc = fail(sub, ND_B200_ENOMEM, "out of host memory while building a column block"); }
        if (rc != ND_B200_OK) { e->err = sub->err; nd_b200_destroy(sub); break; }
        e->blocks.push_back(sub);
      }
      if (rc == ND_B200_OK && cudaMalloc((void**)&e->d_blockacc, sizeof(double) * (size_t)e->nrows_total * (size_t)e->edepth) != cudaSuccess)
        rc = fail(e, ND_B200_ECUDA, "cudaMalloc of the column-block row sums failed");
      if (rc != ND_B200_OK) { g_create_error = e->err; nd_b200_destroy(e); return rc; }
    }
  }
  *out = e;
  return ND_B200_OK;
}

/* Homogeneous network straight from an edge list (the config-5 path): one vertex kind, one edge kind, the
 * flat layout the reference's constructor would produce for it (register_vertices!/register_edges!,
 * src/network_structure.jl:224-258) -- without the caller materialising per-component tables. */
int nd_b200_create_from_edgelist(int32_t device, int64_t nv, int64_t ne, const int64_t* edge_src, const int64_t* edge_dst,
                                 int32_t vertex_kind, int32_t edge_kind, int32_t coupling, int64_t row_begin,
                                 int64_t row_end, int32_t flags, const int64_t* gather_offset, int64_t gather_len,
                                 nd_b200_engine** out) {
  struct VR { int kind, dim, pdim, outdim; };
  static const VR vreg[] = {{ND_B200_V_DIFFUSION, 1, 0, 1}, {ND_B200_V_KURAMOTO_FIRST, 1, 1, 1}, {ND_B200_V_KURAMOTO_SECOND, 2, 3, 1},
                            {ND_B200_V_KURAMOTO_SECOND_BENCH, 2, 1, 1}, {ND_B200_V_SWING_DQ, 2, 4, 2}};
  struct ER { int kind, pdim, odst; };
  static const ER ereg[] = {{ND_B200_E_DIFFUSION, 1, 1}, {ND_B200_E_DIFFUSION_NOP, 0, 1}, {ND_B200_E_KURAMOTO, 1, 1}, {ND_B200_E_LINE_DQ, 3, 2}};
  const VR* v = nullptr;
  const ER* g = nullptr;
  for (const VR& r : vreg) if (r.kind == vertex_kind) v = &r;
  for (const ER& r : ereg) if (r.kind == edge_kind) g = &r;
  if (!out || !v || (ne > 0 && !g) || nv <= 0 || ne < 0 || (ne > 0 && (!edge_src || !edge_dst)))
    return fail(nullptr, ND_B200_EINVAL, "nd_b200_create_from_edgelist: bad arguments or kinds outside the registry");
  if (coupling != ND_B200_ANTISYMMETRIC && coupling != ND_B200_SYMMETRIC && coupling != ND_B200_DIRECTED)
    return fail(nullptr, ND_B200_EINVAL, "nd_b200_create_from_edgelist: unsupported wrapper %d", coupling);
  nd_b200_vbatch vb;
  memset(&vb, 0, sizeof vb);
  vb.kind = v->kind; vb.dim = v->dim; vb.pdim = v->pdim; vb.outdim = v->outdim; vb.count = nv; vb.indices = nullptr;
  vb.state_first = 1; vb.p_first = 1; vb.out_first = 1; vb.aggr_first = 1;
  nd_b200_ebatch eb;
  memset(&eb, 0, sizeof eb);
  const int osrc = (ne > 0 && coupling != ND_B200_DIRECTED) ? g->odst : 0;
  if (ne > 0) {
    eb.kind = g->kind; eb.coupling = coupling; eb.dim = 0; eb.pdim = g->pdim; eb.outdim_src = osrc; eb.outdim_dst = g->odst;
    eb.count = ne; eb.indices = nullptr;
    eb.state_first = nv * v->dim + 1; eb.p_first = nv * v->pdim + 1; eb.out_first = nv * v->outdim + 1; eb.gbuf_first = 1;
  }
  nd_b200_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = ND_B200_ABI_VERSION; d.device = device; d.nv = nv; d.ne = ne; d.edge_src = edge_src; d.edge_dst = edge_dst;
  d.vdepth = v->outdim; d.edepth = ne > 0 ? g->odst : v->outdim;
  d.n_vbatches = 1; d.n_ebatches = ne > 0 ? 1 : 0; d.vbatches = &vb; d.ebatches = &eb;
  d.lastidx_dynamic = nv * v->dim; d.lastidx_p = nv * v->pdim + (ne > 0 ? ne * g->pdim : 0);
  d.lastidx_out = nv * v->outdim + (ne > 0 ? ne * (osrc + g->odst) : 0); d.lastidx_aggr = nv * d.edepth;
  d.row_begin = row_begin; d.row_end = row_end; d.flags = flags;
  d.gather_offset = gather_offset; d.gather_len = gather_len;
  return nd_b200_create(&d, out);
}

void nd_b200_destroy(nd_b200_engine* e) {
  if (!e) return;
  for (nd_b200_engine* sub : e->blocks) nd_b200_destroy(sub);
  e->blocks.clear();
  if (e->host_only) { delete e; return; }
  cudaSetDevice(e->device);
  cudaFree(e->d_blockacc);
  destroy_graph(e);
  cudaFree(e->d_rowptr); cudaFree(e->d_nbr); cudaFree(e->d_epar); cudaFree(e->d_blk_row); cudaFree(e->d_ebid);
  cudaFree(e->d_vb); cudaFree(e->d_eb); cudaFree(e->d_vout[0]); cudaFree(e->d_vout[1]);
  cudaFree(e->d_tiles); cudaFree(e->d_oidx); cudaFree(e->d_es); cudaFree(e->d_et); cudaFree(e->d_eepar); cudaFree(e->d_eooff);
  cudaFree(e->d_eebid); cudaFree(e->d_oedge);
  if (e->c_lib) cudaLibraryUnload(e->c_lib);
  cudaFree(e->d_jslices); cudaFree(e->d_jlanes); cudaFree(e->d_jnbr); cudaFree(e->d_jent); cudaFree(e->d_jebid); cudaFree(e->d_jlong);
  cudaFree(e->d_jwarp);
  for (auto& ob : e->ode) { cudaFree(ob.d_es); cudaFree(ob.d_et); cudaFree(ob.d_ext); }
  for (int* q : e->d_vext) cudaFree(q);
  for (int* q : e->d_ffin) cudaFree(q);
  cudaFree(e->d_ppack);
  for (int* q : e->d_esrc_off) cudaFree(q);
  for (int* q : e->d_edst_off) cudaFree(q);
  cudaFree(e->d_aggrow); cudaFree(e->d_aggidx);
  cudaFree(e->d_tmpA); cudaFree(e->d_tmpB); cudaFree(e->d_ksum); cudaFree(e->d_coop_bar);
  cudaFree(e->d_hu); cudaFree(e->d_hp); cudaFree(e->d_hdu);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->s_copy) cudaStreamDestroy(e->s_copy);
  if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
  for (cudaEvent_t ev : e->ev_pipe) cudaEventDestroy(ev);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  for (cudaEvent_t ev : e->ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->ev_pre) cudaEventDestroy(ev);
  delete e;
}

const char* nd_b200_last_error(const nd_b200_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int nd_b200_rhs(nd_b200_engine* e, double* du, const double* u, const double* p, double t, void* stream) {
  if (int rc = check_call(e, du, u, p)) return rc;
  CUDA_TRY(e, cudaSetDevice(e->device));
  return rhs_impl(e, du, u, p, t, (cudaStream_t)stream, MODE_DU, nullptr);
}

// state ranges [start, len) written by rows [r0, r1] (inclusive), one per vertex batch the rows intersect
static void rows_to_state_ranges(const nd_b200_engine* e, long long r0, long long r1, std::vector<std::pair<long long, long long>>& out) {
  out.clear();
  for (const HostVB& h : e->hvb) {
    const long long lo = std::max<long long>(r0, h.row0), hi = std::min<long long>(r1 + 1, h.row0 + h.count);
    if (lo < hi && h.dim > 0) out.push_back({h.state0 + (lo - h.row0) * h.dim, (hi - lo) * h.dim});
  }
}

int nd_b200_rhs_host(nd_b200_engine* e, double* du_host, const double* u_host, const double* p_host, double t) {
  if (int rc = check_call(e, du_host, u_host, p_host)) return rc;
  if (e->halo_base != INT_MAX) return fail(e, ND_B200_EUNSUPPORTED, "nd_b200_rhs_host on a halo engine");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (!e->own_stream) CUDA_TRY(e, cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  const size_t nb = sizeof(double) * (size_t)e->lastidx_dynamic, pb = sizeof(double) * (size_t)e->lastidx_p;
  if (!e->d_hu) {
    CUDA_TRY(e, cudaMalloc((void**)&e->d_hu, std::max<size_t>(nb, 8)));
    CUDA_TRY(e, cudaMalloc((void**)&e->d_hdu, std::max<size_t>(nb, 8)));
    CUDA_TRY(e, cudaMalloc((void**)&e->d_hp, std::max<size_t>(pb, 8)));
  }
  cudaStream_t st = e->own_stream;
  // ---- pipelined form: the H2D copy of p is cut into pieces, a group of thread blocks starts as soon as the last
  // parameter it reads has landed (blk_pmax; with edges(g) sorted by source the rows of group c read pieces <= c), and
  // the D2H copy of a group's du rows overlaps the next group's kernel and the remaining H2D traffic.
  // piece k covers the fraction 1/2, 1/4, ... of p and of the thread blocks (the last two pieces are equal): what stays
  // exposed after the H2D stream is the LAST group's kernel + D2H, so late pieces are small while early ones keep the
  // per-copy overhead low.  Measured on B200 (cfg2, 48 MB over PCIe per call, profiles/r01_tuning.md section 7): 0.96 ms
  // unpipelined; 0.870 ms with K=3 (1/2, 1/4, 1/4), 0.896 with K=5; the H2D stream alone is ~0.76 ms.
  int K = 3;
  if (const char* s = getenv("ND_B200_HOST_CHUNKS")) K = std::max(1, std::min(10, atoi(s)));
  std::vector<double> cum((size_t)K + 1, 0.0);
  for (int k = 0; k < K; ++k) cum[(size_t)k + 1] = (k == K - 1) ? 1.0 : 1.0 - std::ldexp(1.0, -(k + 1));
  const int nblk = (int)e->blk_pmax.size();
  const double last = std::ldexp(1.0, -(K - 1));   // fraction of the smallest piece
  const bool pipelined = K > 1 && e->gather_from_u && !e->split && !e->jstream && e->ode.empty() && (double)pb * last >= 16384.0 && (double)nblk * last >= 8.0 &&
                         nblk == e->nblocks && (e->row_end - e->row_begin == e->nrows_total);
  if (!pipelined) {
    CUDA_TRY(e, cudaMemcpyAsync(e->d_hu, u_host, nb, cudaMemcpyHostToDevice, st));
    if (pb) CUDA_TRY(e, cudaMemcpyAsync(e->d_hp, p_host, pb, cudaMemcpyHostToDevice, st));
    if (int rc = rhs_impl(e, e->d_hdu, e->d_hu, pb ? e->d_hp : nullptr, t, st, MODE_DU, nullptr)) return rc;
    CUDA_TRY(e, cudaMemcpyAsync(du_host, e->d_hdu, nb, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));
    return ND_B200_OK;
  }
  if (!e->s_copy) CUDA_TRY(e, cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
  if (!e->s_d2h) CUDA_TRY(e, cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
  while ((int)e->ev_pipe.size() < 2 * K + 1) {
    cudaEvent_t ev;
    CUDA_TRY(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->ev_pipe.push_back(ev);
  }
  cudaEvent_t ev_u = e->ev_pipe[0];
  cudaEvent_t* ev_p = &e->ev_pipe[1];
  cudaEvent_t* ev_k = &e->ev_pipe[1 + K];
  // H2D: u, then p piece by piece
  CUDA_TRY(e, cudaMemcpyAsync(e->d_hu, u_host, nb, cudaMemcpyHostToDevice, e->s_copy));
  CUDA_TRY(e, cudaEventRecord(ev_u, e->s_copy));
  std::vector<long long> pend((size_t)K);
  for (int k = 0; k < K; ++k) {
    const long long a = (long long)(e->lastidx_p * cum[(size_t)k]), z = k == K - 1 ? e->lastidx_p : (long long)(e->lastidx_p * cum[(size_t)k + 1]);
    pend[(size_t)k] = z;
    CUDA_TRY(e, cudaMemcpyAsync(e->d_hp + a, p_host + a, sizeof(double) * (size_t)(z - a), cudaMemcpyHostToDevice, e->s_copy));
    CUDA_TRY(e, cudaEventRecord(ev_p[k], e->s_copy));
  }
  // compute: group c = the same fraction of the thread blocks
  KParams P;
  fill_params(e, P);
  P.u = e->d_hu; P.gsrc = e->d_hu; P.p = e->d_hp; P.du = e->d_hdu; P.mode = MODE_DU; P.t = t;
  CUDA_TRY(e, cudaStreamWaitEvent(st, ev_u, 0));
  std::vector<std::pair<long long, long long>> ranges;
  int waited = -1;
  for (int c = 0; c < K; ++c) {
    const int b0 = (int)(nblk * cum[(size_t)c]), b1 = c == K - 1 ? nblk : (int)(nblk * cum[(size_t)c + 1]);
    int pm = 0, rmin = INT_MAX, rmax = -1;
    for (int q = b0; q < b1; ++q) { pm = std::max(pm, e->blk_pmax[(size_t)q]); rmin = std::min(rmin, e->blk_rmin[(size_t)q]); rmax = std::max(rmax, e->blk_rmax[(size_t)q]); }
    int need = -1;
    if (pm > 0) { need = 0; while (need < K - 1 && pend[(size_t)need] < pm) ++need; }
    if (need > waited) { CUDA_TRY(e, cudaStreamWaitEvent(st, ev_p[need], 0)); waited = need; }   // s_copy is in order: piece `need` implies all earlier ones
    P.blk_off = b0;
    e->launch_nblk = b1 - b0;
    cudaError_t ce = launch_fused(e, P, st);
    e->launch_nblk = -1;
    CUDA_TRY(e, ce);
    CUDA_TRY(e, cudaEventRecord(ev_k[c], st));
    if (e->blk_rows_monotone && rmax >= rmin) {
      CUDA_TRY(e, cudaStreamWaitEvent(e->s_d2h, ev_k[c], 0));
      rows_to_state_ranges(e, rmin, rmax, ranges);
      for (const auto& rg : ranges)
        CUDA_TRY(e, cudaMemcpyAsync(du_host + rg.first, e->d_hdu + rg.first, sizeof(double) * (size_t)rg.second, cudaMemcpyDeviceToHost, e->s_d2h));
    }
  }
  if (!e->blk_rows_monotone) {
    CUDA_TRY(e, cudaStreamWaitEvent(e->s_d2h, ev_k[K - 1], 0));
    CUDA_TRY(e, cudaMemcpyAsync(du_host, e->d_hdu, nb, cudaMemcpyDeviceToHost, e->s_d2h));
  }
  CUDA_TRY(e, cudaStreamSynchronize(e->s_d2h));
  CUDA_TRY(e, cudaStreamSynchronize(st));
  return ND_B200_OK;
}

int nd_b200_pack_params(nd_b200_engine* e, const double* p, void* stream) {
  if (!e) return ND_B200_EINVAL;
  if (e->host_only) return fail(e, ND_B200_EUNSUPPORTED, "host-only engine");
  for (nd_b200_engine* sub : e->blocks)       // column blocks keep their own packed copies
    if (int rc = nd_b200_pack_params(sub, p, stream)) { e->err = sub->err; return rc; }
  if (!p) { e->pack_on = false; return ND_B200_OK; }
  if (e->pack_pe <= 0 || e->split) return fail(e, ND_B200_EUNSUPPORTED, "packed edge parameters need ONE registry edge batch with parameters and a fused kernel");
  if (e->jstream ? false : (e->jag ? (e->jag_u != 2 && (e->vdepth != 1 || e->halo_base != INT_MAX)) : (e->block != 128 || e->ept != 4))) return fail(e, ND_B200_EUNSUPPORTED, "packed edge parameters are compiled for the default launch shape only");
  if (e->nentries == 0) return ND_B200_OK;
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long plen = e->jag ? e->jag_len : e->nentries;     // the streamed layout pads its slices
  if (!e->d_ppack) CUDA_TRY(e, cudaMalloc((void**)&e->d_ppack, sizeof(double) * (size_t)plen * (size_t)e->pack_pe));
  const int T = 256;
  const int nb = (int)((plen + T - 1) / T);
  if (e->jag) {
    if (!e->d_jnbr) {   // the PK kernels read a plain neighbour stream: extract it once from the {nbr, epar} pairs
      CUDA_TRY(e, cudaMalloc((void**)&e->d_jnbr, sizeof(int) * (size_t)plen));
      ND_LAUNCH(nb, T, st, ((const int*)e->d_jent, plen, e->d_jnbr), extract_nbr_kernel);
      CUDA_TRY(e, cudaGetLastError());
      e->launches++;
    }
    ND_LAUNCH(nb, T, st, ((const int*)e->d_jent, 2, 1, plen, e->pack_pe, p, e->d_ppack), pack_params_kernel);
  } else {
    ND_LAUNCH(nb, T, st, (e->d_epar, 1, 0, e->nentries, e->pack_pe, p, e->d_ppack), pack_params_kernel);
  }
  CUDA_TRY(e, cudaGetLastError());
  e->launches++;
  e->pack_on = true;
  return ND_B200_OK;
}

int nd_b200_get_buffers(nd_b200_engine* e, double* o, double* aggbuf, const double* u, const double* p, double t,
                        void* stream) {
  if (!e) return ND_B200_EINVAL;
  if (!u || (e->lastidx_p > 0 && !p)) return fail(e, ND_B200_EINVAL, "u or p is NULL");
  if (e->h_esrc_off.size() != e->heb.size()) return fail(e, ND_B200_EUNSUPPORTED, "engine was created with ND_B200_FLAG_NO_EXPORT");
  if (e->halo_base != INT_MAX || e->host_only) return fail(e, ND_B200_EUNSUPPORTED, "nd_b200_get_buffers on a halo / host-only engine");
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const double* gsrc = u;
  if (!e->gather_from_u) {
    CUDA_TRY(e, launch_vout(e, u, p, e->d_vout[0], st, t));
    gsrc = e->d_vout[0];
  }
  if (o) {
    // vertex outputs occupy o[0 .. nv*vdepth) in row order (register_vertices!)
    CUDA_TRY(e, launch_vout(e, u, p, o, st, t));
    if (e->d_esrc_off.empty()) {
      e->d_esrc_off.assign(e->heb.size(), nullptr); e->d_edst_off.assign(e->heb.size(), nullptr);
      for (size_t b = 0; b < e->heb.size(); ++b)
        if (upload(e, &e->d_esrc_off[b], e->h_esrc_off[b]) || upload(e, &e->d_edst_off[b], e->h_edst_off[b])) return ND_B200_ECUDA;
    }
    for (size_t b = 0; b < e->heb.size(); ++b) {
      const HostEB& h = e->heb[b];
      const int T = 256;
      const int nb = (int)((h.count + T - 1) / T);
      e->launches++;
      if (h.dim > 0) {   // edges with states: outputs are StateMask reads
        ND_LAUNCH(nb, T, st, (h.coupling, h.dim, h.osrc, h.odst, h.mask_src, h.mask_dst, h.count, h.state0, h.out0, u, o), edge_mask_out_kernel);
      } else if (e->custom) {
        int kind = h.kind, coupling = h.coupling, pdim = h.pdim, osrc = h.osrc;
        long long count = h.count, p0 = h.p0, out0 = h.out0;
        const int *es = e->d_esrc_off[b], *et = e->d_edst_off[b];
        void* args[] = {&kind, &coupling, &pdim, &osrc, &count, &es, &et, &p0, &out0, &gsrc, &p, &o, &t};
        CUDA_TRY(e, cudaLaunchKernel((const void*)e->c_eout, dim3((unsigned)nb), dim3(T), args, 0, st));
      } else if (e->vdepth == 2)
        ND_LAUNCH(nb, T, st, (h.kind, h.coupling, h.pdim, h.osrc, h.count, e->d_esrc_off[b], e->d_edst_off[b], h.p0, h.out0, gsrc, p, o, t), edge_out_kernel<2, 2>);
      else
        ND_LAUNCH(nb, T, st, (h.kind, h.coupling, h.pdim, h.osrc, h.count, e->d_esrc_off[b], e->d_edst_off[b], h.p0, h.out0, gsrc, p, o, t), edge_out_kernel<1, 1>);
      CUDA_TRY(e, cudaGetLastError());
    }
  }
  if (aggbuf) {
    KParams P;
    fill_params(e, P);
    P.u = u; P.p = p; P.gsrc = gsrc; P.mode = MODE_AGG; P.aggbuf = aggbuf; P.t = t;
    CUDA_TRY(e, launch_fused(e, P, st));
  }
  return ND_B200_OK;
}

int nd_b200_aggregate(nd_b200_engine* e, double* aggbuf, const double* o, void* stream) {
  if (!e) return ND_B200_EINVAL;
  if (!aggbuf || !o) return fail(e, ND_B200_EINVAL, "aggbuf or o is NULL");
  if (e->host_only || e->halo_base != INT_MAX || e->row_end - e->row_begin != e->nrows_total)
    return fail(e, ND_B200_EUNSUPPORTED, "nd_b200_aggregate on a partitioned / host-only engine");
  if (e->h_rowptr.size() != (size_t)e->nrows_total + 1) return fail(e, ND_B200_EUNSUPPORTED, "engine was created with ND_B200_FLAG_NO_EXPORT");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (!e->d_aggrow) {
    std::vector<int> rp(e->h_rowptr.begin(), e->h_rowptr.end());
    std::vector<long long> oi(e->h_aggidx);
    if (oi.empty()) oi.push_back(-1);
    if (upload(e, &e->d_aggrow, rp) || upload(e, &e->d_aggidx, oi)) return ND_B200_ECUDA;
  }
  const long long n = (long long)e->nrows_total * e->edepth;
  if (n == 0) return ND_B200_OK;
  const int T = 256;
  e->launches++;
  ND_LAUNCH((unsigned)((n + T - 1) / T), T, (cudaStream_t)stream, (e->d_aggrow, e->d_aggidx, e->edepth, (long long)e->nrows_total, o, aggbuf), aggregate_kernel);
  CUDA_TRY(e, cudaGetLastError());
  return ND_B200_OK;
}

int nd_b200_rk4(nd_b200_engine* e, double* u, const double* p, double t0, double dt, int64_t nsteps, void* stream) {
  if (int rc = check_call(e, u, u, p)) return rc;
  if (e->row_end - e->row_begin != e->nrows_total || e->halo_base != INT_MAX) return fail(e, ND_B200_EUNSUPPORTED, "nd_b200_rk4 on a partitioned engine: drive the stages from the host (halo exchange between stages)");
  if (nsteps <= 0) return ND_B200_OK;
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = sizeof(double) * (size_t)e->lastidx_dynamic;
  if (!e->d_tmpA) {
    CUDA_TRY(e, cudaMalloc((void**)&e->d_tmpA, nb));
    CUDA_TRY(e, cudaMalloc((void**)&e->d_tmpB, nb));
    CUDA_TRY(e, cudaMalloc((void**)&e->d_ksum, nb));
  }
  // p is constant for the whole call: the edge parameters can be packed once into entry order so that every stage reads
  // them coalesced (contract-free).  ND_B200_RK4_PACK=1 / 0 forces / forbids it.
  struct Unpack { nd_b200_engine* e; bool on; ~Unpack() { if (on) e->pack_on = false; } } unpack{e, false};
  if (!e->pack_on && e->pack_pe > 0) {
    // default: only where the gain does not depend on a measurement -- a parameter vector far beyond the 126 MB L2, whose
    // per-entry reads are isolated DRAM sectors, and enough stages to amortise the one packing pass
    const char* s = getenv("ND_B200_RK4_PACK");
    // measured (profiles/r02d): 207 vs 266 us per step on config 2, 275 vs 329 us on Kuramoto / Erdos-Renyi
    const bool want = s ? atoi(s) > 0 : nsteps >= 4;
    if (want && nd_b200_pack_params(e, p, stream) == ND_B200_OK) unpack.on = true;
  }
  if (!e->gather_from_u && !e->custom) CUDA_TRY(e, launch_vout(e, u, p, e->d_vout[0], st, t0));
#ifndef ND_CUSIM
  // Small graphs (one stage = about one wave of thread blocks): all steps and stages in ONE cooperative launch with grid-wide
  // barriers between the stages (rk4_jag_coop_kernel).  ND_B200_RK4_COOP=1 / 0 forces / forbids it.
  if (e->jag && !e->jaga && !e->jagb && !e->jstream && !e->custom && e->ode.empty() && !e->split && e->jag_u <= 2) {
    const char* s = getenv("ND_B200_RK4_COOP");
    // Opt-in: measured SLOWER than the graph-replayed stages (profiles/r02g_sweep_rk4_coop.jsonl: config 4 57.4 vs 47.2 us per
    // step, config 1 41.5 vs 35.2) -- four grid barriers per step cost more than the launch gaps they remove.
    bool want = s ? atoi(s) > 0 : false;
    if (want) {
      KParams P;
      fill_params(e, P);
      P.p = p; P.mode = MODE_RK; P.u0 = u; P.ksum = e->d_ksum; P.h6 = dt / 6.0;
      CoopArgs R;
      R.u = u; R.tmpA = e->d_tmpA; R.tmpB = e->d_tmpB; R.vout0 = e->d_vout[0]; R.vout1 = e->d_vout[1];
      R.t0 = t0; R.dt = dt; R.nsteps = nsteps;
      int cap = 0;
      CUDA_TRY(e, launch_rk4_coop(e, P, R, st, true, &cap));
      // default: only while every warp of the resident grid walks at most two slices per stage
      if (want && cap > 0) {
        if (!e->d_coop_bar) CUDA_TRY(e, cudaMalloc((void**)&e->d_coop_bar, sizeof(unsigned long long)));
        CUDA_TRY(e, cudaMemsetAsync(e->d_coop_bar, 0, sizeof(unsigned long long), st));
        R.barrier = e->d_coop_bar;
        CUDA_TRY(e, launch_rk4_coop(e, P, R, st));
        e->launches += 1;
        return ND_B200_OK;
      }
    }
  }
#endif
  // The registry models are autonomous, so one captured step can be replayed for every t.  User-supplied kinds may read
  // t: their steps are enqueued one by one with the right stage times.
  if (e->custom) {
    for (int64_t k = 0; k < nsteps; ++k)
      if (int rc = rk4_step_enqueue(e, u, p, t0 + (double)k * dt, dt, st)) return rc;
    return ND_B200_OK;
  }
  const int UNROLL = 8;
  const int per_graph = (int)std::min<int64_t>(UNROLL, nsteps);
  if (!e->graph_exec || e->graph_u != u || e->graph_p != p || e->graph_dt != dt || e->graph_steps != per_graph || e->graph_packed != e->pack_on) {
    destroy_graph(e);
    if (!e->cap_stream) CUDA_TRY(e, cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    CUDA_TRY(e, cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    const long long launches_before = e->launches;
    int rc = ND_B200_OK;
    // the captured stage kernels are chained by programmatic dependencies (rk4_pdl == 2 while capturing): measured
    // (profiles/r02w_sweep_rk4_pdl.jsonl) config 4 48.1 -> 45.6 us per step, config 1 18.1 -> 16.0, config 2 unchanged; the tile
    // kernel on config-4-size graphs loses (47.7 -> 50.8), so it is on for jagged layouts and for small tile grids only.
    // ND_B200_RK4_PDL=1 / 0 forces / forbids it.
    e->rk4_pdl = (e->jag || e->nblocks <= 4 * e->num_sms) ? 1 : 0;
    if (const char* s = getenv("ND_B200_RK4_PDL")) e->rk4_pdl = atoi(s) > 0 ? 1 : 0;
    if (e->rk4_pdl) e->rk4_pdl = 2;
    for (int s = 0; s < per_graph && rc == ND_B200_OK; ++s) rc = rk4_step_enqueue(e, u, p, t0 + s * dt, dt, e->cap_stream);
    if (e->rk4_pdl) e->rk4_pdl = 1;
    cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &e->graph);
    e->launches = launches_before;   // captured, not launched
    if (rc != ND_B200_OK) return rc;
    CUDA_TRY(e, ce);
    CUDA_TRY(e, cudaGraphInstantiate(&e->graph_exec, e->graph, 0));
    e->graph_u = u; e->graph_p = p; e->graph_dt = dt; e->graph_steps = per_graph; e->graph_packed = e->pack_on;
  }
  const long long per_step = 4 * (1 + (long long)e->ode.size());
  int64_t done = 0;
  while (nsteps - done >= per_graph) {
    CUDA_TRY(e, cudaGraphLaunch(e->graph_exec, st));
    e->launches += per_step * per_graph;
    done += per_graph;
  }
  for (; done < nsteps; ++done)
    if (int rc = rk4_step_enqueue(e, u, p, t0 + (double)done * dt, dt, st)) return rc;
  return ND_B200_OK;
}

int nd_b200_export_sizes(const nd_b200_engine* e, int64_t sizes[8]) {
  if (!e || !sizes) return ND_B200_EINVAL;
  sizes[0] = e->row_end - e->row_begin; sizes[1] = e->nentries; sizes[2] = e->nblocks; sizes[3] = e->n_long;
  sizes[4] = e->gather_from_u; sizes[5] = (e->gather_from_u ? 0 : 1) + (e->split ? 2 : 1) + (long long)e->ode.size(); sizes[6] = e->row_begin; sizes[7] = e->row_end;
  return ND_B200_OK;
}

int nd_b200_export_tables(const nd_b200_engine* e, int64_t* rowptr, int64_t* nbr_vertex, int64_t* edge_id, int32_t* side) {
  if (!e) return ND_B200_EINVAL;
  if (e->h_rowptr.empty()) return fail(const_cast<nd_b200_engine*>(e), ND_B200_EUNSUPPORTED, "engine was created with ND_B200_FLAG_NO_EXPORT");
  for (size_t r = 0; r < e->h_rowptr.size(); ++r) rowptr[r] = e->h_rowptr[r];
  for (size_t j = 0; j < (size_t)e->nentries; ++j) {
    nbr_vertex[j] = e->h_nbr_vid[j]; edge_id[j] = e->h_eid[j]; side[j] = e->h_side[j];
  }
  return ND_B200_OK;
}

int nd_b200_export_jag_sizes(const nd_b200_engine* e, int64_t sizes[6]) {
  if (!e || !sizes) return ND_B200_EINVAL;
  sizes[0] = e->jag ? (int64_t)e->nslices : -1; sizes[1] = e->n_jlong; sizes[2] = e->jsplit; sizes[3] = (int64_t)e->h_jorder.size();
  sizes[4] = e->wait_from; sizes[5] = e->halo_base == INT_MAX ? -1 : e->gather_len - e->lastidx_dynamic;
  return ND_B200_OK;
}

int nd_b200_export_jag(const nd_b200_engine* e, int32_t* slices, uint16_t* lanes, int32_t* longs, int32_t* order) {
  if (!e) return ND_B200_EINVAL;
  if (!e->jag || !e->host_only) return fail(const_cast<nd_b200_engine*>(e), ND_B200_EUNSUPPORTED, "jagged tables are kept only by ND_B200_FLAG_HOST_ONLY engines in jag mode");
  for (size_t k = 0; k < e->h_jslices.size(); ++k) { slices[4 * k] = e->h_jslices[k].x; slices[4 * k + 1] = e->h_jslices[k].y; slices[4 * k + 2] = e->h_jslices[k].z; slices[4 * k + 3] = e->h_jslices[k].w; }
  for (size_t k = 0; k < e->h_jlanes.size(); ++k) lanes[k] = e->h_jlanes[k];
  for (size_t k = 0; k < e->h_jlong.size(); ++k) { longs[4 * k] = e->h_jlong[k].x; longs[4 * k + 1] = e->h_jlong[k].y; longs[4 * k + 2] = e->h_jlong[k].z; longs[4 * k + 3] = e->h_jlong[k].w; }
  for (size_t k = 0; k < e->h_jorder.size(); ++k) order[k] = e->h_jorder[k];
  return ND_B200_OK;
}

const char* nd_b200_custom_source(const nd_b200_engine* e) { return (e && e->custom) ? e->custom_src.c_str() : nullptr; }

const char* nd_b200_kernel_name(const nd_b200_engine* e) {
  if (!e) return "";
  if (e->jstream) return "rhs_js_kernel";
  if (e->jagb) return "rhs_jagb_kernel";
  if (e->jaga) return "rhs_jaga_kernel";
  if (e->jag) return "rhs_jag_kernel";
  if (e->split) return "edge_pass_kernel+row_pass_kernel";
  return "rhs_fused_kernel";
}

int64_t nd_b200_launch_count(const nd_b200_engine* e) { return e ? e->launches : 0; }

int nd_b200_set_timing(nd_b200_engine* e, int enabled) {
  if (!e) return ND_B200_EINVAL;
  e->timing = enabled != 0;
  e->ev_used = 0; e->ev_pre_used = 0;
  return ND_B200_OK;
}

int nd_b200_timings(nd_b200_engine* e, double* fused_ms_avg, double* prepass_ms_avg, int64_t* ncalls) {
  if (!e) return ND_B200_EINVAL;
  CUDA_TRY(e, cudaDeviceSynchronize());
  double sum = 0.0, sum_pre = 0.0;
  for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
    float ms = 0.f;
    CUDA_TRY(e, cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
    sum += ms;
  }
  for (size_t i = 0; i + 1 < e->ev_pre_used; i += 2) {
    float ms = 0.f;
    CUDA_TRY(e, cudaEventElapsedTime(&ms, e->ev_pre[i], e->ev_pre[i + 1]));
    sum_pre += ms;
  }
  const int64_t n = (int64_t)(e->ev_used / 2);
  if (fused_ms_avg) *fused_ms_avg = n ? sum / (double)n : 0.0;
  if (prepass_ms_avg) *prepass_ms_avg = e->ev_pre_used ? sum_pre / (double)(e->ev_pre_used / 2) : 0.0;
  if (ncalls) *ncalls = n;
  e->ev_used = 0; e->ev_pre_used = 0;
  return ND_B200_OK;
}

void* nd_b200_host_alloc(int64_t bytes) {
  void* q = nullptr;
  if (cudaHostAlloc(&q, (size_t)std::max<int64_t>(bytes, 8), cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return q;
}
void nd_b200_host_free(void* q) { if (q) cudaFreeHost(q); }

}  // extern "C"

#include "nd_b200_comm.inc"
