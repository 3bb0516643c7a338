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
#ifdef ND_CUSIM
#define ND_LAUNCH(grid, block, stream, args, ...) \
  cusim::launch(dim3((unsigned)(grid)), dim3((unsigned)(block)), (stream), [=]() { __VA_ARGS__ args; })
#else
#define ND_LAUNCH(grid, block, stream, args, ...) __VA_ARGS__<<<(grid), (block), 0, (stream)>>> args
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

// ---- split mode launches --------------------------------------------------------------------------------
template <int VD, int ED, int EK, int PE>
cudaError_t launch_edge_pass_t(const nd_b200_engine* e, const EParams& Q, cudaStream_t st) {
  constexpr int BLOCK = 256, EPT = 2;
  const long long per = BLOCK * EPT;
  const int grid = (int)((e->ne_all + per - 1) / per);
  if (grid == 0) return cudaSuccess;
  ND_LAUNCH(grid, BLOCK, st, (Q), edge_pass_kernel<VD, ED, EK, PE, BLOCK, EPT>);
  return cudaGetLastError();
}
cudaError_t launch_edge_pass(nd_b200_engine* e, const double* gsrc, const double* p, cudaStream_t st) {
  EParams Q;
  memset(&Q, 0, sizeof Q);
  Q.esrc = e->d_es; Q.edst = e->d_et; Q.epar = e->d_eepar; Q.eooff = e->d_eooff; Q.ebid = e->d_eebid; Q.eb = e->d_eb;
  Q.ne = e->ne_all; Q.gsrc = gsrc; Q.p = p; Q.oedge = e->d_oedge;
  if (!e->heb.empty()) { Q.p0 = e->heb[0].p0; Q.coupling0 = e->heb[0].coupling; }
  e->launches += (e->ne_all > 0);
  if (e->vdepth == 2) return launch_edge_pass_t<2, 2, ND_B200_E_LINE_DQ, 3>(e, Q, st);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_edge_pass_t<1, 1, ND_B200_E_DIFFUSION, 1>(e, Q, st);
    case ND_B200_E_DIFFUSION_NOP: return launch_edge_pass_t<1, 1, ND_B200_E_DIFFUSION_NOP, 0>(e, Q, st);
    case ND_B200_E_KURAMOTO: return launch_edge_pass_t<1, 1, ND_B200_E_KURAMOTO, 1>(e, Q, st);
    default: return launch_edge_pass_t<1, 1, EK_GENERIC, 1>(e, Q, st);
  }
}
template <int VD, int ED>
cudaError_t launch_row_pass_t(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  if (e->block == 256 && e->ept == 8) ND_LAUNCH(e->nblocks, 256, st, (P), row_pass_kernel<VD, ED, 256, 8>);
  else if (e->block == 256 && e->ept == 4) ND_LAUNCH(e->nblocks, 256, st, (P), row_pass_kernel<VD, ED, 256, 4>);
  else if (e->block == 128 && e->ept == 8) ND_LAUNCH(e->nblocks, 128, st, (P), row_pass_kernel<VD, ED, 128, 8>);
  else if (e->block == 128 && e->ept == 4) ND_LAUNCH(e->nblocks, 128, st, (P), row_pass_kernel<VD, ED, 128, 4>);
  else return cudaErrorInvalidConfiguration;
  return cudaGetLastError();
}

template <int VD, int ED, int EK, int PE>
cudaError_t launch_shape(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  const int grid = (e->launch_nblk >= 0 ? e->launch_nblk : e->nblocks) + P.n_pub + P.fence;
  if (grid == 0) return cudaSuccess;
  if constexpr (EK != EK_GENERIC) {
    if (e->compact) {   // compact entry words: default launch shape only (plan_tiles)
      if constexpr (PE > 0) {
        if (e->pack_on) {
          if (e->halo_base != INT_MAX) ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, true, true, true>);
          else ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, false, true, true>);
          return cudaGetLastError();
        }
      }
      if (e->halo_base != INT_MAX) ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, true, false, true>);
      else ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, false, false, true>);
      return cudaGetLastError();
    }
  }
  if constexpr (PE > 0 && EK != EK_GENERIC) {
    if (e->pack_on) {   // packed edge parameters: default launch shape only (checked by nd_b200_pack_params)
      if (e->halo_base != INT_MAX) ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, true, true>);
      else ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, false, true>);
      return cudaGetLastError();
    }
  }
  if (e->halo_base != INT_MAX) {   // multi-GPU variant, default launch shape only
    if (e->block != 128 || e->ept != 4) return cudaErrorInvalidConfiguration;
    ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, true>);
  }
  else if (e->block == 256 && e->ept == 8) ND_LAUNCH(grid, 256, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 256, 8, false>);
  else if (e->block == 256 && e->ept == 4) ND_LAUNCH(grid, 256, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 256, 4, false>);
  else if (e->block == 128 && e->ept == 8) ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 8, false>);
  else if (e->block == 128 && e->ept == 4) ND_LAUNCH(grid, 128, st, (P), rhs_fused_kernel<VD, ED, EK, PE, 128, 4, false>);
  else return cudaErrorInvalidConfiguration;
  return cudaGetLastError();
}

// ---- jagged kernel launches ------------------------------------------------------------------------------
template <int VD, int ED, int EK, int PE, int U>
cudaError_t launch_jag_u(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  constexpr int BLOCK = 128;
  const int grid = (e->launch_nblk >= 0 ? e->launch_nblk : e->n_jag_blocks + e->n_jlong) + P.n_pub + P.fence;
  if (grid == 0) return cudaSuccess;
  const int wps = e->jag_wps > 0 ? e->jag_wps : jag_warps_per_sm_default(EK);
  if constexpr (VD == 1 && ED == 1 && EK != EK_GENERIC && U == 2) {
    if (e->jag_win) {   // window mode: block = 128-row window (coalesced own outputs / du / vertex data)
      const bool halo = e->halo_base != INT_MAX;
      if constexpr (PE > 0) {
        if (e->pack_on) {
          if (halo) { if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, true, true, true>); else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, true, true, true>); }
          else if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, false, true, true>);
          else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, false, true, true>);
          return cudaGetLastError();
        }
      }
      if (halo) { if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, true, false, true>); else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, true, false, true>); }
      else if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, false, false, true>);
      else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, false, false, true>);
      return cudaGetLastError();
    }
  }
  if constexpr (VD == 1 && EK != EK_GENERIC && U >= 4) {
    // deep variants (single-GPU, single edge batch): all index / parameter loads of U columns are issued back to back, then
    // all U gathers -- three dependent memory levels per slice instead of 2 per pair of columns
    if (e->halo_base == INT_MAX) {
      if constexpr (PE > 0) {
        if (e->pack_on) {
          if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, false, true>);
          else if (wps >= 32) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, false, true>);
          else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 24, false, true>);
          return cudaGetLastError();
        }
      }
      if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, false>);
      else if (wps >= 32) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, false>);
      else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 24, false>);
      return cudaGetLastError();
    }
  }
  if constexpr (PE > 0 && EK != EK_GENERIC && U == 2) {
    if (e->pack_on) {   // packed edge parameters: U = 2 only (checked by nd_b200_pack_params)
      if (e->halo_base != INT_MAX) {
        if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, true, true>);
        else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, true, true>);
      } else if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, false, true>);
      else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, false, true>);
      return cudaGetLastError();
    }
  }
  if (e->halo_base != INT_MAX) {   // multi-GPU variant: one occupancy setting
    if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, true>);
    else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, true>);
  }
  else if (wps >= 64) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 64, false>);
  else if (wps >= 48) ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 48, false>);
  else ND_LAUNCH(grid, BLOCK, st, (P), rhs_jag_kernel<VD, ED, EK, PE, BLOCK, U, 32, false>);
  return cudaGetLastError();
}
template <int VD, int ED, int EK, int PE>
cudaError_t launch_jag_t(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  if constexpr (VD == 1 && EK != EK_GENERIC) {
    if (e->jag_u >= 8) return launch_jag_u<VD, ED, EK, PE, 8>(e, P, st);
  }
  return e->jag_u >= 4 ? launch_jag_u<VD, ED, EK, PE, 4>(e, P, st) : launch_jag_u<VD, ED, EK, PE, 2>(e, P, st);
}
cudaError_t launch_jag(nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  e->launches += (e->n_jag_blocks + e->n_jlong > 0);
  if (e->vdepth == 2) return launch_jag_t<2, 2, ND_B200_E_LINE_DQ, 3>(e, P, st);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_jag_t<1, 1, ND_B200_E_DIFFUSION, 1>(e, P, st);
    case ND_B200_E_DIFFUSION_NOP: return launch_jag_t<1, 1, ND_B200_E_DIFFUSION_NOP, 0>(e, P, st);
    case ND_B200_E_KURAMOTO: return launch_jag_t<1, 1, ND_B200_E_KURAMOTO, 1>(e, P, st);
    default: return launch_jag_t<1, 1, EK_GENERIC, 1>(e, P, st);
  }
}

// ---- persistent jagged kernel / 64-thread blocks (single GPU, one vertex output, single-batch registry edge kinds) ------
template <int EK, int PE, bool PK>
cudaError_t launch_jag_alt_t(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  const int wps = e->jag_wps > 0 ? e->jag_wps : 32;
  if (e->jag_persist) {
    const int grid = std::min((e->nslices + 3) / 4, e->num_sms * ((wps >= 64 ? 64 : wps >= 48 ? 48 : 32) / 4));
    if (std::max(grid, std::min(e->n_jlong, 1)) == 0) return cudaSuccess;
    const int g = std::max(grid, 1);
    if (e->jag_u >= 4) {
      if (wps >= 48) ND_LAUNCH(g, 128, st, (P), rhs_jag_persist_kernel<1, 1, EK, PE, 128, 4, 48, PK>);
      else ND_LAUNCH(g, 128, st, (P), rhs_jag_persist_kernel<1, 1, EK, PE, 128, 4, 32, PK>);
    } else {
      if (wps >= 64) ND_LAUNCH(g, 128, st, (P), rhs_jag_persist_kernel<1, 1, EK, PE, 128, 2, 64, PK>);
      else if (wps >= 48) ND_LAUNCH(g, 128, st, (P), rhs_jag_persist_kernel<1, 1, EK, PE, 128, 2, 48, PK>);
      else ND_LAUNCH(g, 128, st, (P), rhs_jag_persist_kernel<1, 1, EK, PE, 128, 2, 32, PK>);
    }
    return cudaGetLastError();
  }
  // 64-thread blocks: two slices per block
  KParams Q = P;
  Q.n_jag_blocks = (e->nslices + 1) / 2;
  const int grid = Q.n_jag_blocks + e->n_jlong;
  if (grid == 0) return cudaSuccess;
  if (wps >= 64) ND_LAUNCH(grid, 64, st, (Q), rhs_jag_kernel<1, 1, EK, PE, 64, 2, 64, false, PK>);
  else if (wps >= 48) ND_LAUNCH(grid, 64, st, (Q), rhs_jag_kernel<1, 1, EK, PE, 64, 2, 48, false, PK>);
  else ND_LAUNCH(grid, 64, st, (Q), rhs_jag_kernel<1, 1, EK, PE, 64, 2, 32, false, PK>);
  return cudaGetLastError();
}
template <int EK, int PE>
cudaError_t launch_jag_alt_p(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  if constexpr (PE > 0) {
    if (e->pack_on) return launch_jag_alt_t<EK, PE, true>(e, P, st);
  }
  return launch_jag_alt_t<EK, PE, false>(e, P, st);
}
bool jag_alt_ok(const nd_b200_engine* e, const KParams& P) {
  return (e->jag_persist || e->jag_block == 64) && e->vdepth == 1 && e->edepth == 1 && e->halo_base == INT_MAX && e->launch_nblk < 0 &&
         P.blk_off == 0 && (e->ek == ND_B200_E_DIFFUSION || e->ek == ND_B200_E_DIFFUSION_NOP || e->ek == ND_B200_E_KURAMOTO);
}
cudaError_t launch_jag_alt(nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  e->launches += (e->n_jag_blocks + e->n_jlong > 0);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_jag_alt_p<ND_B200_E_DIFFUSION, 1>(e, P, st);
    case ND_B200_E_DIFFUSION_NOP: return launch_jag_alt_p<ND_B200_E_DIFFUSION_NOP, 0>(e, P, st);
    default: return launch_jag_alt_p<ND_B200_E_KURAMOTO, 1>(e, P, st);
  }
}

// ---- persistent cooperative RK4 (rk4_jag_coop_kernel) ---------------------------------------------------------------------
constexpr int COOP_BLOCK = 1024;   // one block per SM: 148 arrivals per grid barrier
template <int VD, int ED, int EK, int PE, bool PK>
cudaError_t launch_rk4_coop_t(nd_b200_engine* e, const KParams& P, const CoopArgs& R, cudaStream_t st, bool query, int* max_grid) {
#ifndef ND_CUSIM
  auto kern = rk4_jag_coop_kernel<VD, ED, EK, PE, COOP_BLOCK, 2, PK>;
  int per_sm = 0;
  cudaError_t c = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, COOP_BLOCK, 0);
  if (c != cudaSuccess) return c;
  const int cap = per_sm * e->num_sms;
  if (query) { *max_grid = cap; return cudaSuccess; }
  const int want = std::max((e->nslices + COOP_BLOCK / 32 - 1) / (COOP_BLOCK / 32), std::min(e->n_jlong, cap));
  const int grid = std::max(1, std::min(cap, want));
  void* args[] = {const_cast<KParams*>(&P), const_cast<CoopArgs*>(&R)};
  return cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(COOP_BLOCK), args, 0, st);
#else
  (void)e; (void)P; (void)R; (void)st; (void)query; (void)max_grid;
  return cudaErrorInvalidConfiguration;
#endif
}
template <int VD, int ED, int EK, int PE>
cudaError_t launch_rk4_coop_p(nd_b200_engine* e, const KParams& P, const CoopArgs& R, cudaStream_t st, bool query, int* max_grid) {
  if constexpr (PE > 0 && EK != EK_GENERIC) {
    if (e->pack_on) return launch_rk4_coop_t<VD, ED, EK, PE, true>(e, P, R, st, query, max_grid);
  }
  return launch_rk4_coop_t<VD, ED, EK, PE, false>(e, P, R, st, query, max_grid);
}
cudaError_t launch_rk4_coop(nd_b200_engine* e, const KParams& P, const CoopArgs& R, cudaStream_t st, bool query = false, int* max_grid = nullptr) {
  if (e->vdepth == 2) return launch_rk4_coop_p<2, 2, ND_B200_E_LINE_DQ, 3>(e, P, R, st, query, max_grid);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_rk4_coop_p<1, 1, ND_B200_E_DIFFUSION, 1>(e, P, R, st, query, max_grid);
    case ND_B200_E_DIFFUSION_NOP: return launch_rk4_coop_p<1, 1, ND_B200_E_DIFFUSION_NOP, 0>(e, P, R, st, query, max_grid);
    case ND_B200_E_KURAMOTO: return launch_rk4_coop_p<1, 1, ND_B200_E_KURAMOTO, 1>(e, P, R, st, query, max_grid);
    default: return launch_rk4_coop_p<1, 1, EK_GENERIC, 1>(e, P, R, st, query, max_grid);
  }
}

// ---- batched jagged kernel launches ------------------------------------------------------------------------------------
template <int EK, int PE, bool PK>
cudaError_t launch_jagb_t(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  const int grid = (e->launch_nblk >= 0 ? e->launch_nblk : e->n_jag_blocks + e->n_jlong) + P.n_pub + P.fence;
  if (grid == 0) return cudaSuccess;
  const int ch = e->jaga_ch, wps = e->jaga_wps;
  if (e->halo_base != INT_MAX) ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 8, 32, true>);
  else if (ch >= 16) ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 16, 24, false>);
  else if (ch >= 8) {
    if (wps >= 48) ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 8, 48, false>);
    else if (wps >= 40) ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 8, 40, false>);
    else ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 8, 32, false>);
  } else if (ch >= 6) {
    if (wps >= 48) ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 6, 48, false>);
    else ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 6, 40, false>);
  } else {
    if (wps >= 64) ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 4, 64, false>);
    else ND_LAUNCH(grid, 128, st, (P), rhs_jagb_kernel<EK, PE, PK, 4, 48, false>);
  }
  return cudaGetLastError();
}
template <int EK, int PE>
cudaError_t launch_jagb_p(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  if constexpr (PE > 0) {
    if (e->pack_on) return launch_jagb_t<EK, PE, true>(e, P, st);
  }
  return launch_jagb_t<EK, PE, false>(e, P, st);
}
cudaError_t launch_jagb(nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  e->launches += (e->n_jag_blocks + e->n_jlong > 0);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_jagb_p<ND_B200_E_DIFFUSION, 1>(e, P, st);
    case ND_B200_E_DIFFUSION_NOP: return launch_jagb_p<ND_B200_E_DIFFUSION_NOP, 0>(e, P, st);
    case ND_B200_E_KURAMOTO: return launch_jagb_p<ND_B200_E_KURAMOTO, 1>(e, P, st);
    default: return cudaErrorInvalidConfiguration;
  }
}

// ---- asynchronous-gather jagged kernel launches ----------------------------------------------------------------------
template <int EK, int PE, bool PK>
cudaError_t launch_jaga_t(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  const int grid = (e->launch_nblk >= 0 ? e->launch_nblk : e->n_jag_blocks + e->n_jlong) + P.n_pub + P.fence;
  if (grid == 0) return cudaSuccess;
  if (e->halo_base != INT_MAX) ND_LAUNCH(grid, 128, st, (P), rhs_jaga_kernel<EK, PE, PK, 8, 48, true>);
  else if (e->jaga_ch >= 16) ND_LAUNCH(grid, 128, st, (P), rhs_jaga_kernel<EK, PE, PK, 16, 24, false>);
  else if (e->jaga_ch >= 8) {
    if (e->jaga_wps >= 48) ND_LAUNCH(grid, 128, st, (P), rhs_jaga_kernel<EK, PE, PK, 8, 48, false>);
    else ND_LAUNCH(grid, 128, st, (P), rhs_jaga_kernel<EK, PE, PK, 8, 32, false>);
  } else {
    if (e->jaga_wps >= 64) ND_LAUNCH(grid, 128, st, (P), rhs_jaga_kernel<EK, PE, PK, 4, 64, false>);
    else ND_LAUNCH(grid, 128, st, (P), rhs_jaga_kernel<EK, PE, PK, 4, 48, false>);
  }
  return cudaGetLastError();
}
template <int EK, int PE>
cudaError_t launch_jaga_p(const nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  if constexpr (PE > 0) {
    if (e->pack_on) return launch_jaga_t<EK, PE, true>(e, P, st);
  }
  return launch_jaga_t<EK, PE, false>(e, P, st);
}
cudaError_t launch_jaga(nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  e->launches += (e->n_jag_blocks + e->n_jlong > 0);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_jaga_p<ND_B200_E_DIFFUSION, 1>(e, P, st);
    case ND_B200_E_DIFFUSION_NOP: return launch_jaga_p<ND_B200_E_DIFFUSION_NOP, 0>(e, P, st);
    case ND_B200_E_KURAMOTO: return launch_jaga_p<ND_B200_E_KURAMOTO, 1>(e, P, st);
    default: return cudaErrorInvalidConfiguration;
  }
}

// ---- streamed jagged kernel launches -----------------------------------------------------------------------------------
template <int EK, int PE, bool PK, int U, int NST, int MINB>
cudaError_t launch_js_inst(const nd_b200_engine* e, const KParams& P, cudaStream_t st, bool prepare) {
  const int wpb = JS_BLOCK / 32;
  const int smem = js_warp_bytes(PE, PK, NST) * wpb;
  if (prepare) {   // engine construction: opt in to the dynamic shared memory of this instantiation (not a stream operation)
#ifndef ND_CUSIM
    return cudaFuncSetAttribute(rhs_js_kernel<EK, PE, PK, U, NST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
#else
    return cudaSuccess;
#endif
  }
  const int grid = std::max((e->n_jwarps + wpb - 1) / wpb, std::min(e->n_jlong, 148 * 4));
  if (grid == 0) return cudaSuccess;
#ifdef ND_CUSIM
  ND_LAUNCH(grid, JS_BLOCK, st, (P), rhs_js_kernel<EK, PE, PK, U, NST, MINB>);
#else
  rhs_js_kernel<EK, PE, PK, U, NST, MINB><<<grid, JS_BLOCK, smem, st>>>(P);
#endif
  return cudaGetLastError();
}
template <int EK, int PE, bool PK>
cudaError_t launch_js_shape(const nd_b200_engine* e, const KParams& P, cudaStream_t st, bool prepare) {
  // (U, NST): gathers in flight per lane, chunks per ring; MINB caps registers so that js_wps warps fit
  if (e->js_u >= 8) {
    if (e->js_nst >= 8) return launch_js_inst<EK, PE, PK, 8, 8, 4>(e, P, st, prepare);
    return launch_js_inst<EK, PE, PK, 8, 4, 6>(e, P, st, prepare);
  }
  if (e->js_nst >= 8) return launch_js_inst<EK, PE, PK, 4, 8, 4>(e, P, st, prepare);
  return launch_js_inst<EK, PE, PK, 4, 4, 8>(e, P, st, prepare);
}
template <int EK, int PE>
cudaError_t launch_js_t(const nd_b200_engine* e, const KParams& P, cudaStream_t st, bool prepare) {
  if constexpr (PE > 0) {
    if (prepare) {
      cudaError_t c = launch_js_shape<EK, PE, true>(e, P, st, true);
      if (c != cudaSuccess) return c;
    } else if (e->pack_on) return launch_js_shape<EK, PE, true>(e, P, st, false);
  }
  return launch_js_shape<EK, PE, false>(e, P, st, prepare);
}
cudaError_t launch_js(nd_b200_engine* e, const KParams& P, cudaStream_t st, bool prepare = false) {
  if (!prepare) e->launches += (e->n_jwarps + e->n_jlong > 0);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_js_t<ND_B200_E_DIFFUSION, 1>(e, P, st, prepare);
    case ND_B200_E_DIFFUSION_NOP: return launch_js_t<ND_B200_E_DIFFUSION_NOP, 0>(e, P, st, prepare);
    case ND_B200_E_KURAMOTO: return launch_js_t<ND_B200_E_KURAMOTO, 1>(e, P, st, prepare);
    default: return cudaErrorInvalidConfiguration;
  }
}

cudaError_t launch_custom(nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  const int nb = e->jag ? e->n_jag_blocks + e->n_jlong : e->nblocks;
  const int grid = (e->launch_nblk >= 0 ? e->launch_nblk : nb) + P.n_pub + P.fence;
  if (grid == 0) return cudaSuccess;
  e->launches++;
  void* args[] = {const_cast<KParams*>(&P)};
  return cudaLaunchKernel((const void*)(e->jag ? e->c_jag : e->c_fused), dim3((unsigned)grid), dim3(128), args, 0, st);
}

cudaError_t launch_fused(nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  if (e->custom) return launch_custom(e, P, st);
  if (e->jstream) return launch_js(e, P, st);
  if (e->jagb) return launch_jagb(e, P, st);
  if (e->jaga) return launch_jaga(e, P, st);
  if (e->jag && jag_alt_ok(e, P)) return launch_jag_alt(e, P, st);
  if (e->jag) return launch_jag(e, P, st);
  if (e->split) {
    // edge pass (PASS 5) then row pass (aggregate + PASS 6); P.gsrc is the gather source of this evaluation
    cudaError_t c1 = launch_edge_pass(e, P.gsrc, P.p, st);
    if (c1 != cudaSuccess || e->nblocks == 0) return c1;
    e->launches++;
    return e->vdepth == 2 ? launch_row_pass_t<2, 2>(e, P, st) : launch_row_pass_t<1, 1>(e, P, st);
  }
  e->launches += (e->nblocks > 0);
  if (e->vdepth == 2) return launch_shape<2, 2, ND_B200_E_LINE_DQ, 3>(e, P, st);
  switch (e->ek) {
    case ND_B200_E_DIFFUSION: return launch_shape<1, 1, ND_B200_E_DIFFUSION, 1>(e, P, st);
    case ND_B200_E_DIFFUSION_NOP: return launch_shape<1, 1, ND_B200_E_DIFFUSION_NOP, 0>(e, P, st);
    case ND_B200_E_KURAMOTO: return launch_shape<1, 1, ND_B200_E_KURAMOTO, 1>(e, P, st);
    default: return launch_shape<1, 1, EK_GENERIC, 1>(e, P, st);
  }
}

// PASS 1 (+ PASS 3 for feed-forward vertices, which read their hub's output written by the first launch)
cudaError_t launch_vout(nd_b200_engine* e, const double* u, const double* p, double* vout, cudaStream_t st, double t = 0.0) {
  const int T = 256;
  const int nb = (int)((e->nrows_total + T - 1) / T);
  if (nb == 0) return cudaSuccess;
  for (int phase = 0; phase < (e->any_ff ? 2 : 1); ++phase) {
    e->launches++;
    if (e->custom) {
      const VBDev* vb = e->d_vb; int nvb = (int)e->hvb.size(), vd = e->vdepth, nr = (int)e->nrows_total; double t0 = t;
      void* args[] = {&vb, &nvb, &vd, &u, &p, &vout, &nr, &t0, &phase};
      cudaError_t c = cudaLaunchKernel((const void*)e->c_vout, dim3((unsigned)nb), dim3(T), args, 0, st);
      if (c != cudaSuccess) return c;
    } else {
      ND_LAUNCH(nb, T, st, (e->d_vb, (int)e->hvb.size(), e->vdepth, u, p, vout, (int)e->nrows_total, t, phase), vertex_out_kernel);
      cudaError_t c = cudaGetLastError();
      if (c != cudaSuccess) return c;
    }
  }
  return cudaSuccess;
}

// PASS 4 for edge batches with states: du_e = f(u_e, v_src, v_dst, p, t), with the epilogue of the evaluation P describes
cudaError_t launch_edge_f(nd_b200_engine* e, const KParams& P, cudaStream_t st) {
  for (const auto& ob : e->ode) {
    const HostEB& h = e->heb[(size_t)ob.b];
    EFParams Q;
    memset(&Q, 0, sizeof Q);
    // row-partitioned engines (all-gather exchange): every stateful batch is cut in the proportion of the owned row range --
    // contiguous chunks that tile the batch when the ranks' row ranges tile the rows (distributed.py: edge_state_segments)
    long long i0 = 0, i1 = h.count;
    if (e->row_end - e->row_begin != e->nrows_total) {
      i0 = h.count * (long long)e->row_begin / std::max<long long>(e->nrows_total, 1);
      i1 = h.count * (long long)e->row_end / std::max<long long>(e->nrows_total, 1);
    }
    Q.kind = h.kind; Q.dim = h.dim; Q.pdim = h.pdim; Q.count = i1 - i0; Q.state0 = h.state0 + i0 * h.dim; Q.p0 = h.p0 + i0 * h.pdim;
    Q.esrc_off = ob.d_es + i0; Q.edst_off = ob.d_et + i0; Q.ext = ob.d_ext ? ob.d_ext + i0 * ob.extdim : nullptr; Q.extdim = ob.extdim;
    Q.u = P.u; Q.gsrc = P.gsrc; Q.p = P.p; Q.du = P.du; Q.mode = P.mode; Q.stage = P.stage; Q.u0 = P.u0; Q.unext = P.unext;
    Q.ksum = P.ksum; Q.hs = P.hs; Q.h6 = P.h6; Q.t = P.t;
    const int T = 256;
    const int nb = (int)((Q.count + T - 1) / T);
    if (nb == 0) continue;
    e->launches++;
    if (e->custom) {
      void* args[] = {&Q};
      cudaError_t c = cudaLaunchKernel((const void*)e->c_ef, dim3((unsigned)nb), dim3(T), args, 0, st);
      if (c != cudaSuccess) return c;
    } else if (e->vdepth == 2) {
      ND_LAUNCH(nb, T, st, (Q), edge_f_kernel<2>);
    } else {
      ND_LAUNCH(nb, T, st, (Q), edge_f_kernel<1>);
    }
    cudaError_t c = cudaGetLastError();
    if (c != cudaSuccess) return c;
  }
  return cudaSuccess;
}

int ensure_events(nd_b200_engine* e, std::vector<cudaEvent_t>& v, size_t need) {
  while (v.size() < need) {
    cudaEvent_t ev;
    CUDA_TRY(e, cudaEventCreate(&ev));
    v.push_back(ev);
  }
  return 0;
}

// ---- user-supplied kinds: source generation + NVRTC ------------------------------------------------------------------
// nd_b200_kernels.cuh as text (generated next to this file by the build, see _cabi.build): the run-time compiled kernels
// are the SAME templates as the precompiled ones, with the user's functions spliced into the model switches.
static const char* const kKernelHeaderText[] = {
#include "nd_b200_kernels_embed.inc"
};

std::string custom_source(const nd_b200_engine* e, int vdepth) {
  std::string src;
  char buf[512];
  src += "// generated by libnd_b200 (user-supplied component kinds)\n";
  src += "typedef unsigned char uint8_t;\ntypedef unsigned short uint16_t;\ntypedef int int32_t;\ntypedef long long int64_t;\n";
  snprintf(buf, sizeof buf,
           "enum { ND_B200_V_DIFFUSION = %d, ND_B200_V_KURAMOTO_FIRST = %d, ND_B200_V_KURAMOTO_SECOND = %d, ND_B200_V_KURAMOTO_SECOND_BENCH = %d, ND_B200_V_SWING_DQ = %d };\n",
           ND_B200_V_DIFFUSION, ND_B200_V_KURAMOTO_FIRST, ND_B200_V_KURAMOTO_SECOND, ND_B200_V_KURAMOTO_SECOND_BENCH, ND_B200_V_SWING_DQ);
  src += buf;
  snprintf(buf, sizeof buf, "enum { ND_B200_E_DIFFUSION = %d, ND_B200_E_DIFFUSION_NOP = %d, ND_B200_E_KURAMOTO = %d, ND_B200_E_LINE_DQ = %d, ND_B200_E_DIFFUSION_ODE = %d, ND_B200_E_RELAX_ODE = %d, ND_B200_E_DIFFUSION_FID = %d, ND_B200_E_LOOPBACK = %d };\n",
           ND_B200_E_DIFFUSION, ND_B200_E_DIFFUSION_NOP, ND_B200_E_KURAMOTO, ND_B200_E_LINE_DQ, ND_B200_E_DIFFUSION_ODE, ND_B200_E_RELAX_ODE, ND_B200_E_DIFFUSION_FID, ND_B200_E_LOOPBACK);
  src += buf;
  snprintf(buf, sizeof buf, "enum { ND_B200_ANTISYMMETRIC = %d, ND_B200_SYMMETRIC = %d, ND_B200_DIRECTED = %d, ND_B200_FIDUCIAL = %d };\n",
           ND_B200_ANTISYMMETRIC, ND_B200_SYMMETRIC, ND_B200_DIRECTED, ND_B200_FIDUCIAL);
  src += buf;
  snprintf(buf, sizeof buf, "#define ND_MAX_VDIM %d\n#define ND_MAX_VOUT %d\n#define ND_MAX_EDIM %d\n#define ND_MAX_EXT %d\n", std::max(e->c_maxdim, 1), std::max(vdepth, 1), std::max(e->c_maxedim, 2), e->c_maxext);
  src += buf;
  std::string edge_cases, fid_cases, vf_cases, vg_cases, ef_cases;
  src += "namespace ndb_user {\n";
  for (const auto& c : e->customs) {
    const std::string id = std::to_string(c.kind);
    if (c.role == 0) {
      const std::string xa = c.extdim > 0 ? "const double* __restrict__ ext, " : "", xc = c.extdim > 0 ? "ext, " : "";
      src += "__device__ __forceinline__ void vertex_f_" + id + "(double* __restrict__ dv, const double* __restrict__ v, const double* __restrict__ esum, " + xa + "const double* __restrict__ p, double t) {\n" + c.f_body + "\n}\n";
      vf_cases += " case " + id + ": ndb_user::vertex_f_" + id + "(dv, v, acc, " + xc + "pv, t); break;";
      if (!c.g_body.empty()) {
        const std::string ia = c.g_ff ? "const double* __restrict__ ins, " : "", ic = c.g_ff ? "ins, " : "";
        src += "__device__ __forceinline__ void vertex_g_" + id + "(double* __restrict__ out, const double* __restrict__ v, " + ia + "const double* __restrict__ p, double t) {\n" + c.g_body + "\n}\n";
        vg_cases += " case " + id + ": ndb_user::vertex_g_" + id + "(out, v, " + ic + "pv, t); break;";
      }
    } else if (c.dim > 0) {   // edge with states: the body is f; outputs are StateMasks
      const std::string xa = c.extdim > 0 ? "const double* __restrict__ ext, " : "", xc = c.extdim > 0 ? "ext, " : "";
      src += "__device__ __forceinline__ void edge_f_" + id + "(double* __restrict__ de, const double* __restrict__ e, const double* __restrict__ v_src, const double* __restrict__ v_dst, " + xa + "const double* __restrict__ p, double t) {\n" + c.f_body + "\n}\n";
      ef_cases += " case " + id + ": ndb_user::edge_f_" + id + "(de, ue, vs, vd, " + xc + "pe, t); break;";
    } else if (c.two_sided) {
      src += "__device__ __forceinline__ void edge_g_" + id + "(double* __restrict__ e_src, double* __restrict__ e_dst, const double* __restrict__ v_src, const double* __restrict__ v_dst, const double* __restrict__ p, double t) {\n" + c.f_body + "\n}\n";
      fid_cases += " case " + id + ": ndb_user::edge_g_" + id + "(osrc, odst, vs, vd, pe, t); break;";
    } else {
      src += "__device__ __forceinline__ void edge_g_" + id + "(double* __restrict__ e_dst, const double* __restrict__ v_src, const double* __restrict__ v_dst, const double* __restrict__ p, double t) {\n" + c.f_body + "\n}\n";
      edge_cases += " case " + id + ": ndb_user::edge_g_" + id + "(odst, vs, vd, pe, t); break;";
    }
  }
  src += "}  // namespace ndb_user\n";
  src += "#define ND_CUSTOM_EDGE_CASES" + edge_cases + "\n";
  src += "#define ND_CUSTOM_EDGE_FID_CASES" + fid_cases + "\n";
  src += "#define ND_CUSTOM_EDGE_F_CASES" + ef_cases + "\n";
  src += "#define ND_CUSTOM_VERTEX_F_CASES" + vf_cases + "\n";
  src += "#define ND_CUSTOM_VERTEX_G_CASES" + vg_cases + "\n";
  for (const char* part : kKernelHeaderText) src += part;
  return src;
}

// compile the generated source for sm_100a; on success load it (unless host_only) and fetch the kernels
int compile_custom(nd_b200_engine* e, int vdepth, int edepth) {
  e->custom_src = custom_source(e, vdepth);
  // process-wide cache of compiled modules: engines built from the same generated source and kernel instantiations (the
  // ranks' row ranges of one network, a network rebuilt with another aggregator option) share one compilation
  struct Compiled { std::vector<char> cubin; std::string lowered[5]; };
  static std::mutex cache_mu;
  static std::map<std::string, Compiled> cache;
  const char* cache_env = getenv("ND_B200_NVRTC_CACHE");
  const bool use_cache = !(cache_env && atoi(cache_env) == 0) && !e->host_only;
  nvrtcProgram prog = nullptr;
  if (nvrtcCreateProgram(&prog, e->custom_src.c_str(), "nd_b200_custom.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
    return fail(e, ND_B200_ECUDA, "nvrtcCreateProgram failed");
  char nm[5][160];
  const int ek = e->ek, pe = e->c_pe;
  const char* halo = e->halo_base != INT_MAX ? "true" : "false";
  snprintf(nm[0], sizeof nm[0], "ndb::rhs_fused_kernel<%d, %d, %d, %d, 128, 4, %s>", vdepth, edepth, ek, pe, halo);
  snprintf(nm[1], sizeof nm[1], "ndb::rhs_jag_kernel<%d, %d, %d, %d, 128, 2, 48, %s>", vdepth, edepth, ek, pe, halo);
  snprintf(nm[2], sizeof nm[2], "ndb::vertex_out_kernel");
  snprintf(nm[3], sizeof nm[3], "ndb::edge_out_kernel<%d, %d>", vdepth, edepth);
  snprintf(nm[4], sizeof nm[4], "ndb::edge_f_kernel<%d>", vdepth);
  std::string key = e->custom_src;
  for (int k = 0; k < 5; ++k) { key += "\n//"; key += nm[k]; }
  Compiled hit;
  bool have = false;
  if (use_cache) {
    std::lock_guard<std::mutex> lk(cache_mu);
    auto it = cache.find(key);
    if (it != cache.end()) { hit = it->second; have = true; }
  }
  if (have) {
    nvrtcDestroyProgram(&prog);
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaLibraryLoadData(&e->c_lib, hit.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    CUDA_TRY(e, cudaLibraryGetKernel(&e->c_fused, e->c_lib, hit.lowered[0].c_str()));
    CUDA_TRY(e, cudaLibraryGetKernel(&e->c_jag, e->c_lib, hit.lowered[1].c_str()));
    CUDA_TRY(e, cudaLibraryGetKernel(&e->c_vout, e->c_lib, hit.lowered[2].c_str()));
    CUDA_TRY(e, cudaLibraryGetKernel(&e->c_eout, e->c_lib, hit.lowered[3].c_str()));
    CUDA_TRY(e, cudaLibraryGetKernel(&e->c_ef, e->c_lib, hit.lowered[4].c_str()));
    return ND_B200_OK;
  }
  for (int k = 0; k < 5; ++k) nvrtcAddNameExpression(prog, nm[k]);
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", "-default-device"};
  const nvrtcResult rc = nvrtcCompileProgram(prog, 5, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtcGetProgramLog(prog, &log[0]);
    nvrtcDestroyProgram(&prog);
    if (log.size() > 400) log.resize(400);
    return fail(e, ND_B200_EINVAL, "user-supplied component code does not compile: %s", log.c_str());
  }
  if (e->host_only) { nvrtcDestroyProgram(&prog); return ND_B200_OK; }
  size_t nbin = 0;
  nvrtcGetCUBINSize(prog, &nbin);
  std::vector<char> cubin(nbin);
  nvrtcGetCUBIN(prog, cubin.data());
  std::string lowered[5];
  for (int k = 0; k < 5; ++k) {
    const char* ln = nullptr;
    if (nvrtcGetLoweredName(prog, nm[k], &ln) != NVRTC_SUCCESS || !ln) { nvrtcDestroyProgram(&prog); return fail(e, ND_B200_ECUDA, "no lowered name for %s", nm[k]); }
    lowered[k] = ln;
  }
  nvrtcDestroyProgram(&prog);
  if (use_cache) {
    std::lock_guard<std::mutex> lk(cache_mu);
    Compiled& c = cache[key];
    c.cubin = cubin;
    for (int k = 0; k < 5; ++k) c.lowered[k] = lowered[k];
  }
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaLibraryLoadData(&e->c_lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  CUDA_TRY(e, cudaLibraryGetKernel(&e->c_fused, e->c_lib, lowered[0].c_str()));
  CUDA_TRY(e, cudaLibraryGetKernel(&e->c_jag, e->c_lib, lowered[1].c_str()));
  CUDA_TRY(e, cudaLibraryGetKernel(&e->c_vout, e->c_lib, lowered[2].c_str()));
  CUDA_TRY(e, cudaLibraryGetKernel(&e->c_eout, e->c_lib, lowered[3].c_str()));
  CUDA_TRY(e, cudaLibraryGetKernel(&e->c_ef, e->c_lib, lowered[4].c_str()));
  return ND_B200_OK;
}

// Construction of an engine from a descriptor, stage by stage (host side; see the comments of each stage).  The members are
// the tables the stages hand to each other; what the engine keeps is copied / uploaded by the last stage.
struct EngineBuilder {
  nd_b200_engine* e;
  const nd_b200_desc* d;
  // vertices
  std::vector<int> row_of_vertex;            // vertex id - 1 -> aggregation-slot row
  std::vector<int> goff;                     // vertex id - 1 -> offset of its output in the gather source
  long long state_expect = 1, out_expect = 1, p_expect = 1, nrows_owned = 0;
  // edges
  bool any_epar = false, any_ode = false, any_fiducial = false, any_ext = false, any_loopback = false;
  std::vector<int> hub_of_vertex;            // vertex id - 1 -> hub vertex id (1-based) of an injector, else 0
  std::vector<std::vector<int>> vext_codes, eext_codes;   // per batch: resolved external-input sources (see VBDev::ext)
  // CSR over the owned rows, entries in accumulation order
  std::vector<long long> cnt;
  std::vector<int> h_rowptr, h_nbr, h_epar;
  std::vector<int> h_nbr_c;                  // tile layout, compact entry words (engine::compact)
  std::vector<uint8_t> h_ebid;
  bool keep = false, generic_edges = false, want_split = false;
  std::vector<int> h_oidx, h_es, h_et, h_eepar, h_eooff;   // split mode
  std::vector<uint8_t> h_eebid;
  std::vector<char> row_remote;              // owned row reads the halo
  // tile layout
  std::vector<int> blk_row;
  std::vector<VBDev> dvb;
  std::vector<EBDev> deb;
  std::vector<int4> tiles;
  // jagged layout
  std::vector<int4> jslices, jlong;
  std::vector<uint16_t> jlanes;
  std::vector<int> jnbr;
  std::vector<int2> jent;
  std::vector<uint8_t> jebid;
  bool jag_pe = false;

  EngineBuilder(nd_b200_engine* e_, const nd_b200_desc* d_) : e(e_), d(d_) {}

  bool owned(int r) const { return r >= e->row_begin && r < e->row_end; }
  // end of the parameter range an entry / a row reads (host-buffer pipeline, nd_b200_rhs_host)
  int entry_pend(long long j) const {
    if (!any_epar) return 0;
    const int pd = h_ebid.empty() ? e->heb[0].pdim : e->heb[h_ebid[(size_t)j]].pdim;
    return pd > 0 ? h_epar[(size_t)j] + pd : 0;
  }
  int row_pend(long long r, size_t b) const {
    const HostVB& h = e->hvb[b];
    return h.pdim > 0 ? (int)(h.p0 + (r - h.row0 + 1) * h.pdim) : 0;
  }

  // sizes, limits, user-supplied kinds
  int check_descriptor() {
    if (d->abi_version != ND_B200_ABI_VERSION) return fail(e, ND_B200_EINVAL, "descriptor abi_version %d != %d", d->abi_version, ND_B200_ABI_VERSION);
    if (d->nv <= 0) return fail(e, ND_B200_EINVAL, "network needs at least one vertex");
    if (d->n_vbatches <= 0 || d->n_vbatches > MAX_VB) return fail(e, ND_B200_EUNSUPPORTED, "number of vertex batches %d outside 1..%d", d->n_vbatches, MAX_VB);
    if (d->n_ebatches < 0 || d->n_ebatches > MAX_EB) return fail(e, ND_B200_EUNSUPPORTED, "number of edge batches %d outside 0..%d", d->n_ebatches, MAX_EB);
    if (d->ne > 0 && d->n_ebatches == 0) return fail(e, ND_B200_EINVAL, "edges without edge batches");
    if (d->ne < 0 || !d->vbatches || (d->n_ebatches > 0 && !d->ebatches) || (d->ne > 0 && (!d->edge_src || !d->edge_dst)))
      return fail(e, ND_B200_EINVAL, "descriptor with missing tables");
    if (d->vdepth < 1 || (d->ne > 0 && d->edepth < 1)) return fail(e, ND_B200_EINVAL, "vdepth / edepth must be positive");
    e->device = d->device;
#ifndef ND_CUSIM
    if (!(d->flags & ND_B200_FLAG_HOST_ONLY)) {
      int sms = 0;
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device) == cudaSuccess && sms > 0) e->num_sms = sms;
      else cudaGetLastError();
    }
#endif
    e->nv = d->nv; e->ne = d->ne; e->vdepth = d->vdepth; e->edepth = d->ne > 0 ? d->edepth : d->vdepth;
    e->lastidx_dynamic = d->lastidx_dynamic; e->lastidx_p = d->lastidx_p;
    e->lastidx_out = d->lastidx_out; e->lastidx_aggr = d->lastidx_aggr;
    if (d->n_custom < 0 || (d->n_custom > 0 && !d->custom)) return fail(e, ND_B200_EINVAL, "bad custom kind table");
    for (int k = 0; k < d->n_custom; ++k) {
      const nd_b200_custom_kind& c = d->custom[k];
      if (c.kind < ND_B200_CUSTOM_KIND_BASE || (c.role != 0 && c.role != 1) || !c.f_body) return fail(e, ND_B200_EINVAL, "custom kind %d: id must be >= %d, role 0|1, f_body non-NULL", c.kind, ND_B200_CUSTOM_KIND_BASE);
      if (c.dim < 0 || c.dim > 16 || c.pdim < 0 || c.pdim > 64 || c.outdim < 1 || c.outdim > 8) return fail(e, ND_B200_EUNSUPPORTED, "custom kind %d: dims outside dim<=16, pdim<=64, 1<=outdim<=8", c.kind);
      if (c.role == 1) e->c_maxedim = std::max(e->c_maxedim, c.dim);
      if (c.extdim < 0 || c.extdim > 32) return fail(e, ND_B200_EUNSUPPORTED, "custom kind %d: extdim outside 0..32", c.kind);
    e->c_maxext = std::max(e->c_maxext, c.extdim);
    if (c.g_ff && (c.role != 0 || !c.g_body)) return fail(e, ND_B200_EINVAL, "custom kind %d: g_ff needs a vertex kind with a g body", c.kind);
    e->customs.push_back(nd_b200_engine::CustomKind{c.kind, c.role, c.dim, c.pdim, c.outdim, c.two_sided, c.f_body, c.g_body ? c.g_body : "", c.extdim, c.g_ff});
    }
    for (int b = 0; b < d->n_vbatches; ++b) e->custom = e->custom || d->vbatches[b].kind >= ND_B200_CUSTOM_KIND_BASE;
    for (int b = 0; b < d->n_ebatches; ++b) e->custom = e->custom || d->ebatches[b].kind >= ND_B200_CUSTOM_KIND_BASE;
    if (!e->custom && !((d->vdepth == 1 && (d->ne == 0 || d->edepth == 1)) || (d->vdepth == 2 && d->edepth == 2)))
      return fail(e, ND_B200_EUNSUPPORTED, "no precompiled kernel for (vdepth,edepth)=(%d,%d); available: (1,1),(2,2) -- other shapes need user-supplied kinds", d->vdepth, d->edepth);
    if (e->custom && (d->vdepth < 1 || d->vdepth > 8 || (d->ne > 0 && (d->edepth < 1 || d->edepth > 8))))
      return fail(e, ND_B200_EUNSUPPORTED, "(vdepth,edepth)=(%d,%d) outside 1..8", d->vdepth, d->edepth);
    if (d->lastidx_dynamic >= INT_MAX || d->lastidx_p >= INT_MAX || d->lastidx_out >= (long long)INT_MAX * 2)
      return fail(e, ND_B200_EUNSUPPORTED, "network too large for 32-bit offsets");
    e->long_thr = d->long_row_threshold > 0 ? d->long_row_threshold : 128;
    return ND_B200_OK;
  }

  // vertex batches: registry check, contiguity of rows / states (register_vertices!), gather offsets, halo layout
  int register_vertices() {
    // ---- vertex batches: registry check, contiguity of rows/states (register_vertices!) ----------
    row_of_vertex.assign((size_t)d->nv, -1);
    long long row = 0;
    const int ed = d->ne > 0 ? d->edepth : 0;
    bool all_statemask1 = true;
    for (int b = 0; b < d->n_vbatches; ++b) {
      const nd_b200_vbatch& vb = d->vbatches[b];
      std::string why;
      if (!vertex_kind_ok(e, vb, why)) return fail(e, ND_B200_EUNSUPPORTED, "%s (no CPU fallback)", why.c_str());
      if (vb.outdim != d->vdepth) return fail(e, ND_B200_EINVAL, "vertex batch %d outdim %d != vdepth %d", b + 1, vb.outdim, d->vdepth);
      if (vb.count <= 0 || vb.count > d->nv || (!vb.indices && d->n_vbatches != 1)) return fail(e, ND_B200_EINVAL, "vertex batch %d is empty or larger than the graph", b + 1);
      if (vb.state_first != state_expect) return fail(e, ND_B200_EINVAL, "vertex batch %d: statestride.first %lld, expected %lld", b + 1, (long long)vb.state_first, state_expect);
      if (vb.out_first != out_expect) return fail(e, ND_B200_EINVAL, "vertex batch %d: outbufstride.first %lld, expected %lld", b + 1, (long long)vb.out_first, out_expect);
      if (vb.dim < 0 || vb.pdim < 0) return fail(e, ND_B200_EINVAL, "vertex batch %d: negative dimension", b + 1);
      if (vb.pdim > 0 && vb.p_first != p_expect) return fail(e, ND_B200_EINVAL, "vertex batch %d: pstride.first %lld, expected %lld", b + 1, (long long)vb.p_first, p_expect);
      if (ed > 0 && vb.aggr_first != row * ed + 1) return fail(e, ND_B200_EINVAL, "vertex batch %d: inbufstride.first %lld, expected %lld", b + 1, (long long)vb.aggr_first, row * ed + 1);
      const bool ff = vb.kind >= ND_B200_CUSTOM_KIND_BASE && find_custom(e, vb.kind, 0)->g_ff;
      e->any_ff = e->any_ff || ff;
      HostVB h{vb.kind, vb.dim, vb.pdim, vb.outdim, vb.count, vb.state_first - 1, vb.p_first - 1, vb.out_first - 1, row, ff ? 1 : 0};
      e->hvb.push_back(h);
      for (long long i = 0; i < vb.count; ++i) {
        long long vid = vb.indices ? vb.indices[i] : i + 1;
        if (vid < 1 || vid > d->nv || row_of_vertex[(size_t)vid - 1] >= 0) return fail(e, ND_B200_EINVAL, "vertex batch %d: bad or duplicate vertex id %lld", b + 1, vid);
        row_of_vertex[(size_t)vid - 1] = (int)(row + i);
      }
      row += vb.count; state_expect += vb.count * vb.dim; out_expect += vb.count * vb.outdim; p_expect += vb.count * vb.pdim;
      if (vb.kind == ND_B200_V_SWING_DQ) all_statemask1 = false;
      if (vb.kind >= ND_B200_CUSTOM_KIND_BASE && !find_custom(e, vb.kind, 0)->g_body.empty()) all_statemask1 = false;
      if (vb.dim > (e->custom ? 16 : 2)) return fail(e, ND_B200_EUNSUPPORTED, "vertex batch %d: dim %d", b + 1, vb.dim);
      e->c_maxdim = std::max(e->c_maxdim, vb.dim);
    }
    if (row != d->nv) return fail(e, ND_B200_EINVAL, "vertex batches cover %lld of %lld vertices", row, (long long)d->nv);
    e->nrows_total = row;
    e->gather_from_u = (all_statemask1 && d->vdepth == 1) ? 1 : 0;

    e->row_begin = 0; e->row_end = e->nrows_total;
    if (d->row_end > 0 || (d->flags & ND_B200_FLAG_ROW_RANGE)) {
      if (d->row_begin < 0 || d->row_begin > d->row_end || d->row_end > e->nrows_total) return fail(e, ND_B200_EINVAL, "row partition [%lld,%lld) outside 0..%lld", (long long)d->row_begin, (long long)d->row_end, e->nrows_total);
      e->row_begin = d->row_begin; e->row_end = d->row_end;
    }
    nrows_owned = e->row_end - e->row_begin;

    // gather offset of a vertex's output inside the gather source
    goff.assign((size_t)d->nv, 0);
    for (int b = 0; b < d->n_vbatches; ++b) {
      const HostVB& h = e->hvb[b];
      for (long long i = 0; i < h.count; ++i) {
        long long vid = d->vbatches[b].indices ? d->vbatches[b].indices[i] : i + 1;
        goff[(size_t)vid - 1] = e->gather_from_u ? (int)(h.state0 + i * h.dim) : (int)((h.row0 + i) * d->vdepth);
      }
    }
    if (d->gather_offset) {
      // multi-GPU packed halo: remote vertices are read from the halo buffer appended (logically) to the state vector
      if (!e->gather_from_u) return fail(e, ND_B200_EUNSUPPORTED, "gather_offset needs StateMask vertices with one output");
      if (d->gather_len < d->lastidx_dynamic || d->gather_len >= INT_MAX) return fail(e, ND_B200_EINVAL, "gather_len %lld outside [lastidx_dynamic, 2^31)", (long long)d->gather_len);
      for (long long v = 0; v < d->nv; ++v) {
        const long long o = d->gather_offset[v];
        if (o < 0 || o >= d->gather_len) return fail(e, ND_B200_EINVAL, "gather_offset[%lld] = %lld outside [0, %lld)", v + 1, o, (long long)d->gather_len);
        const int r = row_of_vertex[(size_t)v];
        if (r >= e->row_begin && r < e->row_end && o != goff[(size_t)v]) return fail(e, ND_B200_EINVAL, "gather_offset of vertex %lld (an owned row) must be its own state offset", v + 1);
        goff[(size_t)v] = (int)o;
      }
      e->halo_base = (int)d->lastidx_dynamic;
      e->gather_len = d->gather_len;
    }
    return ND_B200_OK;
  }

  // edge batches: registry check, wrappers, strides (register_edges!), kernel family of the network
  int register_edges() {
    // ---- edge batches --------------------------------------------------------------------------
    long long eout_expect = out_expect;
    std::vector<char> edge_seen((size_t)std::max<long long>(d->ne, 1), 0);
    for (int b = 0; b < d->n_ebatches; ++b) {
      const nd_b200_ebatch& eb = d->ebatches[b];
      std::string why;
      if (!edge_kind_ok(e, eb, d->vdepth, why)) return fail(e, ND_B200_EUNSUPPORTED, "%s (no CPU fallback)", why.c_str());
      if (eb.outdim_dst != d->edepth) return fail(e, ND_B200_EINVAL, "edge batch %d: outdim.dst %d != edepth %d", b + 1, eb.outdim_dst, d->edepth);
      e->c_pe = std::max(e->c_pe, eb.pdim);
      if (eb.dim < 0 || eb.dim > (e->custom ? 16 : 2)) return fail(e, ND_B200_EUNSUPPORTED, "edge batch %d: dim %d", b + 1, eb.dim);
      if (eb.dim > 0) {
        // edges with states: outputs are StateMasks over a contiguous range of the edge's own states
        if (d->gather_offset) return fail(e, ND_B200_EUNSUPPORTED, "edges with states on a halo engine");
        if (eb.state_first != state_expect) return fail(e, ND_B200_EINVAL, "edge batch %d: statestride.first %lld, expected %lld", b + 1, (long long)eb.state_first, state_expect);
        if (eb.mask_dst_first < 1 || eb.mask_dst_first + eb.outdim_dst - 1 > eb.dim) return fail(e, ND_B200_EINVAL, "edge batch %d: dst StateMask outside 1..dim", b + 1);
        if (eb.coupling == ND_B200_FIDUCIAL && (eb.mask_src_first < 1 || eb.mask_src_first + eb.outdim_src - 1 > eb.dim)) return fail(e, ND_B200_EINVAL, "edge batch %d: src StateMask outside 1..dim", b + 1);
        state_expect += eb.count * eb.dim;
        any_ode = true;
      }
      if (eb.coupling != ND_B200_ANTISYMMETRIC && eb.coupling != ND_B200_SYMMETRIC && eb.coupling != ND_B200_DIRECTED && eb.coupling != ND_B200_FIDUCIAL)
        return fail(e, ND_B200_EUNSUPPORTED, "edge batch %d: unsupported output wrapper %d", b + 1, eb.coupling);
      const int osrc_expect = eb.coupling == ND_B200_DIRECTED ? 0 : eb.outdim_dst;
      if (eb.outdim_src != osrc_expect) return fail(e, ND_B200_EINVAL, "edge batch %d: outdim.src %d inconsistent with wrapper", b + 1, eb.outdim_src);
      if (eb.out_first != eout_expect) return fail(e, ND_B200_EINVAL, "edge batch %d: outbufstride.first %lld, expected %lld", b + 1, (long long)eb.out_first, eout_expect);
      if (eb.pdim < 0 || eb.pdim > 64) return fail(e, ND_B200_EINVAL, "edge batch %d: pdim %d", b + 1, eb.pdim);
      if (eb.pdim > 0 && eb.p_first != p_expect) return fail(e, ND_B200_EINVAL, "edge batch %d: pstride.first %lld, expected %lld", b + 1, (long long)eb.p_first, p_expect);
      p_expect += eb.count * eb.pdim;
      if (eb.count <= 0 || eb.count > d->ne || (!eb.indices && d->n_ebatches != 1)) return fail(e, ND_B200_EINVAL, "edge batch %d is empty or larger than the graph", b + 1);
      HostEB h{eb.kind, eb.coupling, eb.dim, eb.pdim, eb.outdim_src, eb.outdim_dst, eb.count, eb.p_first - 1, eb.out_first - 1,
               eb.state_first - 1, eb.dim > 0 ? eb.mask_src_first - 1 : 0, eb.dim > 0 ? eb.mask_dst_first - 1 : 0};
      e->heb.push_back(h);
      eout_expect += eb.count * (eb.outdim_src + eb.outdim_dst);
      if (eb.pdim > 0) any_epar = true;
      for (long long i = 0; i < eb.count; ++i) {
        long long eid = eb.indices ? eb.indices[i] : i + 1;
        if (eid < 1 || eid > d->ne || edge_seen[(size_t)eid - 1]) return fail(e, ND_B200_EINVAL, "edge batch %d: bad or duplicate edge id %lld", b + 1, eid);
        edge_seen[(size_t)eid - 1] = 1;
      }
    }
    if (eout_expect - 1 != d->lastidx_out) return fail(e, ND_B200_EINVAL, "lastidx_out %lld inconsistent with batches (%lld)", (long long)d->lastidx_out, eout_expect - 1);
    if (state_expect - 1 != d->lastidx_dynamic) return fail(e, ND_B200_EINVAL, "lastidx_dynamic %lld inconsistent with the batches (%lld)", (long long)d->lastidx_dynamic, state_expect - 1);
    if (p_expect - 1 != d->lastidx_p) return fail(e, ND_B200_EINVAL, "lastidx_p %lld inconsistent with the batches (%lld)", (long long)d->lastidx_p, p_expect - 1);
    e->ek = (d->n_ebatches == 1 && !any_ode) ? d->ebatches[0].kind : EK_GENERIC;   // entries of edges with states: generic kernels only
    // precompiled specialisations exist for the benchmark edge kinds; every other registry kind runs in the generic kernels
    if (!e->custom && d->vdepth == 1 && e->ek != ND_B200_E_DIFFUSION && e->ek != ND_B200_E_DIFFUSION_NOP && e->ek != ND_B200_E_KURAMOTO) e->ek = EK_GENERIC;
    for (int b = 0; b < d->n_ebatches; ++b) any_fiducial = any_fiducial || d->ebatches[b].coupling == ND_B200_FIDUCIAL;
    for (int b = 0; b < d->n_ebatches; ++b) any_loopback = any_loopback || d->ebatches[b].kind == ND_B200_E_LOOPBACK;
    if (any_loopback || e->any_ff) {
      // LoopbackConnection topology (src/construction.jl:52-80): a loopback edge starts at a LEAF (its only edge) -- the
      // injector -- and every feed-forward vertex is such an injector
      if (d->vdepth != d->edepth) return fail(e, ND_B200_EINVAL, "loopback edges need vdepth == edepth");
      if (d->gather_offset) return fail(e, ND_B200_EUNSUPPORTED, "loopback edges on a halo engine");   // row partitions with the complete u are fine
      if (!e->custom && d->vdepth != 1) return fail(e, ND_B200_EUNSUPPORTED, "loopback edges with vdepth %d need user-supplied kinds", d->vdepth);
      std::vector<int> deg((size_t)d->nv, 0);
      for (long long k = 0; k < d->ne; ++k) {
        const long long a = d->edge_src[k], z = d->edge_dst[k];
        if (a < 1 || a > d->nv || z < 1 || z > d->nv) return fail(e, ND_B200_EINVAL, "edge %lld endpoint out of range", k + 1);
        deg[(size_t)a - 1]++; deg[(size_t)z - 1]++;
      }
      hub_of_vertex.assign((size_t)d->nv, 0);
      for (int b = 0; b < d->n_ebatches; ++b) {
        const nd_b200_ebatch& eb = d->ebatches[b];
        if (eb.kind != ND_B200_E_LOOPBACK) continue;
        for (long long i = 0; i < eb.count; ++i) {
          const long long eid = eb.indices ? eb.indices[i] - 1 : i;
          const long long a = d->edge_src[eid], z = d->edge_dst[eid];
          if (deg[(size_t)a - 1] != 1) return fail(e, ND_B200_EINVAL, "all LoopbackConnection edges must originate from leaf nodes (edge %lld)", eid + 1);
          hub_of_vertex[(size_t)a - 1] = (int)z;
        }
      }
      for (int b = 0; b < d->n_vbatches; ++b) {
        if (!e->hvb[(size_t)b].ff) continue;
        for (long long i = 0; i < d->vbatches[b].count; ++i) {
          const long long vid = d->vbatches[b].indices ? d->vbatches[b].indices[i] : i + 1;
          if (!hub_of_vertex[(size_t)vid - 1]) return fail(e, ND_B200_EINVAL, "feed-forward vertex %lld: feed forward vertex models are only allowed as leaf nodes with a single LoopbackConnection to their hub", vid);
        }
      }
      for (long long v = 0; v < d->nv; ++v)      // a hub is not itself a feed-forward vertex: its output must exist after PASS 1
        if (hub_of_vertex[(size_t)v] && e->hvb.size()) {
          const int hr = row_of_vertex[(size_t)hub_of_vertex[(size_t)v] - 1];
          for (const HostVB& h : e->hvb)
            if (hr >= h.row0 && hr < h.row0 + h.count && h.ff) return fail(e, ND_B200_EINVAL, "the hub of injector %lld is a feed-forward vertex", v + 1);
        }
    }
    if (any_ode && (d->lastidx_dynamic >= ND_STATE_ENTRY_BIT || d->nv * (long long)d->vdepth >= ND_STATE_ENTRY_BIT))
      return fail(e, ND_B200_EUNSUPPORTED, "networks with edge states need offsets below 2^30");
    if (any_ode && !e->custom && d->vdepth != 1) return fail(e, ND_B200_EUNSUPPORTED, "edges with states and vdepth %d need user-supplied kinds", d->vdepth);
    e->pack_pe = (!e->custom && e->ek != EK_GENERIC && d->n_ebatches == 1) ? d->ebatches[0].pdim : 0;
    if (!e->custom && d->vdepth == 2 && d->n_ebatches > 1) {
      // all (2,2) batches must be LINE_DQ with one coupling (single templated kernel)
      for (int b = 1; b < d->n_ebatches; ++b)
        if (d->ebatches[b].coupling != d->ebatches[0].coupling) return fail(e, ND_B200_EUNSUPPORTED, "mixed wrappers for dq lines");
    }
    return ND_B200_OK;
  }

  // One entry of the reference's ExtMap (src/external_inputs.jl:1-50) -> where the kernels read it: StateBufIdx = the
  // state vector; OutBufIdx = an output of a component WITHOUT feed forward, i.e. a vertex output (a state for StateMask
  // vertices, else the materialised output block) or a StateMask output of an edge with states (a state, negated on the src
  // side of AntiSymmetric).
  int resolve_ext_source(long long src, int& code) const {
    if (src > 0) {
      if (src > d->lastidx_dynamic) return fail(e, ND_B200_EINVAL, "external input: state index %lld outside 1..%lld", src, (long long)d->lastidx_dynamic);
      code = (int)(src - 1);
      return ND_B200_OK;
    }
    const long long o = -src - 1;     // 0-based position in the output buffer
    if (src == 0 || o >= d->lastidx_out) return fail(e, ND_B200_EINVAL, "external input: output index %lld outside 1..%lld", -src, (long long)d->lastidx_out);
    const long long nvout = d->nv * (long long)d->vdepth;
    if (o < nvout) {
      const long long row = o / d->vdepth, k = o % d->vdepth;
      if (e->gather_from_u) {          // vertex outputs are states (StateMask(1:vdepth))
        size_t b = 0;
        while (b + 1 < e->hvb.size() && row >= e->hvb[b + 1].row0) ++b;
        code = (int)(e->hvb[b].state0 + (row - e->hvb[b].row0) * e->hvb[b].dim + k);
      } else {
        code = (int)(row * d->vdepth + k) | ND_EXT_FROM_VOUT;
      }
      return ND_B200_OK;
    }
    for (const HostEB& h : e->heb) {
      const long long w = h.osrc + h.odst, lo = h.out0, hi = h.out0 + h.count * w;
      if (o < lo || o >= hi) continue;
      if (h.dim == 0) return fail(e, ND_B200_EUNSUPPORTED, "external input: outputs of feed-forward components (static edges) are not allowed (src/external_inputs.jl:42-44)");
      const long long i = (o - lo) / w, c = (o - lo) % w;
      const bool src_side = c < h.osrc;
      const long long comp = src_side ? c : c - h.osrc;
      const long long st = h.state0 + i * h.dim + ((src_side && h.coupling == ND_B200_FIDUCIAL) ? h.mask_src : h.mask_dst) + comp;
      code = (int)st;
      if (src_side && h.coupling == ND_B200_ANTISYMMETRIC) code |= (int)0x80000000u;
      return ND_B200_OK;
    }
    return fail(e, ND_B200_EINVAL, "external input: output index %lld belongs to no component", -src);
  }

  // external inputs of the batches (after both registrations: sources may be any vertex or edge)
  int resolve_externals() {
    vext_codes.assign((size_t)d->n_vbatches, {});
    eext_codes.assign((size_t)std::max(d->n_ebatches, 0), {});
    for (int b = 0; b < d->n_vbatches; ++b) any_ext = any_ext || d->vbatches[b].extdim > 0;
    for (int b = 0; b < d->n_ebatches; ++b) any_ext = any_ext || d->ebatches[b].extdim > 0;
    if (!any_ext) return ND_B200_OK;
    if (d->gather_offset) return fail(e, ND_B200_EUNSUPPORTED, "external inputs on a halo engine");   // row partitions with the complete u are fine
    if (d->lastidx_dynamic >= ND_EXT_FROM_VOUT || d->nv * (long long)d->vdepth >= ND_EXT_FROM_VOUT) return fail(e, ND_B200_EUNSUPPORTED, "networks with external inputs need offsets below 2^30");
    auto one = [&](int extdim, long long count, const int64_t* src, std::vector<int>& out, const char* what, int b) -> int {
      if (extdim <= 0) return ND_B200_OK;
      if (!src) return fail(e, ND_B200_EINVAL, "%s batch %d: extdim %d without ext_src", what, b + 1, extdim);
      out.resize((size_t)(count * extdim));
      for (size_t k = 0; k < out.size(); ++k)
        if (int rc = resolve_ext_source(src[k], out[k])) return rc;
      return ND_B200_OK;
    };
    for (int b = 0; b < d->n_vbatches; ++b)
      if (int rc = one(d->vbatches[b].extdim, d->vbatches[b].count, d->vbatches[b].ext_src, vext_codes[(size_t)b], "vertex", b)) return rc;
    for (int b = 0; b < d->n_ebatches; ++b)
      if (int rc = one(d->ebatches[b].extdim, d->ebatches[b].count, d->ebatches[b].ext_src, eext_codes[(size_t)b], "edge", b)) return rc;
    return ND_B200_OK;
  }

  // destination-sorted CSR over the owned rows in SequentialAggregator order (+ split-mode and get_buffers tables)
  int build_csr() {
    // ---- destination-sorted CSR over owned rows: count, then stable placement ---------------------
    cnt.assign((size_t)nrows_owned + 1, 0);
    for (int b = 0; b < d->n_ebatches; ++b) {
      const nd_b200_ebatch& eb = d->ebatches[b];
      for (long long i = 0; i < eb.count; ++i) {
        const long long eid = eb.indices ? eb.indices[i] - 1 : i;
        const long long s = d->edge_src[eid], t = d->edge_dst[eid];
        if (s < 1 || s > d->nv || t < 1 || t > d->nv) return fail(e, ND_B200_EINVAL, "edge %lld endpoint out of range", eid + 1);
        const int rs = row_of_vertex[(size_t)s - 1], rt = row_of_vertex[(size_t)t - 1];
        if ((eb.outdim_src > 0 || eb.kind == ND_B200_E_LOOPBACK) && owned(rs)) cnt[(size_t)(rs - e->row_begin) + 1]++;   // loopback: the injector's input entry
        if (owned(rt)) cnt[(size_t)(rt - e->row_begin) + 1]++;
      }
    }
    for (long long r = 0; r < nrows_owned; ++r) cnt[(size_t)r + 1] += cnt[(size_t)r];
    e->nentries = cnt[(size_t)nrows_owned];
    if (e->nentries >= INT_MAX) return fail(e, ND_B200_EUNSUPPORTED, "more than 2^31 directed entries on one device");
    h_rowptr.assign((size_t)nrows_owned + 1, 0);
    for (size_t r = 0; r < h_rowptr.size(); ++r) h_rowptr[r] = (int)cnt[r];
    keep = !(d->flags & ND_B200_FLAG_NO_EXPORT);
    h_nbr.assign((size_t)std::max<long long>(e->nentries, 1), 0);
    // the precompiled generic kernels are instantiated with PE = 1 and read the parameter-offset stream even when no edge
    // batch of this network has parameters
    if (e->ek == EK_GENERIC && d->vdepth == 1 && !e->custom) any_epar = true;
    if (any_epar) h_epar.assign((size_t)std::max<long long>(e->nentries, 1), 0);
    if (e->ek == EK_GENERIC && (d->vdepth == 1 || e->custom)) h_ebid.assign((size_t)std::max<long long>(e->nentries, 1), 0);
    if (keep) { e->h_nbr_vid.resize((size_t)e->nentries); e->h_eid.resize((size_t)e->nentries); e->h_side.resize((size_t)e->nentries); e->h_aggidx.resize((size_t)e->nentries); }
    // split mode tables: per entry its position in the edge part of `o`; per edge (in `o` order) the gather offsets
    generic_edges = (e->ek == EK_GENERIC && (d->vdepth == 1 || e->custom));
    e->oedge_base = d->nv * (long long)d->vdepth;
    e->oedge_len = d->lastidx_out - e->oedge_base;
    e->ne_all = d->ne;
    want_split = false;
    if (const char* s = getenv("ND_B200_KERNEL")) want_split = !strcmp(s, "split") && nrows_owned == e->nrows_total && !(d->vdepth == 2 && d->n_ebatches > 1) && !e->custom && !any_ode && !any_fiducial && !any_ext && !any_loopback;
    if (want_split && e->oedge_len >= INT_MAX) return fail(e, ND_B200_EUNSUPPORTED, "edge output buffer exceeds 2^31 scalars on one device");
    h_oidx.assign(want_split ? (size_t)std::max<long long>(e->nentries, 1) : 1, 0);
    h_es.assign(want_split ? (size_t)std::max<long long>(d->ne, 1) : 1, 0);
    h_et.assign(want_split ? (size_t)std::max<long long>(d->ne, 1) : 1, 0);
    if (generic_edges && want_split) { h_eepar.assign((size_t)std::max<long long>(d->ne, 1), 0); h_eooff.assign((size_t)std::max<long long>(d->ne, 1), 0); h_eebid.assign((size_t)std::max<long long>(d->ne, 1), 0); }
    {
      long long kedge = 0;
      std::vector<long long> cur(cnt.begin(), cnt.end() - 1);
      for (int b = 0; b < d->n_ebatches; ++b) {
        const nd_b200_ebatch& eb = d->ebatches[b];
        for (long long i = 0; i < eb.count; ++i) {
          const long long eid = eb.indices ? eb.indices[i] - 1 : i;
          const long long s = d->edge_src[eid], t = d->edge_dst[eid];
          const int rs = row_of_vertex[(size_t)s - 1], rt = row_of_vertex[(size_t)t - 1];
          const int ep = eb.pdim > 0 ? (int)(eb.p_first - 1 + i * eb.pdim) : 0;
          const long long oo = (eb.out_first - 1) + i * (eb.outdim_src + eb.outdim_dst) - e->oedge_base;   // this edge's block in the edge part of o
          if (want_split) { h_es[(size_t)kedge] = goff[(size_t)s - 1]; h_et[(size_t)kedge] = goff[(size_t)t - 1]; }
          if (generic_edges && want_split) { h_eepar[(size_t)kedge] = ep; h_eooff[(size_t)kedge] = (int)oo; h_eebid[(size_t)kedge] = (uint8_t)b; }
          ++kedge;
          // src output precedes dst output in `o` (register_edges!, src/network_structure.jl:244-245)
          // edges with states contribute a StateMask read of their own states: the offset of the output state inside u,
          // flagged with ND_STATE_ENTRY_BIT (state_entry_value in the kernels)
          const int so_src = eb.dim > 0 ? (int)((eb.state_first - 1) + i * eb.dim + (eb.coupling == ND_B200_FIDUCIAL ? eb.mask_src_first - 1 : eb.mask_dst_first - 1)) | ND_STATE_ENTRY_BIT : 0;
          const int so_dst = eb.dim > 0 ? (int)((eb.state_first - 1) + i * eb.dim + (eb.mask_dst_first - 1)) | ND_STATE_ENTRY_BIT : 0;
          if ((eb.outdim_src > 0 || eb.kind == ND_B200_E_LOOPBACK) && owned(rs)) {
            const long long j = cur[(size_t)(rs - e->row_begin)]++;
            if (want_split) h_oidx[(size_t)j] = (int)oo;
            h_nbr[(size_t)j] = eb.dim > 0 ? ~so_src : ~goff[(size_t)t - 1];
            if (any_epar) h_epar[(size_t)j] = ep;
            if (!h_ebid.empty()) h_ebid[(size_t)j] = (uint8_t)b;
            if (keep) { e->h_nbr_vid[(size_t)j] = (int)t; e->h_eid[(size_t)j] = (int)(eid + 1); e->h_side[(size_t)j] = 1; e->h_aggidx[(size_t)j] = eb.outdim_src > 0 ? oo + e->oedge_base : -1; }
          }
          if (owned(rt)) {
            const long long j = cur[(size_t)(rt - e->row_begin)]++;
            if (want_split) h_oidx[(size_t)j] = (int)(oo + eb.outdim_src);
            h_nbr[(size_t)j] = eb.dim > 0 ? so_dst : goff[(size_t)s - 1];
            if (any_epar) h_epar[(size_t)j] = ep;
            if (!h_ebid.empty()) h_ebid[(size_t)j] = (uint8_t)b;
            if (keep) { e->h_nbr_vid[(size_t)j] = (int)s; e->h_eid[(size_t)j] = (int)(eid + 1); e->h_side[(size_t)j] = 0; e->h_aggidx[(size_t)j] = oo + e->oedge_base + eb.outdim_src; }
          }
        }
      }
    }
    if (keep) e->h_rowptr.assign(cnt.begin(), cnt.end());
    // rows that read the halo ("boundary" rows); everything else can run while the halo is in flight
    row_remote.assign((size_t)nrows_owned, 0);
    if (d->gather_offset) {
      for (long long r = 0; r < nrows_owned; ++r)
        for (long long j = cnt[(size_t)r]; j < cnt[(size_t)r + 1]; ++j) {
          const int o = h_nbr[(size_t)j] < 0 ? ~h_nbr[(size_t)j] : h_nbr[(size_t)j];
          if (o >= e->halo_base) { row_remote[(size_t)r] = 1; break; }
        }
    }

    // get_buffers tables: gather offsets per edge in batch order
    if (keep) {
      e->h_esrc_off.resize((size_t)d->n_ebatches); e->h_edst_off.resize((size_t)d->n_ebatches);
      for (int b = 0; b < d->n_ebatches; ++b) {
        const nd_b200_ebatch& eb = d->ebatches[b];
        e->h_esrc_off[(size_t)b].resize((size_t)eb.count); e->h_edst_off[(size_t)b].resize((size_t)eb.count);
        for (long long i = 0; i < eb.count; ++i) {
          const long long eid = eb.indices ? eb.indices[i] - 1 : i;
          e->h_esrc_off[(size_t)b][(size_t)i] = goff[(size_t)d->edge_src[eid] - 1];
          e->h_edst_off[(size_t)b][(size_t)i] = goff[(size_t)d->edge_dst[eid] - 1];
        }
      }
    }
    return ND_B200_OK;
  }

  // launch shape, thread-block row ranges, one descriptor per thread block (tile kernel)
  int plan_tiles() {
    // ---- launch shape + thread-block row ranges ----------------------------------------------------
    e->block = 128; e->ept = 4;   // measured best on B200 for every registry family (profiles/r01_tuning.md)
    if (!e->custom) {             // run-time compiled kernels exist for the default launch shape only
      if (const char* s = getenv("ND_B200_BLOCK")) e->block = atoi(s);
      if (const char* s = getenv("ND_B200_EPT")) e->ept = atoi(s);
    }
    if (!((e->block == 256 || e->block == 128) && (e->ept == 8 || e->ept == 4))) return fail(e, ND_B200_EINVAL, "ND_B200_BLOCK/ND_B200_EPT must be 128|256 / 4|8");
    const int tile = e->block * e->ept;
    e->n_long = 0;
    for (size_t b = 0; b < e->hvb.size(); ++b) {
      const HostVB& h = e->hvb[b];
      VBDev v{h.kind, h.dim, h.pdim, (int)h.row0, (int)h.count, (int)blk_row.size(), h.state0, h.p0, nullptr, d->vbatches[b].extdim, h.ff, nullptr};
      long long r = std::max<long long>(h.row0, e->row_begin);
      const long long rend = std::min<long long>(h.row0 + h.count, e->row_end);
      while (r < rend) {
        blk_row.push_back((int)r);
        const long long deg0 = cnt[(size_t)(r - e->row_begin) + 1] - cnt[(size_t)(r - e->row_begin)];
        if (deg0 > e->long_thr) { e->n_long++; r++; continue; }
        long long rr = r, ents = 0;
        while (rr < rend && rr - r < e->block) {
          const long long deg = cnt[(size_t)(rr - e->row_begin) + 1] - cnt[(size_t)(rr - e->row_begin)];
          if (deg > e->long_thr || ents + deg > tile) break;
          ents += deg; rr++;
        }
        if (rr == r) {   // a single short row that does not fit a tile: only when long_thr >= tile
          return fail(e, ND_B200_EUNSUPPORTED, "row with %lld entries exceeds the %d-entry tile with long rows disabled", deg0, tile);
        }
        r = rr;
      }
      dvb.push_back(v);
    }
    e->nblocks = (int)blk_row.size();
    blk_row.push_back((int)e->row_end);

    // owned state ranges: one per vertex batch the row range intersects
    for (const HostVB& h : e->hvb) {
      const long long lo = std::max<long long>(e->row_begin, h.row0), hi = std::min<long long>(e->row_end, h.row0 + h.count);
      if (lo < hi && h.dim > 0) e->own_segs.push_back({h.state0 + (lo - h.row0) * h.dim, (hi - lo) * h.dim});
    }
    // ---- one 16-byte descriptor per thread block -----------------------------------------------------------
    e->split = 0;   // default: fused kernel (faster on B200 for every config whose state vector fits in L2)
    if (want_split) e->split = 1;
    {
      tiles.reserve((size_t)e->nblocks);
      size_t bi = 0;
      for (int k = 0; k < e->nblocks; ++k) {
        const int r0 = blk_row[(size_t)k], r1 = blk_row[(size_t)k + 1];
        while (bi + 1 < dvb.size() && k >= dvb[bi + 1].blk0) ++bi;
        const long long a = cnt[(size_t)(r0 - e->row_begin)], z = cnt[(size_t)(r1 - e->row_begin)];
        const long long ne = z - a;
        const bool is_long = (r1 - r0 == 1) && ne > e->long_thr;
        int4 t;
        t.x = r0; t.y = (int)a;
        if (is_long) { t.z = (int)ne; t.w = (int)(0x80000000u | ((unsigned)bi << 25) | (1u << 16)); }
        else { t.z = 0; t.w = (int)((unsigned)ne | ((unsigned)(r1 - r0) << 16) | ((unsigned)bi << 25)); }
        tiles.push_back(t);
      }
      e->ntiles = (int)tiles.size();
      e->wait_from = 0;
      if (d->gather_offset) {
        // interior tiles first, tiles that read the halo last (each descriptor is self-contained)
        auto reads_halo = [&](const int4& t) {
          const int nr = (t.w < 0) ? 1 : ((t.w >> 16) & 0x1FF);
          for (int r = 0; r < nr; ++r)
            if (row_remote[(size_t)(t.x + r - e->row_begin)]) return true;
          return false;
        };
        auto mid = std::stable_partition(tiles.begin(), tiles.end(), [&](const int4& t) { return !reads_halo(t); });
        e->wait_from = (int)(mid - tiles.begin());
      }
      e->blk_pmax.assign(tiles.size(), 0); e->blk_rmin.assign(tiles.size(), 0); e->blk_rmax.assign(tiles.size(), 0);
      for (size_t k = 0; k < tiles.size(); ++k) {
        const int4& t = tiles[k];
        const bool lg = t.w < 0;
        const int nr = lg ? 1 : ((t.w >> 16) & 0x1FF), ne = lg ? t.z : (t.w & 0xFFFF);
        const size_t b = (size_t)((t.w >> 25) & 0x3F);
        int pm = row_pend(t.x + nr - 1, b);
        for (long long j = t.y; j < (long long)t.y + ne; ++j) pm = std::max(pm, entry_pend(j));
        e->blk_pmax[k] = pm; e->blk_rmin[k] = t.x; e->blk_rmax[k] = t.x + nr - 1;
      }
    }
    for (const HostEB& h : e->heb) deb.push_back(EBDev{h.kind, h.coupling, h.pdim, h.dim});
    // compact entry words for the specialised tile kernels: offset | local row << 23 | side << 30
    {
      const long long max_off = std::max<long long>(e->gather_from_u ? d->lastidx_dynamic : d->nv * (long long)d->vdepth, d->gather_offset ? d->gather_len : 0);
      const bool specialised = !e->custom && !e->split && (d->vdepth == 2 ? true : e->ek != EK_GENERIC);
      e->compact = specialised && e->block == 128 && e->ept == 4 && max_off < (1LL << ND_CR_OFF_BITS) && !getenv("ND_B200_NO_COMPACT");
      if (e->compact) {
        h_nbr_c.assign(h_nbr.size(), 0);
        for (const int4& t : tiles) {
          const bool lg = t.w < 0;
          const int nr = lg ? 1 : ((t.w >> 16) & 0x1FF);
          for (int r = 0; r < nr; ++r) {
            const long long row = (long long)t.x + r - e->row_begin;
            for (long long j = cnt[(size_t)row]; j < cnt[(size_t)row + 1]; ++j) {
              const int w = h_nbr[(size_t)j];
              const int side = w < 0, off = side ? ~w : w;
              h_nbr_c[(size_t)j] = off | (r << ND_CR_OFF_BITS) | (side << 30);
            }
          }
        }
      }
    }
    return ND_B200_OK;
  }

  // kernel family choice + the jagged (warp-slice) layout
  int build_jagged() {
    // ---- jagged layout: 32-lane slices, column-major compacted entries (rhs_jag_kernel) ------------------------
    // ND_B200_KERNEL=jag|fused|split overrides the automatic choice.  A strictly sequential long_row_threshold beyond
    // what one lane can hold (63 entries) needs the tile kernel.
    // Kernel family (measured on B200, profiles/r01c_sweep_fused_vs_jag.jsonl): the tile kernel wins whenever degrees
    // vary (idle lanes in the jagged walk: ER cfg2 72 vs 76 us, BA cfg3 91 vs 148 us) or the graph is small (latency of the
    // per-lane walk: cfg1); the jagged kernel wins on large regular-degree graphs (cfg4 grid: RK4 step 45 vs 55 us).
    // auto = jagged iff lane utilisation of the walk >= 0.8 and there are enough rows to fill the machine.
    int auto_window = 32;
    {
      long long sum_max = 0;
      for (long long r = 0; r < nrows_owned; r += 32) {
        long long m = 0;
        for (long long q = r; q < std::min<long long>(r + 32, nrows_owned); ++q) m = std::max(m, cnt[(size_t)q + 1] - cnt[(size_t)q]);
        sum_max += m;
      }
      const double util = sum_max > 0 ? (double)e->nentries / (32.0 * (double)sum_max) : 0.0;
      e->jag = (util >= 0.8 && nrows_owned >= 65536) ? 1 : 0;
      // Irregular graphs WITHOUT hubs (Erdos-Renyi: configs 2 and 5): degree-bucketed slices over windows of 128 rows.
      // Measured on B200 (profiles/r02d_sweep_jag_pipelined.jsonl, config 2): same speed as the tile kernel with live
      // parameters (72.7 vs 72.1 us) and faster once the edge parameters are packed (58.4 vs 65 us); with power-law hubs
      // (config 3) the tile kernel stays ahead (95 vs 122-137 us).
      if (!e->jag && nrows_owned >= 65536 && d->vdepth == 1 && !e->custom && e->ek != EK_GENERIC) {
        long long maxdeg = 0, sum128 = 0;
        for (long long r = 0; r < nrows_owned; ++r) maxdeg = std::max(maxdeg, cnt[(size_t)r + 1] - cnt[(size_t)r]);
        if (maxdeg <= 64) {
          std::vector<long long> deg;
          for (long long w0 = 0; w0 < nrows_owned; w0 += 128) {
            deg.clear();
            for (long long q = w0; q < std::min<long long>(w0 + 128, nrows_owned); ++q) deg.push_back(cnt[(size_t)q + 1] - cnt[(size_t)q]);
            std::sort(deg.begin(), deg.end(), std::greater<long long>());
            for (size_t k = 0; k < deg.size(); k += 32) sum128 += deg[k];     // longest lane of each slice
          }
          const double util128 = sum128 > 0 ? (double)e->nentries / (32.0 * (double)sum128) : 0.0;
          if (util128 >= 0.7) { e->jag = 1; auto_window = 128; }
        }
      }
    }
    // streamed jagged kernel (rhs_js_kernel): single-batch benchmark edge kinds, one vertex output, no halo layout
    const bool js_ok = !e->custom && d->vdepth == 1 && e->edepth == 1 && e->gather_from_u && !d->gather_offset && !any_ode &&
                       (e->ek == ND_B200_E_DIFFUSION || e->ek == ND_B200_E_DIFFUSION_NOP || e->ek == ND_B200_E_KURAMOTO) && e->c_maxdim <= ND_MAX_VDIM;
    e->jstream = 0;
    // asynchronous-gather jagged kernel (rhs_jaga_kernel): single-batch benchmark edge kinds, one vertex output
    const bool jaga_ok = !e->custom && d->vdepth == 1 && e->edepth == 1 && !any_ode &&
                         (e->ek == ND_B200_E_DIFFUSION || e->ek == ND_B200_E_DIFFUSION_NOP || e->ek == ND_B200_E_KURAMOTO) && e->c_maxdim <= ND_MAX_VDIM;
    e->jaga = 0; e->jagb = 0;
    if (const char* s = getenv("ND_B200_KERNEL")) {
      if (!strcmp(s, "jag")) e->jag = 1;
      else if (!strcmp(s, "jaga") && jaga_ok) { e->jag = 1; e->jaga = 1; }
      else if (!strcmp(s, "jagb") && jaga_ok) { e->jag = 1; e->jagb = 1; }
      else if (!strcmp(s, "js") && js_ok) { e->jag = 1; e->jstream = 1; }
      else if (!strcmp(s, "fused") || !strcmp(s, "split") || !strcmp(s, "v1")) e->jag = 0;
    }
    if (e->split) { e->jag = 0; e->jstream = 0; }
    if (d->long_row_threshold > 63 * 32) { e->jag = 0; e->jstream = 0; }
    for (const HostVB& h : e->hvb) if (h.pdim > 4) e->jstream = 0;      // the kernel prefetches up to 4 vertex parameters
    if (!e->jag) { e->jstream = 0; e->jaga = 0; e->jagb = 0; }
    if (const char* s = getenv("ND_B200_JAG_PERSIST")) e->jag_persist = atoi(s) > 0;
    if (const char* s = getenv("ND_B200_JAG_BLOCK")) e->jag_block = atoi(s) == 64 ? 64 : 128;
    // L2 prefetch of the entry streams, one cp.async.bulk.prefetch.L2 per slice and stream, 600 K entries (about half a wave of
    // resident warps) ahead: measured 1.5-3 % on the large Erdos-Renyi configs (profiles/r02l_sweep_l2_prefetch.jsonl)
    e->pf_dist = (e->jag && !e->custom && e->nentries >= 2000000) ? 600000 : 0;
    if (const char* s = getenv("ND_B200_PF_DIST")) e->pf_dist = std::max(0, atoi(s));
    if (const char* s = getenv("ND_B200_JAGA_CH")) e->jaga_ch = atoi(s);
    if (const char* s = getenv("ND_B200_JAGA_WPS")) e->jaga_wps = atoi(s);
    e->jag_u = 2;
    e->jag_wps = 32;   // spill-free register budget of the software-pipelined walk, best measured (profiles/r02d)
    jag_pe = any_epar || (generic_edges && !e->custom);   // kernels instantiated with PE > 0 read {nbr, epar} pairs
    if (e->jag) {
      e->jsplit = 32;
      if (const char* s = getenv("ND_B200_JAG_SPLIT")) e->jsplit = std::min(63, std::max(1, atoi(s)));
      if (const char* s = getenv("ND_B200_JAG_U")) e->jag_u = atoi(s);
      if (const char* s = getenv("ND_B200_JAG_WPS")) e->jag_wps = atoi(s);
      int jwindow = e->jstream ? 128 : auto_window;
      if (const char* s = getenv("ND_B200_JAG_WINDOW")) { const int w = atoi(s); if (w == 32 || w == 64 || w == 128) jwindow = w; }
      if (e->jstream) {
        if (const char* s = getenv("ND_B200_JS_U")) e->js_u = atoi(s) >= 8 ? 8 : 4;
        if (const char* s = getenv("ND_B200_JS_NST")) e->js_nst = atoi(s) >= 8 ? 8 : 4;
        // resident warps per SM of the chosen instantiation (launch_js_shape: MINB blocks of 4 warps)
        e->js_wps = e->js_u >= 8 ? (e->js_nst >= 8 ? 16 : 24) : (e->js_nst >= 8 ? 16 : 32);
        if (const char* s = getenv("ND_B200_JS_WPS")) e->js_wps = std::max(1, std::min(64, atoi(s)));
      }
      // rows longer than this are reduced by a whole block: explicit threshold if the caller gave one, else what a
      // slice can hold
      const long long block_thr = d->long_row_threshold > 0 ? std::min<long long>(d->long_row_threshold, 32LL * e->jsplit) : 32LL * e->jsplit;
      // window mode (rhs_jag_kernel<..., WIN>): 128-row windows, one vertex output, registry kinds; ND_B200_JAG_WIN=0 disables
      // Opt-in (ND_B200_JAG_WIN=1): measured on config 2 (profiles/r02m_sweep_window_mode.jsonl) it moves 0.7 M fewer sectors
      // through the L1 but the two block barriers cost more than that (66.8 vs 59.0 us at 32 warps per SM, 57.5 at 48).
      const bool pad_windows = jwindow == 128 && !e->jstream && d->vdepth == 1 && e->edepth == 1 && !e->custom && !any_ode &&
                               getenv("ND_B200_JAG_WIN") && atoi(getenv("ND_B200_JAG_WIN")) > 0;
      e->jag_win = pad_windows ? 1 : 0;
      std::vector<int> order;
      order.reserve((size_t)e->nentries);
      struct Lane { int rowrel, len, head; long long start; };
      std::vector<Lane> lanes;
      std::vector<std::pair<long long, int>> long_rows;   // (row, batch)
      int jag_wait_from = 0;
      const int nclasses = d->gather_offset ? 2 : 1;   // class 0: interior rows, class 1: rows that read the halo
      for (int cls = 0; cls < nclasses; ++cls) {
      if (cls == 1) jag_wait_from = (int)jslices.size();
      for (size_t b = 0; b < e->hvb.size(); ++b) {
        const HostVB& h = e->hvb[b];
        const long long lo = std::max<long long>(h.row0, e->row_begin), hi = std::min<long long>(h.row0 + h.count, e->row_end);
        long long row0 = -1;
        int maxparts = 1;
        auto flush = [&]() {
          if (lanes.empty()) return;
          int maxlen = 0;
          for (const Lane& L : lanes) maxlen = std::max(maxlen, L.len);
          const int e0 = (int)order.size();
          for (int j = 0; j < maxlen; ++j)
            for (const Lane& L : lanes)
              if (L.len > j) order.push_back((int)(L.start + j));
          if (e->jstream) while (order.size() & 3) order.push_back(-1);   // 16-byte granularity of the bulk copies
          jslices.push_back(make_int4(e0, (int)row0, (int)b, maxparts));
          for (int l = 0; l < 32; ++l) {
            uint16_t v = 0;
            if (l < (int)lanes.size()) v = (uint16_t)(lanes[(size_t)l].len | (lanes[(size_t)l].rowrel << 6) | (lanes[(size_t)l].head << 13) | (1 << 14));
            jlanes.push_back(v);
          }
          lanes.clear(); row0 = -1; maxparts = 1;
        };
        if (jwindow > 32) {
          // degree-bucketed slices (ND_B200_JAG_WINDOW = 64 | 128): the rows of a window of consecutive rows are dealt to
          // the lanes in order of decreasing degree, so the 32 rows that share a slice have (nearly) equal length and the
          // lane walk wastes no iterations on short rows next to long ones; lanes address their row relative to the window
          // start (7 bits).  The row's own u / du / vertex parameters stay within the window (<= 1 KB of each vector).
          std::vector<long long> wrows;
          for (long long w0 = lo; w0 < hi; w0 += jwindow) {
            wrows.clear();
            for (long long r = w0; r < std::min<long long>(w0 + jwindow, hi); ++r) {
              if (nclasses == 2 && row_remote[(size_t)(r - e->row_begin)] != cls) continue;
              const long long deg = cnt[(size_t)(r - e->row_begin) + 1] - cnt[(size_t)(r - e->row_begin)];
              const int nparts = (int)std::max<long long>(1, (deg + e->jsplit - 1) / e->jsplit);
              if (deg > block_thr || nparts > 32) { long_rows.push_back({r, (int)b}); continue; }
              wrows.push_back(r);
            }
            std::stable_sort(wrows.begin(), wrows.end(), [&](long long x, long long y) {
              return cnt[(size_t)(x - e->row_begin) + 1] - cnt[(size_t)(x - e->row_begin)] > cnt[(size_t)(y - e->row_begin) + 1] - cnt[(size_t)(y - e->row_begin)];
            });
            for (long long r : wrows) {
              const long long a = cnt[(size_t)(r - e->row_begin)], deg = cnt[(size_t)(r - e->row_begin) + 1] - a;
              const int nparts = (int)std::max<long long>(1, (deg + e->jsplit - 1) / e->jsplit);
              if ((int)lanes.size() + nparts > 32) flush();
              if (lanes.empty()) row0 = w0;
              for (int k = 0; k < nparts; ++k) {
                const long long len = std::min<long long>(e->jsplit, deg - (long long)k * e->jsplit);
                lanes.push_back(Lane{(int)(r - w0), (int)std::max<long long>(len, 0), k == 0, a + (long long)k * e->jsplit});
              }
              maxparts = std::max(maxparts, nparts);
            }
            flush();
            if (pad_windows) {   // window mode: a thread block (four slices) never spans two windows
              while (jslices.size() & 3) {
                jslices.push_back(make_int4((int)order.size(), (int)w0, (int)b, 1));
                for (int l = 0; l < 32; ++l) jlanes.push_back(0);
              }
            }
          }
        } else {
        for (long long r = lo; r < hi; ++r) {
          if (nclasses == 2 && row_remote[(size_t)(r - e->row_begin)] != cls) continue;
          const long long a = cnt[(size_t)(r - e->row_begin)], deg = cnt[(size_t)(r - e->row_begin) + 1] - a;
          const int nparts = (int)std::max<long long>(1, (deg + e->jsplit - 1) / e->jsplit);
          if (deg > block_thr || nparts > 32) { long_rows.push_back({r, (int)b}); continue; }
          if ((int)lanes.size() + nparts > 32 || (row0 >= 0 && r - row0 >= 32)) flush();
          if (lanes.empty()) row0 = r;
          for (int k = 0; k < nparts; ++k) {
            const long long len = std::min<long long>(e->jsplit, deg - (long long)k * e->jsplit);
            lanes.push_back(Lane{(int)(r - row0), (int)std::max<long long>(len, 0), k == 0, a + (long long)k * e->jsplit});
          }
          maxparts = std::max(maxparts, nparts);
        }
        flush();
        }
      }
      }
      if (nclasses == 1) jag_wait_from = 0;
      for (const auto& lr : long_rows) {
        const long long a = cnt[(size_t)(lr.first - e->row_begin)], deg = cnt[(size_t)(lr.first - e->row_begin) + 1] - a;
        jlong.push_back(make_int4((int)order.size(), (int)lr.first, (int)deg, lr.second));
        for (long long j = 0; j < deg; ++j) order.push_back((int)(a + j));
      }
      {
        long long real = 0;
        for (int o : order) real += o >= 0;
        if (real != e->nentries) return fail(e, ND_B200_EINVAL, "internal: jagged layout holds %lld of %lld entries", real, e->nentries);
      }
      e->jag_len = (long long)order.size();
      // padding slots (-1, streamed layout only) are never addressed by a lane; they hold offset 0
      if (jag_pe) {
        jent.resize(std::max<size_t>(order.size(), 1));
        for (size_t k = 0; k < order.size(); ++k) jent[k] = order[k] < 0 ? make_int2(0, 0) : make_int2(h_nbr[(size_t)order[k]], any_epar ? h_epar[(size_t)order[k]] : 0);
      } else {
        jnbr.resize(std::max<size_t>(order.size(), 1));
        for (size_t k = 0; k < order.size(); ++k) jnbr[k] = order[k] < 0 ? 0 : h_nbr[(size_t)order[k]];
      }
      if (!h_ebid.empty()) {
        jebid.resize(std::max<size_t>(order.size(), 1));
        for (size_t k = 0; k < order.size(); ++k) jebid[k] = order[k] < 0 ? 0 : h_ebid[(size_t)order[k]];
      }
      e->host_only = (d->flags & ND_B200_FLAG_HOST_ONLY) != 0;
      if (e->host_only) { e->h_jslices = jslices; e->h_jlong = jlong; e->h_jlanes = jlanes; e->h_jorder = order; }
      {
        const size_t nb = (jslices.size() + 3) / 4;
        e->blk_pmax.assign(nb + jlong.size(), 0); e->blk_rmin.assign(nb + jlong.size(), INT_MAX); e->blk_rmax.assign(nb + jlong.size(), -1);
        for (size_t sidx = 0; sidx < jslices.size(); ++sidx) {
          const size_t k = sidx / 4;
          const int4& S = jslices[sidx];
          const long long eend = sidx + 1 < jslices.size() ? jslices[sidx + 1].x : (jlong.empty() ? (long long)order.size() : jlong[0].x);
          int pm = e->blk_pmax[k];
          for (long long q = S.x; q < eend; ++q) if (order[(size_t)q] >= 0) pm = std::max(pm, entry_pend(order[(size_t)q]));
          for (int l = 0; l < 32; ++l) {
            const uint16_t v = jlanes[sidx * 32 + (size_t)l];
            if (!((v >> 14) & 1)) break;
            const int r = S.y + ((v >> 6) & 127);
            e->blk_rmin[k] = std::min(e->blk_rmin[k], r); e->blk_rmax[k] = std::max(e->blk_rmax[k], r);
            pm = std::max(pm, row_pend(r, (size_t)S.z));
          }
          e->blk_pmax[k] = pm;
        }
        for (size_t q = 0; q < jlong.size(); ++q) {
          const int4& Lr = jlong[q];
          int pm = row_pend(Lr.y, (size_t)Lr.w);
          for (long long j = Lr.x; j < (long long)Lr.x + Lr.z; ++j) pm = std::max(pm, entry_pend(order[(size_t)j]));
          e->blk_pmax[nb + q] = pm; e->blk_rmin[nb + q] = Lr.y; e->blk_rmax[nb + q] = Lr.y;
        }
      }
      e->wait_from = jag_wait_from;
      e->nslices = (int)jslices.size();
      e->n_jag_blocks = (e->nslices + 3) / 4;   // BLOCK = 128: four slices per thread block
      e->n_jlong = (int)jlong.size();
      if (e->jstream) {
        // persistent warps: js_wps warps on each SM, every warp owns a contiguous range of slices = one contiguous piece
        // of the entry stream.  Ranges are balanced by cost (entries + a fixed share per slice for descriptors, own data
        // and the vertex phase).  The slice table ends with a sentinel carrying the end of the last slice's entries.
        const long long slice_end = jlong.empty() ? (long long)order.size() : (long long)jlong[0].x;
        int nsm = 148;
#ifndef ND_CUSIM
        if (!e->host_only && !(d->flags & ND_B200_FLAG_HOST_ONLY)) {
          int v = 0;
          if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, e->device) == cudaSuccess && v > 0) nsm = v;
        }
#endif
        const int nw = std::max(1, std::min(nsm * e->js_wps, e->nslices));
        const long long fixed = 48;
        long long total = 0;
        for (int k = 0; k < e->nslices; ++k) total += ((k + 1 < e->nslices ? jslices[(size_t)k + 1].x : slice_end) - jslices[(size_t)k].x) + fixed;
        e->h_jwarp.clear();
        long long accum = 0;
        int k0 = 0;
        for (int wi = 0; wi < nw && e->nslices > 0; ++wi) {
          const long long target = total * (wi + 1) / nw;
          int k1 = k0;
          while (k1 < e->nslices && (accum < target || k1 == k0) && (e->nslices - k1) > (nw - 1 - wi)) {
            accum += ((k1 + 1 < e->nslices ? jslices[(size_t)k1 + 1].x : slice_end) - jslices[(size_t)k1].x) + fixed;
            ++k1;
          }
          if (wi == nw - 1) k1 = e->nslices;
          e->h_jwarp.push_back(make_int2(k0, k1));
          k0 = k1;
        }
        e->n_jwarps = (int)e->h_jwarp.size();
        jslices.push_back(make_int4((int)slice_end, 0, 0, 1));
      }
      e->nblocks = e->n_jag_blocks + e->n_jlong;
      e->n_long = e->n_jlong;
    }

    e->blk_rows_monotone = true;
    for (size_t k = 0; k + 1 < e->blk_rmin.size(); ++k)
      if (e->blk_rmin[k + 1] <= e->blk_rmax[k]) { e->blk_rows_monotone = false; break; }
    return ND_B200_OK;
  }

  // run-time compilation of user-supplied kinds, uploads
  int finish() {
    if (d->flags & ND_B200_FLAG_HOST_ONLY) e->host_only = true;
    if (e->custom) {
      if (e->jag) { e->jag_wps = 48; e->jag_u = 2; }   // the one jagged instantiation that is compiled for user-supplied kinds
      if (int rc = compile_custom(e, d->vdepth, e->edepth)) return rc;
    }
    if (e->host_only) return ND_B200_OK;
    CUDA_TRY(e, cudaSetDevice(e->device));
    e->d_ffin.assign((size_t)d->n_vbatches, nullptr);
    for (int b = 0; b < d->n_vbatches; ++b) {
      if (!e->hvb[(size_t)b].ff) continue;
      std::vector<int> ffin((size_t)d->vbatches[b].count);
      for (long long i = 0; i < d->vbatches[b].count; ++i) {
        const long long vid = d->vbatches[b].indices ? d->vbatches[b].indices[i] : i + 1;
        ffin[(size_t)i] = goff[(size_t)hub_of_vertex[(size_t)vid - 1] - 1];     // offset of the hub's output in the vertex-output block
      }
      if (upload(e, &e->d_ffin[(size_t)b], ffin)) return ND_B200_ECUDA;
      dvb[(size_t)b].ffin = e->d_ffin[(size_t)b];
    }
    e->d_vext.assign((size_t)d->n_vbatches, nullptr);
    for (int b = 0; b < d->n_vbatches; ++b) {
      if (d->vbatches[b].extdim <= 0) continue;
      if (upload(e, &e->d_vext[(size_t)b], vext_codes[(size_t)b])) return ND_B200_ECUDA;
      dvb[(size_t)b].ext = e->d_vext[(size_t)b];
    }
    for (int b = 0; b < d->n_ebatches; ++b) {
      const nd_b200_ebatch& eb = d->ebatches[b];
      if (eb.dim == 0) continue;
      std::vector<int> es((size_t)eb.count), et((size_t)eb.count);
      for (long long i = 0; i < eb.count; ++i) {
        const long long eid = eb.indices ? eb.indices[i] - 1 : i;
        es[(size_t)i] = goff[(size_t)d->edge_src[eid] - 1];
        et[(size_t)i] = goff[(size_t)d->edge_dst[eid] - 1];
      }
      nd_b200_engine::OdeBatch ob{b, nullptr, nullptr, nullptr, eb.extdim};
      if (upload(e, &ob.d_es, es) || upload(e, &ob.d_et, et)) return ND_B200_ECUDA;
      if (eb.extdim > 0 && upload(e, &ob.d_ext, eext_codes[(size_t)b])) return ND_B200_ECUDA;
      e->ode.push_back(ob);
    }
    if (e->jag) {
      if (upload(e, &e->d_vb, dvb) || upload(e, &e->d_eb, deb) || upload(e, &e->d_jslices, jslices) || upload(e, &e->d_jlanes, jlanes) ||
          upload(e, &e->d_jlong, jlong))
        return ND_B200_ECUDA;
      if (jag_pe ? upload(e, &e->d_jent, jent) : upload(e, &e->d_jnbr, jnbr)) return ND_B200_ECUDA;
      if (!jebid.empty() && upload(e, &e->d_jebid, jebid)) return ND_B200_ECUDA;
      if (e->jstream) {
        if (upload(e, &e->d_jwarp, e->h_jwarp)) return ND_B200_ECUDA;
        KParams P0;
        fill_params(e, P0);
        CUDA_TRY(e, launch_js(e, P0, nullptr, true));
      }
      if (!e->gather_from_u) {
        for (int k = 0; k < 2; ++k) CUDA_TRY(e, cudaMalloc((void**)&e->d_vout[k], sizeof(double) * (size_t)(e->nrows_total * e->vdepth)));
      }
      return ND_B200_OK;
    }
    if (upload(e, &e->d_rowptr, h_rowptr) || upload(e, &e->d_nbr, e->compact ? h_nbr_c : h_nbr) || upload(e, &e->d_blk_row, blk_row) ||
        upload(e, &e->d_vb, dvb) || upload(e, &e->d_eb, deb))
      return ND_B200_ECUDA;
    if (any_epar && upload(e, &e->d_epar, h_epar)) return ND_B200_ECUDA;
    if (!h_ebid.empty() && upload(e, &e->d_ebid, h_ebid)) return ND_B200_ECUDA;
    if (!e->gather_from_u) {
      for (int k = 0; k < 2; ++k) CUDA_TRY(e, cudaMalloc((void**)&e->d_vout[k], sizeof(double) * (size_t)(e->nrows_total * e->vdepth)));
    }
    if (upload(e, &e->d_tiles, tiles)) return ND_B200_ECUDA;
    if (e->split) {
      if (upload(e, &e->d_oidx, h_oidx) || upload(e, &e->d_es, h_es) || upload(e, &e->d_et, h_et)) return ND_B200_ECUDA;
      if (generic_edges && (upload(e, &e->d_eepar, h_eepar) || upload(e, &e->d_eooff, h_eooff) || upload(e, &e->d_eebid, h_eebid))) return ND_B200_ECUDA;
      CUDA_TRY(e, cudaMalloc((void**)&e->d_oedge, sizeof(double) * (size_t)std::max<long long>(e->oedge_len, 2)));
      // the fused kernel's per-entry arrays are not needed
      cudaFree(e->d_nbr); cudaFree(e->d_epar); cudaFree(e->d_ebid);
      e->d_nbr = nullptr; e->d_epar = nullptr; e->d_ebid = nullptr;
    }
    return ND_B200_OK;
    return ND_B200_OK;
  }

  int run() {
    if (int rc = check_descriptor()) return rc;
    if (int rc = register_vertices()) return rc;
    if (int rc = register_edges()) return rc;
    if (int rc = resolve_externals()) return rc;
    if (int rc = build_csr()) return rc;
    if (int rc = plan_tiles()) return rc;
    if (int rc = build_jagged()) return rc;
    return finish();
  }
};

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
  if (e->host_only) { delete e; return; }
  cudaSetDevice(e->device);
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
    for (int s = 0; s < per_graph && rc == ND_B200_OK; ++s) rc = rk4_step_enqueue(e, u, p, t0 + s * dt, dt, e->cap_stream);
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

// ---- multi-GPU exchange object ------------------------------------------------------------------------------------------
// Packed halo over NVLink peer memory.  Every rank owns ONE device allocation [halo parity 0 | halo parity 1 | flags]
// that the peers map through CUDA IPC.  halo = the outputs of exactly the remote vertices this rank's rows read, grouped
// by owner rank (ascending) and sorted by state offset inside a group; the engine was built with gather offsets that
// point into it (nd_b200_desc.gather_offset).
struct nd_b200_comm {
  int device = 0, rank = 0, world = 1;
  long long halo_len = 0;                 // doubles in this rank's halo buffer
  size_t halo_bytes = 0;                  // bytes of ONE parity of the LARGEST halo among the ranks (same layout everywhere), 256-aligned
  unsigned char* base[HALO_MAX_WORLD] = {nullptr};   // [r]: rank r's shared block (own: cudaMalloc, peers: IPC mapping)
  int* d_send_idx[HALO_MAX_WORLD] = {nullptr};       // [r]: state offsets of the outputs rank r reads from this rank
  long long send_n[HALO_MAX_WORLD] = {0};
  long long dst_off[HALO_MAX_WORLD] = {0};           // [r]: where this rank's block starts inside rank r's halo
  unsigned int* d_done = nullptr;
  int* d_timeout = nullptr;
  int* h_timeout = nullptr;               // the sticky time-out mark again, in pinned (mapped) host memory: readable without a sync
  long long timeout_clocks = 60000000000LL;   // spin budget of a waiting tile, clock64 ticks (~30 s; ND_B200_HALO_TIMEOUT_MS)
  unsigned long long seq = 0;
  std::string err;
  double* halo(int r, int parity) const { return reinterpret_cast<double*>(base[r] + (size_t)parity * halo_bytes); }
  unsigned long long* flags(int r) const { return reinterpret_cast<unsigned long long*>(base[r] + 2 * halo_bytes); }
};

namespace {
int cfail(nd_b200_comm* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}
#define COMM_TRY(c, call)                                                                        \
  do {                                                                                           \
    cudaError_t _c = (call);                                                                     \
    if (_c != cudaSuccess) return cfail(c, ND_B200_ECUDA, "%s: %s", #call, cudaGetErrorString(_c)); \
  } while (0)
}  // namespace

extern "C" {

int nd_b200_comm_create(int32_t device, int32_t rank, int32_t world, int64_t halo_len, int64_t max_halo_len,
                        nd_b200_comm** out) {
  if (!out || world < 1 || world > HALO_MAX_WORLD || rank < 0 || rank >= world || halo_len < 0 || max_halo_len < halo_len)
    return cfail(nullptr, ND_B200_EINVAL, "nd_b200_comm_create: bad arguments (world must be 1..%d, 0 <= halo_len <= max_halo_len)", HALO_MAX_WORLD);
  nd_b200_comm* c = new (std::nothrow) nd_b200_comm();
  if (!c) return cfail(nullptr, ND_B200_ENOMEM, "out of host memory");
  c->device = device; c->rank = rank; c->world = world; c->halo_len = halo_len;
  c->halo_bytes = ((size_t)std::max<int64_t>(max_halo_len, 1) * sizeof(double) + 255) / 256 * 256;
  const size_t total = 2 * c->halo_bytes + 256;
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&c->base[rank], total);
  if (ce == cudaSuccess) ce = cudaMemset(c->base[rank], 0, total);
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&c->d_done, sizeof(unsigned int));
  if (ce == cudaSuccess) ce = cudaMemset(c->d_done, 0, sizeof(unsigned int));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&c->d_timeout, sizeof(int));
  if (ce == cudaSuccess) ce = cudaMemset(c->d_timeout, 0, sizeof(int));
#ifdef ND_CUSIM
  if (ce == cudaSuccess) ce = cudaHostAlloc((void**)&c->h_timeout, sizeof(int), cudaHostAllocDefault);
#else
  if (ce == cudaSuccess) ce = cudaHostAlloc((void**)&c->h_timeout, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable);
#endif
  if (ce == cudaSuccess) *c->h_timeout = 0;
  if (const char* s = getenv("ND_B200_HALO_TIMEOUT_MS")) {
    const double ms = atof(s);
    if (ms > 0) c->timeout_clocks = (long long)(ms * 2.0e6);      // ~2 GHz SM clock
  }
  if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) {
    cfail(nullptr, ND_B200_ECUDA, "nd_b200_comm_create: %s", cudaGetErrorString(ce));
    nd_b200_comm_destroy(c);
    return ND_B200_ECUDA;
  }
  *out = c;
  return ND_B200_OK;
}

int nd_b200_comm_export(nd_b200_comm* c, void* handle_out) {
  if (!c || !handle_out) return ND_B200_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) <= ND_B200_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  COMM_TRY(c, cudaSetDevice(c->device));
  COMM_TRY(c, cudaIpcGetMemHandle(&h, c->base[c->rank]));
  memset(handle_out, 0, ND_B200_IPC_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof h);
  return ND_B200_OK;
}

int nd_b200_comm_open_peer(nd_b200_comm* c, int32_t peer, const void* handle) {
  if (!c || !handle || peer < 0 || peer >= c->world) return ND_B200_EINVAL;
  if (peer == c->rank) return ND_B200_OK;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  COMM_TRY(c, cudaSetDevice(c->device));
  void* q = nullptr;
  COMM_TRY(c, cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
  c->base[peer] = static_cast<unsigned char*>(q);
  return ND_B200_OK;
}

int nd_b200_comm_set_send(nd_b200_comm* c, int32_t peer, const int64_t* state_offsets, int64_t n, int64_t dst_offset) {
  if (!c || peer < 0 || peer >= c->world || peer == c->rank || n < 0 || dst_offset < 0 || (n > 0 && !state_offsets))
    return cfail(c, ND_B200_EINVAL, "nd_b200_comm_set_send: bad arguments");
  COMM_TRY(c, cudaSetDevice(c->device));
  cudaFree(c->d_send_idx[peer]);
  c->d_send_idx[peer] = nullptr;
  c->send_n[peer] = n; c->dst_off[peer] = dst_offset;
  if (n == 0) return ND_B200_OK;
  std::vector<int> idx((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (state_offsets[i] < 0 || state_offsets[i] >= INT_MAX) return cfail(c, ND_B200_EINVAL, "send offset %lld out of range", (long long)state_offsets[i]);
    idx[(size_t)i] = (int)state_offsets[i];
  }
  COMM_TRY(c, cudaMalloc((void**)&c->d_send_idx[peer], sizeof(int) * (size_t)n));
  COMM_TRY(c, cudaMemcpy(c->d_send_idx[peer], idx.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
  return ND_B200_OK;
}

int nd_b200_comm_status(nd_b200_comm* c, int32_t* timed_out) {
  if (!c || !timed_out) return ND_B200_EINVAL;
  int v = 0;
  COMM_TRY(c, cudaSetDevice(c->device));
  COMM_TRY(c, cudaMemcpy(&v, c->d_timeout, sizeof v, cudaMemcpyDeviceToHost));
  if (c->h_timeout && *(volatile int*)c->h_timeout) v = 1;
  *timed_out = v;
  return ND_B200_OK;
}

const char* nd_b200_comm_last_error(const nd_b200_comm* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

void nd_b200_comm_destroy(nd_b200_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r) {
    cudaFree(c->d_send_idx[r]);
    if (!c->base[r]) continue;
    if (r == c->rank) cudaFree(c->base[r]); else cudaIpcCloseMemHandle(c->base[r]);
  }
  cudaFree(c->d_done); cudaFree(c->d_timeout);
  if (c->h_timeout) cudaFreeHost(c->h_timeout);
  delete c;
}

}  // extern "C"

namespace {
int check_exchange(nd_b200_engine* e, nd_b200_comm* c, const char* who) {
  if (!c) return fail(e, ND_B200_EINVAL, "%s: comm is NULL", who);
  if (!e->gather_from_u || e->split) return fail(e, ND_B200_EUNSUPPORTED, "%s needs StateMask vertices and a fused kernel", who);
  if (e->halo_base == INT_MAX) return fail(e, ND_B200_EINVAL, "the engine was created without gather_offset (no halo layout)");
  if (e->gather_len - e->lastidx_dynamic != c->halo_len) return fail(e, ND_B200_EINVAL, "comm halo holds %lld outputs, the engine expects %lld", c->halo_len, e->gather_len - e->lastidx_dynamic);
  for (int r = 0; r < c->world; ++r)
    if (!c->base[r]) return fail(e, ND_B200_EINVAL, "peer %d has not been opened", r);
  // sticky: a tile of an earlier launch gave up waiting for a peer -- every result since then is poisoned (NaN)
  if (c->h_timeout && *(volatile int*)c->h_timeout)
    return fail(e, ND_B200_ETIMEOUT, "%s: an earlier exchange timed out waiting for a peer's boundary outputs (rank skew beyond the spin budget, "
                "ND_B200_HALO_TIMEOUT_MS, or a dead peer); results since then are NaN -- recreate the comm", who);
  return ND_B200_OK;
}
// one exchange = the next sequence number: what the publishing blocks of this launch send (outputs packed from `src`) and
// what its halo-reading tiles wait for
WaitSpec next_exchange(nd_b200_comm* c, const double* src, HaloParams& H) {
  const unsigned long long seq = ++c->seq;
  const int parity = (int)(seq & 1ull);
  memset(&H, 0, sizeof H);
  long long total = 0;
  for (int r = 0; r < c->world; ++r) {
    H.halo[r] = c->halo(r, parity); H.flags[r] = c->flags(r);
    H.send_idx[r] = c->d_send_idx[r]; H.send_n[r] = c->send_n[r]; H.dst_off[r] = c->dst_off[r];
    total += c->send_n[r];
  }
  H.world = c->world; H.rank = c->rank; H.seq = seq; H.src = src; H.done_counter = c->d_done;
  // publishing blocks: 128 threads each, ~8 outputs per thread, at least one (it raises the flags even when nothing is sent)
  const int n_pub = (int)std::max<long long>(1, std::min<long long>(148 * 16, (total + 1023) / 1024));
  return WaitSpec{c->halo(c->rank, parity), c->flags(c->rank), seq, c->world, c->d_timeout, &H, n_pub, c->h_timeout, c->timeout_clocks};
}
}  // namespace

extern "C" {

int nd_b200_rhs_exchange(nd_b200_engine* e, nd_b200_comm* c, double* du, const double* u, const double* p, double t,
                         void* stream) {
  if (int rc = check_call(e, du, u, p)) return rc;
  if (int rc = check_exchange(e, c, "nd_b200_rhs_exchange")) return rc;
  CUDA_TRY(e, cudaSetDevice(e->device));
  HaloParams H;
  const WaitSpec w = next_exchange(c, u, H);
  return rhs_impl(e, du, u, p, t, (cudaStream_t)stream, MODE_DU, nullptr, &w);
}

/* Classical RK4 on a row-partitioned engine: four exchanging launches per step and nothing else -- every stage packs the
 * boundary outputs of ITS input vector for the peers, evaluates the owned rows and applies the fused stage update to the
 * owned states (same operation order as nd_b200_rk4 / the oracle).  Only the owned states of u are read and advanced. */
int nd_b200_rk4_exchange(nd_b200_engine* e, nd_b200_comm* c, double* u, const double* p, double t0, double dt, int64_t nsteps,
                         void* stream) {
  if (int rc = check_call(e, u, u, p)) return rc;
  if (int rc = check_exchange(e, c, "nd_b200_rk4_exchange")) return rc;
  if (nsteps <= 0) return ND_B200_OK;
  CUDA_TRY(e, cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = sizeof(double) * (size_t)e->lastidx_dynamic;
  if (!e->d_tmpA) {
    CUDA_TRY(e, cudaMalloc((void**)&e->d_tmpA, nb));
    CUDA_TRY(e, cudaMalloc((void**)&e->d_tmpB, nb));
    CUDA_TRY(e, cudaMalloc((void**)&e->d_ksum, nb));
  }
  const double h2 = 0.5 * dt;
  const double* in[4] = {u, e->d_tmpA, e->d_tmpB, e->d_tmpA};
  double* out[4] = {e->d_tmpA, e->d_tmpB, e->d_tmpA, u};
  const double hs[4] = {h2, h2, dt, 0.0};
  for (int64_t k = 0; k < nsteps; ++k) {
    const double t = t0 + (double)k * dt;
    const double ts[4] = {t, t + h2, t + h2, t + dt};
    for (int s = 0; s < 4; ++s) {
      HaloParams H;
      const WaitSpec w = next_exchange(c, in[s], H);
      KParams P;
      fill_params(e, P);
      P.p = p; P.mode = MODE_RK; P.u0 = u; P.ksum = e->d_ksum; P.h6 = dt / 6.0;
      P.stage = s + 1; P.u = in[s]; P.gsrc = in[s]; P.unext = out[s]; P.hs = hs[s]; P.t = ts[s]; P.vout_next = nullptr;
      P.halo = w.halo; P.wait_flags = w.flags; P.wait_seq = w.seq; P.wait_world = w.world; P.wait_timeout = w.timeout;
      P.wait_timeout_host = w.timeout_host; P.wait_budget = w.budget;
      P.H = H; P.n_pub = w.n_pub;
      const bool waits = e->jag ? (e->wait_from < e->nslices || e->n_jlong > 0) : (e->wait_from < e->nblocks);
      P.fence = waits ? 0 : 1;
      CUDA_TRY(e, launch_fused(e, P, st));
    }
  }
  return ND_B200_OK;
}

/* timing aid: the owned rows evaluated on whatever the halo buffer currently holds -- no publish, no wait */
int nd_b200_rhs_local(nd_b200_engine* e, nd_b200_comm* c, double* du, const double* u, const double* p, double t, void* stream) {
  if (int rc = check_call(e, du, u, p)) return rc;
  if (!c || e->halo_base == INT_MAX) return fail(e, ND_B200_EINVAL, "nd_b200_rhs_local needs a halo engine and its comm");
  CUDA_TRY(e, cudaSetDevice(e->device));
  WaitSpec w{c->halo(c->rank, (int)(c->seq & 1ull)), nullptr, 0, 0, nullptr, nullptr, 0, nullptr, 0};
  return rhs_impl(e, du, u, p, t, (cudaStream_t)stream, MODE_DU, nullptr, &w);
}

}  // extern "C"
