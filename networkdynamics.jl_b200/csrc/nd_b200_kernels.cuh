// nd_b200_kernels.cuh -- sm_100a kernels of the fused network RHS.
//
// One fused kernel replaces the reference's per-call pipeline
//   fill!(du) / fill!(o,NaN) / fill!(aggbuf)            src/coreloop.jl:24-30
//   PASS 1 vertex g (StateMask)                         src/coreloop.jl:39
//   gather!(gbuf, o)                                    src/coreloop.jl:67, src/gbufs.jl:25
//   PASS 5 edge g                                       src/coreloop.jl:78
//   aggregate!(aggbuf, o)                               src/coreloop.jl:90, src/aggregators.jl:140-151
//   PASS 6 vertex f                                     src/coreloop.jl:97
// over a destination-sorted CSR whose rows are the aggregation slots and whose entries are in the
// reference's accumulation order (ascending position in `o`).  Neither `o`, `gbuf` nor `aggbuf`
// exist in memory: edge values live in shared memory, row sums in registers.
//
// Compiled with -fmad=false: the reference (Julia) never contracts a*b+c, so neither do we.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nd_b200.h"

namespace ndb {

struct VBDev {          // one vertex ComponentBatch, 0-based offsets
  int kind, dim, pdim;
  int row0, nrows;      // aggregation-slot rows [row0, row0+nrows) of this batch (global numbering)
  int blk0;             // first thread block working on this batch
  long long state0, p0; // offsets into u / p
};
struct EBDev {          // one edge ComponentBatch
  int kind, coupling, pdim;
};

enum { MODE_DU = 0, MODE_AGG = 1, MODE_RK = 2 };
constexpr int EK_GENERIC = -1;   // several edge batches: per-entry batch id lookup

struct KParams {
  const int* __restrict__ rowptr;      // [nrows_owned+1], entries of owned rows, relative to row_base
  const int* __restrict__ nbr;         // per entry: offset of the neighbour's output in gsrc; ~offset when this row is the edge's src
  const int* __restrict__ epar;        // per entry: offset of the edge's parameter block in p (nullptr when no edge has parameters)
  const uint8_t* __restrict__ ebid;    // per entry: edge batch id (EK_GENERIC only)
  const int* __restrict__ blk_row;     // [nblocks+1] first (global) row of each thread block
  const VBDev* __restrict__ vb;
  const EBDev* __restrict__ eb;
  int n_vb, n_eb;
  int row_base;                        // first owned row
  int long_thr;                        // rows with more entries are reduced by the whole block
  int gather_from_u;                   // 1: vertex outputs are read straight from u (all StateMask, vdepth 1)
  int mode;
  const double* __restrict__ u;        // state vector the vertex models read
  const double* __restrict__ gsrc;     // gather source: u (gather_from_u) or the materialised vertex outputs
  const double* __restrict__ p;
  double* __restrict__ du;             // MODE_DU
  double* __restrict__ aggbuf;         // MODE_AGG
  // MODE_RK: fused classical RK4 stage (see nd_b200_rk4)
  int stage;                           // 1..4
  const double* __restrict__ u0;       // state at the start of the step
  double* __restrict__ unext;          // stage 1-3: next stage input; stage 4: u0 itself
  double* __restrict__ ksum;           // running k1 + 2k2 + 2k3
  double* __restrict__ vout_next;      // materialised outputs of unext (when !gather_from_u)
  double hs;                           // dt/2, dt/2, dt for stages 1..3
  double h6;                           // dt/6
  double t;
  // pipelined kernel (rhs_pipe_kernel): tile-padded copies of the entry arrays
  const int4* __restrict__ tiles;      // per tile {row0, e0 (multiple of 4), rp0 (multiple of 8) | ne of a long tile, ne | nrows<<16 | batch<<25 | long<<31}
  int ntiles;
  const unsigned short* __restrict__ rp16;  // per tile nrows+1 row offsets relative to e0
  const int* __restrict__ nbr2;
  const int* __restrict__ epar2;
  const uint8_t* __restrict__ ebid2;
};

// ------------------------------------------------------------------------------------------------
// model arithmetic -- expression order exactly as the cited reference source (and as oracle/nd_oracle.c)
// ------------------------------------------------------------------------------------------------

// inner edge function: writes the DST output, g(odst, vsrc, vdst, p, t)
template <int VD, int ED>
__device__ __forceinline__ void edge_g_dst(int kind, double* odst, const double* vs, const double* vd,
                                           const double* __restrict__ pe) {
  if constexpr (VD == 1 && ED == 1) {
    switch (kind) {
      case ND_B200_E_DIFFUSION:      // test/ComponentLibrary.jl:8-10
        odst[0] = pe[0] * (vs[0] - vd[0]);
        break;
      case ND_B200_E_DIFFUSION_NOP:  // benchmark/benchmark_models.jl:5-8
        odst[0] = vs[0] - vd[0];
        break;
      case ND_B200_E_KURAMOTO:       // test/ComponentLibrary.jl:51-53
        odst[0] = pe[0] * sin(vs[0] - vd[0]);
        break;
      default: odst[0] = 0.0;
    }
  } else if constexpr (VD == 2 && ED == 2) {
    // ND_B200_E_LINE_DQ, test/ComponentLibrary.jl:212-245: idst = active*1/Z*(Vsrc-Vdst), Z = R+jX
    double R = pe[0], X = pe[1], active = pe[2];
    double dr = vs[0] - vd[0];
    double di = vs[1] - vd[1];
    double den = R * R + X * X;
    odst[0] = active * ((R * dr + X * di) / den);
    odst[1] = active * ((R * di - X * dr) / den);
  }
}

// value an entry contributes to ITS row: the dst output if the row is the edge's dst, else the src
// output produced by the wrapper (AntiSymmetric: -odst, Symmetric: odst; src/component_functions.jl:117-152)
template <int VD, int ED>
__device__ __forceinline__ void entry_value(int kind, int coupling, int side, const double* self,
                                            const double* xn, const double* __restrict__ pe, double* val) {
  const double* vs = side ? self : xn;
  const double* vd = side ? xn : self;
  edge_g_dst<VD, ED>(kind, val, vs, vd, pe);
  if (side && coupling == ND_B200_ANTISYMMETRIC) {
#pragma unroll
    for (int d = 0; d < ED; ++d) val[d] = -val[d];
  }
}

// vertex g for non-StateMask models: NoFeedForward g(out,u,p,t)
__device__ __forceinline__ void vertex_g(int kind, int outdim, double* out, const double* v,
                                         const double* __restrict__ pv) {
  if (kind == ND_B200_V_SWING_DQ) {   // test/ComponentLibrary.jl:158-159
    double V = pv[3];
    out[0] = V * cos(v[0]);
    out[1] = V * sin(v[0]);
  } else {                            // StateMask(1:outdim), src/component_functions.jl:81-99
    for (int k = 0; k < outdim; ++k) out[k] = v[k];
  }
}

// vertex f: f(dv, v, acc, p, t).  selfout = this vertex's own outputs (only used by SWING_DQ, whose f
// recomputes u_r,u_i with the same expressions as its g).
template <int VD, int ED>
__device__ __forceinline__ void vertex_f(int kind, double* dv, const double* v, const double* acc,
                                         const double* __restrict__ pv, const double* selfout) {
  switch (kind) {
    case ND_B200_V_DIFFUSION:            // test/ComponentLibrary.jl:42-45
      dv[0] = acc[0];
      break;
    case ND_B200_V_KURAMOTO_FIRST:       // test/ComponentLibrary.jl:69-71
      dv[0] = pv[0] + acc[0];
      break;
    case ND_B200_V_KURAMOTO_SECOND: {    // test/ComponentLibrary.jl:59-63
      double M = pv[0], D = pv[1], Pm = pv[2];
      dv[0] = v[1];
      dv[1] = 1.0 / M * (Pm - D * v[1] + acc[0]);
    } break;
    case ND_B200_V_KURAMOTO_SECOND_BENCH: {  // benchmark/benchmark_models.jl:37-41
      double P = pv[0];
      dv[0] = v[1];
      dv[1] = P - 1.0 * v[1];
      dv[1] += acc[0];
    } break;
    case ND_B200_V_SWING_DQ: {           // test/ComponentLibrary.jl:139-161
      if constexpr (VD == 2 && ED == 2) {
        double M = pv[0], D = pv[1], Pmech = pv[2];
        double Pel = selfout[0] * acc[0] + selfout[1] * acc[1];
        double Pdamping = -D * v[1];
        dv[0] = v[1];
        dv[1] = 1.0 / M * (Pmech + Pdamping + Pel);
      }
    } break;
  }
}

// PASS 6 for one row + the epilogue selected by mode.  v = the vertex's states, pv = its parameters
// (pointers into global or shared memory).
template <int VD, int ED>
__device__ __forceinline__ void vertex_phase(const KParams& P, const VBDev& B, int row, const double* acc,
                                             const double* selfout, const double* vin, const double* pv) {
  if (P.mode == MODE_AGG) {
#pragma unroll
    for (int d = 0; d < ED; ++d) P.aggbuf[(long long)row * ED + d] = acc[d];
    return;
  }
  const long long i = row - B.row0;
  const long long s = B.state0 + i * B.dim;
  const bool two = B.dim == 2;          // registry: dim is 1 or 2
  const double v[2] = {vin[0], two ? vin[1] : 0.0};
  double dv[2] = {0.0, 0.0};
  vertex_f<VD, ED>(B.kind, dv, v, acc, pv, selfout);
  if (P.mode == MODE_DU) {
    if (two && ((s & 1) == 0)) {
      *reinterpret_cast<double2*>(P.du + s) = make_double2(dv[0], dv[1]);
    } else {
      P.du[s] = dv[0];
      if (two) P.du[s + 1] = dv[1];
    }
    return;
  }
  // MODE_RK: classical RK4, operation order of the CPU restatement's rk4:
  //   u <- u + (dt/6)*(((k1 + 2k2) + 2k3) + k4)
  double un[2] = {0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (c == 1 && !two) break;
    const long long idx = s + c;
    if (P.stage == 1) {
      P.ksum[idx] = dv[c];
      un[c] = v[c] + P.hs * dv[c];          // u == u0 in stage 1
    } else if (P.stage < 4) {
      P.ksum[idx] = P.ksum[idx] + 2.0 * dv[c];
      un[c] = P.u0[idx] + P.hs * dv[c];
    } else {
      un[c] = P.u0[idx] + P.h6 * (P.ksum[idx] + dv[c]);
    }
    P.unext[idx] = un[c];
  }
  if (!P.gather_from_u) {
    double out[VD];
    vertex_g(B.kind, VD, out, un, pv);
#pragma unroll
    for (int k = 0; k < VD; ++k) P.vout_next[(long long)row * VD + k] = out[k];
  }
}

// loads the (<= 2) states of a row from global memory
__device__ __forceinline__ void load_vertex_state(const KParams& P, const VBDev& B, int row, double* v) {
  const long long s = B.state0 + (long long)(row - B.row0) * B.dim;
  v[0] = P.u[s];
  v[1] = (B.dim == 2) ? P.u[s + 1] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// fused gather -> edge -> ordered row reduce -> vertex kernel
// ------------------------------------------------------------------------------------------------
// Occupancy is the lever for this kernel (it is bound by the L2 sector bandwidth of the random gathers and hides
// latency with many independent blocks), so registers are capped through the min-blocks launch bound.
// Measured on B200 (profiles/r01_tuning.md): 64 resident warps/SM (32 registers) is best for the arithmetic-free
// diffusion kernels, 48 warps/SM (40 registers) for the kernels that evaluate sin / complex division.
constexpr int fused_warps_per_sm(int ek) {
  return (ek == ND_B200_E_DIFFUSION || ek == ND_B200_E_DIFFUSION_NOP) ? 64 : 48;
}
template <int VD, int ED, int EK, int PE, int BLOCK, int EPT>
__global__ void __launch_bounds__(BLOCK, (fused_warps_per_sm(EK) * 32) / BLOCK) rhs_fused_kernel(const __grid_constant__ KParams P) {
  constexpr int TILE = BLOCK * EPT;
  static_assert(BLOCK <= 256, "row ids are stored as uint8");
  __shared__ double s_val[TILE * ED];
  __shared__ double s_self[BLOCK * VD];
  __shared__ int s_rp[BLOCK + 1];
  __shared__ uint8_t s_rowid[TILE];

  const int tid = threadIdx.x;
  // one 16-byte descriptor per thread block: {row0, e0, ne (long rows), ne | nrows<<16 | batch<<25 | long<<31}
  const int4 d = __ldg(&P.tiles[blockIdx.x]);
  const int r0 = d.x, e0 = d.y;
  const bool is_long = d.w < 0;
  const int nrows = (d.w >> 16) & 0x1FF;
  const int ne = is_long ? d.z : (d.w & 0xFFFF);
  const VBDev B = P.vb[(d.w >> 25) & 0x3F];
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;

  // ---------------- long row: whole block reduces one row with a fixed-shape tree -----------------
  if (is_long) {
    double self[VD];
    const long long sidx = P.gather_from_u ? (B.state0 + (long long)(r0 - B.row0) * B.dim) : (long long)r0 * VD;
#pragma unroll
    for (int k = 0; k < VD; ++k) self[k] = P.gsrc[sidx + k];
    double part[ED];
#pragma unroll
    for (int q = 0; q < ED; ++q) part[q] = 0.0;
    for (int jj = tid; jj < ne; jj += BLOCK) {
      int nb = P.nbr[e0 + jj];
      const int side = nb < 0;
      nb = side ? ~nb : nb;
      double xn[VD];
#pragma unroll
      for (int k = 0; k < VD; ++k) xn[k] = P.gsrc[(long long)nb + k];
      const double* pe = P.p;
      if constexpr (PE > 0) pe = P.p + P.epar[e0 + jj];
      int kind = EK, coupling = coupling0;
      if constexpr (EK == EK_GENERIC) {
        const EBDev E = P.eb[P.ebid[e0 + jj]];
        kind = E.kind; coupling = E.coupling;
      }
      double val[ED];
      entry_value<VD, ED>(kind, coupling, side, self, xn, pe, val);
#pragma unroll
      for (int q = 0; q < ED; ++q) part[q] = part[q] + val[q];
    }
#pragma unroll
    for (int q = 0; q < ED; ++q) s_val[tid * ED + q] = part[q];
    __syncthreads();
    for (int s = BLOCK / 2; s > 0; s >>= 1) {
      if (tid < s) {
#pragma unroll
        for (int q = 0; q < ED; ++q) s_val[tid * ED + q] = s_val[tid * ED + q] + s_val[(tid + s) * ED + q];
      }
      __syncthreads();
    }
    if (tid == 0) {
      double acc[ED], v[2];
#pragma unroll
      for (int q = 0; q < ED; ++q) acc[q] = s_val[q];
      load_vertex_state(P, B, r0, v);
      vertex_phase<VD, ED>(P, B, r0, acc, self, v, P.p + B.p0 + (long long)(r0 - B.row0) * B.pdim);
    }
    return;
  }

  // ---------------- regular tile: <= BLOCK rows, <= TILE entries ---------------------------------
  // Register budget matters more than load hoisting here: the kernel is bound by L2 sector bandwidth of the
  // random gathers, and occupancy (many independent blocks in different phases) is what keeps L2 busy.
  // (1) coalesced index loads, issued first so they overlap the row bookkeeping
  int nb[EPT], ep[EPT];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    nb[k] = 0; ep[k] = 0;
    if (jj < ne) {
      nb[k] = P.nbr[e0 + jj];
      if constexpr (PE > 0) ep[k] = P.epar[e0 + jj];
    }
  }
  // (2) row pointers + own outputs of the block's rows
  if (tid < nrows) {
    s_rp[tid] = P.rowptr[(r0 - P.row_base) + tid] - e0;
    const long long sidx = P.gather_from_u ? (B.state0 + (long long)(r0 + tid - B.row0) * B.dim)
                                           : (long long)(r0 + tid) * VD;
#pragma unroll
    for (int k = 0; k < VD; ++k) s_self[tid * VD + k] = P.gsrc[sidx + k];
  }
  if (tid == 0) s_rp[nrows] = ne;
  __syncthreads();
  // (3) entry -> local row map
  if (tid < nrows) {
    const int a = s_rp[tid], z = s_rp[tid + 1];
    for (int jj = a; jj < z; ++jj) s_rowid[jj] = (uint8_t)tid;
  }
  // (4) the gather: EPT independent random reads per thread
  double xn[EPT][VD];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    const int off = nb[k] < 0 ? ~nb[k] : nb[k];
#pragma unroll
    for (int q = 0; q < VD; ++q) xn[k][q] = 0.0;
    if (jj < ne) {
      if constexpr (VD == 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(P.gsrc + off);
        xn[k][0] = t2.x; xn[k][1] = t2.y;
      } else {
        xn[k][0] = P.gsrc[off];
      }
    }
  }
  __syncthreads();
  // (5) edge evaluation, one entry per thread per k
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    if (jj < ne) {
      const int side = nb[k] < 0;
      const int r = s_rowid[jj];
      double self[VD];
#pragma unroll
      for (int q = 0; q < VD; ++q) self[q] = s_self[r * VD + q];
      int kind = EK, coupling = coupling0;
      if constexpr (EK == EK_GENERIC) {
        const EBDev E = P.eb[P.ebid[e0 + jj]];
        kind = E.kind; coupling = E.coupling;
      }
      double val[ED];
      entry_value<VD, ED>(kind, coupling, side, self, xn[k], P.p + ep[k], val);
#pragma unroll
      for (int q = 0; q < ED; ++q) s_val[jj * ED + q] = val[q];
    }
  }
  __syncthreads();
  // (6) ordered per-row accumulation (the reference's sequential order) + vertex model
  if (tid < nrows) {
    double acc[ED];
#pragma unroll
    for (int q = 0; q < ED; ++q) acc[q] = 0.0;
    const int a = s_rp[tid], z = s_rp[tid + 1];
    for (int jj = a; jj < z; ++jj) {
#pragma unroll
      for (int q = 0; q < ED; ++q) acc[q] = acc[q] + s_val[jj * ED + q];
    }
    double self[VD], v[2];
#pragma unroll
    for (int q = 0; q < VD; ++q) self[q] = s_self[tid * VD + q];
    load_vertex_state(P, B, r0 + tid, v);
    vertex_phase<VD, ED>(P, B, r0 + tid, acc, self, v, P.p + B.p0 + (long long)(r0 + tid - B.row0) * B.pdim);
  }
}

// PASS 1 for networks whose vertex outputs are not plain state copies: vout[row*VD + k] = g_v(u, p)
__global__ void vertex_out_kernel(const VBDev* __restrict__ vb, int n_vb, int vd, const double* __restrict__ u,
                                  const double* __restrict__ p, double* __restrict__ vout, int nrows_total) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_total) return;
  int b = 0;
  for (int i = 1; i < n_vb; ++i)
    if (row >= vb[i].row0) b = i;
  const VBDev B = vb[b];
  const long long i = row - B.row0;
  double v[2] = {0.0, 0.0}, out[2] = {0.0, 0.0};
  for (int c = 0; c < B.dim && c < 2; ++c) v[c] = u[B.state0 + i * B.dim + c];
  vertex_g(B.kind, vd, out, v, p + B.p0 + i * B.pdim);
  for (int k = 0; k < vd; ++k) vout[(long long)row * vd + k] = out[k];
}

// get_buffers support: edge outputs into the reference's `o` layout (src range first, dst right after)
template <int VD, int ED>
__global__ void edge_out_kernel(int kind, int coupling, int pdim, int osrc, long long count,
                                const int* __restrict__ esrc_off, const int* __restrict__ edst_off,
                                long long p0, long long out0, const double* __restrict__ gsrc,
                                const double* __restrict__ p, double* __restrict__ o) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double vs[VD], vd[VD], val[ED];
#pragma unroll
  for (int k = 0; k < VD; ++k) { vs[k] = gsrc[(long long)esrc_off[i] + k]; vd[k] = gsrc[(long long)edst_off[i] + k]; }
  edge_g_dst<VD, ED>(kind, val, vs, vd, p + p0 + i * pdim);
  double* oo = o + out0 + i * (osrc + ED);
  if (osrc) {
#pragma unroll
    for (int d = 0; d < ED; ++d) oo[d] = (coupling == ND_B200_ANTISYMMETRIC) ? -val[d] : val[d];
  }
#pragma unroll
  for (int d = 0; d < ED; ++d) oo[osrc + d] = val[d];
}


// ------------------------------------------------------------------------------------------------
// v2: persistent, software-pipelined version of the fused kernel.
//
// The v1 kernel is bound by dependent memory phases (index -> gather -> parameters -> reduce), each
// exposed to the full L1TEX/L2 queueing latency (ncu: long_scoreboard dominant, L1TEX 57 % busy).
// Here every CTA is persistent and walks its tiles with a 3-deep pipeline built on cp.async
// (LDGSTS), so the index stream of tile i+2 and the random gathers of tile i+1 are in flight while
// tile i is evaluated and reduced:
//   stage A(i+2): cp.async.cg 16 B   nbr / epar / row offsets  -> shared        (coalesced stream)
//   stage B(i+1): cp.async.ca  8 B   gsrc[nbr], p[epar], row self/state/params -> shared (gathers)
//   stage C(i)  : edge values in place in shared memory, ordered per-row sums, vertex model, store
// Accumulation order per row is unchanged (strictly sequential, the reference's order).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VD, int ED, int EK, int PE, int BLOCK, int EPT>
struct PipeSmem {
  static constexpr int TILE = BLOCK * EPT;
  int4 desc[4];
  double xn[2][TILE * VD];                 // gathered neighbour outputs; overwritten in place by the edge values
  double pe[2][PE > 0 ? TILE * PE : 2];    // gathered edge parameters
  double self[2][BLOCK * VD];              // own outputs of the tile's rows
  double vu[2][BLOCK * 2];                 // states of the tile's rows (dim <= 2)
  double vp[2][BLOCK * 4];                 // parameters of the tile's rows (pdim <= 4)
  int nbr[2][TILE];
  int epar[2][PE > 0 ? TILE : 4];
  unsigned short rp[2][BLOCK + 8];
  uint8_t rowid[2][TILE];
  uint8_t ebid[2][EK == EK_GENERIC ? TILE : 16];
};

__device__ __forceinline__ int tile_ne(const int4& d) { return d.w & 0xFFFF; }
__device__ __forceinline__ int tile_nrows(const int4& d) { return (d.w >> 16) & 0x1FF; }
__device__ __forceinline__ int tile_batch(const int4& d) { return (d.w >> 25) & 0x3F; }
__device__ __forceinline__ bool tile_long(const int4& d) { return d.w < 0; }

template <int VD, int ED, int EK, int PE, int BLOCK, int EPT>
__global__ void __launch_bounds__(BLOCK) rhs_pipe_kernel(const __grid_constant__ KParams P) {
  static_assert(VD == ED, "edge values overwrite the gathered neighbour values in place");
  static_assert(BLOCK <= 256 && EPT % 4 == 0, "row ids are uint8; index chunks are 16 bytes");
  using Smem = PipeSmem<VD, ED, EK, PE, BLOCK, EPT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  const int first = blockIdx.x;
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;
  auto tile_of = [&](int i) { return first + i * G; };

  // prologue: descriptors of the first three tiles of this CTA
  if (tid < 3) {
    const int t = tile_of(tid);
    S.desc[tid] = t < P.ntiles ? P.tiles[t] : make_int4(0, 0, 0, 0);
  }
  __syncthreads();

  unsigned side_cur = 0, side_nxt = 0;   // per-thread side bits of its EPT entries (tile i / tile i+1)
  unsigned rng_cur = 0, rng_nxt = 0;     // row-thread entry range (a | z << 16)

  const int my_tiles = first < P.ntiles ? (P.ntiles - first + G - 1) / G : 0;
  for (int i = -2; i < my_tiles; ++i) {
    // ---- S1: stage A of tile i+1 has landed; everybody is done with tile i-1 ----------------------
    cp_async_wait<1>();
    __syncthreads();

    // ---- stage D(i+3) + A(i+2): descriptor and index stream -------------------------------------
    {
      const int t3 = tile_of(i + 3);
      if (tid == 0 && i + 3 >= 3 && t3 < P.ntiles) cp_async16(&S.desc[(i + 3) & 3], &P.tiles[t3]);
      const int ia = i + 2;
      if (ia >= 0 && ia < my_tiles) {
        const int4 d = S.desc[ia & 3];
        if (!tile_long(d)) {
          const int b = ia & 1;
          const int ne = tile_ne(d), nr = tile_nrows(d);
          const int nchunk = (ne + 3) >> 2;
          for (int c = tid; c < nchunk; c += BLOCK) {
            cp_async16(&S.nbr[b][c * 4], P.nbr2 + d.y + c * 4);
            if constexpr (PE > 0) cp_async16(&S.epar[b][c * 4], P.epar2 + d.y + c * 4);
          }
          if constexpr (EK == EK_GENERIC) {
            const int nc16 = (ne + 15) >> 4;
            for (int c = tid; c < nc16; c += BLOCK) cp_async16(&S.ebid[b][c * 16], P.ebid2 + d.y + c * 16);
          }
          const int nrc = (nr + 1 + 7) >> 3;
          if (tid < nrc) cp_async16(&S.rp[b][tid * 8], P.rp16 + d.z + tid * 8);
        }
      }
      cp_async_commit();
    }

    // ---- stage B(i+1): gathers ---------------------------------------------------------------------
    {
      const int ib = i + 1;
      side_nxt = 0; rng_nxt = 0;
      if (ib >= 0 && ib < my_tiles) {
        const int4 d = S.desc[ib & 3];
        if (!tile_long(d)) {
          const int b = ib & 1;
          const int ne = tile_ne(d), nr = tile_nrows(d);
#pragma unroll
          for (int k = 0; k < EPT; ++k) {
            const int jj = k * BLOCK + tid;
            if (jj < ne) {
              const int nb = S.nbr[b][jj];
              const int off = nb < 0 ? ~nb : nb;
              side_nxt |= (nb < 0 ? 1u : 0u) << k;
              if constexpr (VD == 2) cp_async16(&S.xn[b][jj * 2], P.gsrc + off);
              else cp_async8(&S.xn[b][jj], P.gsrc + off);
              if constexpr (PE > 0) {
                const int ep = S.epar[b][jj];
#pragma unroll
                for (int q = 0; q < PE; ++q) cp_async8(&S.pe[b][jj * PE + q], P.p + ep + q);
              }
            }
          }
          if (tid < nr) {
            const VBDev B = P.vb[tile_batch(d)];
            const int row = d.x + tid;
            const long long li = row - B.row0;
            const long long s = B.state0 + li * B.dim;
            const long long sidx = P.gather_from_u ? s : (long long)row * VD;
            if constexpr (VD == 2) cp_async16(&S.self[b][tid * 2], P.gsrc + sidx);
            else cp_async8(&S.self[b][tid], P.gsrc + sidx);
            if (P.mode != MODE_AGG) {
              for (int c = 0; c < B.dim; ++c) cp_async8(&S.vu[b][tid * 2 + c], P.u + s + c);
              for (int c = 0; c < B.pdim; ++c) cp_async8(&S.vp[b][tid * 4 + c], P.p + B.p0 + li * B.pdim + c);
            }
            const unsigned a = S.rp[b][tid], z = S.rp[b][tid + 1];
            rng_nxt = a | (z << 16);
            for (unsigned jj = a; jj < z; ++jj) S.rowid[b][jj] = (uint8_t)tid;
          }
        }
      }
      cp_async_commit();
    }

    // ---- S2: the gathers of tile i have landed ------------------------------------------------------
    cp_async_wait<2>();
    __syncthreads();

    if (i >= 0) {
      const int4 d = S.desc[i & 3];
      const int b = i & 1;
      const VBDev B = P.vb[tile_batch(d)];
      if (tile_long(d)) {
        // ---- long row: the whole CTA reduces one row with a fixed-shape tree (direct loads) ---------
        const int row = d.x, e0 = d.y, ne = d.z;
        double self[VD];
        const long long sidx = P.gather_from_u ? (B.state0 + (long long)(row - B.row0) * B.dim) : (long long)row * VD;
#pragma unroll
        for (int k = 0; k < VD; ++k) self[k] = P.gsrc[sidx + k];
        double part[ED];
#pragma unroll
        for (int q = 0; q < ED; ++q) part[q] = 0.0;
        for (int jj = tid; jj < ne; jj += BLOCK) {
          int nb = P.nbr2[e0 + jj];
          const int side = nb < 0;
          nb = side ? ~nb : nb;
          double xn[VD];
#pragma unroll
          for (int k = 0; k < VD; ++k) xn[k] = P.gsrc[(long long)nb + k];
          const double* pe = P.p;
          if constexpr (PE > 0) pe = P.p + P.epar2[e0 + jj];
          int kind = EK, coupling = coupling0;
          if constexpr (EK == EK_GENERIC) {
            const EBDev E = P.eb[P.ebid2[e0 + jj]];
            kind = E.kind; coupling = E.coupling;
          }
          double val[ED];
          entry_value<VD, ED>(kind, coupling, side, self, xn, pe, val);
#pragma unroll
          for (int q = 0; q < ED; ++q) part[q] = part[q] + val[q];
        }
#pragma unroll
        for (int q = 0; q < ED; ++q) S.xn[b][tid * ED + q] = part[q];
        __syncthreads();
        for (int s = BLOCK / 2; s > 0; s >>= 1) {
          if (tid < s) {
#pragma unroll
            for (int q = 0; q < ED; ++q) S.xn[b][tid * ED + q] = S.xn[b][tid * ED + q] + S.xn[b][(tid + s) * ED + q];
          }
          __syncthreads();
        }
        if (tid == 0) {
          double acc[ED], v[2];
#pragma unroll
          for (int q = 0; q < ED; ++q) acc[q] = S.xn[b][q];
          load_vertex_state(P, B, row, v);
          vertex_phase<VD, ED>(P, B, row, acc, self, v, P.p + B.p0 + (long long)(row - B.row0) * B.pdim);
        }
      } else {
        // ---- stage C(i): edge values in place ----------------------------------------------------------
        const int ne = tile_ne(d), nr = tile_nrows(d);
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          const int jj = k * BLOCK + tid;
          if (jj < ne) {
            const int r = S.rowid[b][jj];
            double self[VD], xn[VD], val[ED];
#pragma unroll
            for (int q = 0; q < VD; ++q) { self[q] = S.self[b][r * VD + q]; xn[q] = S.xn[b][jj * VD + q]; }
            int kind = EK, coupling = coupling0;
            if constexpr (EK == EK_GENERIC) {
              const EBDev E = P.eb[S.ebid[b][jj]];
              kind = E.kind; coupling = E.coupling;
            }
            entry_value<VD, ED>(kind, coupling, (side_cur >> k) & 1u, self, xn, &S.pe[b][PE > 0 ? jj * PE : 0], val);
#pragma unroll
            for (int q = 0; q < ED; ++q) S.xn[b][jj * ED + q] = val[q];
          }
        }
        __syncthreads();
        // ---- ordered per-row accumulation (reference order) + vertex model -------------------------------
        if (tid < nr) {
          double acc[ED];
#pragma unroll
          for (int q = 0; q < ED; ++q) acc[q] = 0.0;
          const int a = rng_cur & 0xFFFF, z = rng_cur >> 16;
          for (int jj = a; jj < z; ++jj) {
#pragma unroll
            for (int q = 0; q < ED; ++q) acc[q] = acc[q] + S.xn[b][jj * ED + q];
          }
          double self[VD];
#pragma unroll
          for (int q = 0; q < VD; ++q) self[q] = S.self[b][tid * VD + q];
          vertex_phase<VD, ED>(P, B, d.x + tid, acc, self, &S.vu[b][tid * 2], &S.vp[b][tid * 4]);
        }
      }
    }
    side_cur = side_nxt;
    rng_cur = rng_nxt;
  }
  cp_async_wait<0>();
}

}  // namespace ndb
