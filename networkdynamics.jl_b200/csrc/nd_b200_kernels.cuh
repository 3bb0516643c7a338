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
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#include "nd_b200.h"
#endif
// When this file is compiled at run time (NVRTC, networks with user-supplied component kinds -- see custom_source() in
// nd_b200.cu) the generator has already emitted: the integer typedefs, the registry enums of nd_b200.h, the user's
// component functions in namespace ndb_user, and the macros below that splice them into the model switches.
#ifndef ND_MAX_VDIM
#define ND_MAX_VDIM 2            // largest vertex state dimension of the network (registry models: 2)
#endif
#ifndef ND_CUSTOM_EDGE_CASES
#define ND_CUSTOM_EDGE_CASES     // case <kind>: ndb_user::edge_g_<kind>(odst, vs, vd, pe, t); break;
#endif
#ifndef ND_CUSTOM_EDGE_FID_CASES
#define ND_CUSTOM_EDGE_FID_CASES // case <kind>: ndb_user::edge_g_<kind>(osrc, odst, vs, vd, pe, t); break;   (Fiducial)
#endif
#ifndef ND_CUSTOM_EDGE_F_CASES
#define ND_CUSTOM_EDGE_F_CASES   // case <kind>: ndb_user::edge_f_<kind>(de, ue, vs, vd, pe, t); break;   (edges with states)
#endif
#ifndef ND_MAX_EDIM
#define ND_MAX_EDIM 2            // largest edge state dimension of the network (registry models: 2)
#endif
#ifndef ND_MAX_EXT
#define ND_MAX_EXT 0             // largest number of external inputs of a component (src/external_inputs.jl); 0: none
#endif
#ifndef ND_CUSTOM_VERTEX_F_CASES
#define ND_CUSTOM_VERTEX_F_CASES // case <kind>: ndb_user::vertex_f_<kind>(dv, v, acc, pv, t); break;
#endif
#ifndef ND_CUSTOM_VERTEX_G_CASES
#define ND_CUSTOM_VERTEX_G_CASES // case <kind>: ndb_user::vertex_g_<kind>(out, v, pv, t); break;
                                 // feed-forward vertices (injectors): ndb_user::vertex_g_<kind>(out, v, ins, pv, t)
#endif

namespace ndb {

struct VBDev {          // one vertex ComponentBatch, 0-based offsets
  int kind, dim, pdim;
  int row0, nrows;      // aggregation-slot rows [row0, row0+nrows) of this batch (global numbering)
  int blk0;             // first thread block working on this batch
  long long state0, p0; // offsets into u / p
  // external inputs (src/external_inputs.jl): component i of the batch reads extdim scalars, ext[i*extdim + k] = where
  // from: low 30 bits = offset, bit 30 = from the materialised vertex outputs instead of u, bit 31 = negated
  const int* ext;
  int extdim;
  // feed-forward vertices ("injector" leaves behind a LoopbackConnection, src/post_utils.jl:110-190): g also reads the
  // vertex's input = the output of its hub; ffin[i] = offset of that hub's output in the materialised vertex outputs
  int ff;
  const int* ffin;
};
constexpr int ND_EXT_FROM_VOUT = 1 << 30;
// collect_externals! for one component (src/coreloop.jl:61, src/external_inputs.jl:52-66)
__device__ __forceinline__ void gather_ext(const int* __restrict__ codes, int n, const double* __restrict__ u,
                                           const double* __restrict__ vout, double* ext) {
#if ND_MAX_EXT > 0
#pragma unroll
  for (int k = 0; k < ND_MAX_EXT; ++k) {
    ext[k] = 0.0;
    if (k < n) {
      const int c = codes[k];
      const double x = ((c & ND_EXT_FROM_VOUT) ? vout : u)[c & (ND_EXT_FROM_VOUT - 1)];
      ext[k] = c < 0 ? -x : x;
    }
  }
#else
  (void)codes; (void)n; (void)u; (void)vout; (void)ext;
#endif
}
struct EBDev {          // one edge ComponentBatch
  int kind, coupling, pdim;
  int dim;              // > 0: edge with states; its outputs are StateMasks of those states (no g arithmetic)
};
// entries of edges with states carry this bit in their offset: the offset then addresses the edge's own output state in u
// instead of a neighbour's output in the gather source (generic kernels only; such networks keep offsets below 2^30)
constexpr int ND_STATE_ENTRY_BIT = 1 << 30;

enum { MODE_DU = 0, MODE_AGG = 1, MODE_RK = 2 };
constexpr int EK_GENERIC = -1;   // several edge batches: per-entry batch id lookup

constexpr int HALO_MAX_WORLD = 8;
struct HaloParams {
  double* halo[HALO_MAX_WORLD];                    // [peer] -> that rank's halo buffer (this sequence's parity), peer-mapped
  unsigned long long* flags[HALO_MAX_WORLD];       // [peer] -> that rank's arrival-flag array
  const int* send_idx[HALO_MAX_WORLD];             // [peer] -> offsets (into src) of the outputs that peer reads, ascending
  long long send_n[HALO_MAX_WORLD];                //          how many
  long long dst_off[HALO_MAX_WORLD];               //          where this rank's block starts inside the peer's halo buffer
  int world, rank;
  unsigned long long seq;
  const double* src;                                // the owner's state vector (full layout, owned ranges valid)
  unsigned int* done_counter;                       // local: blocks finished (reset by the last block)
};

struct KParams {
  const int* __restrict__ rowptr;      // [nrows_owned+1], entries of owned rows, relative to row_base
  const int* __restrict__ nbr;         // per entry: offset of the neighbour's output in gsrc; ~offset when this row is the edge's src
  const int* __restrict__ epar;        // per entry: offset of the edge's parameter block in p (nullptr when no edge has parameters)
  const double* __restrict__ ppack;    // PK kernels: the edge parameters of entry j at ppack[j*PE ..], in the entry order of the
                                       // layout in use (filled by pack_params_kernel; nd_b200_pack_params)
  const uint8_t* __restrict__ ebid;    // per entry: edge batch id (EK_GENERIC only)
  const int* __restrict__ blk_row;     // [nblocks+1] first (global) row of each thread block
  const VBDev* __restrict__ vb;
  const EBDev* __restrict__ eb;
  int n_vb, n_eb;
  int row_base;                        // first owned row
  int long_thr;                        // rows with more entries are reduced by the whole block
  int gather_from_u;                   // 1: vertex outputs are read straight from u (all StateMask, vdepth 1)
  int state_edges;                     // 1: some edge batch has states (entries flagged with ND_STATE_ENTRY_BIT exist)
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
  const int4* __restrict__ tiles;      // per thread block {row0, e0, ne of a long row, ne | nrows<<16 | batch<<25 | long<<31}
  int ntiles;
  // split mode (edge pass + row pass): per entry the position of its value in the edge-output buffer
  const int* __restrict__ oidx;
  const double* __restrict__ oedge;    // edge part of the reference's output buffer `o` (src range, dst range per edge)
  // multi-GPU: the gather source is this rank's replica of the state vector, filled by the peers' publish kernels
  // through NVLink; the kernel waits until every peer's arrival flag has reached wait_seq (see halo_publish_kernel)
  const unsigned long long* wait_flags;
  unsigned long long wait_seq;
  int wait_world;
  int* wait_timeout;                   // set to 1 if a flag did not arrive within the spin budget
  int* wait_timeout_host;              // the same mark in mapped host memory (the entry points refuse further exchanges)
  long long wait_budget;               // spin budget in clock64 ticks
  // jagged ("warp-sliced") layout, see rhs_jag_kernel
  const int4* __restrict__ jslices;    // per 32-lane slice {entry base, first row, vertex batch, max parts of a split row}
  const uint16_t* __restrict__ jlanes; // per lane: len | rowrel<<6 (7 bits) | head<<13 | valid<<14
  const int* __restrict__ jnbr;        // per entry, slice-local column-major compacted order (PE == 0 kernels)
  const int2* __restrict__ jent;       // per entry {nbr, epar} (PE > 0 kernels)
  const uint8_t* __restrict__ jebid;   // per entry edge batch id (EK_GENERIC)
  const int4* __restrict__ jlong;      // rows reduced by a whole block {entry base, row, entries, vertex batch}
  int nslices, n_jag_blocks;
  // streamed jagged kernel (rhs_js_kernel): every warp of the (persistent) grid owns the contiguous slices [x, y)
  const int2* __restrict__ jwarp;
  int n_jwarps, n_jlong;
  // multi-GPU packed halo (jagged kernel only): gather offsets >= halo_base address the halo buffer that the peers'
  // publish kernels fill with exactly the remote vertex outputs this rank's rows read; slices >= wait_from_slice
  // (the ones that read remote outputs) wait for the arrival flags, interior slices run while the halo is in flight
  const double* halo;
  int halo_base;
  int wait_from;                       // first slice (jagged) / tile (tile kernel) that reads the halo
  int blk_off;                         // this launch covers thread blocks [blk_off, blk_off + gridDim.x) of the full grid
  // jagged kernels: every slice asks the L2 (cp.async.bulk.prefetch.L2, off the LSU path) for the piece of the entry streams
  // that lies pf_dist entries ahead of its own piece; 0 = off.  The slices' pieces tile the streams, so do the prefetches.
  int pf_dist;
  long long jag_len;                   // length of the jagged entry streams
  // column-blocked evaluation (graphs whose vertex outputs exceed the L2, see nd_b200_create): this launch adds the entries of ONE
  // block of neighbour columns; acc_in (nullable) holds the rows' sums over the blocks before it, [row * edepth + d]
  const double* acc_in;
  // multi-GPU: the first n_pub thread blocks of the grid pack this rank's boundary outputs into the peers' halo buffers
  // (NVLink stores) and raise the arrival flags; interior tiles follow, tiles that read the halo come last
  int n_pub;
  // 1: this rank has no tile / slice that waits for the peers' flags.  The double buffering of the halo relies on every
  // rank's call s+1 not COMPLETING before all peers have started theirs (a peer that races two calls ahead would overwrite
  // the buffer parity a slower reader is still gathering from), so such a launch carries one extra, last block that only
  // waits for the arrival flags.
  int fence;
  HaloParams H;
};

// parameters of the edge pass (split mode)
struct EParams {
  const int* __restrict__ esrc;        // per edge (in `o` order): gather offset of the src vertex output
  const int* __restrict__ edst;        // ... of the dst vertex output
  const int* __restrict__ epar;        // generic only: offset of the edge's parameters in p
  const int* __restrict__ eooff;       // generic only: offset of the edge's block in oedge
  const uint8_t* __restrict__ ebid;    // generic only: edge batch id
  const EBDev* __restrict__ eb;
  long long ne;
  long long p0;                        // single batch: parameters of edge k at p0 + k*PE
  int coupling0;                       // single batch: wrapper
  const double* __restrict__ gsrc;
  const double* __restrict__ p;
  double* __restrict__ oedge;
};

// ------------------------------------------------------------------------------------------------
// model arithmetic -- expression order exactly as the cited reference source (and as oracle/nd_oracle.c)
// ------------------------------------------------------------------------------------------------

// inner edge function: writes the DST output, g(odst, vsrc, vdst, p, t)
template <int VD, int ED>
__device__ __forceinline__ void edge_g_dst(int kind, double* odst, const double* vs, const double* vd,
                                           const double* __restrict__ pe, double t) {
  (void)t;
  switch (kind) {
    case ND_B200_E_DIFFUSION:      // test/ComponentLibrary.jl:8-10
      if constexpr (VD == 1 && ED == 1) odst[0] = pe[0] * (vs[0] - vd[0]);
      break;
    case ND_B200_E_DIFFUSION_NOP:  // benchmark/benchmark_models.jl:5-8
      if constexpr (VD == 1 && ED == 1) odst[0] = vs[0] - vd[0];
      break;
    case ND_B200_E_KURAMOTO:       // test/ComponentLibrary.jl:51-53
      if constexpr (VD == 1 && ED == 1) odst[0] = pe[0] * sin(vs[0] - vd[0]);
      break;
    case ND_B200_E_LINE_DQ:        // test/ComponentLibrary.jl:212-245: idst = active*1/Z*(Vsrc-Vdst), Z = R+jX
      if constexpr (VD == 2 && ED == 2) {
        double R = pe[0], X = pe[1], active = pe[2];
        double dr = vs[0] - vd[0];
        double di = vs[1] - vd[1];
        double den = R * R + X * X;
        odst[0] = active * ((R * dr + X * di) / den);
        odst[1] = active * ((R * di - X * dr) / den);
      }
      break;
    case ND_B200_E_LOOPBACK:       // LOOPBACK_G, src/post_utils.jl:105-108: outdst .= -1 .* insrc
      if constexpr (VD == ED) {
#pragma unroll
        for (int d = 0; d < ED; ++d) odst[d] = -1.0 * vs[d];
      }
      break;
    ND_CUSTOM_EDGE_CASES
    default:
#pragma unroll
      for (int d = 0; d < ED; ++d) odst[d] = 0.0;
  }
}

// value an entry contributes to ITS row: the dst output if the row is the edge's dst, else the src
// output produced by the wrapper (AntiSymmetric: -odst, Symmetric: odst; src/component_functions.jl:117-152;
// Fiducial: the edge's own two-sided g(osrc, odst, ...), :189-203 -- user-supplied kinds only)
template <int VD, int ED>
__device__ __forceinline__ void entry_value(int kind, int coupling, int side, const double* self,
                                            const double* xn, const double* __restrict__ pe, double t, double* val) {
  const double* vs = side ? self : xn;
  const double* vd = side ? xn : self;
  if (kind == ND_B200_E_LOOPBACK && side) {
    // apply_loopback! (src/coreloop.jl:47, src/post_utils.jl:213-234): the injector's input IS its hub's output.  The
    // engine stores it as the one entry of the injector's row (the loopback edge has no src output in `o`).
    if constexpr (VD == ED) {
#pragma unroll
      for (int d = 0; d < ED; ++d) val[d] = vd[d];
    }
    return;
  }
  if (coupling == ND_B200_FIDUCIAL) {
    double osrc[ED], odst[ED];
#pragma unroll
    for (int d = 0; d < ED; ++d) { osrc[d] = 0.0; odst[d] = 0.0; }
    switch (kind) {
      case ND_B200_E_DIFFUSION_FID:   // test/ComponentLibrary.jl:22-25
        if constexpr (VD == 1 && ED == 1) { odst[0] = pe[0] * (vs[0] - vd[0]); osrc[0] = -odst[0]; }
        break;
      ND_CUSTOM_EDGE_FID_CASES
      default: break;
    }
#pragma unroll
    for (int d = 0; d < ED; ++d) val[d] = side ? osrc[d] : odst[d];
    return;
  }
  edge_g_dst<VD, ED>(kind, val, vs, vd, pe, t);
  if (side && coupling == ND_B200_ANTISYMMETRIC) {
#pragma unroll
    for (int d = 0; d < ED; ++d) val[d] = -val[d];
  }
}

// entries of edges WITH states: the contribution is a StateMask read of the edge's own states (PASS 2, src/coreloop.jl:41;
// apply_compg(::PureStateMap), :230-233; wrappers src/component_functions.jl:117-203).  `off` addresses, inside u, the first
// state of the output this side of the edge shows (Fiducial(src=..., dst=...): two different masks; the other wrappers
// read the dst mask on both sides).
template <int ED>
__device__ __forceinline__ void state_entry_value(const double* __restrict__ u, int coupling, int side, int off, double* val) {
#pragma unroll
  for (int d = 0; d < ED; ++d) {
    const double x = u[off + d];
    val[d] = (side && coupling == ND_B200_ANTISYMMETRIC) ? -x : x;
  }
}

// edge f (PASS 4, edges with states): f(de, e, vsrc, vdst, p, t), src/coreloop.jl:76,194-218
template <int VD>
__device__ __forceinline__ void edge_f(int kind, double* de, const double* ue, const double* vs, const double* vd,
                                       const double* __restrict__ pe, double t, const double* ext = nullptr) {
  (void)t; (void)ext;
  switch (kind) {
    case ND_B200_E_DIFFUSION_ODE:   // test/ComponentLibrary.jl:30-34
      if constexpr (VD == 1) {
        const double tau = pe[0];
        de[0] = 1.0 / tau * (sin(vs[0] - vd[0]) - ue[0]);
        de[1] = 1.0 / tau * (sin(vd[0] - vs[0]) - ue[1]);
      }
      break;
    case ND_B200_E_RELAX_ODE:       // test/diffusion_test.jl:96-100
      if constexpr (VD == 1) {
        de[0] = vs[0] - vd[0] - ue[0];
        de[1] = vd[0] - vs[0] - ue[1];
      }
      break;
    ND_CUSTOM_EDGE_F_CASES
    default: break;
  }
}

// vertex g for non-StateMask models: NoFeedForward g(out,u,p,t)
__device__ __forceinline__ void vertex_g(int kind, int outdim, double* out, const double* v,
                                         const double* __restrict__ pv, double t, const double* ins = nullptr) {
  (void)t; (void)ins;
  switch (kind) {
    case ND_B200_V_SWING_DQ: {          // test/ComponentLibrary.jl:158-159
      double V = pv[3];
      out[0] = V * cos(v[0]);
      out[1] = V * sin(v[0]);
    } break;
    ND_CUSTOM_VERTEX_G_CASES
    default:                            // StateMask(1:outdim), src/component_functions.jl:81-99
      for (int k = 0; k < outdim; ++k) out[k] = v[k];
  }
}

// vertex f: f(dv, v, acc, p, t).  selfout = this vertex's own outputs (only used by SWING_DQ, whose f
// recomputes u_r,u_i with the same expressions as its g).
template <int VD, int ED>
__device__ __forceinline__ void vertex_f(int kind, double* dv, const double* v, const double* acc,
                                         const double* __restrict__ pv, const double* selfout, double t,
                                         const double* ext = nullptr) {
  (void)t; (void)selfout; (void)ext;
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
    ND_CUSTOM_VERTEX_F_CASES
    default: break;
  }
}

// PASS 6 for one row + the epilogue selected by mode.  v = the vertex's states, pv = its parameters
// (pointers into global or shared memory).
template <int VD, int ED>
__device__ __forceinline__ void vertex_phase(const KParams& P, const VBDev& B, int row, const double* acc_in,
                                             const double* selfout, const double* vin, const double* pv) {
  // multi-GPU: after a halo time-out the row sums were computed from stale or partial boundary outputs -- poison them so
  // that the result cannot be mistaken for a valid one
  double acc[ED];
#pragma unroll
  for (int d = 0; d < ED; ++d) acc[d] = acc_in[d];
  if (P.wait_timeout != nullptr) {
    if (*(volatile const int*)P.wait_timeout) {
#pragma unroll
      for (int d = 0; d < ED; ++d) acc[d] = acc[d] * (0.0 / 0.0) + (0.0 / 0.0);
    }
  }
  if (P.mode == MODE_AGG) {
#pragma unroll
    for (int d = 0; d < ED; ++d) P.aggbuf[(long long)row * ED + d] = acc[d];
    return;
  }
  const long long i = row - B.row0;
  const long long s = B.state0 + i * B.dim;
  const int dim = B.dim;                // <= ND_MAX_VDIM
  double v[ND_MAX_VDIM], dv[ND_MAX_VDIM];
#pragma unroll
  for (int c = 0; c < ND_MAX_VDIM; ++c) { v[c] = c < dim ? vin[c] : 0.0; dv[c] = 0.0; }
#if ND_MAX_EXT > 0
  double ext[ND_MAX_EXT];
  gather_ext(B.ext + i * B.extdim, B.extdim, P.u, P.gsrc, ext);
  vertex_f<VD, ED>(B.kind, dv, v, acc, pv, selfout, P.t, ext);
#else
  vertex_f<VD, ED>(B.kind, dv, v, acc, pv, selfout, P.t);
#endif
  if (P.mode == MODE_DU) {
    if (ND_MAX_VDIM == 2 && dim == 2 && ((s & 1) == 0)) {
      *reinterpret_cast<double2*>(P.du + s) = make_double2(dv[0], dv[ND_MAX_VDIM > 1 ? 1 : 0]);
    } else {
#pragma unroll
      for (int c = 0; c < ND_MAX_VDIM; ++c)
        if (c < dim) P.du[s + c] = dv[c];
    }
    return;
  }
  // MODE_RK: classical RK4, operation order of the CPU restatement's rk4:
  //   u <- u + (dt/6)*(((k1 + 2k2) + 2k3) + k4)
  double un[ND_MAX_VDIM];
#pragma unroll
  for (int c = 0; c < ND_MAX_VDIM; ++c) {
    un[c] = 0.0;
    if (c >= dim) continue;
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
  if (!P.gather_from_u && P.vout_next) {      // absent: the next stage runs the output pre-pass itself
    double out[VD];
    vertex_g(B.kind, VD, out, un, pv, P.t);
#pragma unroll
    for (int k = 0; k < VD; ++k) P.vout_next[(long long)row * VD + k] = out[k];
  }
}

// loads the (<= ND_MAX_VDIM) states of a row from global memory
__device__ __forceinline__ void load_vertex_state(const KParams& P, const VBDev& B, int row, double* v) {
  const long long s = B.state0 + (long long)(row - B.row0) * B.dim;
#pragma unroll
  for (int c = 0; c < ND_MAX_VDIM; ++c) v[c] = c < B.dim ? P.u[s + c] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// fused gather -> edge -> ordered row reduce -> vertex kernel
// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
// multi-GPU halo exchange over NVLink peer memory (no NCCL on the data path)
//
// Every rank owns contiguous state ranges.  halo_publish_kernel copies the owner's ranges into the state replica of
// EVERY rank (its own and, through peer-mapped pointers, all others: plain st.global over NVLink 5 / NVSwitch), then
// -- after a system-scope fence -- the last block to finish stores the sequence number into slot [owner] of every
// rank's arrival-flag array.  The consuming RHS kernel (next in the peer's stream) spins on its LOCAL flag array until
// all `world` slots have reached the sequence number, then gathers from its LOCAL replica.  Replicas are double-buffered
// by sequence parity, which is enough: a rank can only publish sequence s+2 after it has consumed every peer's s+1,
// which those peers published after finishing their own reads of s.
// ------------------------------------------------------------------------------------------------
#ifdef ND_CUSIM   // tests/cusim (CPU emulation of these kernels for the CPU test suite; never part of the product build)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// Programmatic dependent launch (the stage kernels of nd_b200_rk4's captured graph): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may become resident while its predecessor in the stream is still running;
// pdl_wait() (griddepcontrol.wait) returns once the predecessor has completed and its writes are visible -- it must precede the
// first read of anything the predecessor wrote (states, outputs, k-sums); pdl_launch_dependents() lets the successor start
// launching.  Both are no-ops in a launch without the attribute.
#ifdef ND_CUSIM
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// pack + publish, executed by the first n_pub thread blocks of the RHS grid: for every peer, gather the outputs it
// needs from the owner's state vector (ascending offsets: the reads are nearly coalesced) and store them contiguously
// into the peer's halo buffer (coalesced NVLink stores).  The last publishing block to finish raises this rank's arrival
// flag on every rank.  Publishing blocks never wait on anything and are scheduled before every tile of the same grid,
// so tiles that spin on the peers' flags cannot starve them.
__device__ __forceinline__ void publish_block(const HaloParams& H, int bid, int nblocks) {
  const long long tid = (long long)bid * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)nblocks * blockDim.x;
  // peers in rotated order (rank+1, rank+2, ...): at every moment each destination receives from ONE source instead of all
  // ranks storing into rank 0 first, then rank 1, ... (N=8, config 2: 100 us of exposed halo for 46 us of wire time before)
  for (int k = 1; k < H.world; ++k) {
    const int r = (H.rank + k) % H.world;
    const long long n = H.send_n[r];
    const int* __restrict__ idx = H.send_idx[r];
    double* __restrict__ dst = H.halo[r] + H.dst_off[r];
    long long i = tid;
    for (; i + 3 * nthreads < n; i += 4 * nthreads) {   // four independent gathers in flight per thread
      const int i0 = idx[i], i1 = idx[i + nthreads], i2 = idx[i + 2 * nthreads], i3 = idx[i + 3 * nthreads];
      const double v0 = H.src[i0], v1 = H.src[i1], v2 = H.src[i2], v3 = H.src[i3];
      dst[i] = v0; dst[i + nthreads] = v1; dst[i + 2 * nthreads] = v2; dst[i + 3 * nthreads] = v3;
    }
    for (; i < n; i += nthreads) dst[i] = H.src[idx[i]];
  }
  // make this block's peer stores visible system-wide, then count the block as done
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(H.done_counter, 1u);
    if (prev == (unsigned int)nblocks - 1u) {
      *H.done_counter = 0;
      __threadfence_system();
      for (int r = 0; r < H.world; ++r) st_release_sys(H.flags[r] + H.rank, H.seq);
    }
  }
}

// called at the top of the consuming kernels
__device__ __forceinline__ void halo_wait(const KParams& P) {
  if (P.wait_flags == nullptr) return;
  if (threadIdx.x < P.wait_world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(P.wait_flags + threadIdx.x) < P.wait_seq) {
      // budget exhausted (default 30 s, ND_B200_HALO_TIMEOUT_MS): a peer died or the ranks' call sequences diverged; do not
      // hang the GPU.  The mark is sticky: once set, later tiles and calls give up at once, every row of such a launch
      // writes NaN (vertex_phase) and the exchanging entry points return ND_B200_ETIMEOUT.
      if (*(volatile int*)P.wait_timeout || clock64() - t0 > P.wait_budget) {
        *P.wait_timeout = 1;
        if (P.wait_timeout_host) *(volatile int*)P.wait_timeout_host = 1;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
}

// warp-level variant for the jagged kernel: only the warps whose slice reads remote outputs wait
__device__ __forceinline__ void halo_wait_warp(const KParams& P) {
  if (P.wait_flags == nullptr) return;
  const int lane = threadIdx.x & 31;
  if (lane < P.wait_world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(P.wait_flags + lane) < P.wait_seq) {
      if (*(volatile int*)P.wait_timeout || clock64() - t0 > P.wait_budget) {
        *P.wait_timeout = 1;
        if (P.wait_timeout_host) *(volatile int*)P.wait_timeout_host = 1;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncwarp();
}

// gather source of an offset: the state vector (or materialised vertex outputs) below halo_base, the halo buffer above
template <bool HALO>
__device__ __forceinline__ const double* gather_ptr(const KParams& P, int off) {
  if constexpr (HALO) return (off >= P.halo_base ? P.halo - P.halo_base : P.gsrc) + off;
  else return P.gsrc + off;
}

// Occupancy is the lever for this kernel (it is bound by the L2 sector bandwidth of the random gathers and hides
// latency with many independent blocks), so registers are capped through the min-blocks launch bound.
// Measured on B200 (profiles/r01_tuning.md): 64 resident warps/SM (32 registers) is best for the arithmetic-free
// diffusion kernels, 48 warps/SM (40 registers) for the kernels that evaluate sin / complex division.
__host__ __device__ constexpr int fused_warps_per_sm(int ek, bool pk = false) {
  // packed-parameter kernels hold their EPT parameters in registers from step (1) on (the stream's DRAM latency overlaps
  // the gather instead of following the second barrier): 40 registers
  return ((ek == ND_B200_E_DIFFUSION || ek == ND_B200_E_DIFFUSION_NOP) && !pk) ? 64 : 48;
}
// HALO = true: the multi-GPU variant (publishing blocks, flag waits, gathers from [u | halo]); the single-GPU variant
// carries none of it (the 32-register diffusion kernels lose 9-19 % when that code shares their register allocation).
// PK = true: edge parameters come from the engine's packed per-entry copy (coalesced with the index stream; no parameter
// offsets are read) -- valid while the caller guarantees p is unchanged since nd_b200_pack_params (nd_b200_rk4 packs per call).
// CR = true ("compact rows", networks whose gather offsets stay below 2^23, single edge batch): an entry word is
//   offset | local row << 23 | side << 30
// so the entry threads know their row without the per-tile row-id table in shared memory (its byte scatter by the row
// threads and the per-entry look-up were 1.7 M of the 5.5 M shared-memory wavefronts on config 2, profiles/r02c).
constexpr int ND_CR_OFF_BITS = 23;
template <int VD, int ED, int EK, int PE, int BLOCK, int EPT, bool HALO, bool PK = false, bool CR = false>
__global__ void __launch_bounds__(BLOCK, (fused_warps_per_sm(EK, PK && PE == 1) * 32) / BLOCK) rhs_fused_kernel(const __grid_constant__ KParams P) {
  static_assert(!CR || (EK != EK_GENERIC && BLOCK <= 128), "compact rows: 7 bits of local row, no state-entry flag");
  constexpr int TILE = BLOCK * EPT;
  static_assert(BLOCK <= 256, "row ids are stored as uint8");
  // entry jj of local row r is staged at slot jj + r: one slot of padding per row.  The ordered row sums of step (6) read
  // slot a_r + j + r from lane r; with row starts a_r ~ 8 r (mean degree 8) the unpadded layout puts the 16 lanes of a
  // half warp on two 8-byte banks (8-way conflict, profiles/r01b: a third of the kernel's L1TEX wavefronts were shared
  // memory); the skew spreads them over all 16.
  // (only in the 40-register packed-parameter kernels: in the 32-register ones the extra index arithmetic spills)
  constexpr int SKEW = (PK && PE == 1) ? 1 : 0;
  __shared__ double s_val[(TILE + SKEW * BLOCK) * ED];
  __shared__ double s_self[BLOCK * VD];
  __shared__ int s_rp[BLOCK + 1];
  __shared__ uint8_t s_rowid[CR ? 4 : TILE];

  const int tid = threadIdx.x;
  if constexpr (HALO) {
    if ((int)blockIdx.x < P.n_pub) { publish_block(P.H, blockIdx.x, P.n_pub); return; }
    if (P.fence && blockIdx.x == gridDim.x - 1) { halo_wait(P); return; }
  }
  // one 16-byte descriptor per thread block: {row0, e0, ne (long rows), ne | nrows<<16 | batch<<25 | long<<31}
  const int bid = (int)blockIdx.x - (HALO ? P.n_pub : 0) + P.blk_off;
  const int4 d = __ldg(&P.tiles[bid]);
  const int r0 = d.x, e0 = d.y;
  const bool is_long = d.w < 0;
  const int nrows = (d.w >> 16) & 0x1FF;
  const int ne = is_long ? d.z : (d.w & 0xFFFF);
  const VBDev B = P.vb[(d.w >> 25) & 0x3F];
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;
  if constexpr (HALO) {
    if (bid >= P.wait_from) halo_wait(P);   // this tile reads the halo (block-uniform)
  } else {
    pdl_launch_dependents();
    pdl_wait();                             // everything above reads engine-owned tables only
  }

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
      int side;
      if constexpr (CR) { side = (nb >> 30) & 1; nb &= (1 << ND_CR_OFF_BITS) - 1; }
      else { side = nb < 0; nb = side ? ~nb : nb; }
      bool st = false;
      if constexpr (EK == EK_GENERIC) { st = P.state_edges && (nb & ND_STATE_ENTRY_BIT); nb &= ~ND_STATE_ENTRY_BIT; }
      double xn[VD];
#pragma unroll
      for (int k = 0; k < VD; ++k) xn[k] = 0.0;
      if (!st) {
        const double* gp = gather_ptr<HALO>(P, nb);
#pragma unroll
        for (int k = 0; k < VD; ++k) xn[k] = gp[k];
      }
      const double* pe = P.p;
      if constexpr (PE > 0) pe = PK ? P.ppack + (long long)(e0 + jj) * PE : P.p + P.epar[e0 + jj];
      int kind = EK, coupling = coupling0;
      if constexpr (EK == EK_GENERIC) {
        const EBDev E = P.eb[P.ebid[e0 + jj]];
        kind = E.kind; coupling = E.coupling;
      }
      double val[ED];
      if (st) state_entry_value<ED>(P.u, coupling, side, nb, val);
      else entry_value<VD, ED>(kind, coupling, side, self, xn, pe, P.t, val);
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
      double acc[ED], v[ND_MAX_VDIM];
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
  double pk1[(PK && PE == 1) ? EPT : 1];       // packed single parameters: loaded with the indices
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    nb[k] = 0; ep[k] = 0;
    if constexpr (PK && PE == 1) pk1[k] = 0.0;
    if (jj < ne) {
      nb[k] = P.nbr[e0 + jj];
      if constexpr (PE > 0 && !PK) ep[k] = P.epar[e0 + jj];
      if constexpr (PK && PE == 1) pk1[k] = P.ppack[e0 + jj];
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
  // (3) entry -> local row map (compact rows: carried by the entry word)
  if constexpr (!CR) {
    if (tid < nrows) {
      const int a = s_rp[tid], z = s_rp[tid + 1];
      for (int jj = a; jj < z; ++jj) s_rowid[jj] = (uint8_t)tid;
    }
  }
  // (4) the gather: EPT independent random reads per thread
  double xn[EPT][VD];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    const int off = CR ? (nb[k] & ((1 << ND_CR_OFF_BITS) - 1)) : (nb[k] < 0 ? ~nb[k] : nb[k]);
#pragma unroll
    for (int q = 0; q < VD; ++q) xn[k][q] = 0.0;
    bool st = false;   // entry of an edge with states: read in step (5) from u
    if constexpr (EK == EK_GENERIC) st = P.state_edges && (off & ND_STATE_ENTRY_BIT);
    if (jj < ne && !st) {
      const double* gp = gather_ptr<HALO>(P, off);
      if constexpr (VD == 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(gp);
        xn[k][0] = t2.x; xn[k][1] = t2.y;
      } else {
#pragma unroll
        for (int q = 0; q < VD; ++q) xn[k][q] = gp[q];
      }
    }
  }
  __syncthreads();
  // (5) edge evaluation, one entry per thread per k
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    if (jj < ne) {
      const int side = CR ? ((nb[k] >> 30) & 1) : (nb[k] < 0);
      const int r = CR ? ((nb[k] >> ND_CR_OFF_BITS) & 127) : (int)s_rowid[jj];
      double self[VD];
#pragma unroll
      for (int q = 0; q < VD; ++q) self[q] = s_self[r * VD + q];
      int kind = EK, coupling = coupling0;
      bool st = false;
      int off = CR ? 0 : (side ? ~nb[k] : nb[k]);
      if constexpr (EK == EK_GENERIC) {
        const EBDev E = P.eb[P.ebid[e0 + jj]];
        kind = E.kind; coupling = E.coupling;
        st = P.state_edges && (off & ND_STATE_ENTRY_BIT);
        off &= ~ND_STATE_ENTRY_BIT;
      }
      double val[ED];
      if (st) state_entry_value<ED>(P.u, coupling, side, off, val);
      else entry_value<VD, ED>(kind, coupling, side, self, xn[k], (PK && PE == 1) ? &pk1[k] : (PK ? P.ppack + (long long)(e0 + jj) * PE : P.p + ep[k]), P.t, val);
#pragma unroll
      for (int q = 0; q < ED; ++q) s_val[(jj + SKEW * r) * ED + q] = val[q];
    }
  }
  __syncthreads();
  // (6) ordered per-row accumulation (the reference's sequential order) + vertex model
  if (tid < nrows) {
    double acc[ED];
#pragma unroll
    for (int q = 0; q < ED; ++q) acc[q] = 0.0;
    const int a = s_rp[tid] + SKEW * tid, z = s_rp[tid + 1] + SKEW * tid;
    for (int jj = a; jj < z; ++jj) {
#pragma unroll
      for (int q = 0; q < ED; ++q) acc[q] = acc[q] + s_val[jj * ED + q];
    }
    double self[VD], v[ND_MAX_VDIM];
#pragma unroll
    for (int q = 0; q < VD; ++q) self[q] = s_self[tid * VD + q];
    load_vertex_state(P, B, r0 + tid, v);
    vertex_phase<VD, ED>(P, B, r0 + tid, acc, self, v, P.p + B.p0 + (long long)(r0 + tid - B.row0) * B.pdim);
  }
}

// ppack[j*pe + k] = p[off_j + k]: the edge parameters in the entry order of the layout in use.  `offs` is the layout's own
// per-entry offset stream (tile layout: epar, stride 1; jagged layout: the .y of the {nbr, epar} pairs, stride 2).
__global__ void pack_params_kernel(const int* __restrict__ offs, int stride, int first, long long n, int pe,
                                   const double* __restrict__ p, double* __restrict__ ppack) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long o = offs[j * stride + first];
  for (int k = 0; k < pe; ++k) ppack[j * pe + k] = p[o + k];
}

// Standalone aggregate!(aggregator, aggbuf, o) (src/aggregators.jl:140-151): every aggregation slot adds its entries of the
// edge-output block of `o` in ascending `o` order, starting from the slot's present content (additive, test/
// aggregators_test.jl:69-79).  One thread per (row, component): the left-to-right order of the sequential sweep.
__global__ void aggregate_kernel(const int* __restrict__ rowptr, const long long* __restrict__ oidx, int ed, long long nrows,
                                 const double* __restrict__ o, double* __restrict__ aggbuf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * ed) return;
  const long long row = i / ed;
  const int q = (int)(i - row * ed);
  double acc = aggbuf[i];
  for (int j = rowptr[row]; j < rowptr[row + 1]; ++j) {
    const long long oi = oidx[j];
    if (oi >= 0) acc = acc + o[oi + q];
  }
  aggbuf[i] = acc;
}

__global__ void extract_nbr_kernel(const int* __restrict__ pairs, long long n, int* __restrict__ nbr) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) nbr[j] = pairs[2 * j];
}

// PASS 1 for networks whose vertex outputs are not plain state copies: vout[row*VD + k] = g_v(u, p)
#ifndef ND_MAX_VOUT
#define ND_MAX_VOUT 2            // largest vertex output dimension (vdepth)
#endif
// phase 0: vertices without feed forward (src/coreloop.jl:39); phase 1: feed-forward vertices, after the loopback copy
// (src/coreloop.jl:47,55) -- their g reads the hub's output that phase 0 wrote
__global__ void vertex_out_kernel(const VBDev* __restrict__ vb, int n_vb, int vd, const double* __restrict__ u,
                                  const double* __restrict__ p, double* __restrict__ vout, int nrows_total, double t, int phase) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows_total) return;
  int b = 0;
  for (int i = 1; i < n_vb; ++i)
    if (row >= vb[i].row0) b = i;
  const VBDev B = vb[b];
  if ((B.ff != 0) != (phase != 0)) return;
  const long long i = row - B.row0;
  double v[ND_MAX_VDIM], out[ND_MAX_VOUT], ins[ND_MAX_VOUT];
  for (int c = 0; c < ND_MAX_VDIM; ++c) v[c] = c < B.dim ? u[B.state0 + i * B.dim + c] : 0.0;
  for (int k = 0; k < ND_MAX_VOUT; ++k) { out[k] = 0.0; ins[k] = 0.0; }
  if (B.ff)
    for (int k = 0; k < vd; ++k) ins[k] = vout[(long long)B.ffin[i] + k];
  vertex_g(B.kind, vd, out, v, p + B.p0 + i * B.pdim, t, ins);
  for (int k = 0; k < vd; ++k) vout[(long long)row * vd + k] = out[k];
}

// get_buffers support: edge outputs into the reference's `o` layout (src range first, dst right after)
template <int VD, int ED>
__global__ void edge_out_kernel(int kind, int coupling, int pdim, int osrc, long long count,
                                const int* __restrict__ esrc_off, const int* __restrict__ edst_off,
                                long long p0, long long out0, const double* __restrict__ gsrc,
                                const double* __restrict__ p, double* __restrict__ o, double t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double vs[VD], vd[VD], val[ED];
#pragma unroll
  for (int k = 0; k < VD; ++k) { vs[k] = gsrc[(long long)esrc_off[i] + k]; vd[k] = gsrc[(long long)edst_off[i] + k]; }
  double* oo = o + out0 + i * (osrc + ED);
  if (coupling == ND_B200_FIDUCIAL) {     // the edge's own two-sided g: both outputs through entry_value
    double sv[ED];
    entry_value<VD, ED>(kind, coupling, 1, vs, vd, p + p0 + i * pdim, t, sv);    // side 1: self = src, neighbour = dst
    entry_value<VD, ED>(kind, coupling, 0, vd, vs, p + p0 + i * pdim, t, val);   // side 0: self = dst, neighbour = src
#pragma unroll
    for (int d = 0; d < ED; ++d) { oo[d] = sv[d]; oo[osrc + d] = val[d]; }
    return;
  }
  edge_g_dst<VD, ED>(kind, val, vs, vd, p + p0 + i * pdim, t);
  if (osrc) {
#pragma unroll
    for (int d = 0; d < ED; ++d) oo[d] = (coupling == ND_B200_ANTISYMMETRIC) ? -val[d] : val[d];
  }
#pragma unroll
  for (int d = 0; d < ED; ++d) oo[osrc + d] = val[d];
}


// ------------------------------------------------------------------------------------------------
// edges with states ("ODE edges", src/coreloop.jl:41,76): their outputs are StateMasks of their own states -- the row
// kernels read them straight from u (state_entry_value) -- and their f runs here, one thread per edge of a batch:
//   du_e = f(u_e, v_src, v_dst, p_e, t)     with the same epilogues as the vertex phase (du / fused RK4 stage)
// States and parameters of a batch are contiguous (coalesced), the two vertex outputs are gathered.
// ------------------------------------------------------------------------------------------------
struct EFParams {
  int kind, dim, pdim;
  long long count, state0, p0;
  const int* __restrict__ esrc_off;    // per edge of the batch: gather offset of the src / dst vertex output
  const int* __restrict__ edst_off;
  const int* __restrict__ ext;         // external inputs of the batch (see VBDev), extdim codes per edge
  int extdim;
  const double* __restrict__ u;        // state vector (stage input)
  const double* __restrict__ gsrc;     // gather source of vertex outputs
  const double* __restrict__ p;
  double* __restrict__ du;
  int mode, stage;
  const double* __restrict__ u0;
  double* __restrict__ unext;
  double* __restrict__ ksum;
  double hs, h6, t;
};

template <int VD>
__global__ void __launch_bounds__(256) edge_f_kernel(const __grid_constant__ EFParams Q) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q.count) return;
  double vs[VD], vd[VD], ue[ND_MAX_EDIM], de[ND_MAX_EDIM];
#pragma unroll
  for (int k = 0; k < VD; ++k) { vs[k] = Q.gsrc[(long long)Q.esrc_off[i] + k]; vd[k] = Q.gsrc[(long long)Q.edst_off[i] + k]; }
  const long long s = Q.state0 + i * Q.dim;
#pragma unroll
  for (int c = 0; c < ND_MAX_EDIM; ++c) { ue[c] = c < Q.dim ? Q.u[s + c] : 0.0; de[c] = 0.0; }
#if ND_MAX_EXT > 0
  double ext[ND_MAX_EXT];
  gather_ext(Q.ext + i * Q.extdim, Q.extdim, Q.u, Q.gsrc, ext);
  edge_f<VD>(Q.kind, de, ue, vs, vd, Q.p + Q.p0 + i * Q.pdim, Q.t, ext);
#else
  edge_f<VD>(Q.kind, de, ue, vs, vd, Q.p + Q.p0 + i * Q.pdim, Q.t);
#endif
#pragma unroll
  for (int c = 0; c < ND_MAX_EDIM; ++c) {
    if (c >= Q.dim) continue;
    const long long idx = s + c;
    if (Q.mode == MODE_DU) { Q.du[idx] = de[c]; continue; }
    // MODE_RK: same stage algebra as vertex_phase
    double un;
    if (Q.stage == 1) {
      Q.ksum[idx] = de[c];
      un = ue[c] + Q.hs * de[c];
    } else if (Q.stage < 4) {
      Q.ksum[idx] = Q.ksum[idx] + 2.0 * de[c];
      un = Q.u0[idx] + Q.hs * de[c];
    } else {
      un = Q.u0[idx] + Q.h6 * (Q.ksum[idx] + de[c]);
    }
    Q.unext[idx] = un;
  }
}

// get_buffers support: outputs of a batch of edges with states in the reference's `o` layout (src range, dst range)
__global__ void edge_mask_out_kernel(int coupling, int dim, int osrc, int odst, int mask_src, int mask_dst, long long count,
                                     long long state0, long long out0, const double* __restrict__ u, double* __restrict__ o) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const double* ue = u + state0 + i * dim;
  double* oo = o + out0 + i * (osrc + odst);
  for (int d = 0; d < odst; ++d) oo[osrc + d] = ue[mask_dst + d];
  for (int d = 0; d < osrc; ++d)
    oo[d] = coupling == ND_B200_FIDUCIAL ? ue[mask_src + d] : (coupling == ND_B200_ANTISYMMETRIC ? -ue[mask_dst + d] : ue[mask_dst + d]);
}

// ------------------------------------------------------------------------------------------------
// split mode ("edge once"): PASS 5 and the aggregation as two kernels around the reference's own
// edge-output buffer.
//
// Optional (ND_B200_KERNEL=split).  The fused kernel evaluates every edge from both endpoints: per undirected edge
// 2 random neighbour reads + 1 random parameter read and 2 evaluations of g.  Here each edge is evaluated once.
// Measured on B200 (profiles/r01_tuning.md): slower than the fused kernel whenever u fits in L2 (the random reads
// move from the 8 MB state vector to the 64 MB edge-output buffer), ~9 % faster when nothing fits (config 5 scale).
//   edge_pass_kernel : one thread per edge in `o` order; src output nearly coalesced (edges are sorted by
//                      src), dst output random, parameters coalesced; writes [osrc | odst] exactly like
//                      PASS 5 (src/coreloop.jl:78) -> 1 random request and 1 evaluation of g per edge;
//   row_pass_kernel  : ordered per-row sum over the entries' positions in `o` (the SequentialAggregator
//                      sweep restricted to one row, src/aggregators.jl:140-151) + vertex model; entries that
//                      are this row's src outputs are consecutive in `o` -> coalesce; dst outputs random
//                      -> 1 random request per edge.
// ------------------------------------------------------------------------------------------------
template <int VD, int ED, int EK, int PE, int BLOCK, int EPT>
__global__ void __launch_bounds__(BLOCK) edge_pass_kernel(const __grid_constant__ EParams Q) {
  const long long base = (long long)blockIdx.x * (BLOCK * EPT) + threadIdx.x;
  int so[EPT], to[EPT];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const long long e = base + (long long)k * BLOCK;
    so[k] = 0; to[k] = 0;
    if (e < Q.ne) { so[k] = Q.esrc[e]; to[k] = Q.edst[e]; }
  }
  double vs[EPT][VD], vd[EPT][VD];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const long long e = base + (long long)k * BLOCK;
#pragma unroll
    for (int q = 0; q < VD; ++q) { vs[k][q] = 0.0; vd[k][q] = 0.0; }
    if (e < Q.ne) {
      if constexpr (VD == 2) {
        const double2 a = *reinterpret_cast<const double2*>(Q.gsrc + so[k]);
        const double2 b = *reinterpret_cast<const double2*>(Q.gsrc + to[k]);
        vs[k][0] = a.x; vs[k][1] = a.y; vd[k][0] = b.x; vd[k][1] = b.y;
      } else {
        vs[k][0] = Q.gsrc[so[k]]; vd[k][0] = Q.gsrc[to[k]];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const long long e = base + (long long)k * BLOCK;
    if (e >= Q.ne) continue;
    int kind = EK, coupling = Q.coupling0;
    const double* pe = Q.p + Q.p0 + e * PE;
    long long ooff;
    if constexpr (EK == EK_GENERIC) {
      const EBDev E = Q.eb[Q.ebid[e]];
      kind = E.kind; coupling = E.coupling;
      pe = Q.p + Q.epar[e];
      ooff = Q.eooff[e];
    } else {
      ooff = e * (coupling == ND_B200_DIRECTED ? ED : 2 * ED);
    }
    double val[ED];
    edge_g_dst<VD, ED>(kind, val, vs[k], vd[k], pe, 0.0);
    double* oo = Q.oedge + ooff;
    if (coupling == ND_B200_DIRECTED) {
#pragma unroll
      for (int q = 0; q < ED; ++q) oo[q] = val[q];
    } else {
      // AntiSymmetric: osrc = -odst; Symmetric: osrc = odst (src/component_functions.jl:117-152)
      double sv[ED];
#pragma unroll
      for (int q = 0; q < ED; ++q) sv[q] = (coupling == ND_B200_ANTISYMMETRIC) ? -val[q] : val[q];
      if constexpr (EK == EK_GENERIC) {     // mixed wrappers: blocks of width ED (Directed) shift the alignment
#pragma unroll
        for (int q = 0; q < ED; ++q) { oo[q] = sv[q]; oo[ED + q] = val[q]; }
      } else if constexpr (ED == 1) {
        *reinterpret_cast<double2*>(oo) = make_double2(sv[0], val[0]);
      } else {
        *reinterpret_cast<double2*>(oo) = make_double2(sv[0], sv[1]);
        *reinterpret_cast<double2*>(oo + 2) = make_double2(val[0], val[1]);
      }
    }
  }
}

template <int VD, int ED, int BLOCK, int EPT>
__global__ void __launch_bounds__(BLOCK, 2048 / BLOCK) row_pass_kernel(const __grid_constant__ KParams P) {
  constexpr int TILE = BLOCK * EPT;
  __shared__ double s_val[TILE * ED];
  __shared__ int s_rp[BLOCK + 1];
  const int tid = threadIdx.x;
  const int4 d = __ldg(&P.tiles[blockIdx.x]);
  const int r0 = d.x, e0 = d.y;
  const bool is_long = d.w < 0;
  const int nrows = (d.w >> 16) & 0x1FF;
  const int ne = is_long ? d.z : (d.w & 0xFFFF);
  const VBDev B = P.vb[(d.w >> 25) & 0x3F];

  if (is_long) {   // hub row: strided partial sums + fixed-shape tree
    double part[ED];
#pragma unroll
    for (int q = 0; q < ED; ++q) part[q] = 0.0;
    for (int jj = tid; jj < ne; jj += BLOCK) {
      const int oi = P.oidx[e0 + jj];
#pragma unroll
      for (int q = 0; q < ED; ++q) part[q] = part[q] + P.oedge[(long long)oi + q];
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
      double acc[ED], v[ND_MAX_VDIM], self[VD];
#pragma unroll
      for (int q = 0; q < ED; ++q) acc[q] = s_val[q];
#pragma unroll
      for (int q = 0; q < VD; ++q) self[q] = P.gather_from_u ? 0.0 : P.gsrc[(long long)r0 * VD + q];
      load_vertex_state(P, B, r0, v);
      vertex_phase<VD, ED>(P, B, r0, acc, self, v, P.p + B.p0 + (long long)(r0 - B.row0) * B.pdim);
    }
    return;
  }

  int oi[EPT];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    oi[k] = jj < ne ? P.oidx[e0 + jj] : 0;
  }
  if (tid < nrows) s_rp[tid] = P.rowptr[(r0 - P.row_base) + tid] - e0;
  if (tid == 0) s_rp[nrows] = ne;
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int jj = k * BLOCK + tid;
    if (jj < ne) {
      if constexpr (ED == 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(P.oedge + oi[k]);
        s_val[jj * 2] = t2.x; s_val[jj * 2 + 1] = t2.y;
      } else {
        s_val[jj] = P.oedge[oi[k]];
      }
    }
  }
  __syncthreads();
  if (tid < nrows) {
    double acc[ED];
#pragma unroll
    for (int q = 0; q < ED; ++q) acc[q] = 0.0;
    const int a = s_rp[tid], z = s_rp[tid + 1];
    for (int jj = a; jj < z; ++jj) {
#pragma unroll
      for (int q = 0; q < ED; ++q) acc[q] = acc[q] + s_val[jj * ED + q];
    }
    double self[VD], v[ND_MAX_VDIM];
#pragma unroll
    for (int q = 0; q < VD; ++q) self[q] = P.gather_from_u ? 0.0 : P.gsrc[(long long)(r0 + tid) * VD + q];
    load_vertex_state(P, B, r0 + tid, v);
    vertex_phase<VD, ED>(P, B, r0 + tid, acc, self, v, P.p + B.p0 + (long long)(r0 + tid - B.row0) * B.pdim);
  }
}

// ------------------------------------------------------------------------------------------------
// jagged ("warp-sliced") fused kernel: gather -> edge -> ordered row sum -> vertex model with NO shared memory
// and NO block barrier on the regular path.
//
// Why: ncu on rhs_fused_kernel (profiles/r01b_fused_cfg2_ncu_summary.txt) shows the L1TEX data pipe as the busiest
// unit (78-84 %), one third of its wavefronts being the shared-memory round trip of the edge values (entry -> row map,
// value store, conflict-prone per-row reads) -- traffic that exists only to turn an entry-parallel gather into a
// row-ordered sum.  Here one LANE owns one row (thread-per-row, so the sum is a register accumulated in the reference's
// sequential order, src/aggregators.jl:140-151) and the index stream is laid out so that the lanes' loads still
// coalesce: a slice = 32 lanes; its entries are stored column-major and COMPACTED -- column j holds the j-th entry of
// every lane that has more than j entries, in lane order.  A lane finds its slot with one ballot + popc per column.
// No padding is stored, rows keep their natural order (coalesced u / du / vertex parameters).
//
// Rows with more than `jsplit` (<= 63, default 32) entries are cut into consecutive parts of jsplit entries that sit
// on consecutive lanes of the same slice; the head lane adds the parts' sums in order with shuffles (deterministic;
// differs from the sequential sum only in association, like the block tree of the long rows).  Rows that would need
// more than 32 lanes go to the whole-block path (same code as rhs_fused_kernel's long rows).
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int jag_warps_per_sm_default(int ek) {
  return (ek == ND_B200_E_DIFFUSION || ek == ND_B200_E_DIFFUSION_NOP) ? 64 : 48;
}

template <int VD, int ED, int EK, int PE, int BLOCK, bool HALO, bool PK>
__device__ __forceinline__ void long_row_block(const KParams& P, const int4 d, double* s_val) {
  const int tid = threadIdx.x;
  const int e0 = d.x, r0 = d.y, ne = d.z;
  const VBDev B = P.vb[d.w];
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;
  double self[VD];
  const long long sidx = P.gather_from_u ? (B.state0 + (long long)(r0 - B.row0) * B.dim) : (long long)r0 * VD;
#pragma unroll
  for (int k = 0; k < VD; ++k) self[k] = P.gsrc[sidx + k];
  double part[ED];
#pragma unroll
  for (int q = 0; q < ED; ++q) part[q] = 0.0;
  for (int jj = tid; jj < ne; jj += BLOCK) {
    int nb, ep = 0;
    if constexpr (PE > 0 && !PK) { const int2 t2 = P.jent[e0 + jj]; nb = t2.x; ep = t2.y; }
    else nb = P.jnbr[e0 + jj];
    const int side = nb < 0;
    nb = side ? ~nb : nb;
    bool st = false;
    if constexpr (EK == EK_GENERIC) { st = P.state_edges && (nb & ND_STATE_ENTRY_BIT); nb &= ~ND_STATE_ENTRY_BIT; }
    double xn[VD];
#pragma unroll
    for (int k = 0; k < VD; ++k) xn[k] = 0.0;
    if (!st) {
      const double* gp = gather_ptr<HALO>(P, nb);
#pragma unroll
      for (int k = 0; k < VD; ++k) xn[k] = gp[k];
    }
    int kind = EK, coupling = coupling0, pd = PE;
    if constexpr (EK == EK_GENERIC) {
      const EBDev E = P.eb[P.jebid[e0 + jj]];
      kind = E.kind; coupling = E.coupling; pd = E.pdim;
    }
    double pl[PE > 0 ? PE : 1];
    pl[0] = 0.0;
    if constexpr (PE > 0) {
#pragma unroll
      for (int k = 0; k < PE; ++k) pl[k] = PK ? P.ppack[(long long)(e0 + jj) * PE + k] : (k < pd ? P.p[(long long)ep + k] : 0.0);
    }
    double val[ED];
    if (st) state_entry_value<ED>(P.u, coupling, side, nb, val);
    else entry_value<VD, ED>(kind, coupling, side, self, xn, pl, P.t, val);
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
    double acc[ED], v[ND_MAX_VDIM];
#pragma unroll
    for (int q = 0; q < ED; ++q) acc[q] = s_val[q];
    if constexpr (!HALO) {
      if (P.acc_in != nullptr) {
#pragma unroll
        for (int q = 0; q < ED; ++q) acc[q] = P.acc_in[(long long)r0 * ED + q] + acc[q];
      }
    }
    load_vertex_state(P, B, r0, v);
    vertex_phase<VD, ED>(P, B, r0, acc, self, v, P.p + B.p0 + (long long)(r0 - B.row0) * B.pdim);
  }
}

// L2 prefetch of [p + first, p + first + count) elements, issued by one lane (16-byte granules; cp.async.bulk.prefetch.L2)
template <class T>
__device__ __forceinline__ void l2_prefetch_range(const T* p, long long first, int count, long long limit) {
#ifndef ND_CUSIM
  if (first >= limit || count <= 0) return;
  if (first + count > limit) count = (int)(limit - first);
  const unsigned long long a0 = (unsigned long long)(p + first) & ~15ULL;
  const unsigned long long a1 = ((unsigned long long)(p + first + count) + 15ULL) & ~15ULL;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
#else
  (void)p; (void)first; (void)count; (void)limit;
#endif
}
__device__ __forceinline__ int warp_sum(int v) {
#ifndef ND_CUSIM
  return __reduce_add_sync(0xffffffffu, v);
#else
  return v;
#endif
}

// one 32-lane slice of the jagged layout, executed by one warp (body of rhs_jag_kernel and of the persistent RK4 kernel)
// WIN = true (rhs_jag_kernel's window mode): the block's four slices hold the rows of ONE 128-row window; the rows' own
// outputs come from the block's shared-memory copy of the window (loaded coalesced) and the row sums are handed back
// through shared memory, so that the vertex phase runs one thread per row in natural row order.
template <int VD, int ED, int EK, int PE, int U, bool HALO, bool PK, bool WIN = false>
__device__ __forceinline__ void jag_slice(const KParams& P, const int4 S, const unsigned desc, const int lane,
                                          const double* s_self = nullptr, double* s_acc = nullptr, unsigned char* s_flag = nullptr) {
  const int len = desc & 63;
  const int row = S.y + ((desc >> 6) & 127);
  const bool head = (desc >> 13) & 1, valid = (desc >> 14) & 1;
  const VBDev B = P.vb[S.z];
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;
  const unsigned lt = (1u << lane) - 1u;
  if (P.pf_dist > 0) {                   // warp-uniform
    const int es = warp_sum(len);
    if (lane == 0) {
      if constexpr (PE > 0 && !PK) l2_prefetch_range(P.jent, (long long)S.x + P.pf_dist, es, P.jag_len);
      else {
        l2_prefetch_range(P.jnbr, (long long)S.x + P.pf_dist, es, P.jag_len);
        if constexpr (PE > 0) l2_prefetch_range(P.ppack, ((long long)S.x + P.pf_dist) * PE, es * PE, P.jag_len * PE);
      }
    }
  }

  double acc[ED];
#pragma unroll
  for (int q = 0; q < ED; ++q) acc[q] = 0.0;
  if constexpr (!HALO && !WIN) {
    if (P.acc_in != nullptr && head) {     // the row's sum over the earlier column blocks comes first (entry order is kept)
#pragma unroll
      for (int q = 0; q < ED; ++q) acc[q] = P.acc_in[(long long)row * ED + q];
    }
  }

  // The walk is software-pipelined: the index (and packed-parameter) loads of columns j+U .. j+2U-1 are issued BEFORE the
  // gathers of columns j .. j+U-1 are waited for, so an iteration exposes one memory latency (the gather) instead of two
  // dependent ones (index, then gather) -- profiles/r02b_jag128_packed: 22 % of the stall samples sat on the index load,
  // 32 % on the gather.
  int base = S.x;
  bool actA[U];
  int nbA[U], epA[U], posA[U];
  double plA[U][PE > 0 ? PE : 1];
  // stage A: slots of this lane in columns j .. j+U-1 and their coalesced index / packed-parameter loads
  auto fetch = [&](int j) -> bool {
    const unsigned m0 = __ballot_sync(0xffffffffu, j < len);
    if (m0 == 0u) return false;
#pragma unroll
    for (int q = 0; q < U; ++q) {
      actA[q] = (j + q) < len;
      const unsigned m = q == 0 ? m0 : __ballot_sync(0xffffffffu, actA[q]);
      posA[q] = base + __popc(m & lt);
      base += __popc(m);
      nbA[q] = 0; epA[q] = 0;
#pragma unroll
      for (int k = 0; k < (PE > 0 ? PE : 1); ++k) plA[q][k] = 0.0;
      if (actA[q]) {
        if constexpr (PE > 0 && !PK) { const int2 t2 = __ldcs(&P.jent[posA[q]]); nbA[q] = t2.x; epA[q] = t2.y; }
        else nbA[q] = __ldcs(&P.jnbr[posA[q]]);
        if constexpr (PE > 0 && PK) {
#pragma unroll
          for (int k = 0; k < PE; ++k) plA[q][k] = __ldcs(&P.ppack[(long long)posA[q] * PE + k]);
        }
      }
    }
    return true;
  };
  bool more = fetch(0);
  // everything up to here read engine-owned tables (and the packed parameter copy) only: under programmatic dependent launch
  // the index loads above are in flight while the previous stage kernel drains
  if constexpr (!HALO && !WIN) pdl_wait();
  double self[VD];
#pragma unroll
  for (int k = 0; k < VD; ++k) self[k] = 0.0;
  if (valid) {
    if constexpr (WIN) self[0] = s_self[(desc >> 6) & 127];
    else {
      const long long sidx = P.gather_from_u ? (B.state0 + (long long)(row - B.row0) * B.dim) : (long long)row * VD;
#pragma unroll
      for (int k = 0; k < VD; ++k) self[k] = P.gsrc[sidx + k];
    }
  }
  for (int j = 0; more; j += U) {
    // stage B operands of this iteration
    bool act[U];
    int nb[U], ep[U], pos[U];
    double pl[U][PE > 0 ? PE : 1];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      act[q] = actA[q]; nb[q] = nbA[q]; ep[q] = epA[q]; pos[q] = posA[q];
#pragma unroll
      for (int k = 0; k < (PE > 0 ? PE : 1); ++k) pl[q][k] = plA[q][k];
    }
    // (3) the gathers and the live edge parameters: up to 2U independent random reads per lane
    double xn[U][VD];
    int kind[U], coupling[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int off = nb[q] < 0 ? ~nb[q] : nb[q];
#pragma unroll
      for (int k = 0; k < VD; ++k) xn[q][k] = 0.0;
      kind[q] = EK; coupling[q] = coupling0;
      bool st = false;   // entry of an edge with states: read in step (4) from u
      if constexpr (EK == EK_GENERIC) st = P.state_edges && (off & ND_STATE_ENTRY_BIT);
      if (act[q]) {
        if (!st) {
          const double* gp = gather_ptr<HALO>(P, off);
          if constexpr (VD == 2) {
            const double2 t2 = *reinterpret_cast<const double2*>(gp);
            xn[q][0] = t2.x; xn[q][1] = t2.y;
          } else {
#pragma unroll
            for (int k = 0; k < VD; ++k) xn[q][k] = gp[k];
          }
        }
        int pd = PE;
        if constexpr (EK == EK_GENERIC) {
          const EBDev E = P.eb[P.jebid[pos[q]]];
          kind[q] = E.kind; coupling[q] = E.coupling; pd = E.pdim;
        }
        if constexpr (PE > 0 && !PK) {
#pragma unroll
          for (int k = 0; k < PE; ++k) pl[q][k] = k < pd ? P.p[(long long)ep[q] + k] : 0.0;
        }
      }
    }
    // stage A of the NEXT iteration, in flight while this iteration's gathers return (one output per vertex; the dq
    // kernels' two-component gathers leave no registers for it: measured slower, profiles/r02d)
    if constexpr (VD == 1) more = fetch(j + U);
    // (4) edge model + sequential accumulation in entry order
#pragma unroll
    for (int q = 0; q < U; ++q) {
      if (act[q]) {
        double val[ED];
        bool st = false;
        int off = nb[q] < 0 ? ~nb[q] : nb[q];
        if constexpr (EK == EK_GENERIC) { st = P.state_edges && (off & ND_STATE_ENTRY_BIT); off &= ~ND_STATE_ENTRY_BIT; }
        if (st) state_entry_value<ED>(P.u, coupling[q], nb[q] < 0, off, val);
        else entry_value<VD, ED>(kind[q], coupling[q], nb[q] < 0, self, xn[q], pl[q], P.t, val);
#pragma unroll
        for (int d = 0; d < ED; ++d) acc[d] = acc[d] + val[d];
      }
    }
    if constexpr (VD != 1) more = fetch(j + U);
  }
  // (5) rows cut into several lanes: the head lane adds the parts in order
  if (S.w > 1) {
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned hmask = __ballot_sync(0xffffffffu, head);
    const unsigned cont = vmask & ~hmask;
    // number of continuation lanes directly above this lane
    const unsigned above = lane == 31 ? 0u : (~cont >> (lane + 1));
    const int nparts = 1 + (lane == 31 ? 0 : (above ? __ffs(above) - 1 : 31 - lane));
    for (int k = 1; k < S.w; ++k) {
#pragma unroll
      for (int d = 0; d < ED; ++d) {
        const double v = __shfl_down_sync(0xffffffffu, acc[d], k);
        if (head && k < nparts) acc[d] = acc[d] + v;
      }
    }
  }
  // (6) vertex model
  if (head) {
    if constexpr (WIN) { s_acc[(desc >> 6) & 127] = acc[0]; s_flag[(desc >> 6) & 127] = 1; }
    else {
      double v[ND_MAX_VDIM];
      load_vertex_state(P, B, row, v);
      vertex_phase<VD, ED>(P, B, row, acc, self, v, P.p + B.p0 + (long long)(row - B.row0) * B.pdim);
    }
  }
}

// WIN = true (degree-bucketed layout with 128-row windows, one vertex output; the engine pads every window to four slices):
// a block IS a window.  Bucketing deals the window's rows to the lanes by degree, so a lane's own output, its du and its
// vertex data sit anywhere in the window: profiles/r02h_cfg2_ncu_source.csv counts 681 K L2 sectors for each of the three
// accesses where 250 K would do -- and the L2 sector rate is what bounds this kernel.  Here the block loads
// the window's outputs once, coalesced, and runs the vertex phase one thread per row in natural order.
template <int VD, int ED, int EK, int PE, int BLOCK, int U, int WPS, bool HALO, bool PK = false, bool WIN = false>
__global__ void __launch_bounds__(BLOCK, (WPS * 32) / BLOCK) rhs_jag_kernel(const __grid_constant__ KParams P) {
  __shared__ double s_val[BLOCK * ED];   // long rows; WIN: the rows' sums
  __shared__ double s_self[WIN ? BLOCK : 1];
  __shared__ unsigned char s_flag[WIN ? BLOCK : 1];
  if constexpr (HALO) {
    if ((int)blockIdx.x < P.n_pub) { publish_block(P.H, blockIdx.x, P.n_pub); return; }
    if (P.fence && blockIdx.x == gridDim.x - 1) { halo_wait(P); return; }
  }
  const int bid = (int)blockIdx.x - (HALO ? P.n_pub : 0) + P.blk_off;
  if constexpr (!HALO) pdl_launch_dependents();
  if (bid >= P.n_jag_blocks) {
    if constexpr (HALO) halo_wait(P);
    const int4 dl = __ldg(&P.jlong[bid - P.n_jag_blocks]);
    if constexpr (!HALO) pdl_wait();
    long_row_block<VD, ED, EK, PE, BLOCK, HALO, PK>(P, dl, s_val);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int sl = bid * (BLOCK / 32) + (threadIdx.x >> 5);
  if constexpr (WIN) {
    static_assert(!WIN || (VD == 1 && ED == 1 && BLOCK == 128), "window mode: one vertex output, 128-row windows");
    // the engine pads the slice table to whole blocks in this mode: every warp has a slice
    const int4 S = __ldg(&P.jslices[sl]);
    const unsigned desc = __ldg(&P.jlanes[(long long)sl * 32 + lane]);
    const VBDev& B = P.vb[S.z];
    const int t = threadIdx.x;
    const int rowt = S.y + t;             // S.y = first row of the window, the same in the block's four slices
    s_flag[t] = 0;
    if constexpr (!HALO) pdl_wait();      // descriptors above are engine-owned tables
    if (rowt < B.row0 + B.nrows) s_self[t] = P.gsrc[P.gather_from_u ? (B.state0 + (long long)(rowt - B.row0) * B.dim) : (long long)rowt];
    __syncthreads();
    if constexpr (HALO) {
      if (sl >= P.wait_from) halo_wait_warp(P);   // this slice reads the halo
    }
    jag_slice<VD, ED, EK, PE, U, HALO, PK, true>(P, S, desc, lane, s_self, s_val, s_flag);
    __syncthreads();
    if (s_flag[t]) {
      const double acc = s_val[t], self = s_self[t];
      double v[ND_MAX_VDIM];
      load_vertex_state(P, B, rowt, v);
      vertex_phase<VD, ED>(P, B, rowt, &acc, &self, v, P.p + B.p0 + (long long)(rowt - B.row0) * B.pdim);
    }
  } else {
    if (sl >= P.nslices) return;           // warp-uniform
    if constexpr (HALO) {
      if (sl >= P.wait_from) halo_wait_warp(P);   // this slice reads the halo
    }
    jag_slice<VD, ED, EK, PE, U, HALO, PK>(P, __ldg(&P.jslices[sl]), __ldg(&P.jlanes[(long long)sl * 32 + lane]), lane);   // calls pdl_wait()
  }
}

// persistent variant (ND_B200_JAG_PERSIST=1, single GPU): the grid is sized to the machine, warps stride over the slices and
// load the NEXT slice's descriptors before walking the current one.  rhs_jag_kernel's achieved occupancy on config 2 is
// 56 % of the SM's warp slots against 75 % allocated (profiles/r02b_jag128_packed_cfg2_ncu_summary.txt): a block's slot
// stays occupied until its slowest warp ends, and every fresh warp starts with three dependent loads.
template <int VD, int ED, int EK, int PE, int BLOCK, int U, int WPS, bool PK = false>
__global__ void __launch_bounds__(BLOCK, (WPS * 32) / BLOCK) rhs_jag_persist_kernel(const __grid_constant__ KParams P) {
  __shared__ double s_val[BLOCK * ED];   // long rows only
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)gridDim.x * (BLOCK / 32);
  int sl = (int)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
  if (sl < P.nslices) {
    int4 S = __ldg(&P.jslices[sl]);
    unsigned desc = __ldg(&P.jlanes[(long long)sl * 32 + lane]);
    while (true) {
      const int nx = sl + nwarps;
      int4 Sn = S;
      unsigned dn = 0;
      if (nx < P.nslices) { Sn = __ldg(&P.jslices[nx]); dn = __ldg(&P.jlanes[(long long)nx * 32 + lane]); }
      jag_slice<VD, ED, EK, PE, U, false, PK>(P, S, desc, lane);
      if (nx >= P.nslices) break;
      sl = nx; S = Sn; desc = dn;
    }
  }
  if (P.n_jlong > 0) {
    __syncthreads();
    for (int b = (int)blockIdx.x; b < P.n_jlong; b += (int)gridDim.x) {
      long_row_block<VD, ED, EK, PE, BLOCK, false, PK>(P, __ldg(&P.jlong[b]), s_val);
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// persistent cooperative RK4 (rk4_jag_coop_kernel): ALL steps and stages of nd_b200_rk4 in ONE launch.
//
// Why: on graphs of config-1 / config-4 size a stage kernel is one wave of thread blocks whose duration is a chain of
// dependent memory latencies (~6 us) plus the launch gap of the next graph node -- 11.6 us per stage on config 4
// (profiles/r02d_sweep_defaults.jsonl: 46.5 us per RK4 step) for 26 MB of L2-resident traffic.  Here the grid is launched
// once (cudaLaunchCooperativeKernel: every block resident), warps stride over the slices of the jagged layout, and the four
// stages of every step are separated by a grid-wide barrier (one atomic arrival per block on a monotonic counter, acquire
// spin by thread 0, gpu-scope fences on both sides -- the fence also invalidates the SM's L1, so the next stage's gathers
// see the other blocks' stage outputs).  Stage algebra, buffers and operation order are those of rk4_step_enqueue /
// vertex_phase: results are bit-identical to the graph-replayed stages.
// ------------------------------------------------------------------------------------------------
struct CoopArgs {
  double* u;                       // state vector (in: u(t0), out: u(t0 + nsteps*dt))
  double* tmpA;                    // stage inputs 2 and 4
  double* tmpB;                    // stage input 3
  double* vout0;                   // materialised vertex outputs, ping-pong (networks whose vertex g is not a state copy)
  double* vout1;
  double t0, dt;
  long long nsteps;
  unsigned long long* barrier;     // monotonic arrival counter, zeroed before the launch
};
#ifdef ND_CUSIM
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
#else
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#endif
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1ULL);
    while (ld_acquire_gpu(ctr) < target) {}
    __threadfence();
  }
  __syncthreads();
}

template <int VD, int ED, int EK, int PE, int BLOCK, int U, bool PK>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK >= 1 ? 1024 / BLOCK : 1) rk4_jag_coop_kernel(const __grid_constant__ KParams P0, const __grid_constant__ CoopArgs R) {
  __shared__ double s_val[BLOCK * ED];   // long rows only
  KParams P = P0;
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)gridDim.x * (BLOCK / 32);
  const int gw = (int)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
  const double h2 = 0.5 * R.dt;
  unsigned long long target = 0;
  for (long long step = 0; step < R.nsteps; ++step) {
    const double t = R.t0 + (double)step * R.dt;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      P.stage = s + 1;
      P.u = s == 0 ? R.u : (s == 2 ? R.tmpB : R.tmpA);
      P.unext = s == 0 ? R.tmpA : (s == 1 ? R.tmpB : (s == 2 ? R.tmpA : R.u));
      P.hs = s == 2 ? R.dt : (s == 3 ? 0.0 : h2);
      P.t = s == 0 ? t : (s == 3 ? t + R.dt : t + h2);
      if (P0.gather_from_u) { P.gsrc = P.u; P.vout_next = nullptr; }
      else { P.gsrc = (s & 1) ? R.vout1 : R.vout0; P.vout_next = (s & 1) ? R.vout0 : R.vout1; }
      for (int sl = gw; sl < P0.nslices; sl += nwarps)
        jag_slice<VD, ED, EK, PE, U, false, PK>(P, __ldg(&P0.jslices[sl]), __ldg(&P0.jlanes[(long long)sl * 32 + lane]), lane);
      for (int b = (int)blockIdx.x; b < P0.n_jlong; b += (int)gridDim.x) {
        long_row_block<VD, ED, EK, PE, BLOCK, false, PK>(P, __ldg(&P0.jlong[b]), s_val);
        __syncthreads();
      }
      target += gridDim.x;
      grid_barrier(R.barrier, target);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// asynchronous-gather jagged kernel (rhs_jaga_kernel): the layout of rhs_jag_kernel, the gathers as LDGSTS.
//
// Why (profiles/r02b_jag128_packed_cfg2_ncu_summary.txt): rhs_jag_kernel keeps the L1TEX pipe only 45 % busy on config 2 --
// it is bound by the number of gathers IN FLIGHT, not by the gather rate: a lane holds U = 2 gathered values in registers per
// iteration, ~2 K gathers per SM, while tools/gather_bench.cu needs >= 8 K per SM to reach the tag-stage rate
// (profiles/r02a_gather_bench_raw.txt: ILP1 47 us, ILP4 35 us, ILP8 33 us).  Registers are the limit of the register-gather
// form (U = 8 was measured slower, profiles/r02c_sweep_jag_deep_unroll.jsonl).  Here the gathered values never wait in
// registers: for a chunk of CH columns of its slice the warp
//   (1) counts the chunk's entries (CH ballots) -- they are one contiguous piece of the slice's entry stream;
//   (2) walks that piece 32 entries at a time: coalesced index (and packed-parameter position) loads, then ONE cp.async
//       (LDGSTS, 8 bytes) per entry that copies the neighbour's output -- and the entry's edge parameter -- straight into
//       the warp's panel in shared memory; up to 32*CH gathers per warp in flight at no register cost;
//   (3) waits for the group (cp.async.wait_group + __syncwarp), then every lane reads ITS row's entries from the panel
//       (consecutive lanes read consecutive words: conflict-free) and adds them in the reference's sequential order
//       (src/aggregators.jl:140-151) -- bit-identical to rhs_jag_kernel.
// Single vertex output, single-batch registry edge kinds (vdepth = edepth = 1), rows split over lanes and whole-block rows
// exactly as in rhs_jag_kernel.
// ------------------------------------------------------------------------------------------------
#ifdef ND_CUSIM
__device__ __forceinline__ void cp_async_f64(double* dst, const double* src) { *dst = *src; }
__device__ __forceinline__ void cp_async_commit() {}
__device__ __forceinline__ void cp_async_wait_all() {}
#else
__device__ __forceinline__ void cp_async_f64(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

template <int EK, int PE, bool PK, int CH, int WPS, bool HALO>
__global__ void __launch_bounds__(128, (WPS * 32) / 128) rhs_jaga_kernel(const __grid_constant__ KParams P) {
  constexpr int BLOCK = 128;
  static_assert(PE <= 1, "one parameter per edge");
  __shared__ double s_x[BLOCK / 32][CH * 32];                    // gathered neighbour outputs of the chunk
  __shared__ double s_p[BLOCK / 32][PE > 0 ? CH * 32 : 1];       // the entries' edge parameters
  __shared__ unsigned char s_side[BLOCK / 32][CH * 32];          // 1: the row is the edge's src
  if constexpr (HALO) {
    if ((int)blockIdx.x < P.n_pub) { publish_block(P.H, blockIdx.x, P.n_pub); return; }
    if (P.fence && blockIdx.x == gridDim.x - 1) { halo_wait(P); return; }
  }
  const int bid = (int)blockIdx.x - (HALO ? P.n_pub : 0) + P.blk_off;
  if (bid >= P.n_jag_blocks) {
    if constexpr (HALO) halo_wait(P);
    static_assert(CH * 32 * (BLOCK / 32) >= BLOCK, "panel too small for the block tree");
    long_row_block<1, 1, EK, PE, BLOCK, HALO, PK>(P, __ldg(&P.jlong[bid - P.n_jag_blocks]), &s_x[0][0]);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sl = bid * (BLOCK / 32) + warp;
  if (sl >= P.nslices) return;           // warp-uniform
  if constexpr (HALO) {
    if (sl >= P.wait_from) halo_wait_warp(P);   // this slice reads the halo
  }
  const int4 S = __ldg(&P.jslices[sl]);
  const unsigned desc = __ldg(&P.jlanes[(long long)sl * 32 + lane]);
  const int len = desc & 63;
  const int row = S.y + ((desc >> 6) & 127);
  const bool head = (desc >> 13) & 1, valid = (desc >> 14) & 1;
  const VBDev& B = P.vb[S.z];
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;
  const unsigned lt = (1u << lane) - 1u;
  double* sx = s_x[warp];
  double* sp = s_p[warp];
  unsigned char* ss = s_side[warp];

  double self = 0.0;
  if (valid) self = P.gsrc[P.gather_from_u ? (B.state0 + (long long)(row - B.row0) * B.dim) : (long long)row];
  double acc = 0.0;
  int cb = S.x;
  for (int c0 = 0;; c0 += CH) {
    // (1) this lane's positions inside the chunk, the chunk's entry count
    const unsigned m0 = __ballot_sync(0xffffffffu, c0 < len);
    if (m0 == 0u) break;
    unsigned posw[(CH + 1) / 2];         // positions inside the chunk (< 32*CH <= 65536), two per word
    int n_c = 0;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const unsigned m = q == 0 ? m0 : __ballot_sync(0xffffffffu, c0 + q < len);
      const unsigned pq = (unsigned)(n_c + __popc(m & lt));
      if (q & 1) posw[q >> 1] |= pq << 16; else posw[q >> 1] = pq;
      n_c += __popc(m);
    }
    // (2) the chunk's entries, 32 at a time: coalesced index loads, asynchronous gathers into the panel
    int nbq[CH], epq[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int i = q * 32 + lane;
      nbq[q] = 0; epq[q] = 0;
      if (i < n_c) {
        if constexpr (PE > 0 && !PK) { const int2 t2 = __ldcs(&P.jent[cb + i]); nbq[q] = t2.x; epq[q] = t2.y; }
        else nbq[q] = __ldcs(&P.jnbr[cb + i]);
      }
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int i = q * 32 + lane;
      if (i < n_c) {
        const int side = nbq[q] < 0;
        const int off = side ? ~nbq[q] : nbq[q];
        cp_async_f64(sx + i, gather_ptr<HALO>(P, off));
        if constexpr (PE > 0) cp_async_f64(sp + i, PK ? P.ppack + (cb + i) : P.p + epq[q]);
        ss[i] = (unsigned char)side;
      }
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    // (3) edge model + sequential accumulation in entry order
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      if (c0 + q < len) {
        const int pq = (int)((posw[q >> 1] >> ((q & 1) * 16)) & 0xffffu);
        const double xn = sx[pq];
        double pl = 0.0;
        if constexpr (PE > 0) pl = sp[pq];
        double val;
        entry_value<1, 1>(EK, coupling0, ss[pq], &self, &xn, &pl, P.t, &val);
        acc = acc + val;
      }
    }
    __syncwarp();       // the next chunk overwrites the panel
    cb += n_c;
  }
  // rows cut into several lanes: the head lane adds the parts in order (as in rhs_jag_kernel)
  if (S.w > 1) {
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned hmask = __ballot_sync(0xffffffffu, head);
    const unsigned cont = vmask & ~hmask;
    const unsigned above = lane == 31 ? 0u : (~cont >> (lane + 1));
    const int nparts = 1 + (lane == 31 ? 0 : (above ? __ffs(above) - 1 : 31 - lane));
    for (int k = 1; k < S.w; ++k) {
      const double v = __shfl_down_sync(0xffffffffu, acc, k);
      if (head && k < nparts) acc = acc + v;
    }
  }
  if (head) {
    double v[ND_MAX_VDIM];
    load_vertex_state(P, B, row, v);
    vertex_phase<1, 1>(P, B, row, &acc, &self, v, P.p + B.p0 + (long long)(row - B.row0) * B.pdim);
  }
}

// ------------------------------------------------------------------------------------------------
// batched jagged kernel (rhs_jagb_kernel): the layout of rhs_jag_kernel, three dependent memory round trips per chunk.
//
// Why (profiles/r02h_cfg2_ncu_source.csv, rhs_jag_kernel on config 2): 39 % of the warp samples sit on the first use of a
// gathered value, 8 % on the index of the next pair of columns, 6 % on the slice descriptor; a slice walks 5.1 iterations of
// two columns, each one an exposed gather round trip, with 25 resident warps x 64 gathers = ~600 gathers in flight per SM on
// average where the tag stage needs ~1500 (tools/path_gather_bench.cu: 0.84 per clock x ~1750 clocks of loaded latency).
// Here a warp takes CH columns of its slice at once:
//   (1) CH ballots give every lane its positions and the chunk's entry count (one contiguous piece of the entry stream);
//   (2) the piece is loaded 32 entries at a time -- CH coalesced index loads and CH coalesced packed-parameter loads per
//       lane, all independent -- and parked in the warp's panel in shared memory (plain stores);
//   (3) every lane reads ITS entries' indices back (consecutive lanes read consecutive words), issues up to CH gathers
//       back to back into registers, then adds the edge values in the reference's sequential order
//       (src/aggregators.jl:140-151) -- bit-identical to rhs_jag_kernel.
// Rows of up to CH entries cost descriptor -> stream -> gather round trips in total, with 32*CH gathers per warp in flight.
// ------------------------------------------------------------------------------------------------
template <int EK, int PE, bool PK, int CH, int WPS, bool HALO>
__global__ void __launch_bounds__(128, (WPS * 32) / 128) rhs_jagb_kernel(const __grid_constant__ KParams P) {
  constexpr int BLOCK = 128;
  static_assert(PE <= 1, "one parameter per edge");
  __shared__ int s_nb[BLOCK / 32][CH * 32];                      // the chunk's index words
  __shared__ double s_p[BLOCK / 32][PE > 0 ? CH * 32 : 1];       // PK: the entries' edge parameters; else their offsets in p
  __shared__ double s_val[BLOCK];                                // long rows only
  if constexpr (HALO) {
    if ((int)blockIdx.x < P.n_pub) { publish_block(P.H, blockIdx.x, P.n_pub); return; }
    if (P.fence && blockIdx.x == gridDim.x - 1) { halo_wait(P); return; }
  }
  const int bid = (int)blockIdx.x - (HALO ? P.n_pub : 0) + P.blk_off;
  if (bid >= P.n_jag_blocks) {
    if constexpr (HALO) halo_wait(P);
    long_row_block<1, 1, EK, PE, BLOCK, HALO, PK>(P, __ldg(&P.jlong[bid - P.n_jag_blocks]), s_val);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sl = bid * (BLOCK / 32) + warp;
  if (sl >= P.nslices) return;           // warp-uniform
  if constexpr (HALO) {
    if (sl >= P.wait_from) halo_wait_warp(P);   // this slice reads the halo
  }
  const int4 S = __ldg(&P.jslices[sl]);
  const unsigned desc = __ldg(&P.jlanes[(long long)sl * 32 + lane]);
  const int len = desc & 63;
  const int row = S.y + ((desc >> 6) & 127);
  const bool head = (desc >> 13) & 1, valid = (desc >> 14) & 1;
  const VBDev& B = P.vb[S.z];
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;
  const unsigned lt = (1u << lane) - 1u;
  if (P.pf_dist > 0) {                   // warp-uniform
    const int es = warp_sum(len);
    if (lane == 0) {
      if constexpr (PE > 0 && !PK) l2_prefetch_range(P.jent, (long long)S.x + P.pf_dist, es, P.jag_len);
      else {
        l2_prefetch_range(P.jnbr, (long long)S.x + P.pf_dist, es, P.jag_len);
        if constexpr (PE > 0) l2_prefetch_range(P.ppack, ((long long)S.x + P.pf_dist) * PE, es * PE, P.jag_len * PE);
      }
    }
  }
  int* snb = s_nb[warp];
  double* sp = s_p[warp];

  double self = 0.0;
  if (valid) self = P.gsrc[P.gather_from_u ? (B.state0 + (long long)(row - B.row0) * B.dim) : (long long)row];
  double acc = 0.0;
  int cb = S.x;
  for (int c0 = 0;; c0 += CH) {
    // (1) this lane's positions inside the chunk, the chunk's entry count
    const unsigned m0 = __ballot_sync(0xffffffffu, c0 < len);
    if (m0 == 0u) break;
    unsigned posw[(CH + 1) / 2];         // positions inside the chunk (< 32*CH), two per word
    int n_c = 0;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const unsigned m = q == 0 ? m0 : __ballot_sync(0xffffffffu, c0 + q < len);
      const unsigned pq = (unsigned)(n_c + __popc(m & lt));
      if (q & 1) posw[q >> 1] |= pq << 16; else posw[q >> 1] = pq;
      n_c += __popc(m);
    }
    // (2) the chunk's piece of the entry stream -> panel
    {
      int nbq[CH];
      double plq[PE > 0 ? CH : 1];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int i = q * 32 + lane;
        nbq[q] = 0;
        if constexpr (PE > 0) plq[q] = 0.0;
        if (i < n_c) {
          if constexpr (PE > 0 && !PK) { const int2 t2 = __ldcs(&P.jent[cb + i]); nbq[q] = t2.x; plq[q] = (double)t2.y; }   // offsets < 2^31: exact
          else nbq[q] = __ldcs(&P.jnbr[cb + i]);
          if constexpr (PE > 0 && PK) plq[q] = __ldcs(&P.ppack[cb + i]);
        }
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int i = q * 32 + lane;
        if (i < n_c) {
          snb[i] = nbq[q];
          if constexpr (PE > 0) sp[i] = plq[q];
        }
      }
    }
    __syncwarp();
    // (3) this lane's gathers, all in flight together; then edge model + sequential accumulation in entry order
    double xn[CH];
    double plv[(PE > 0 && !PK) ? CH : 1];      // live parameters: a second gather per entry, issued with the first
    unsigned sides = 0;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      xn[q] = 0.0;
      if (c0 + q < len) {
        const int pq = (int)((posw[q >> 1] >> ((q & 1) * 16)) & 0xffffu);
        const int nb = snb[pq];
        const int side = nb < 0;
        sides |= (unsigned)side << q;
        xn[q] = *gather_ptr<HALO>(P, side ? ~nb : nb);
        if constexpr (PE > 0 && !PK) plv[q] = P.p[(long long)sp[pq]];
      }
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      if (c0 + q < len) {
        const int pq = (int)((posw[q >> 1] >> ((q & 1) * 16)) & 0xffffu);
        double pl = 0.0;
        if constexpr (PE > 0) {
          if constexpr (PK) pl = sp[pq]; else pl = plv[q];
        }
        double val;
        entry_value<1, 1>(EK, coupling0, (int)((sides >> q) & 1u), &self, &xn[q], &pl, P.t, &val);
        acc = acc + val;
      }
    }
    __syncwarp();       // the next chunk overwrites the panel
    cb += n_c;
  }
  // rows cut into several lanes: the head lane adds the parts in order (as in rhs_jag_kernel)
  if (S.w > 1) {
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned hmask = __ballot_sync(0xffffffffu, head);
    const unsigned cont = vmask & ~hmask;
    const unsigned above = lane == 31 ? 0u : (~cont >> (lane + 1));
    const int nparts = 1 + (lane == 31 ? 0 : (above ? __ffs(above) - 1 : 31 - lane));
    for (int k = 1; k < S.w; ++k) {
      const double v = __shfl_down_sync(0xffffffffu, acc, k);
      if (head && k < nparts) acc = acc + v;
    }
  }
  if (head) {
    double v[ND_MAX_VDIM];
    load_vertex_state(P, B, row, v);
    vertex_phase<1, 1>(P, B, row, &acc, &self, v, P.p + B.p0 + (long long)(row - B.row0) * B.pdim);
  }
}

// ------------------------------------------------------------------------------------------------
// streamed jagged kernel: the jagged warp-slice walk of rhs_jag_kernel, fed by TMA bulk copies.
//
// Why (profiles/r02b_jag128_packed_cfg2_ncu_summary.txt, B200): rhs_jag_kernel issues only 47 K L1TEX wavefronts per SM on
// config 2 (the tile kernel: 100 K) but keeps the L1TEX data pipe 45 % busy: every warp lives for ONE slice and walks a
// chain of dependent long-latency loads -- slice descriptor -> lane descriptor -> own state, then per iteration index ->
// gather -- so the kernel is bound by exposed latency (long scoreboard 21 per issue, 0.5 eligible warps per scheduler),
// not by the gather rate (profiles/r02a_*: 0.83 random gathers per clock per SM, the LSU tag stage).  Here
//   * warps are persistent: a warp owns a contiguous range of slices, i.e. ONE contiguous piece of the index / parameter
//     streams;
//   * that piece is streamed into a per-warp ring in shared memory by cp.async.bulk (TMA, SASS UBLKCP) in chunks of
//     JS_CHUNK entries, NST chunks deep, completion on one mbarrier per chunk: the coalesced streams never pass through
//     the LSU pipe as global loads and are in flight many iterations ahead of their use;
//   * the index of a gather is then a 30-cycle shared-memory read, so the only long-latency load on the critical path
//     is the gather itself, U of them in flight per lane;
//   * descriptors, own state and vertex parameters of the NEXT slice are prefetched into registers.
// Layout = the degree-bucketed jagged layout (window 128), every slice padded to a multiple of 4 entries (16-byte
// alignment of the bulk copies).  Row sums stay in registers, in the reference's sequential order
// (src/aggregators.jl:140-151).  Instantiated for the single-batch benchmark edge kinds with vdepth = edepth = 1.
// ------------------------------------------------------------------------------------------------
constexpr int JS_BLOCK = 128;
constexpr int JS_CHUNK = 128;
#ifdef ND_CUSIM
#define ND_DYN_SMEM(name) static thread_local __attribute__((aligned(128))) unsigned char name[232448]
__device__ __forceinline__ void js_bar_init(unsigned long long* bar) { *bar = 0; }
__device__ __forceinline__ void js_fence_init() {}
__device__ __forceinline__ void js_expect(unsigned long long*, unsigned) {}
__device__ __forceinline__ void js_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void js_wait(unsigned long long*, unsigned) {}
#else
#define ND_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
__device__ __forceinline__ unsigned js_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void js_bar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(js_smem_addr(bar)), "r"(1) : "memory");
}
__device__ __forceinline__ void js_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void js_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(js_smem_addr(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA); bytes and both addresses are multiples of 16
__device__ __forceinline__ void js_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(js_smem_addr(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(js_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void js_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = js_smem_addr(bar);
  unsigned ok = 0;
  do {
    asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
#endif

__host__ __device__ constexpr int js_imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int js_imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int js_warp_bytes(int pe, bool pk, int nst) {
  return nst * JS_CHUNK * ((pe > 0 && !pk) ? 8 : 4) + (pk ? nst * JS_CHUNK * 8 * pe : 0) + nst * 8;
}

template <int EK, int PE, bool PK, int U, int NST, int MINB>
__global__ void __launch_bounds__(JS_BLOCK, MINB) rhs_js_kernel(const __grid_constant__ KParams P) {
  constexpr bool LIVE = PE > 0 && !PK;          // parameters re-read from the caller's p through per-entry offsets
  constexpr int RING = NST * JS_CHUNK;
  constexpr int WB = js_warp_bytes(PE, PK, NST);
  constexpr int IDXB = LIVE ? 8 : 4;
  static_assert((NST & (NST - 1)) == 0 && NST >= 4 && U * 32 <= 2 * JS_CHUNK, "ring / unroll shape");
  static_assert(PE <= 1, "one edge parameter at most (diffusion, Kuramoto)");
  ND_DYN_SMEM(js_smem);
  __shared__ double s_val[JS_BLOCK];            // long rows only
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = js_smem + wib * WB;
  const int* r_nbr = reinterpret_cast<const int*>(wbase);
  const int2* r_ent = reinterpret_cast<const int2*>(wbase);
  const double* r_pk = reinterpret_cast<const double*>(wbase + RING * IDXB);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(wbase + WB - NST * 8);
  const int w = (int)blockIdx.x * (JS_BLOCK / 32) + wib;
  int sb = 0, se = 0;
  if (w < P.n_jwarps) { const int2 r = __ldg(&P.jwarp[w]); sb = r.x; se = r.y; }
  const int coupling0 = P.n_eb > 0 ? P.eb[0].coupling : 0;
  const unsigned lt = (1u << lane) - 1u;

  if (sb < se) {
    const int E0 = __ldg(&P.jslices[sb]).x, E1 = __ldg(&P.jslices[se]).x;     // jslices ends with a sentinel
    const int etot = E1 - E0, nchunks = (etot + JS_CHUNK - 1) / JS_CHUNK;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NST; ++k) js_bar_init(bars + k);
      js_fence_init();
    }
    __syncwarp();
    int issued = 0, ready = 0;
    // lane 0 starts the bulk copies of chunks [issued, lim); every lane tracks the count
    auto issue_upto = [&](int lim) {
      if (lane == 0) {
        for (int c = issued; c < lim; ++c) {
          const int st = c & (NST - 1);
          const int n = js_imin(JS_CHUNK, etot - c * JS_CHUNK);       // multiple of 4
          js_expect(bars + st, (unsigned)(n * IDXB + (PK ? n * 8 * PE : 0)));
          const void* src = LIVE ? (const void*)(P.jent + E0 + (long long)c * JS_CHUNK) : (const void*)(P.jnbr + E0 + (long long)c * JS_CHUNK);
          js_bulk_g2s(wbase + st * JS_CHUNK * IDXB, src, (unsigned)(n * IDXB), bars + st);
          if constexpr (PK)
            js_bulk_g2s(wbase + RING * IDXB + st * JS_CHUNK * 8 * PE, P.ppack + (long long)(E0 + (long long)c * JS_CHUNK) * PE, (unsigned)(n * 8 * PE), bars + st);
        }
      }
      issued = js_imax(issued, lim);
    };
    issue_upto(js_imin(nchunks, NST));

    // register prefetch: descriptors two slices ahead, own state / vertex parameters one slice ahead
    int4 S0 = __ldg(&P.jslices[sb]);
    unsigned d0 = __ldg(&P.jlanes[(long long)sb * 32 + lane]);
    int4 S1 = S0; unsigned d1 = 0;
    if (sb + 1 < se) { S1 = __ldg(&P.jslices[sb + 1]); d1 = __ldg(&P.jlanes[(long long)(sb + 1) * 32 + lane]); }
    int curb = S0.z;
    VBDev B = P.vb[curb];
    double v0[ND_MAX_VDIM], pv0[4], v1[ND_MAX_VDIM], pv1[4];
    auto load_own = [&](const VBDev& Bx, const int4& Sx, unsigned dx, double* v, double* pv) {
#pragma unroll
      for (int c = 0; c < ND_MAX_VDIM; ++c) v[c] = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) pv[c] = 0.0;
      if ((dx >> 14) & 1) {
        const long long i = (long long)(Sx.y + (int)((dx >> 6) & 127)) - Bx.row0;
        const long long so = Bx.state0 + i * Bx.dim;
#pragma unroll
        for (int c = 0; c < ND_MAX_VDIM; ++c) if (c < Bx.dim) v[c] = P.u[so + c];
        if ((dx >> 13) & 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c) if (c < Bx.pdim) pv[c] = P.p[Bx.p0 + i * Bx.pdim + c];
        }
      }
    };
    load_own(B, S0, d0, v0, pv0);

    for (int s = sb; s < se; ++s) {
      // prefetch: descriptors of s+2, own data of s+1 (same vertex batch; a batch change reloads at the top of s+1)
      int4 S2 = S1; unsigned d2 = 0;
      if (s + 2 < se) { S2 = __ldg(&P.jslices[s + 2]); d2 = __ldg(&P.jlanes[(long long)(s + 2) * 32 + lane]); }
      const bool pre1 = (s + 1 < se) && S1.z == curb;
      if (pre1) load_own(B, S1, d1, v1, pv1);

      const int len = d0 & 63;
      const int row = S0.y + (int)((d0 >> 6) & 127);
      const bool head = (d0 >> 13) & 1, valid = (d0 >> 14) & 1;
      const double self = v0[0];
      double acc = 0.0;
      int base = S0.x - E0;
      for (int j = 0;; j += U) {
        const unsigned m0 = __ballot_sync(0xffffffffu, j < len);
        if (m0 == 0u) break;
        int pos[U];
        bool act[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          act[q] = (j + q) < len;
          const unsigned m = q == 0 ? m0 : __ballot_sync(0xffffffffu, act[q]);
          pos[q] = (base + __popc(m & lt)) & (RING - 1);
          base += __popc(m);
        }
        // the chunks that hold entries below `base` must have landed
        const int need = (base + JS_CHUNK - 1) / JS_CHUNK;
        while (ready < need) { js_wait(bars + (ready & (NST - 1)), (unsigned)((ready / NST) & 1)); ++ready; }
        int nb[U], ep[U];
        double pl[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          nb[q] = 0; ep[q] = 0; pl[q] = 0.0;
          if (act[q]) {
            if constexpr (LIVE) { const int2 t2 = r_ent[pos[q]]; nb[q] = t2.x; ep[q] = t2.y; }
            else nb[q] = r_nbr[pos[q]];
            if constexpr (PK) pl[q] = r_pk[pos[q] * PE];
          }
        }
        double xn[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          xn[q] = 0.0;
          if (act[q]) {
            xn[q] = P.gsrc[nb[q] < 0 ? ~nb[q] : nb[q]];
            if constexpr (LIVE) pl[q] = P.p[ep[q]];
          }
        }
        // ring slots below `base` rounded down to a chunk are free again: refill them
        __syncwarp();
        issue_upto(js_imin(nchunks, base / JS_CHUNK + NST));
#pragma unroll
        for (int q = 0; q < U; ++q) {
          if (act[q]) {
            double val;
            entry_value<1, 1>(EK, coupling0, nb[q] < 0, &self, &xn[q], &pl[q], P.t, &val);
            acc = acc + val;
          }
        }
      }
      // rows cut into several lanes: the head lane adds the parts in order
      if (S0.w > 1) {
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned hmask = __ballot_sync(0xffffffffu, head);
        const unsigned cont = vmask & ~hmask;
        const unsigned above = lane == 31 ? 0u : (~cont >> (lane + 1));
        const int nparts = 1 + (lane == 31 ? 0 : (above ? __ffs(above) - 1 : 31 - lane));
        for (int k = 1; k < S0.w; ++k) {
          const double vv = __shfl_down_sync(0xffffffffu, acc, k);
          if (head && k < nparts) acc = acc + vv;
        }
      }
      if (head) vertex_phase<1, 1>(P, B, row, &acc, &self, v0, pv0);
      // rotate the prefetch registers
      S0 = S1; d0 = d1; S1 = S2; d1 = d2;
      if (s + 1 < se) {
        if (pre1) {
#pragma unroll
          for (int c = 0; c < ND_MAX_VDIM; ++c) v0[c] = v1[c];
#pragma unroll
          for (int c = 0; c < 4; ++c) pv0[c] = pv1[c];
        } else {
          curb = S0.z;
          B = P.vb[curb];
          load_own(B, S0, d0, v0, pv0);
        }
      }
    }
  }
  // rows longer than a slice can hold: whole-block reduction, blocks take them round robin
  if (P.n_jlong > 0) {
    __syncthreads();
    for (int q = (int)blockIdx.x; q < P.n_jlong; q += (int)gridDim.x) {
      long_row_block<1, 1, EK, PE, JS_BLOCK, false, PK>(P, __ldg(&P.jlong[q]), s_val);
      __syncthreads();
    }
  }
}

}  // namespace ndb
