/*
 * nd_oracle.c -- TEST INFRASTRUCTURE ONLY.  See nd_oracle.h for the contract
 * and the parity-pinning statement.  Compile with -O2 -ffp-contract=off
 * (Julia does not contract a*b+c into FMA; neither may this file).
 *
 * Every function cites the reference file:line (relative to the
 * NetworkDynamics.jl tree) whose behaviour it restates.
 */
#include "nd_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static char g_err[512] = "";
const char* ndo_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return -1; } while (0)

typedef struct {
  int32_t spec;        /* index into vspecs / especs */
  int64_t len;
  int64_t* indices;    /* 1-based component ids, ascending */
  /* BatchStride firsts (1-based), src/network_structure.jl:232-238,250-256 */
  int64_t state_first, p_first, in_first, out_first;
} ndo_batch;

struct ndo_network {
  int64_t nv, ne;
  int64_t *esrc, *edst;
  int32_t n_vspecs, n_especs;
  ndo_vspec* vspecs;
  ndo_espec* especs;
  int32_t *vtype, *etype;
  int32_t vdepth, edepth;
  int64_t lastidx_dynamic, lastidx_p, lastidx_out, lastidx_aggr, lastidx_gbuf;
  int64_t *v_data, *v_out, *v_para, *v_aggr;
  int64_t *e_data, *e_out_src, *e_out_dst, *e_para, *e_gbuf_src, *e_gbuf_dst;
  int32_t n_vb, n_eb;
  ndo_batch *vb, *eb;
  int64_t* gbufmap;            /* EagerGBufProvider.map, src/gbufs.jl:13-22 */
  int64_t aggr_first, aggr_len;/* AggregationMap.range, src/aggregators.jl:80-86 */
  int64_t* aggmap;             /* AggregationMap.map */
  /* caches (DiffCache stand-ins, src/construction.jl:205-208) */
  double *o, *aggbuf, *gbuf;
  /* inverse aggregation map for the threaded aggregator, src/aggregators.jl:203-230 */
  int64_t *inv_ptr, *inv_src;
  /* rk4 scratch */
  double *k1, *k2, *k3, *k4, *tmp;
};

/* ---- batching: find_identical, src/utils.jl:197-217 --------------------------------
 * Components are "identical" when _component_hash agrees (src/construction.jl:245-256);
 * here that is "same spec index".  Batches are ordered by first occurrence, members
 * ascending.  (The single-model shortcut of src/construction.jl:156-168 yields the same
 * one batch 1:n.) */
static int make_batches(int64_t n, const int32_t* type, int32_t nspecs, int32_t* nb_out, ndo_batch** b_out) {
  int32_t* spec2batch = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nspecs > 0 ? nspecs : 1));
  for (int32_t s = 0; s < nspecs; ++s) spec2batch[s] = -1;
  int64_t* counts = (int64_t*)calloc((size_t)(nspecs > 0 ? nspecs : 1), sizeof(int64_t));
  int32_t nb = 0;
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nspecs > 0 ? nspecs : 1));
  for (int64_t i = 0; i < n; ++i) {
    int32_t s = type[i];
    if (s < 0 || s >= nspecs) { free(spec2batch); free(counts); free(order); FAIL("component %lld has invalid spec %d", (long long)i + 1, s); }
    if (spec2batch[s] < 0) { spec2batch[s] = nb; order[nb] = s; nb++; }
    counts[spec2batch[s]]++;
  }
  ndo_batch* b = (ndo_batch*)calloc((size_t)(nb > 0 ? nb : 1), sizeof(ndo_batch));
  for (int32_t k = 0; k < nb; ++k) {
    b[k].spec = order[k];
    b[k].len = 0;
    b[k].indices = (int64_t*)malloc(sizeof(int64_t) * (size_t)counts[k]);
  }
  for (int64_t i = 0; i < n; ++i) {
    ndo_batch* bb = &b[spec2batch[type[i]]];
    bb->indices[bb->len++] = i + 1;
  }
  free(spec2batch); free(counts); free(order);
  *nb_out = nb; *b_out = b;
  return 0;
}

/* _nextrange, src/network_structure.jl:289 : returns first index of the new range */
static inline int64_t nextrange(int64_t* last, int64_t n) { int64_t f = *last + 1; *last += n; return f; }

ndo_network* ndo_build(int64_t nv, int64_t ne, const int64_t* esrc, const int64_t* edst,
                       int32_t n_vspecs, const ndo_vspec* vspecs, const int32_t* vtype,
                       int32_t n_especs, const ndo_espec* especs, const int32_t* etype) {
  g_err[0] = 0;
  ndo_network* nw = (ndo_network*)calloc(1, sizeof *nw);
  nw->nv = nv; nw->ne = ne;
  nw->esrc = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ne ? ne : 1));
  nw->edst = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ne ? ne : 1));
  memcpy(nw->esrc, esrc, sizeof(int64_t) * (size_t)ne);
  memcpy(nw->edst, edst, sizeof(int64_t) * (size_t)ne);
  nw->n_vspecs = n_vspecs; nw->n_especs = n_especs;
  nw->vspecs = (ndo_vspec*)malloc(sizeof(ndo_vspec) * (size_t)n_vspecs);
  nw->especs = (ndo_espec*)malloc(sizeof(ndo_espec) * (size_t)(n_especs ? n_especs : 1));
  memcpy(nw->vspecs, vspecs, sizeof(ndo_vspec) * (size_t)n_vspecs);
  memcpy(nw->especs, especs, sizeof(ndo_espec) * (size_t)n_especs);
  nw->vtype = (int32_t*)malloc(sizeof(int32_t) * (size_t)nv);
  nw->etype = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ne ? ne : 1));
  memcpy(nw->vtype, vtype, sizeof(int32_t) * (size_t)nv);
  memcpy(nw->etype, etype, sizeof(int32_t) * (size_t)ne);

  for (int64_t i = 0; i < ne; ++i)
    if (esrc[i] < 1 || esrc[i] > nv || edst[i] < 1 || edst[i] > nv) {
      snprintf(g_err, sizeof g_err, "edge %lld endpoint out of range", (long long)i + 1);
      ndo_free(nw); return NULL;
    }

  /* uniform depths, src/construction.jl:99-106 */
  nw->vdepth = vspecs[vtype[0]].outdim;
  for (int64_t i = 0; i < nv; ++i)
    if (vspecs[vtype[i]].outdim != nw->vdepth) { snprintf(g_err, sizeof g_err, "vertices have different outdim"); ndo_free(nw); return NULL; }
  nw->edepth = ne ? especs[etype[0]].outdim_dst : 0;
  for (int64_t i = 0; i < ne; ++i) {
    const ndo_espec* s = &especs[etype[i]];
    if (s->outdim_dst != nw->edepth) { snprintf(g_err, sizeof g_err, "edges have different outdim.dst"); ndo_free(nw); return NULL; }
    if (s->outdim_src != 0 && s->outdim_src != nw->edepth) { snprintf(g_err, sizeof g_err, "outdim.src must be 0 or edepth"); ndo_free(nw); return NULL; }
  }

  if (make_batches(nv, vtype, n_vspecs, &nw->n_vb, &nw->vb) || make_batches(ne, etype, n_especs, &nw->n_eb, &nw->eb)) {
    ndo_free(nw); return NULL;
  }

  size_t NV = (size_t)nv, NE = (size_t)(ne ? ne : 1);
  nw->v_data = (int64_t*)malloc(8 * NV); nw->v_out = (int64_t*)malloc(8 * NV);
  nw->v_para = (int64_t*)malloc(8 * NV); nw->v_aggr = (int64_t*)malloc(8 * NV);
  nw->e_data = (int64_t*)malloc(8 * NE); nw->e_out_src = (int64_t*)malloc(8 * NE);
  nw->e_out_dst = (int64_t*)malloc(8 * NE); nw->e_para = (int64_t*)malloc(8 * NE);
  nw->e_gbuf_src = (int64_t*)malloc(8 * NE); nw->e_gbuf_dst = (int64_t*)malloc(8 * NE);

  /* register_vertices!, src/network_structure.jl:224-239 : all vertex batches first
   * (src/construction.jl:171-184) */
  for (int32_t b = 0; b < nw->n_vb; ++b) {
    ndo_batch* B = &nw->vb[b];
    const ndo_vspec* s = &vspecs[B->spec];
    for (int64_t k = 0; k < B->len; ++k) {
      int64_t i = B->indices[k] - 1;
      nw->v_data[i] = nextrange(&nw->lastidx_dynamic, s->dim);
      nw->v_out[i] = nextrange(&nw->lastidx_out, s->outdim);
      nw->v_para[i] = nextrange(&nw->lastidx_p, s->pdim);
      nw->v_aggr[i] = nextrange(&nw->lastidx_aggr, nw->edepth);
    }
    int64_t f = B->indices[0] - 1;
    B->state_first = nw->v_data[f]; B->p_first = nw->v_para[f];
    B->in_first = nw->v_aggr[f]; B->out_first = nw->v_out[f];
  }
  /* register_edges!, src/network_structure.jl:240-258 : src out range first, dst right after */
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    ndo_batch* B = &nw->eb[b];
    const ndo_espec* s = &especs[B->spec];
    for (int64_t k = 0; k < B->len; ++k) {
      int64_t i = B->indices[k] - 1;
      nw->e_data[i] = nextrange(&nw->lastidx_dynamic, s->dim);
      nw->e_out_src[i] = nextrange(&nw->lastidx_out, s->outdim_src);
      nw->e_out_dst[i] = nextrange(&nw->lastidx_out, s->outdim_dst);
      nw->e_para[i] = nextrange(&nw->lastidx_p, s->pdim);
      nw->e_gbuf_src[i] = nextrange(&nw->lastidx_gbuf, nw->vdepth);
      nw->e_gbuf_dst[i] = nextrange(&nw->lastidx_gbuf, nw->vdepth);
    }
    int64_t f = B->indices[0] - 1;
    B->state_first = nw->e_data[f]; B->p_first = nw->e_para[f];
    B->in_first = nw->e_gbuf_src[f]; B->out_first = nw->e_out_src[f];
  }

  /* EagerGBufProvider map, src/gbufs.jl:13-18 */
  nw->gbufmap = (int64_t*)calloc((size_t)(nw->lastidx_gbuf ? nw->lastidx_gbuf : 1), 8);
  for (int64_t i = 0; i < ne; ++i)
    for (int32_t k = 0; k < nw->vdepth; ++k) {
      nw->gbufmap[nw->e_gbuf_src[i] - 1 + k] = nw->v_out[esrc[i] - 1] + k;
      nw->gbufmap[nw->e_gbuf_dst[i] - 1 + k] = nw->v_out[edst[i] - 1] + k;
    }

  /* AggregationMap, src/aggregators.jl:58-86 */
  int64_t* full = (int64_t*)calloc((size_t)(nw->lastidx_out ? nw->lastidx_out : 1), 8);
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    ndo_batch* B = &nw->eb[b];
    const ndo_espec* s = &especs[B->spec];
    for (int64_t k = 0; k < B->len; ++k) {
      int64_t i = B->indices[k] - 1;
      for (int32_t d = 0; d < s->outdim_dst; ++d) full[nw->e_out_dst[i] - 1 + d] = nw->v_aggr[edst[i] - 1] + d;
      for (int32_t d = 0; d < s->outdim_src; ++d) full[nw->e_out_src[i] - 1 + d] = nw->v_aggr[esrc[i] - 1] + d;
    }
  }
  int64_t first = -1, last = -1;
  for (int64_t k = 0; k < nw->lastidx_out; ++k) if (full[k] != 0) { if (first < 0) first = k; last = k; }
  if (first < 0) { nw->aggr_first = 1; nw->aggr_len = 0; nw->aggmap = (int64_t*)malloc(8); }
  else {
    nw->aggr_first = first + 1; nw->aggr_len = last - first + 1;
    nw->aggmap = (int64_t*)malloc(8 * (size_t)nw->aggr_len);
    memcpy(nw->aggmap, full + first, 8 * (size_t)nw->aggr_len);
  }
  free(full);

  nw->o = (double*)malloc(8 * (size_t)(nw->lastidx_out ? nw->lastidx_out : 1));
  nw->aggbuf = (double*)malloc(8 * (size_t)(nw->lastidx_aggr ? nw->lastidx_aggr : 1));
  nw->gbuf = (double*)malloc(8 * (size_t)(nw->lastidx_gbuf ? nw->lastidx_gbuf : 1));
  return nw;
}

static void free_batches(ndo_batch* b, int32_t n) {
  if (!b) return;
  for (int32_t i = 0; i < n; ++i) free(b[i].indices);
  free(b);
}
void ndo_free(ndo_network* nw) {
  if (!nw) return;
  free(nw->esrc); free(nw->edst); free(nw->vspecs); free(nw->especs); free(nw->vtype); free(nw->etype);
  free(nw->v_data); free(nw->v_out); free(nw->v_para); free(nw->v_aggr);
  free(nw->e_data); free(nw->e_out_src); free(nw->e_out_dst); free(nw->e_para);
  free(nw->e_gbuf_src); free(nw->e_gbuf_dst);
  free_batches(nw->vb, nw->n_vb); free_batches(nw->eb, nw->n_eb);
  free(nw->gbufmap); free(nw->aggmap); free(nw->o); free(nw->aggbuf); free(nw->gbuf);
  free(nw->inv_ptr); free(nw->inv_src);
  free(nw->k1); free(nw->k2); free(nw->k3); free(nw->k4); free(nw->tmp);
  free(nw);
}

int64_t ndo_size(const ndo_network* nw, int which) {
  switch (which) {
    case 0: return nw->lastidx_dynamic; case 1: return nw->lastidx_p; case 2: return nw->lastidx_out;
    case 3: return nw->lastidx_aggr; case 4: return nw->lastidx_gbuf; case 5: return nw->n_vb;
    case 6: return nw->n_eb; case 7: return nw->vdepth; case 8: return nw->edepth;
    case 9: return nw->aggr_first; case 10: return nw->aggr_len;
  }
  return -1;
}
const int64_t* ndo_table(const ndo_network* nw, int which) {
  switch (which) {
    case 0: return nw->v_data; case 1: return nw->v_out; case 2: return nw->v_para; case 3: return nw->v_aggr;
    case 4: return nw->e_data; case 5: return nw->e_out_src; case 6: return nw->e_out_dst; case 7: return nw->e_para;
    case 8: return nw->e_gbuf_src; case 9: return nw->e_gbuf_dst; case 10: return nw->gbufmap; case 11: return nw->aggmap;
  }
  return NULL;
}
int64_t ndo_batch_len(const ndo_network* nw, int kind, int b) { return kind ? nw->eb[b].len : nw->vb[b].len; }
const int64_t* ndo_batch_indices(const ndo_network* nw, int kind, int b) { return kind ? nw->eb[b].indices : nw->vb[b].indices; }
int32_t ndo_batch_spec(const ndo_network* nw, int kind, int b) { return kind ? nw->eb[b].spec : nw->vb[b].spec; }

/* ---------------------------------------------------------------------------------------
 * Model arithmetic.  Expression order exactly as written in the cited Julia source,
 * evaluated left to right, no FMA contraction.
 * ------------------------------------------------------------------------------------- */

/* vertex g (PASS 1).  StateMask: apply_compg(::PureStateMap), src/coreloop.jl:230-233 +
 * src/component_functions.jl:81-99 : out[k] = u[idxs[k]] with idxs = 1:outdim.
 * SWING_DQ: NoFeedForward g(out,u,p,t), u_r = V cos(theta), u_i = V sin(theta)
 * (test/ComponentLibrary.jl:158-159; hand-written equivalent of the MTK-generated g). */
static inline void vertex_g(int kind, int outdim, double* out, const double* u, const double* p) {
  switch (kind) {
    case NDO_V_SWING_DQ: {
      double V = p[3];
      out[0] = V * cos(u[0]);
      out[1] = V * sin(u[0]);
    } break;
    default:
      for (int k = 0; k < outdim; ++k) out[k] = u[k];
  }
}

/* edge g inner function writes the dst output (AntiSymmetric/Symmetric/Directed wrappers call
 * g(odst, vsrc, vdst, p, t), src/component_functions.jl:117-175). */
static inline void edge_g_dst(int kind, double* odst, const double* vs, const double* vd, const double* p) {
  switch (kind) {
    case NDO_E_DIFFUSION:      /* test/ComponentLibrary.jl:8-10 : e .= p * (v_s[1] .- v_d[1]) */
      odst[0] = p[0] * (vs[0] - vd[0]);
      break;
    case NDO_E_DIFFUSION_NOP:  /* benchmark/benchmark_models.jl:5-8 : e[1] = v_s[1] - v_d[1] */
      odst[0] = vs[0] - vd[0];
      break;
    case NDO_E_KURAMOTO:       /* test/ComponentLibrary.jl:51-53, benchmark_models.jl:27-29 */
      odst[0] = p[0] * sin(vs[0] - vd[0]);
      break;
    case NDO_E_LOOPBACK:       /* src/post_utils.jl:105-108 : outdst .= -1 .* insrc (caller passes the vertex depth) */
      break;
    case NDO_E_LINE_DQ: {      /* test/ComponentLibrary.jl:212-245: idst = active*1/Z*(Vsrc-Vdst), Z=R+jX.
                                  Hand-written equivalent (MTK codegen order is not recoverable):
                                  1/Z = (R - jX)/(R^2+X^2). */
      double R = p[0], X = p[1], active = p[2];
      double dr = vs[0] - vd[0];
      double di = vs[1] - vd[1];
      double den = R * R + X * X;
      odst[0] = active * ((R * dr + X * di) / den);
      odst[1] = active * ((R * di - X * dr) / den);
    } break;
  }
}

/* two-sided static g (no wrapper): g(osrc, odst, vsrc, vdst, p, t), src/coreloop.jl:208,220-222 */
static inline void edge_g_two_sided(int kind, double* osrc, double* odst, const double* vs, const double* vd, const double* p) {
  switch (kind) {
    case NDO_E_DIFFUSION_FID:  /* test/ComponentLibrary.jl:22-25 : e_d[1] = p*(v_s[1] .- v_d[1]); e_s[1] = -e_d[1] */
      odst[0] = p[0] * (vs[0] - vd[0]);
      osrc[0] = -odst[0];
      break;
  }
}

/* edge f (PASS 4, edges with states): f(de, e, vsrc, vdst, p, t), src/coreloop.jl:194-211,213-218 */
static inline void edge_f(int kind, double* de, const double* e, const double* vs, const double* vd, const double* p) {
  switch (kind) {
    case NDO_E_DIFFUSION_ODE: { /* test/ComponentLibrary.jl:30-34 : de[1] = 1/tau*(sin(v_s[1]-v_d[1]) - e[1]) ... */
      double tau = p[0];
      de[0] = 1.0 / tau * (sin(vs[0] - vd[0]) - e[0]);
      de[1] = 1.0 / tau * (sin(vd[0] - vs[0]) - e[1]);
    } break;
    case NDO_E_RELAX_ODE:       /* test/diffusion_test.jl:96-100 : de[1] = v_s[1]-v_d[1]-e[1]; de[2] = v_d[1]-v_s[1]-e[2] */
      de[0] = vs[0] - vd[0] - e[0];
      de[1] = vd[0] - vs[0] - e[1];
      break;
  }
}
static inline int edge_kind_has_states(int kind) { return kind == NDO_E_DIFFUSION_ODE || kind == NDO_E_RELAX_ODE; }

/* vertex f (PASS 6): f(du, u, agg, p, t), src/coreloop.jl:176-192,213-218 */
static inline void vertex_f(int kind, double* dv, const double* v, const double* acc, const double* p) {
  switch (kind) {
    case NDO_V_DIFFUSION:      /* test/ComponentLibrary.jl:42-45 : dv[1] = acc[1] */
      dv[0] = acc[0];
      break;
    case NDO_V_KURAMOTO_FIRST: /* test/ComponentLibrary.jl:69-71 : dθ[1] = ω + esum[1] */
      dv[0] = p[0] + acc[0];
      break;
    case NDO_V_KURAMOTO_SECOND: { /* test/ComponentLibrary.jl:59-63 */
      double M = p[0], D = p[1], Pm = p[2];
      dv[0] = v[1];
      dv[1] = 1.0 / M * (Pm - D * v[1] + acc[0]);
    } break;
    case NDO_V_KURAMOTO_SECOND_BENCH: { /* benchmark/benchmark_models.jl:37-41 */
      double P = p[0];
      dv[0] = v[1];
      dv[1] = P - 1.0 * v[1];
      dv[1] += acc[0];
    } break;
    case NDO_V_SWING_DQ: {     /* test/ComponentLibrary.jl:139-161 (hand-written equivalent):
                                  Dt(θ)=ω; Dt(ω)=1/M*(Pmech + Pdamping + Pel),
                                  Pdamping=-D*ω, Pel=u_r*i_r+u_i*i_i, u_r=V cos θ, u_i=V sin θ */
      double M = p[0], D = p[1], Pmech = p[2], V = p[3];
      double ur = V * cos(v[0]);
      double ui = V * sin(v[0]);
      double Pel = ur * acc[0] + ui * acc[1];
      double Pdamping = -D * v[1];
      dv[0] = v[1];
      dv[1] = 1.0 / M * (Pmech + Pdamping + Pel);
    } break;
  }
}

static int check_supported(const ndo_network* nw) {
  for (int32_t b = 0; b < nw->n_vb; ++b) {
    const ndo_vspec* s = &nw->vspecs[nw->vb[b].spec];
    if (s->kind < 0 || s->kind > NDO_V_SWING_DQ) FAIL("vertex kind %d has no RHS in the oracle", s->kind);
  }
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_espec* s = &nw->especs[nw->eb[b].spec];
    if (s->kind < 0 || s->kind > NDO_E_LOOPBACK) FAIL("edge kind %d has no RHS in the oracle", s->kind);
    if (s->kind == NDO_E_LOOPBACK && (s->coupling != NDO_DIRECTED || s->pdim != 0 || s->dim != 0 || nw->vdepth != nw->edepth)) FAIL("a loopback edge is Directed(LOOPBACK_G) with vdepth == edepth");
    if ((s->dim != 0) != edge_kind_has_states(s->kind)) FAIL("edge kind %d: dim %d does not fit the model", s->kind, s->dim);
    if (s->dim != 0) {   /* outputs are StateMasks over the edge's own states */
      if (s->mask_dst < 1 || s->mask_dst + s->outdim_dst - 1 > s->dim) FAIL("edge StateMask (dst) outside the states");
      if (s->coupling == NDO_FIDUCIAL && (s->mask_src < 1 || s->mask_src + s->outdim_src - 1 > s->dim)) FAIL("edge StateMask (src) outside the states");
    } else if ((s->coupling == NDO_FIDUCIAL) != (s->kind == NDO_E_DIFFUSION_FID)) FAIL("two-sided g and the Fiducial coupling go together");
  }
  return 0;
}

/* one vertex batch, component k: g pass */
static inline void vb_g(ndo_network* nw, const ndo_batch* B, const ndo_vspec* s, int64_t k, const double* u, const double* p) {
  /* BatchStride ranges, src/utils.jl:21-29 : start = first + (i-1)*stride */
  const double* uu = u + (B->state_first - 1) + k * s->dim;
  const double* pp = p ? p + (B->p_first - 1) + k * s->pdim : NULL;
  double* out = nw->o + (B->out_first - 1) + k * s->outdim;
  vertex_g(s->kind, s->outdim, out, uu, pp);
}
/* PASS 2: g of edges WITHOUT feed forward = edges whose outputs are StateMasks of their own states
 * (apply_compg(::PureStateMap), src/coreloop.jl:230-233; wrappers src/component_functions.jl:117-203) */
static inline void eb_g_state(ndo_network* nw, const ndo_batch* B, const ndo_espec* s, int64_t k, const double* u) {
  const double* ue = u + (B->state_first - 1) + k * s->dim;
  double* osrc = nw->o + (B->out_first - 1) + k * (s->outdim_src + s->outdim_dst);
  double* odst = osrc + s->outdim_src;
  for (int d = 0; d < s->outdim_dst; ++d) odst[d] = ue[s->mask_dst - 1 + d];
  if (s->coupling == NDO_ANTISYMMETRIC) for (int d = 0; d < s->outdim_src; ++d) osrc[d] = -odst[d];
  else if (s->coupling == NDO_SYMMETRIC) for (int d = 0; d < s->outdim_src; ++d) osrc[d] = odst[d];
  else if (s->coupling == NDO_FIDUCIAL) for (int d = 0; d < s->outdim_src; ++d) osrc[d] = ue[s->mask_src - 1 + d];
}
/* PASS 4: f of edges without feed forward, src/coreloop.jl:76 */
static inline void eb_f(ndo_network* nw, const ndo_batch* B, const ndo_espec* s, int64_t k, double* du, const double* u, const double* p) {
  int vd = nw->vdepth;
  const double* vsrc = nw->gbuf + (B->in_first - 1) + k * 2 * vd;
  const double* vdst = vsrc + vd;
  const double* pp = p ? p + (B->p_first - 1) + k * s->pdim : NULL;
  edge_f(s->kind, du + (B->state_first - 1) + k * s->dim, u + (B->state_first - 1) + k * s->dim, vsrc, vdst, pp);
}
static inline void eb_g(ndo_network* nw, const ndo_batch* B, const ndo_espec* s, int64_t k, const double* p) {
  int vd = nw->vdepth;
  const double* vsrc = nw->gbuf + (B->in_first - 1) + k * 2 * vd;      /* get_src_dst, src/gbufs.jl:27-31 */
  const double* vdst = vsrc + vd;
  const double* pp = p ? p + (B->p_first - 1) + k * s->pdim : NULL;
  double* osrc = nw->o + (B->out_first - 1) + k * (s->outdim_src + s->outdim_dst);
  double* odst = osrc + s->outdim_src;
  if (s->coupling == NDO_FIDUCIAL) { edge_g_two_sided(s->kind, osrc, odst, vsrc, vdst, pp); return; }
  if (s->kind == NDO_E_LOOPBACK) { for (int d = 0; d < s->outdim_dst; ++d) odst[d] = -1.0 * vsrc[d]; return; }
  edge_g_dst(s->kind, odst, vsrc, vdst, pp);
  if (s->coupling == NDO_ANTISYMMETRIC)      /* src/component_functions.jl:117-127 */
    for (int d = 0; d < s->outdim_src; ++d) osrc[d] = -odst[d];
  else if (s->coupling == NDO_SYMMETRIC)     /* :142-152 */
    for (int d = 0; d < s->outdim_src; ++d) osrc[d] = odst[d];
  /* Directed: osrc is empty (:167-175) */
}
static inline void vb_f(ndo_network* nw, const ndo_batch* B, const ndo_vspec* s, int64_t k, double* du, const double* u, const double* p) {
  const double* uu = u + (B->state_first - 1) + k * s->dim;
  double* dd = du + (B->state_first - 1) + k * s->dim;
  const double* pp = p ? p + (B->p_first - 1) + k * s->pdim : NULL;
  const double* acc = nw->aggbuf + (B->in_first - 1) + k * nw->edepth;
  vertex_f(s->kind, dd, uu, acc, pp);
}

/* gen_loopback_map + _apply_loopback!, src/post_utils.jl:213-234: for every loopback edge copy the dst (hub) vertex's
 * output into the src (injector) vertex's aggregation slot */
static void apply_loopback(ndo_network* nw) {
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_batch* B = &nw->eb[b];
    if (nw->especs[B->spec].kind != NDO_E_LOOPBACK) continue;
    for (int64_t k = 0; k < B->len; ++k) {
      int64_t i = B->indices[k] - 1;
      const double* ov = nw->o + nw->v_out[nw->edst[i] - 1] - 1;
      double* av = nw->aggbuf + nw->v_aggr[nw->esrc[i] - 1] - 1;
      for (int d = 0; d < nw->vdepth; ++d) av[d] = ov[d];
    }
  }
}

/* (nw::Network)(du,u,p,t), src/coreloop.jl:1-102, SequentialExecution{true} (:111-119),
 * SequentialAggregator (src/aggregators.jl:140-151). */
int ndo_rhs_sequential(ndo_network* nw, double* du, const double* u, const double* p, double t) {
  (void)t;
  if (check_supported(nw)) return -1;
  for (int64_t i = 0; i < nw->lastidx_dynamic; ++i) du[i] = 0.0;          /* coreloop.jl:24 */
  for (int64_t i = 0; i < nw->lastidx_out; ++i) nw->o[i] = NAN;           /* network_structure.jl:153 */
  for (int64_t i = 0; i < nw->lastidx_aggr; ++i) nw->aggbuf[i] = 0.0;     /* coreloop.jl:30 */
  /* PASS 1: vertex g without ff, coreloop.jl:39 */
  for (int32_t b = 0; b < nw->n_vb; ++b) {
    const ndo_batch* B = &nw->vb[b]; const ndo_vspec* s = &nw->vspecs[B->spec];
    for (int64_t k = 0; k < B->len; ++k) vb_g(nw, B, s, k, u, p);
  }
  /* PASS 2: g of edges without ff (edges with states), coreloop.jl:41 */
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_batch* B = &nw->eb[b]; const ndo_espec* s = &nw->especs[B->spec];
    if (s->dim == 0) continue;
    for (int64_t k = 0; k < B->len; ++k) eb_g_state(nw, B, s, k, u);
  }
  /* apply_loopback!, coreloop.jl:47 + post_utils.jl:213-234 : injector input <- hub output */
  apply_loopback(nw);
  /* PASS 3 (ff vertices), externals: empty for the restated models */
  /* gather!, coreloop.jl:67 + gbufs.jl:25 */
  for (int64_t k = 0; k < nw->lastidx_gbuf; ++k) nw->gbuf[k] = nw->o[nw->gbufmap[k] - 1];
  /* PASS 4: f of edges without ff, coreloop.jl:76 */
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_batch* B = &nw->eb[b]; const ndo_espec* s = &nw->especs[B->spec];
    if (s->dim == 0) continue;
    for (int64_t k = 0; k < B->len; ++k) eb_f(nw, B, s, k, du, u, p);
  }
  /* PASS 5: fg of ff edges (static edges: g only), coreloop.jl:78 */
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_batch* B = &nw->eb[b]; const ndo_espec* s = &nw->especs[B->spec];
    if (s->dim != 0) continue;
    for (int64_t k = 0; k < B->len; ++k) eb_g(nw, B, s, k, p);
  }
  /* aggregate!, aggregators.jl:140-151 : single ascending sweep */
  {
    const double* dat = nw->o + (nw->aggr_first - 1);
    for (int64_t k = 0; k < nw->aggr_len; ++k) {
      int64_t dst = nw->aggmap[k];
      if (dst != 0) nw->aggbuf[dst - 1] = nw->aggbuf[dst - 1] + dat[k];
    }
  }
  /* PASS 6: vertex f, coreloop.jl:97 */
  for (int32_t b = 0; b < nw->n_vb; ++b) {
    const ndo_batch* B = &nw->vb[b]; const ndo_vspec* s = &nw->vspecs[B->spec];
    for (int64_t k = 0; k < B->len; ++k) vb_f(nw, B, s, k, du, u, p);
  }
  return 0;
}

int ndo_rhs_sequential_bufs(ndo_network* nw, double* du, const double* u, const double* p, double t,
                            double* o_out, double* aggbuf_out) {
  int rc = ndo_rhs_sequential(nw, du, u, p, t);
  if (rc) return rc;
  if (o_out) memcpy(o_out, nw->o, 8 * (size_t)nw->lastidx_out);
  if (aggbuf_out) memcpy(aggbuf_out, nw->aggbuf, 8 * (size_t)nw->lastidx_aggr);
  return 0;
}

/* _inv_aggregation_map, src/aggregators.jl:203-230: per aggbuf slot, the o indices that feed it,
 * ascending.  Stored CSR-style instead of Vector{Tuple{Int,Vector{Int}}}. */
static void build_inverse(ndo_network* nw) {
  if (nw->inv_ptr) return;
  int64_t ns = nw->lastidx_aggr;
  nw->inv_ptr = (int64_t*)calloc((size_t)ns + 2, 8);
  for (int64_t k = 0; k < nw->aggr_len; ++k) if (nw->aggmap[k]) nw->inv_ptr[nw->aggmap[k] + 1]++;
  for (int64_t s = 0; s < ns + 1; ++s) nw->inv_ptr[s + 1] += nw->inv_ptr[s];
  nw->inv_src = (int64_t*)malloc(8 * (size_t)(nw->inv_ptr[ns + 1] ? nw->inv_ptr[ns + 1] : 1));
  int64_t* cur = (int64_t*)malloc(8 * ((size_t)ns + 2));
  memcpy(cur, nw->inv_ptr, 8 * ((size_t)ns + 2));
  for (int64_t k = 0; k < nw->aggr_len; ++k) {
    int64_t dst = nw->aggmap[k];
    if (dst) nw->inv_src[cur[dst]++] = (nw->aggr_first - 1) + k;   /* 0-based o index */
  }
  free(cur);
}

int ndo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ThreadedExecution{true}: Threads.@threads over each batch (src/coreloop.jl:121-129; equal
 * static chunks) + ThreadedAggregator (src/aggregators.jl:192-201; parallel over slots, each
 * slot summed in ascending o order -> same result as sequential). */
int ndo_rhs_threaded(ndo_network* nw, double* du, const double* u, const double* p, double t, int nthreads) {
  (void)t;
  if (check_supported(nw)) return -1;
  build_inverse(nw);
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  /* fill! calls are serial in the reference (coreloop.jl:24,30; network_structure.jl:153) */
  memset(du, 0, 8 * (size_t)nw->lastidx_dynamic);
  for (int64_t i = 0; i < nw->lastidx_out; ++i) nw->o[i] = NAN;
  memset(nw->aggbuf, 0, 8 * (size_t)nw->lastidx_aggr);
  for (int32_t b = 0; b < nw->n_vb; ++b) {
    const ndo_batch* B = &nw->vb[b]; const ndo_vspec* s = &nw->vspecs[B->spec];
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < B->len; ++k) vb_g(nw, B, s, k, u, p);
  }
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_batch* B = &nw->eb[b]; const ndo_espec* s = &nw->especs[B->spec];
    if (s->dim == 0) continue;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < B->len; ++k) eb_g_state(nw, B, s, k, u);
  }
  apply_loopback(nw);   /* serial in the reference (CPU version of _apply_loopback!) */
  /* NNlib.gather! on Vectors is multithreaded over the destination */
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t k = 0; k < nw->lastidx_gbuf; ++k) nw->gbuf[k] = nw->o[nw->gbufmap[k] - 1];
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_batch* B = &nw->eb[b]; const ndo_espec* s = &nw->especs[B->spec];
    if (s->dim == 0) continue;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < B->len; ++k) eb_f(nw, B, s, k, du, u, p);
  }
  for (int32_t b = 0; b < nw->n_eb; ++b) {
    const ndo_batch* B = &nw->eb[b]; const ndo_espec* s = &nw->especs[B->spec];
    if (s->dim != 0) continue;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < B->len; ++k) eb_g(nw, B, s, k, p);
  }
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t s = 1; s <= nw->lastidx_aggr; ++s) {
    double acc = nw->aggbuf[s - 1];
    for (int64_t j = nw->inv_ptr[s]; j < nw->inv_ptr[s + 1]; ++j) acc = acc + nw->o[nw->inv_src[j]];
    nw->aggbuf[s - 1] = acc;
  }
  for (int32_t b = 0; b < nw->n_vb; ++b) {
    const ndo_batch* B = &nw->vb[b]; const ndo_vspec* s = &nw->vspecs[B->spec];
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < B->len; ++k) vb_f(nw, B, s, k, du, u, p);
  }
  return 0;
}

/* Classical RK4.  The reference has no fixed-step integrator of its own (OrdinaryDiffEq's RK4 is
 * an external package) -> operation order defined here and mirrored by the engine:
 *   k1=f(u,t); k2=f(u+(dt/2)k1,t+dt/2); k3=f(u+(dt/2)k2,t+dt/2); k4=f(u+dt*k3,t+dt);
 *   u <- u + (dt/6)*(((k1+2k2)+2k3)+k4) */
int ndo_rk4(ndo_network* nw, double* u, const double* p, double t0, double dt, int64_t nsteps, int nthreads) {
  int64_t n = nw->lastidx_dynamic;
  if (!nw->k1) {
    nw->k1 = (double*)malloc(8 * (size_t)n); nw->k2 = (double*)malloc(8 * (size_t)n);
    nw->k3 = (double*)malloc(8 * (size_t)n); nw->k4 = (double*)malloc(8 * (size_t)n);
    nw->tmp = (double*)malloc(8 * (size_t)n);
  }
  double h2 = 0.5 * dt, h6 = dt / 6.0;
#define RHS(du, uu, tt) ((nthreads > 1) ? ndo_rhs_threaded(nw, du, uu, p, tt, nthreads) : ndo_rhs_sequential(nw, du, uu, p, tt))
  for (int64_t s = 0; s < nsteps; ++s) {
    double t = t0 + (double)s * dt;
    if (RHS(nw->k1, u, t)) return -1;
    for (int64_t i = 0; i < n; ++i) nw->tmp[i] = u[i] + h2 * nw->k1[i];
    if (RHS(nw->k2, nw->tmp, t + h2)) return -1;
    for (int64_t i = 0; i < n; ++i) nw->tmp[i] = u[i] + h2 * nw->k2[i];
    if (RHS(nw->k3, nw->tmp, t + h2)) return -1;
    for (int64_t i = 0; i < n; ++i) nw->tmp[i] = u[i] + dt * nw->k3[i];
    if (RHS(nw->k4, nw->tmp, t + dt)) return -1;
    for (int64_t i = 0; i < n; ++i)
      u[i] = u[i] + h6 * (((nw->k1[i] + 2.0 * nw->k2[i]) + 2.0 * nw->k3[i]) + nw->k4[i]);
  }
#undef RHS
  return 0;
}
