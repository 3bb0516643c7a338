/*
 * nd_oracle.h -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * CPU restatement of the NetworkDynamics.jl network right-hand side
 * `nw(du, u, p, t)` with SequentialExecution{true} + SequentialAggregator(+),
 * plus an OpenMP restatement of ThreadedExecution{true} + ThreadedAggregator
 * (the reference's CPU path, used only as the timed CPU baseline).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load this library.  The product (networkdynamics.jl_b200/) never does.
 *
 * Parity pinning: the reference is Julia and cannot run in this environment
 * (no julia binary, no depot, no network).  The restatement is pinned against
 * the reference's own Julia-free known answers:
 *   - du == -L*x for the diffusion network      (test/diffusion_test.jl:80-90)
 *   - find_identical batching order             (test/utils_test.jl:43-61)
 *   - flat state/parameter layout pins          (test/symbolicindexing_test.jl:27-62,98-111)
 *   - the hand-worked 3-vertex layout           
 *   - edges with states on their constraint     (test/diffusion_test.jl:96-129)
 *   - loopback identities                       (src/post_utils.jl:105-234)
 * The reference ships no stored golden vectors and could not be run to make any.
 * PARITY UNPINNED for two things: the last-bit behaviour of Julia's Base.sin/cos
 * (libm is used here) and the operation order of the fixed-step RK4 (the
 * reference's stepper is OrdinaryDiffEq's, an external package; classical RK4 is
 * restated).  Both sit inside the stated tolerances (1e-12 per RHS, 1e-9 per
 * 1000-step trajectory).
 *
 * All indices in the emitted tables are 1-based int64, exactly as the
 * reference's IndexManager holds them.
 */
#ifndef ND_ORACLE_H
#define ND_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* vertex kinds (model arithmetic: see nd_oracle.c for file:line citations) */
enum {
  NDO_V_DIFFUSION = 0,          /* dv = acc                                   */
  NDO_V_KURAMOTO_FIRST = 1,     /* dth = omega + esum                         */
  NDO_V_KURAMOTO_SECOND = 2,    /* p=(M,D,Pm)                                 */
  NDO_V_KURAMOTO_SECOND_BENCH = 3, /* p=(P,), benchmark variant               */
  NDO_V_SWING_DQ = 4,           /* dq swing bus, NoFeedForward g              */
  NDO_V_OPAQUE = 100            /* layout only (index pins); RHS rejects it   */
};
/* edge kinds */
enum {
  NDO_E_DIFFUSION = 0,          /* e = p*(vs - vd)                            */
  NDO_E_DIFFUSION_NOP = 1,      /* e = vs - vd                                */
  NDO_E_KURAMOTO = 2,           /* e = K*sin(ths - thd)                       */
  NDO_E_LINE_DQ = 3,            /* static dq line, p=(R,X,active)             */
  NDO_E_DIFFUSION_ODE = 4,      /* ODE edge, dim 2, p=(tau,): de = 1/tau*(sin(..)-e); outputs are states  */
  NDO_E_RELAX_ODE = 5,          /* ODE edge, dim 2, no p: de1 = vs-vd-e1, de2 = vd-vs-e2                 */
  NDO_E_DIFFUSION_FID = 6,      /* static two-sided g: e_d = p*(vs-vd); e_s = -e_d                       */
  NDO_E_LOOPBACK = 7,           /* LoopbackConnection: Directed(LOOPBACK_G), out_dst = -in_src; the src vertex's
                                   input is the dst vertex's output (apply_loopback!)                    */
  NDO_E_OPAQUE = 100
};
/* edge output wrappers, src/component_functions.jl:117-203 */
enum { NDO_ANTISYMMETRIC = 0, NDO_SYMMETRIC = 1, NDO_DIRECTED = 2, NDO_FIDUCIAL = 3 };

typedef struct {
  int32_t kind;
  int32_t dim, pdim, outdim;
} ndo_vspec;

typedef struct {
  int32_t kind;
  int32_t coupling;
  int32_t dim, pdim, outdim_src, outdim_dst;
  /* edges with states (dim > 0): the outputs are StateMasks (src/component_functions.jl:81-99) -- dst output k is state
   * mask_dst + k, and for Fiducial(src=..., dst=...) src output k is state mask_src + k (1-based firsts; 0 = unused) */
  int32_t mask_src, mask_dst;
} ndo_espec;

typedef struct ndo_network ndo_network;

/* Build the reference index tables.
 *  vtype[i] / etype[i]: index into vspecs / especs of the model of vertex i / edge i
 *  (the "model object identity" that _component_hash distinguishes).
 *  single_vmodel / single_emodel != 0 restates the single-model shortcut
 *  (src/construction.jl:156-168).
 *  esrc/edst: 1-based, already in Graphs.jl edges(g) order.
 * Returns NULL on error (message via ndo_last_error). */
ndo_network* ndo_build(int64_t nv, int64_t ne, const int64_t* esrc, const int64_t* edst,
                       int32_t n_vspecs, const ndo_vspec* vspecs, const int32_t* vtype,
                       int32_t n_especs, const ndo_espec* especs, const int32_t* etype);
void ndo_free(ndo_network*);
const char* ndo_last_error(void);

/* scalar sizes: which = 0 lastidx_dynamic, 1 lastidx_p, 2 lastidx_out, 3 lastidx_aggr,
 * 4 lastidx_gbuf, 5 n_vbatches, 6 n_ebatches, 7 vdepth, 8 edepth,
 * 9 aggmap range first (1-based), 10 aggmap length */
int64_t ndo_size(const ndo_network*, int which);

/* per-component range starts (1-based); which:
 * 0 v_data 1 v_out 2 v_para 3 v_aggr            (length nv)
 * 4 e_data 5 e_out_src 6 e_out_dst 7 e_para 8 e_gbuf_src 9 e_gbuf_dst (length ne)
 * 10 gbuf map (length lastidx_gbuf)  11 aggregation map (length aggmap length) */
const int64_t* ndo_table(const ndo_network*, int which);

/* batches: kind 0 = vertex, 1 = edge */
int64_t ndo_batch_len(const ndo_network*, int kind, int b);
const int64_t* ndo_batch_indices(const ndo_network*, int kind, int b);
int32_t ndo_batch_spec(const ndo_network*, int kind, int b);

/* SequentialExecution{true} + SequentialAggregator(+): src/coreloop.jl:1-102 */
int ndo_rhs_sequential(ndo_network*, double* du, const double* u, const double* p, double t);
/* same, but also hands back o / aggbuf (RET=:buf_init plus the f pass) */
int ndo_rhs_sequential_bufs(ndo_network*, double* du, const double* u, const double* p, double t,
                            double* o_out, double* aggbuf_out);
/* ThreadedExecution{true} + ThreadedAggregator(+): src/coreloop.jl:121-129,
 * src/aggregators.jl:159-235.  nthreads<=0: omp default. */
int ndo_rhs_threaded(ndo_network*, double* du, const double* u, const double* p, double t, int nthreads);
/* classical fixed-step RK4 driven by ndo_rhs_sequential (threaded if nthreads>1) */
int ndo_rk4(ndo_network*, double* u, const double* p, double t0, double dt, int64_t nsteps, int nthreads);
int ndo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
