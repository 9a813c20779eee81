/*
 * include/ipmz.h -- C ABI of the B200-native ipm-zoo numerical interior-point hot path.
 *
 * The reference (albfre/ipm-zoo) has no FFI: its boundary for this path is the C++ API of
 * namespace NumericalOptimization.  This header is the thin C layer a C++ adapter with the
 * reference's own class/function names binds to (ipm-zoo_b200/host/):
 *
 *   reference interface (file:line)                                   replaced by
 *   ----------------------------------------------------------------  -------------------------
 *   struct Data                EnvironmentBuilder.h:7-17              ipmz_problem
 *   SymbolicOptimization::Settings  SymbolicOptimization.h:58-64      ipmz_problem.{ineq_bounds,
 *                                                                      var_bounds,equalities}
 *   build_environment (initial point) EnvironmentBuilder.cpp:34-73    ipmz_create (device-side)
 *   Optimizer::Optimizer       Optimizer.h:15-18, Optimizer.cpp:27-61 ipmz_create
 *   Optimizer::solve           Optimizer.h:20, Optimizer.cpp:63-220   ipmz_solve
 *   Environment as in/out state Evaluation.h:22                       ipmz_set_iterate / ipmz_get_iterate
 *   stdout "iter:" / "b:" lines Optimizer.cpp:131-132, :359           ipmz_result + ipmz_get_trace
 *   LinearSolvers::ldlt_decomposition     LinearSolvers.h:11          ipmz_ldlt_decomposition
 *   LinearSolvers::overwriting_solve_ldlt LinearSolvers.h:16-17       ipmz_overwriting_solve_ldlt
 *   LinearSolvers::symmetric_indefinite_factorization  LinearSolvers.h:24-25  ipmz_symmetric_indefinite_factorization
 *   LinearSolvers::overwriting_solve_bunch_kaufman     LinearSolvers.h:29-31  ipmz_overwriting_solve_bunch_kaufman
 *   (batched independent QPs: no reference counterpart, north_star)   ipmz_batch_*
 *
 * Conventions: every entry point returns 0 on success and a non-zero ipmz_status otherwise;
 * ipmz_last_error() gives the message.  The C++ adapter turns non-zero into
 * Utils::AssertionError-compatible std::logic_error (reference convention,
 * include/Utils/Assert.h:7-12).  No exceptions or C++ types cross this boundary.  Host
 * buffers are caller-owned; device buffers are library-owned and live until *_destroy.
 * All matrices are dense row-major IEEE FP64.  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with IPMZ_ERR_CUDA.
 *
 * Packed iterate layout (length ipmz_iterate_len = 5n + 6 m_ineq + 6 m_eq), names as in
 * SymbolicOptimization.h:5-26:
 *   x[n] lamA[mi] s[mi] lamg[mi] lamh[mi] g[mi] h[mi]
 *   lamC[me] t[me] lamv[me] lamw[me] v[me] w[me]  lamy[n] lamz[n] y[n] z[n]
 */
#ifndef IPMZ_H
#define IPMZ_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  IPMZ_OK = 0,
  IPMZ_ERR_ARG = 1,        /* bad argument / unsupported Settings combination */
  IPMZ_ERR_CUDA = 2,       /* CUDA runtime failure or no device */
  IPMZ_ERR_ALLOC = 3,
  IPMZ_ERR_INDEFINITE = 4, /* reference: solve_indefinite_() == ASSERT(false), Optimizer.cpp:75 */
  IPMZ_ERR_BOUNDS = 5      /* reference: ASSERT(l < u), EnvironmentBuilder.cpp:10-17 */
} ipmz_status;

enum { IPMZ_BOUNDS_NONE = 0, IPMZ_BOUNDS_LOWER = 1, IPMZ_BOUNDS_UPPER = 2, IPMZ_BOUNDS_BOTH = 3 };

/* ipmz_problem.equalities: Settings::equalities + Settings::equality_handling (SymbolicOptimization.h:42-64). */
enum {
  IPMZ_EQ_OFF = 0,            /* Settings::equalities == false */
  IPMZ_EQ_SLACKED_SLACKS = 1, /* EqualityHandling::SlackedSlacks: C x - t = 0, t - v = d, t + w = d (quasi-definite) */
  IPMZ_EQ_NONE = 2,           /* EqualityHandling::None: C x = d with multiplier lambda_C only
                                 (SymbolicOptimization.cpp:137-140).  The augmented system gets a zero diagonal
                                 block, which the reference routes to solve_indefinite_() == ASSERT(false)
                                 (Optimizer.cpp:63-75); here it is factorized with Bunch-Kaufman pivoting
                                 (LinearSolvers.cpp:76-318 on the device).  AUGMENTED reduction only. */
  IPMZ_EQ_REGULARIZATION = 3, /* EqualityHandling::Regularization (SymbolicOptimization.cpp:184-192): objective +
                                 1/2 p^T p, rows C x - d + delta p = 0 with delta = ipmz_options.delta_eq.  Eliminating p
                                 leaves the scalar block -delta^2 I on the diagonal, which the reference's evaluator
                                 cannot assemble (Evaluation.cpp:53-60); here the rows are quasi-definite like any other.
                                 p travels in the `t` slot of the packed iterate.  AUGMENTED or NORMAL reduction. */
  IPMZ_EQ_PENALTY = 4         /* EqualityHandling::PenaltyFunction and PenaltyFunctionWithExtraDual
                                 (SymbolicOptimization.cpp:173-183): both give the Newton rows  C dx - mu dlambda_C =
                                 d + mu lambda_C - C x  (get_newton_system), i.e. the scalar block -mu I that depends on the
                                 barrier parameter and that the reference's evaluator cannot assemble (Evaluation.cpp:53-60).
                                 mu follows the reference's loop (Optimizer.cpp:138-181): the matrix is assembled with the
                                 value the previous iteration left in the environment (sigma mu; 1 before the first iteration,
                                 EnvironmentBuilder.cpp:48), the predictor's residual with mu = 0, the corrector's with the new
                                 sigma mu.  Only the multiplier exists (no slacks).  AUGMENTED reduction only (condensing -mu I
                                 onto dx would put C^T C / mu, mu -> 0, into the matrix). */
};

/* Which reduction of the Newton system is assembled and factorized (north_star). */
enum {
  IPMZ_REDUCTION_AUGMENTED = 0, /* quasi-definite [[Hx, M^T],[M, -W^-1]], LDL^T, N = n+m   */
  IPMZ_REDUCTION_NORMAL = 1,    /* primal condensed Hx + M^T W M, root-free Cholesky, N = n */
  IPMZ_REDUCTION_FULL = 2,      /* un-reduced Newton system (SymbolicOptimization.cpp:417-433), symmetrised and
                                   ordered so that unpivoted LDL^T performs the reference's block elimination;
                                   N = n + 2m + 2n*(sides of the box) + 2m*(sides of the rows) <= 5n + 6m */
  IPMZ_REDUCTION_DUAL_NORMAL = 3 /* the reference's own "normal equations" (get_normal_equations,
                                   SymbolicOptimization.cpp:465-478: row 0 = dx eliminated): root-free Cholesky of Hx
                                   (n x n), S = W^-1 + M Hx^-1 M^T (m x m) formed by the panel solves + DMMA updates of
                                   the rows of M, root-free Cholesky of S,  S dlam = M Hx^-1 b0 - b1,
                                   dx = Hx^-1 (b0 - M^T dlam).  ipmz_assemble returns S (N_out = m). */
};

typedef struct {
  int n;       /* variables */
  int m_ineq;  /* rows of A_ineq */
  int m_eq;    /* rows of A_eq (handled as EqualityHandling::SlackedSlacks) */
  const double* Q;   /* n x n, symmetric */
  const double* c;   /* n */
  const double* A;   /* m_ineq x n */
  const double* l_A; /* m_ineq */
  const double* u_A; /* m_ineq */
  const double* C;   /* m_eq x n */
  const double* d;   /* m_eq */
  const double* l_x; /* n */
  const double* u_x; /* n */
  int ineq_bounds;   /* Settings::inequalities      (IPMZ_BOUNDS_*) */
  int var_bounds;    /* Settings::variable_bounds   (IPMZ_BOUNDS_*) */
  int equalities;    /* IPMZ_EQ_* (any non-zero value other than IPMZ_EQ_NONE means SlackedSlacks) */
} ipmz_problem;

typedef struct {
  double tolerance;            /* 1e-8   Optimizer.cpp:124 */
  int max_iter;                /* 100    Optimizer.cpp:125 */
  double fraction_to_boundary; /* 0.995  Optimizer.cpp:216 */
  double sigma_power;          /* 3      Optimizer.cpp:178 */
  int reduction;               /* IPMZ_REDUCTION_* */
  int device;                  /* CUDA device ordinal */
  int record_steps;            /* keep every iteration's solved Newton steps for ipmz_get_trace */
  int refine_steps;            /* normal reduction: iterative-refinement steps against the augmented
                                  residual per Newton solve; -1 = default policy by each problem's mu (none while
                                  mu >= 1e-3, then 1; 2 for a single QP once mu < 1e-6); ignored for AUGMENTED */
  double delta_eq;             /* 1e-4   EnvironmentBuilder.cpp:48 (IPMZ_EQ_REGULARIZATION only) */
} ipmz_options;

typedef struct {
  int iterations; /* Newton steps taken */
  int converged;  /* 1 iff res < tol && mu < tol was met */
  double f;       /* objective at the last evaluated iterate */
  double res;     /* ||full Newton RHS at mu=0||_2  (Optimizer.cpp:240-247) */
  double mu;      /* mean |complementarity|         (Optimizer.cpp:249-268) */
  double solve_ms;        /* device time of the loop, CUDA events */
  double factor_flops;    /* sum over iterations of N^3/3 */
} ipmz_result;

typedef struct ipmz_solver_s* ipmz_handle;
typedef struct ipmz_batch_s* ipmz_batch_handle;
typedef struct ipmz_factor_s* ipmz_factor_handle;

const char* ipmz_last_error(void);
const char* ipmz_version(void);
int ipmz_device_count(void);
void ipmz_default_options(ipmz_options* opt);
int ipmz_iterate_len(const ipmz_problem* p);
/* Host-only: offsets of the unknown groups of the FULL reduction, dy dz dsl dsu dlam_y dlam_z dlam_l dlam_u ds dx dlam,
 * followed by the system size N (12 ints; absent groups have zero width).  Only sizes and Settings of *p are read. */
int ipmz_full_layout(const ipmz_problem* p, int* offsets12);
/* Page-locked host buffers for the end-to-end path (H2D / D2H at full PCIe rate). */
void* ipmz_host_alloc(size_t bytes);
void ipmz_host_free(void* p);
/* Kernels launched by this library so far (bench.py reports the delta as gpu_launches). */
unsigned long long ipmz_launch_count(void);
/* FP64 tensor-pipe (DMMA) issue-rate ceiling of `device`, measured with a register-resident probe. */
int ipmz_fp64_peak_probe(int device, double* tflops);

/* ---- single QP, iterate device-resident across iterations ---- */
int ipmz_create(const ipmz_problem* p, const ipmz_options* opt, ipmz_handle* out);
int ipmz_destroy(ipmz_handle h);
int ipmz_set_iterate(ipmz_handle h, const double* packed);
int ipmz_get_iterate(ipmz_handle h, double* packed);
int ipmz_reset_iterate(ipmz_handle h); /* the reference's initial point */
int ipmz_solve(ipmz_handle h, ipmz_result* res);
/* One predictor-corrector iteration at the current iterate WITHOUT moving it: writes the
 * solved augmented steps [dx; dlam] (length n + m_ineq + m_eq) of the predictor and the
 * corrector, as the reference prints them at Optimizer.cpp:359; any pointer may be NULL. */
int ipmz_newton_step(ipmz_handle h, double* step_aff, double* step_cor, double* alpha_aff,
                     double* sigma, double* alpha);
/* What the reference leaves in the Environment besides the iterate (Optimizer.cpp:147-157, :188-209, :369, :377): the
 * LAST Newton iteration's corrector direction (`\Delta v` keys), affine direction (`\Delta v_affine` keys), shorthand
 * residuals r_* (corrector values: complementarity rows include the affine second-order term) -- each in the packed
 * iterate layout, the entry of variable v at v's slot -- and the centred barrier parameter sigma*mu.  Any pointer may
 * be NULL.  Meaningful after an ipmz_solve that took at least one iteration. */
int ipmz_get_last_iteration(ipmz_handle h, double* delta, double* delta_affine, double* residuals, double* mu_centered);
/* Per-iteration log of the last ipmz_solve: f/res/mu have iterations+1 entries (cap = how
 * many the caller's arrays hold); steps need record_steps and hold cap x (n+m) doubles. */
int ipmz_get_trace(ipmz_handle h, int cap, double* f, double* res, double* mu, double* step_aff,
                   double* step_cor, double* alpha_aff, double* sigma, double* alpha);
/* Dense copy (N x N, row-major, full symmetric) of the reduced matrix assembled at the
 * current iterate, N = n+m (augmented), n (normal) or the full-system size (full; unknowns ordered
 * dy dz dsl dsu | dlam_y dlam_z dlam_l dlam_u | ds | dx | dlam, absent groups skipped). */
int ipmz_assemble(ipmz_handle h, double* K_host, int* N_out);

/* Diagnostics for bench.py (HBM roofline of the streaming kernels): CUDA-event time per launch (reps back to back on
 * the library stream) and algorithmic bytes per launch at the current iterate, 7 slots: k_matvec Q x | k_matvec M x |
 * k_matvec M^T lambda | k_assemble | k_residuals_rhs<0> | k_backsub_step<0> | k_update (alpha forced to 0). */
int ipmz_probe_kernels(ipmz_handle h, int reps, double* ms_per_launch, double* bytes_per_launch);

/* ---- mirror of LinearSolvers (host buffers in/out, as the reference's free functions) ---- */
/* L: n x n unit lower triangular (zeros above the diagonal), D: n pivots; a zero pivot is
 * replaced by 1e-8 (LinearSolvers.cpp:28). */
int ipmz_ldlt_decomposition(int n, const double* A, double* L, double* D);
int ipmz_overwriting_solve_ldlt(int n, const double* L, const double* D, double* b);

/* Bunch-Kaufman (LinearSolvers.cpp:76-318).  LD: n x n copy of A whose lower triangle holds L and the 1x1 / 2x2
 * blocks of D (upper triangle = A's, as in the reference); ipiv[k] >= 0: 1x1 pivot interchanged with row ipiv[k];
 * ipiv[k] = ipiv[k+1] = -kp: 2x2 pivot whose second row was interchanged with row kp.  The factorization is
 * bit-exact against the reference (same pivots, same bits); the solve agrees to rounding. */
int ipmz_symmetric_indefinite_factorization(int n, const double* A, double* LD, int* ipiv);
int ipmz_overwriting_solve_bunch_kaufman(int n, const double* LD, const int* ipiv, double* b);
/* Device time (CUDA events) of one Bunch-Kaufman factorization of A, averaged over reps runs from a pristine device copy. */
int ipmz_bk_factor_time(int n, const double* A, int reps, double* ms_per_factorization);

/* ---- device-resident factor + solve (bench / roofline path) ---- */
int ipmz_factor_create(int n, int device, ipmz_factor_handle* out);
int ipmz_factor_destroy(ipmz_factor_handle h);
int ipmz_factor_set_matrix(ipmz_factor_handle h, const double* A_host); /* pristine copy kept */
int ipmz_factor_set_rhs(ipmz_factor_handle h, const double* b_host);
/* reps x { L,D <- LDL^T(pristine A) ; nrhs solves from the pristine rhs }, timed with CUDA
 * events on the library stream; ms_total is the device time of all reps. */
int ipmz_factor_run(ipmz_factor_handle h, int reps, int nrhs, double* ms_total);
/* One factorization with CUDA events around every launch: ms3[0..2] = device time of the
 * diagonal-block, panel and trailing-update (DMMA) kernels; flops_syrk / n_syrk = algorithmic
 * flops and launch count of the trailing updates (the roofline numerator of bench.py). */
int ipmz_factor_profile(ipmz_factor_handle h, double* ms3, double* flops_syrk, int* n_syrk);
/* Which factorization path the handle runs: *dataflow = 1 for the persistent single-launch
 * dataflow kernel (n >= IPMZ_DATAFLOW_MIN_N, default 512), with its task count and the
 * makespan of the host-simulated list schedule; 0 for the multi-kernel schedule. */
int ipmz_factor_info(ipmz_factor_handle h, int* dataflow, int* ntasks, double* simulated_us);
/* Host-only (no GPU needed): build the dataflow task list of an n x n matrix for `workers` SMs and
 * check that it is a topological order covering every tile exactly once (the invariant that makes
 * the device-side ticket queue deadlock-free).  counts3 = DIAG / TRSM / UPD tasks.  0 = valid. */
int ipmz_schedule_check(int n, int workers, int* counts3, double* makespan_us, double* work_us);
/* Host-only: the UPD task list of the condensed assembly M^T W M (n x n, inner dimension m) on the dataflow kernel
 * covers every lower tile's K range exactly once and in order; *ntasks = 0 when that path does not apply. */
int ipmz_assembly_schedule_check(int n, int m, int* ntasks);
int ipmz_factor_get_solution(ipmz_factor_handle h, double* x_host);
int ipmz_factor_get_ld(ipmz_factor_handle h, double* L_host, double* D_host);

/* ---- batch of independent QPs with one shared shape/Settings (cfg4) ---- */
/* Arrays hold `count` problems back to back (Q: count*n*n, A: count*m_ineq*n, ...). */
int ipmz_batch_create(int count, const ipmz_problem* shape_and_data, const ipmz_options* opt,
                      ipmz_batch_handle* out);
int ipmz_batch_destroy(ipmz_batch_handle h);
/* Upload (again) the problem data from host arrays laid out as at create (same shape and Settings: checked); part of
 * the end-to-end timed region of bench.py.  The copies are asynchronous: page-locked host arrays must stay untouched
 * until the next ipmz_batch_solve* / ipmz_batch_get_* call on the handle has returned. */
int ipmz_batch_upload(ipmz_batch_handle h, const ipmz_problem* data);
/* Every solve restarts the per-problem counters and continues from the handle's current iterates (after an upload: the
 * reference's initial point; after a solve: warm start, 0 iterations when already converged). */
int ipmz_batch_solve(ipmz_batch_handle h, ipmz_result* per_problem /* count entries or NULL */,
                     double* ms_total);
/* Upload + solve pipelined: the persistent batch kernel starts first and takes problems as the copy stream delivers the
 * `chunks` groups of host data (laid out as for ipmz_batch_upload); one launch, upload and solve overlap. */
int ipmz_batch_solve_streamed(ipmz_batch_handle h, const ipmz_problem* data, int chunks,
                              ipmz_result* per_problem /* count entries or NULL */, double* ms_total);
/* g batches of ONE device solved concurrently, one host thread + CUDA stream per handle (latency-bound
 * kernels of one sub-batch overlap throughput-bound kernels of another); ms_total = device time from the
 * earliest start to the latest end. Per-problem results: ipmz_batch_get_x / _get_iterates per handle. */
int ipmz_batch_solve_group(int g, ipmz_batch_handle* handles, double* ms_total);
/* Per-problem outcome (iterations, converged, f, res, mu) of the handle's last solve; count entries. */
int ipmz_batch_results(ipmz_batch_handle h, ipmz_result* per_problem);
int ipmz_batch_get_iterates(ipmz_batch_handle h, double* packed /* count x iterate_len */);
int ipmz_batch_get_x(ipmz_batch_handle h, double* x /* count x n */);

#ifdef __cplusplus
}
#endif
#endif
