/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Shared C declarations for the two CPU checkers of the ipm-zoo numerical
 * interior-point hot path:
 *
 *   1. oracle/_ref/libipmzoo_ref.so  ("reference"): the UNMODIFIED reference sources
 *      compiled where they lie under /root/reference by oracle/Makefile and driven
 *      by oracle/ref_harness.cpp through the reference's own public C++ API
 *      (build_environment + get_newton_system + Optimizer::solve,
 *      LinearSolvers::*).  Entry points are prefixed ref_.
 *   2. oracle/_build/libipmzoo_oracle.so ("port"): oracle/ipm_oracle.c, a plain-C
 *      restatement of the same algorithm.  Entry points are prefixed orc_.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load either library, and only as the checker or as the timed
 * CPU baseline.  Nothing under ipm-zoo_b200/ includes this header.
 *
 * Iterate layout ("packed iterate", length orc_iterate_len(p) = 5n + 6mi + 6me):
 *   x[n] lamA[mi] s[mi] lamg[mi] lamh[mi] g[mi] h[mi]
 *   lamC[me] t[me] lamv[me] lamw[me] v[me] w[me]
 *   lamy[n] lamz[n] y[n] z[n]
 * (names follow /root/reference/include/SymbolicOptimization.h:5-26; groups that the
 * Settings do not create keep their slots and are never touched).
 */
#ifndef IPMZ_ORACLE_H
#define IPMZ_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_BOUNDS_NONE = 0, ORC_BOUNDS_LOWER = 1, ORC_BOUNDS_UPPER = 2, ORC_BOUNDS_BOTH = 3 };

/* Mirrors NumericalOptimization::Data (EnvironmentBuilder.h:7-17) + the subset of
 * SymbolicOptimization::Settings (SymbolicOptimization.h:58-64) that the reference can
 * solve numerically (inequality_handling = SlackedSlacks always; equality_handling =
 * SlackedSlacks when equalities != 0). All matrices dense row-major. */
typedef struct {
  int n;       /* variables */
  int m_ineq;  /* rows of A_ineq (0 when ineq_bounds == NONE) */
  int m_eq;    /* rows of A_eq   (0 when equalities == 0) */
  const double* Q;   /* n x n */
  const double* c;   /* n */
  const double* A;   /* m_ineq x n */
  const double* l_A; /* m_ineq */
  const double* u_A; /* m_ineq */
  const double* C;   /* m_eq x n */
  const double* d;   /* m_eq */
  const double* l_x; /* n (always required: EnvironmentBuilder.cpp:37-41) */
  const double* u_x; /* n */
  int ineq_bounds;   /* Settings::inequalities */
  int var_bounds;    /* Settings::variable_bounds */
  int equalities;    /* Settings::equalities (SlackedSlacks handling) */
} orc_problem;

/* Per-solve trace. Arrays are caller-allocated; any pointer may be NULL to skip it.
 * cap_iters bounds how many iterations are recorded (and, for ref_solve with
 * stop_after_cap != 0, how many the reference is allowed to run). */
typedef struct {
  int cap_iters;
  int stop_after_cap;
  /* outputs */
  int iterations;   /* Newton steps taken (= number of "iter:" lines - 1 when converged) */
  int converged;    /* 1 iff the loop left through the tolerance test */
  int n_logged;     /* entries valid in f/res/mu (<= cap_iters + 1) */
  double* f;        /* [cap_iters+1] objective at the start of iteration k  (Optimizer.cpp:128) */
  double* res;      /* [cap_iters+1] residual norm                         (Optimizer.cpp:129) */
  double* mu;       /* [cap_iters+1] mean complementarity ("gap")          (Optimizer.cpp:130) */
  double* rhs_aff;  /* [cap_iters x N] augmented RHS, predictor  ("b0:" line, Optimizer.cpp:356) */
  double* step_aff; /* [cap_iters x N] solved augmented step, predictor ("b:", Optimizer.cpp:359) */
  double* rhs_cor;  /* [cap_iters x N] augmented RHS, corrector */
  double* step_cor; /* [cap_iters x N] solved augmented step, corrector */
  double* alpha_aff;/* [cap_iters] (port only; the reference never prints it) */
  double* sigma;    /* [cap_iters] (port only) */
  double* alpha;    /* [cap_iters] (port only) */
  double* iterate;  /* in/out packed iterate; on input used iff use_initial_iterate */
  int use_initial_iterate;
  double seconds;   /* wall time of ctor+solve (ref) or of the loop (port) */
} orc_trace;

int orc_iterate_len(const orc_problem* p);
int orc_aug_dim(const orc_problem* p); /* N = n + m_ineq + m_eq */

/* ---- port (ipm_oracle.c) ---- */
int orc_initial_iterate(const orc_problem* p, double* iterate);
int orc_solve(const orc_problem* p, orc_trace* tr);
/* Assemble the augmented KKT (N x N row-major, full symmetric) and, if rhs != NULL, the
 * affine (mu=0) augmented RHS at the given packed iterate. */
int orc_assemble_kkt(const orc_problem* p, const double* iterate, double* K, double* rhs);
/* LinearSolvers.cpp:14-42 and :44-74 restated on flat row-major arrays. */
int orc_ldlt(int n, const double* A, double* L, double* D);
int orc_solve_ldlt(int n, const double* L, const double* D, double* b);
/* LinearSolvers.cpp:76-207 and :209-318 restated (Bunch-Kaufman, lower, LAPACK ipiv). */
int orc_bk_factor(int n, const double* A, double* LD, int* ipiv);
int orc_bk_solve(int n, const double* LD, const int* ipiv, double* b);

/* ---- reference harness (ref_harness.cpp, linked against the unmodified reference) ---- */
int ref_solve(const orc_problem* p, orc_trace* tr, int quiet /* 1: cout failbit, no trace */);
int ref_ldlt(int n, const double* A, double* L, double* D);
int ref_solve_ldlt(int n, const double* L, const double* D, double* b);
int ref_bk_factor(int n, const double* A, double* LD, int* ipiv);
int ref_bk_solve(int n, const double* LD, const int* ipiv, double* b);
const char* ref_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
