/*
 * oracle/ipm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement ("port") of the numerical interior-point hot path of
 * albfre/ipm-zoo, used ONLY as the checker in tests/, __graft_entry__.smoke() and as the
 * timed CPU baseline of bench.py.  Nothing under ipm-zoo_b200/ links or loads it.
 *
 * PARITY PINNING: the reference holds no test, golden vector or fixture for this path
 * (SURVEY.md section 0.5 / 8c), so this port is pinned against outputs of the reference
 * itself: tests/test_oracle_vs_reference.py drives the UNMODIFIED reference
 * (oracle/_ref/libipmzoo_ref.so, built by oracle/Makefile) and this file on the same
 * inputs and compares every iteration's f / res / gap, both Newton steps and the final
 * iterate; tests/golden/ holds reference-generated vectors (tests/golden/make_golden.py)
 * so the same comparison runs where /root/reference is absent.
 *
 * What is restated (reference file:line):
 *   - the Mehrotra predictor-corrector loop   src/NumericalOptimization/Optimizer.cpp:77-220
 *   - residual norm, mean complementarity     Optimizer.cpp:240-268
 *   - single primal/dual step length          Optimizer.cpp:270-342
 *   - search direction + back-substitution    Optimizer.cpp:344-380
 *   - unpivoted LDL^T and its solve           src/NumericalOptimization/LinearSolvers.cpp:14-74
 *   - Bunch-Kaufman factor / solve            LinearSolvers.cpp:76-318
 *   - initial point                           src/NumericalOptimization/EnvironmentBuilder.cpp:34-73
 *   - evaluator semantics: 1/0 guard, matvec as row dot products, Matrix+Diag on the
 *     diagonal                                src/NumericalOptimization/Evaluation.cpp:23-50,202-271
 *   - the block formulas of the augmented system that the symbolic layer derives
 *     (src/SymbolicOptimization.cpp:451-527); the printed forms are in DESIGN.md.
 *
 * The reference evaluates these through a tree-walking interpreter; here each symbolic
 * formula is written out by hand with the same association order, so results agree with
 * the reference to a few ulps (not merely to the 1e-9 tolerance of the GPU parity tests).
 */
#define _POSIX_C_SOURCE 200809L
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* Evaluation.cpp:267-271 */
static double inv_guard(double v) { return v == 0.0 ? sqrt(DBL_MAX) : 1.0 / v; }

/* std::inner_product order: sequential, starting from 0.0 (Evaluation.cpp:18-21). */
static double dot_seq(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

int orc_aug_dim(const orc_problem* p) { return p->n + p->m_ineq + p->m_eq; }
int orc_iterate_len(const orc_problem* p) { return 5 * p->n + 6 * p->m_ineq + 6 * p->m_eq; }

/* One block of constraint rows (inequalities A with slack s/g/h, or equalities C with
 * slack t/v/w, both handled as "SlackedSlacks": SymbolicOptimization.cpp:86-103,150-162). */
typedef struct {
  int m;
  int lo, up;          /* lower / upper side present */
  const double* M;     /* m x n */
  const double *lb, *ub;
  /* iterate views */
  double *lam, *sv, *laml, *lamu, *sl, *su;
  /* search-direction views */
  double *dlam, *dsv, *dlaml, *dlamu, *dsl, *dsu;
  /* shorthand residuals */
  double *r_lam, *r_sv, *r_laml, *r_lamu, *r_sl, *r_su;
  double *winv;        /* the (G^-1 L_g + H^-1 L_h)^-1 style diagonal, without the sign */
  double *tmp;         /* the bracket shared by b[1] and Delta s */
  int row0;            /* first row of this block inside the augmented system */
} rows_t;

typedef struct {
  int n, N;
  int ylo, zup;        /* variable lower / upper bound slacks present */
  const orc_problem* p;
  double *x, *lamy, *lamz, *y, *z;
  double *dx, *dlamy, *dlamz, *dy, *dz;
  double *r_x, *r_lamy, *r_lamz, *r_y, *r_z;
  double *Qx;
  rows_t g[2];
  int ng;
  double* work;        /* everything below is carved out of one allocation */
} state_t;

static void bind_iterate(state_t* st, double* it, int delta) {
  const orc_problem* p = st->p;
  const int n = p->n, mi = p->m_ineq, me = p->m_eq;
  double* q = it;
  double* xs = q; q += n;
  double* v_ineq[6]; for (int k = 0; k < 6; ++k) { v_ineq[k] = q; q += mi; }
  double* v_eq[6];   for (int k = 0; k < 6; ++k) { v_eq[k] = q; q += me; }
  double* lamy = q; q += n;
  double* lamz = q; q += n;
  double* y = q; q += n;
  double* z = q;
  if (!delta) { st->x = xs; st->lamy = lamy; st->lamz = lamz; st->y = y; st->z = z; }
  else { st->dx = xs; st->dlamy = lamy; st->dlamz = lamz; st->dy = y; st->dz = z; }
  int gi = 0;
  if (mi > 0 && p->ineq_bounds != ORC_BOUNDS_NONE) {
    rows_t* r = &st->g[gi++];
    if (!delta) { r->lam = v_ineq[0]; r->sv = v_ineq[1]; r->laml = v_ineq[2]; r->lamu = v_ineq[3]; r->sl = v_ineq[4]; r->su = v_ineq[5]; }
    else { r->dlam = v_ineq[0]; r->dsv = v_ineq[1]; r->dlaml = v_ineq[2]; r->dlamu = v_ineq[3]; r->dsl = v_ineq[4]; r->dsu = v_ineq[5]; }
  }
  if (me > 0 && p->equalities) {
    rows_t* r = &st->g[gi++];
    if (!delta) { r->lam = v_eq[0]; r->sv = v_eq[1]; r->laml = v_eq[2]; r->lamu = v_eq[3]; r->sl = v_eq[4]; r->su = v_eq[5]; }
    else { r->dlam = v_eq[0]; r->dsv = v_eq[1]; r->dlaml = v_eq[2]; r->dlamu = v_eq[3]; r->dsl = v_eq[4]; r->dsu = v_eq[5]; }
  }
}

static void setup(state_t* st, const orc_problem* p) {
  memset(st, 0, sizeof(*st));
  st->p = p;
  st->n = p->n;
  st->N = orc_aug_dim(p);
  st->ylo = (p->var_bounds == ORC_BOUNDS_LOWER || p->var_bounds == ORC_BOUNDS_BOTH);
  st->zup = (p->var_bounds == ORC_BOUNDS_UPPER || p->var_bounds == ORC_BOUNDS_BOTH);
  int row0 = p->n;
  if (p->m_ineq > 0 && p->ineq_bounds != ORC_BOUNDS_NONE) {
    rows_t* r = &st->g[st->ng++];
    r->m = p->m_ineq; r->M = p->A; r->lb = p->l_A; r->ub = p->u_A;
    r->lo = (p->ineq_bounds == ORC_BOUNDS_LOWER || p->ineq_bounds == ORC_BOUNDS_BOTH);
    r->up = (p->ineq_bounds == ORC_BOUNDS_UPPER || p->ineq_bounds == ORC_BOUNDS_BOTH);
    r->row0 = row0; row0 += r->m;
  }
  if (p->m_eq > 0 && p->equalities) {
    rows_t* r = &st->g[st->ng++];
    r->m = p->m_eq; r->M = p->C; r->lb = p->d; r->ub = p->d; r->lo = 1; r->up = 1;
    r->row0 = row0; row0 += r->m;
  }
  st->N = row0;
  /* scratch: 6 n-vectors + per block 8 m-vectors */
  size_t need = (size_t)6 * p->n;
  for (int k = 0; k < st->ng; ++k) need += (size_t)8 * st->g[k].m;
  st->work = (double*)calloc(need ? need : 1, sizeof(double));
  double* q = st->work;
  st->r_x = q; q += p->n; st->r_lamy = q; q += p->n; st->r_lamz = q; q += p->n;
  st->r_y = q; q += p->n; st->r_z = q; q += p->n; st->Qx = q; q += p->n;
  for (int k = 0; k < st->ng; ++k) {
    rows_t* r = &st->g[k];
    r->r_lam = q; q += r->m; r->r_sv = q; q += r->m; r->r_laml = q; q += r->m;
    r->r_lamu = q; q += r->m; r->r_sl = q; q += r->m; r->r_su = q; q += r->m;
    r->winv = q; q += r->m; r->tmp = q; q += r->m;
  }
}

/* EnvironmentBuilder.cpp:34-73: x and s start at the mid-point of their bounds, every
 * other variable (slacks, duals, and also t for equalities) starts at 1. */
int orc_initial_iterate(const orc_problem* p, double* it) {
  const int len = orc_iterate_len(p);
  for (int i = 0; i < len; ++i) it[i] = 1.0;
  for (int i = 0; i < p->n; ++i) it[i] = 0.5 * (p->l_x[i] + p->u_x[i]);
  double* s = it + p->n + p->m_ineq;
  for (int i = 0; i < p->m_ineq; ++i) s[i] = 0.5 * (p->l_A[i] + p->u_A[i]);
  return 0;
}

/* Shorthand residuals r_* (SymbolicOptimization.cpp:480-492; forms listed in SURVEY 3.2),
 * evaluated with barrier parameter `mu` (0 for the predictor, sigma*mu for the corrector). */
static void residuals(state_t* st, double mu) {
  const orc_problem* p = st->p;
  const int n = st->n;
  for (int i = 0; i < n; ++i) st->Qx[i] = dot_seq(p->Q + (size_t)i * n, st->x, n);
  for (int i = 0; i < n; ++i) {
    double acc = p->c[i];
    if (st->zup) acc += st->lamz[i];
    acc += st->Qx[i];
    for (int k = 0; k < st->ng; ++k) {
      const rows_t* r = &st->g[k];
      double t = 0.0; /* row i of the materialised transpose times lambda */
      for (int j = 0; j < r->m; ++j) t += r->M[(size_t)j * n + i] * r->lam[j];
      acc += t;
    }
    if (st->ylo) acc += -st->lamy[i];
    st->r_x[i] = acc;
    if (st->ylo) {
      st->r_lamy[i] = (p->l_x[i] + st->y[i]) + -st->x[i];
      st->r_y[i] = st->y[i] * st->lamy[i] + -(mu * 1.0);
    }
    if (st->zup) {
      st->r_lamz[i] = (st->x[i] + st->z[i]) + -p->u_x[i];
      st->r_z[i] = st->z[i] * st->lamz[i] + -(mu * 1.0);
    }
  }
  for (int k = 0; k < st->ng; ++k) {
    rows_t* r = &st->g[k];
    for (int j = 0; j < r->m; ++j) {
      r->r_lam[j] = dot_seq(r->M + (size_t)j * n, st->x, n) + -r->sv[j];
      if (r->lo && r->up) r->r_sv[j] = -((r->lam[j] + r->laml[j]) + -r->lamu[j]);
      else if (r->lo) r->r_sv[j] = -(r->lam[j] + r->laml[j]);
      else r->r_sv[j] = r->lamu[j] + -r->lam[j];
      if (r->lo) {
        /* inequalities: l_A + g - s ; equalities: v + d - t (SymbolicOptimization.cpp:93,153) */
        r->r_laml[j] = (r->M == p->A) ? (r->lb[j] + r->sl[j]) + -r->sv[j]
                                      : (r->sl[j] + r->lb[j]) + -r->sv[j];
        r->r_sl[j] = r->sl[j] * r->laml[j] + -(mu * 1.0);
      }
      if (r->up) {
        /* h + s - u_A ; t + w - d */
        r->r_lamu[j] = (r->M == p->A) ? (r->su[j] + r->sv[j]) + -r->ub[j]
                                      : (r->sv[j] + r->su[j]) + -r->ub[j];
        r->r_su[j] = r->su[j] * r->lamu[j] + -(mu * 1.0);
      }
    }
  }
}

/* Optimizer.cpp:128 with the evaluation order of Evaluation.cpp:154-173. */
static double objective(state_t* st) {
  const orc_problem* p = st->p;
  const int n = st->n;
  double quad = 0.0, lin = 0.0;
  for (int i = 0; i < n; ++i) quad += (0.5 * st->x[i]) * dot_seq(p->Q + (size_t)i * n, st->x, n);
  for (int i = 0; i < n; ++i) lin += p->c[i] * st->x[i];
  return quad + lin;
}

/* Optimizer.cpp:240-268: ||full Newton RHS at mu=0||_2 and the mean |complementarity|.
 * Rows are visited in the full-system variable order x, lam_rows, slacks, bound duals,
 * non-negative slacks (SymbolicOptimization.cpp:370-374). Uses the r_* of residuals(st,0). */
static void res_and_mu(state_t* st, double* res, double* mu) {
  const int n = st->n;
  double ss = 0.0, cs = 0.0;
  long cnt = 0;
  ss += dot_seq(st->r_x, st->r_x, n);
  for (int k = 0; k < st->ng; ++k) { const rows_t* r = &st->g[k]; for (int j = 0; j < r->m; ++j) ss += r->r_lam[j] * r->r_lam[j]; }
  for (int k = 0; k < st->ng; ++k) { const rows_t* r = &st->g[k]; for (int j = 0; j < r->m; ++j) ss += r->r_sv[j] * r->r_sv[j]; }
  for (int k = 0; k < st->ng; ++k) {
    const rows_t* r = &st->g[k];
    if (r->lo) for (int j = 0; j < r->m; ++j) ss += r->r_laml[j] * r->r_laml[j];
    if (r->up) for (int j = 0; j < r->m; ++j) ss += r->r_lamu[j] * r->r_lamu[j];
  }
  if (st->ylo) for (int i = 0; i < n; ++i) ss += st->r_lamy[i] * st->r_lamy[i];
  if (st->zup) for (int i = 0; i < n; ++i) ss += st->r_lamz[i] * st->r_lamz[i];
  for (int k = 0; k < st->ng; ++k) {
    const rows_t* r = &st->g[k];
    if (r->lo) for (int j = 0; j < r->m; ++j) { ss += r->r_sl[j] * r->r_sl[j]; cs += fabs(r->r_sl[j]); ++cnt; }
    if (r->up) for (int j = 0; j < r->m; ++j) { ss += r->r_su[j] * r->r_su[j]; cs += fabs(r->r_su[j]); ++cnt; }
  }
  if (st->ylo) for (int i = 0; i < n; ++i) { ss += st->r_y[i] * st->r_y[i]; cs += fabs(st->r_y[i]); ++cnt; }
  if (st->zup) for (int i = 0; i < n; ++i) { ss += st->r_z[i] * st->r_z[i]; cs += fabs(st->r_z[i]); ++cnt; }
  *res = sqrt(ss);
  *mu = cnt ? cs / (double)cnt : 0.0;
}

/* Mean complementarity only, at the current iterate (Optimizer.cpp:177 path). */
static double mu_only(state_t* st) {
  double cs = 0.0;
  long cnt = 0;
  for (int k = 0; k < st->ng; ++k) {
    const rows_t* r = &st->g[k];
    if (r->lo) for (int j = 0; j < r->m; ++j) { cs += fabs(r->sl[j] * r->laml[j] + -(0.0 * 1.0)); ++cnt; }
    if (r->up) for (int j = 0; j < r->m; ++j) { cs += fabs(r->su[j] * r->lamu[j] + -(0.0 * 1.0)); ++cnt; }
  }
  if (st->ylo) for (int i = 0; i < st->n; ++i) { cs += fabs(st->y[i] * st->lamy[i] + -(0.0 * 1.0)); ++cnt; }
  if (st->zup) for (int i = 0; i < st->n; ++i) { cs += fabs(st->z[i] * st->lamz[i] + -(0.0 * 1.0)); ++cnt; }
  return cnt ? cs / (double)cnt : 0.0;
}

/* Augmented KKT [[Q + Y^-1 L_y + Z^-1 L_z, M^T],[M, -W^-1]] as a dense N x N row-major
 * matrix (Optimizer.cpp:387-391, :441-501; Evaluation.cpp:53-77, :126-140, :212-241). */
static void assemble(state_t* st, double* K) {
  const orc_problem* p = st->p;
  const int n = st->n, N = st->N;
  memset(K, 0, sizeof(double) * (size_t)N * N);
  for (int i = 0; i < n; ++i) {
    memcpy(K + (size_t)i * N, p->Q + (size_t)i * n, sizeof(double) * n);
    double dii = K[(size_t)i * N + i];
    if (st->ylo) dii = dii + inv_guard(st->y[i]) * st->lamy[i];
    if (st->zup) dii = dii + inv_guard(st->z[i]) * st->lamz[i];
    K[(size_t)i * N + i] = dii;
  }
  for (int k = 0; k < st->ng; ++k) {
    rows_t* r = &st->g[k];
    for (int j = 0; j < r->m; ++j) {
      const int row = r->row0 + j;
      for (int i = 0; i < n; ++i) {
        const double a = r->M[(size_t)j * n + i];
        K[(size_t)row * N + i] = a;
        K[(size_t)i * N + row] = a;
      }
      double wi;
      if (r->lo && r->up) wi = inv_guard(inv_guard(r->sl[j]) * r->laml[j] + inv_guard(r->su[j]) * r->lamu[j]);
      else if (r->lo) wi = inv_guard(r->laml[j]) * r->sl[j];
      else wi = inv_guard(r->lamu[j]) * r->su[j];
      r->winv[j] = wi;
      K[(size_t)row * N + row] = -wi;
    }
  }
}

/* Augmented right-hand side from the stored r_* (formulas: SURVEY 3.2 / DESIGN.md). */
static void aug_rhs(state_t* st, double* b) {
  const int n = st->n;
  for (int i = 0; i < n; ++i) {
    double tz = 0.0, ty = 0.0;
    if (st->zup) tz = inv_guard(st->z[i]) * (st->r_z[i] + -(st->lamz[i] * st->r_lamz[i]));
    if (st->ylo) ty = inv_guard(st->y[i]) * (st->r_y[i] + -(st->lamy[i] * st->r_lamy[i]));
    if (st->zup && st->ylo) b[i] = (tz + -st->r_x[i]) + -ty;
    else if (st->zup) b[i] = tz + -st->r_x[i];
    else if (st->ylo) b[i] = -(st->r_x[i] + ty);
    else b[i] = -st->r_x[i];
  }
  for (int k = 0; k < st->ng; ++k) {
    rows_t* r = &st->g[k];
    for (int j = 0; j < r->m; ++j) {
      double v;
      if (r->lo && r->up) {
        const double th = inv_guard(r->su[j]) * (r->r_su[j] + -(r->lamu[j] * r->r_lamu[j]));
        const double tg = inv_guard(r->sl[j]) * (r->r_sl[j] + -(r->laml[j] * r->r_laml[j]));
        r->tmp[j] = (th + -r->r_sv[j]) + -tg; /* bracket without Delta lambda */
        v = r->winv[j] * r->tmp[j] + -r->r_lam[j];
      } else if (r->lo) {
        v = -((r->r_lam[j] + inv_guard(r->laml[j]) * (r->r_sl[j] + r->sl[j] * r->r_sv[j])) + -r->r_laml[j]);
      } else {
        v = (inv_guard(r->lamu[j]) * (r->r_su[j] + -(r->su[j] * r->r_sv[j])) + -r->r_lam[j]) + -r->r_lamu[j];
      }
      b[r->row0 + j] = v;
    }
  }
}

/* Optimizer.cpp:361-378: split the solved augmented vector, then evaluate the eliminated
 * variables' Delta definitions in reverse elimination order (Delta s first, then the bound
 * duals, then the non-negative slacks). */
static void back_substitute(state_t* st, const double* sol) {
  const int n = st->n;
  memcpy(st->dx, sol, sizeof(double) * n);
  for (int k = 0; k < st->ng; ++k) {
    rows_t* r = &st->g[k];
    memcpy(r->dlam, sol + r->row0, sizeof(double) * r->m);
    for (int j = 0; j < r->m; ++j) {
      if (r->lo && r->up) {
        const double th = inv_guard(r->su[j]) * (r->r_su[j] + -(r->lamu[j] * r->r_lamu[j]));
        const double tg = inv_guard(r->sl[j]) * (r->r_sl[j] + -(r->laml[j] * r->r_laml[j]));
        r->dsv[j] = r->winv[j] * (((r->dlam[j] + th) + -r->r_sv[j]) + -tg);
      } else if (r->lo) {
        const double tg = inv_guard(r->sl[j]) * (r->r_sl[j] + -(r->laml[j] * r->r_laml[j]));
        r->dsv[j] = (inv_guard(r->laml[j]) * r->sl[j]) * ((r->dlam[j] + -r->r_sv[j]) + -tg);
      } else {
        const double th = inv_guard(r->su[j]) * (r->r_su[j] + -(r->lamu[j] * r->r_lamu[j]));
        r->dsv[j] = (inv_guard(r->lamu[j]) * r->su[j]) * ((r->dlam[j] + th) + -r->r_sv[j]);
      }
    }
    for (int j = 0; j < r->m; ++j) {
      if (r->lo) {
        r->dlaml[j] = -((inv_guard(r->sl[j]) * r->laml[j]) *
                        ((r->dsv[j] + inv_guard(r->laml[j]) * r->r_sl[j]) + -r->r_laml[j]));
        r->dsl[j] = -(inv_guard(r->laml[j]) * (r->r_sl[j] + r->sl[j] * r->dlaml[j]));
      }
      if (r->up) {
        r->dlamu[j] = -((inv_guard(r->su[j]) * r->lamu[j]) *
                        ((inv_guard(r->lamu[j]) * r->r_su[j] + -r->r_lamu[j]) + -r->dsv[j]));
        r->dsu[j] = -(inv_guard(r->lamu[j]) * (r->r_su[j] + r->su[j] * r->dlamu[j]));
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    if (st->ylo) {
      st->dlamy[i] = -((inv_guard(st->y[i]) * st->lamy[i]) *
                       ((st->dx[i] + inv_guard(st->lamy[i]) * st->r_y[i]) + -st->r_lamy[i]));
      st->dy[i] = -(inv_guard(st->lamy[i]) * (st->r_y[i] + st->y[i] * st->dlamy[i]));
    }
    if (st->zup) {
      st->dlamz[i] = -((inv_guard(st->z[i]) * st->lamz[i]) *
                       ((inv_guard(st->lamz[i]) * st->r_z[i] + -st->r_lamz[i]) + -st->dx[i]));
      st->dz[i] = -(inv_guard(st->lamz[i]) * (st->r_z[i] + st->z[i] * st->dlamz[i]));
    }
  }
}

static void ratio(const double* v, const double* d, int n, double* step) {
  for (int i = 0; i < n; ++i)
    if (d[i] < 0.0) { const double a = -v[i] / d[i]; if (a < *step) *step = a; }
}

/* Optimizer.cpp:270-342: one step length for all variables, capped at 1; when the system
 * has no inequality slacks g/h, x is additionally kept inside [l_x, u_x] directly. */
static double max_step(state_t* st) {
  double a = 1.0;
  const int n = st->n;
  int have_gh = 0;
  for (int k = 0; k < st->ng; ++k) {
    rows_t* r = &st->g[k];
    if (r->M == st->p->A) have_gh = 1;
    if (r->lo) { ratio(r->sl, r->dsl, r->m, &a); ratio(r->laml, r->dlaml, r->m, &a); }
    if (r->up) { ratio(r->su, r->dsu, r->m, &a); ratio(r->lamu, r->dlamu, r->m, &a); }
  }
  if (st->ylo) { ratio(st->y, st->dy, n, &a); ratio(st->lamy, st->dlamy, n, &a); }
  if (st->zup) { ratio(st->z, st->dz, n, &a); ratio(st->lamz, st->dlamz, n, &a); }
  if (!have_gh) {
    for (int i = 0; i < n; ++i) {
      const double d = st->dx[i], v = st->x[i];
      if (d < 0.0) { const double t = (st->p->l_x[i] - v) / d; if (t < a) a = t; }
      if (d > 0.0) { const double t = (st->p->u_x[i] - v) / d; if (t < a) a = t; }
    }
  }
  return a;
}

/* ------------------------------------------------------------------------------------ */
/* LinearSolvers.cpp:14-42: row-oriented unpivoted LDL^T, zero pivot replaced by 1e-8.  */
int orc_ldlt(int n, const double* A, double* L, double* D) {
  memset(L, 0, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; ++i) {
    double sd = A[(size_t)i * n + i];
    const double* Li = L + (size_t)i * n;
    for (int j = 0; j < i; ++j) sd -= Li[j] * Li[j] * D[j];
    D[i] = sd == 0.0 ? 1e-8 : sd;
    for (int j = i + 1; j < n; ++j) {
      double s = A[(size_t)j * n + i];
      const double* Lj = L + (size_t)j * n;
      for (int k = 0; k < i; ++k) s -= Lj[k] * Li[k] * D[k];
      L[(size_t)j * n + i] = s / D[i];
    }
    L[(size_t)i * n + i] = 1.0;
  }
  return 0;
}

/* LinearSolvers.cpp:44-74 */
int orc_solve_ldlt(int n, const double* L, const double* D, double* b) {
  for (int i = 0; i < n; ++i) b[i] -= dot_seq(L + (size_t)i * n, b, i);
  for (int i = 0; i < n; ++i) b[i] /= D[i];
  for (int i = n - 1; i >= 0; --i) {
    double s = 0.0;
    for (int j = i + 1; j < n; ++j) s += L[(size_t)j * n + i] * b[j];
    b[i] -= s;
  }
  return 0;
}

/* LinearSolvers.cpp:76-207: unblocked Bunch-Kaufman on the lower triangle (the LAPACK
 * dsytf2 'L' scheme with alpha = (1+sqrt 17)/8); ipiv[k] >= 0 marks a 1x1 pivot swapped
 * with row ipiv[k], a negative pair marks a 2x2 pivot whose second row was swapped with
 * row -ipiv[k]. */
static void bk_absmax(const double* A, int n, int from, int to, int fixed, int down_column,
                      int* arg, double* val) {
  *arg = 0; *val = 0.0;
  for (int i = from; i < to; ++i) {
    const double v = fabs(down_column ? A[(size_t)i * n + fixed] : A[(size_t)fixed * n + i]);
    if (v > *val) { *val = v; *arg = i; }
  }
}

int orc_bk_factor(int n, const double* Ain, double* A, int* ipiv) {
  if (A != Ain) memcpy(A, Ain, sizeof(double) * (size_t)n * n);
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  int info = 0;
  for (int i = 0; i < n; ++i) ipiv[i] = 0;
#define E(i, j) A[(size_t)(i) * n + (j)]
  int k = 0;
  while (k < n) {
    int width = 1, kp = 0, imax; double colmax;
    const double akk = fabs(E(k, k));
    bk_absmax(A, n, k + 1, n, k, 1, &imax, &colmax);
    if (akk == 0.0 && colmax == 0.0) {
      if (info == 0) { info = k; kp = k; }
    } else {
      if (akk >= alpha * colmax) {
        kp = k;
      } else {
        int dummy; double r1, r2;
        bk_absmax(A, n, k, imax, imax, 0, &dummy, &r1);
        bk_absmax(A, n, imax + 1, n, imax, 1, &dummy, &r2);
        const double rowmax = r1 > r2 ? r1 : r2;
        if (akk * rowmax >= alpha * colmax * colmax) kp = k;
        else if (fabs(E(imax, imax)) >= alpha * rowmax) kp = imax;
        else { kp = imax; width = 2; }
      }
      const int kk = k + width - 1;
      if (kp != kk) {
        double t;
        for (int i = kp + 1; i < n; ++i) { t = E(i, kp); E(i, kp) = E(i, kk); E(i, kk) = t; }
        for (int j = kk + 1; j < kp; ++j) { t = E(kp, j); E(kp, j) = E(j, kk); E(j, kk) = t; }
        t = E(kp, kp); E(kp, kp) = E(kk, kk); E(kk, kk) = t;
        if (width == 2) { t = E(kk, k); E(kk, k) = E(kp, k); E(kp, k) = t; }
      }
      if (width == 1) {
        const double rp = 1.0 / E(k, k);
        for (int j = k + 1; j < n; ++j) {
          const double sf = rp * E(j, k);
          for (int i = j; i < n; ++i) E(i, j) -= sf * E(i, k);
          E(j, k) *= rp;
        }
      } else if (k < n - 1) {
        double d21 = E(k + 1, k);
        const double d11 = E(k + 1, k + 1) / d21;
        const double d22 = E(k, k) / d21;
        const double t = 1.0 / (d11 * d22 - 1.0);
        d21 = t / d21;
        for (int j = k + 2; j < n; ++j) {
          const double wk = d21 * (d11 * E(j, k) - E(j, k + 1));
          const double wk1 = d21 * (d22 * E(j, k + 1) - E(j, k));
          for (int i = j; i < n; ++i) E(i, j) -= (E(i, k) * wk + E(i, k + 1) * wk1);
          E(j, k) = wk;
          E(j, k + 1) = wk1;
        }
      }
    }
    if (width == 1) ipiv[k] = kp;
    else { ipiv[k] = -kp; ipiv[k + 1] = -kp; }
    k += width;
  }
#undef E
  return 0;
}

/* LinearSolvers.cpp:209-318 */
int orc_bk_solve(int n, const double* L, const int* ipiv, double* b) {
#define E(i, j) L[(size_t)(i) * n + (j)]
  int k = 0;
  while (k < n) {
    if (ipiv[k] >= 0) {
      const int kp = ipiv[k];
      if (kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      const double mlt = -b[k];
      for (int i = k + 1; i < n; ++i) b[i] += E(i, k) * mlt;
      b[k] /= E(k, k);
      k += 1;
    } else {
      const int kp = -ipiv[k];
      if (kp != k + 1) { const double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
      if (k < n - 1) {
        const double m0 = -b[k];
        for (int i = k + 2; i < n; ++i) b[i] += E(i, k) * m0;
        const double m1 = -b[k + 1];
        for (int i = k + 2; i < n; ++i) b[i] += E(i, k + 1) * m1;
      }
      const double off = E(k + 1, k);
      const double a0 = E(k, k) / off, a1 = E(k + 1, k + 1) / off;
      const double den = a0 * a1 - 1.0;
      const double b0 = b[k] / off, b1 = b[k + 1] / off;
      b[k] = (a1 * b0 - b1) / den;
      b[k + 1] = (a0 * b1 - b0) / den;
      k += 2;
    }
  }
  k = n - 1;
  while (k >= 0) {
    if (ipiv[k] >= 0) {
      if (k < n - 1) { double s = 0.0; for (int i = k + 1; i < n; ++i) s += E(i, k) * b[i]; b[k] -= s; }
      const int kp = ipiv[k];
      if (kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k -= 1;
    } else {
      if (k < n - 1) {
        double s = 0.0; for (int i = k + 1; i < n; ++i) s += E(i, k) * b[i]; b[k] -= s;
        s = 0.0; for (int i = k + 1; i < n; ++i) s += E(i, k - 1) * b[i]; b[k - 1] -= s;
      }
      const int kp = -ipiv[k];
      if (kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k -= 2;
    }
  }
#undef E
  return 0;
}

/* ------------------------------------------------------------------------------------ */
int orc_assemble_kkt(const orc_problem* p, const double* iterate, double* K, double* rhs) {
  state_t st;
  setup(&st, p);
  double* it = (double*)malloc(sizeof(double) * (size_t)orc_iterate_len(p));
  memcpy(it, iterate, sizeof(double) * (size_t)orc_iterate_len(p));
  bind_iterate(&st, it, 0);
  assemble(&st, K);
  if (rhs) { residuals(&st, 0.0); aug_rhs(&st, rhs); }
  free(it);
  free(st.work);
  return 0;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void axpy_all(double* v, double a, const double* d, int len) {
  for (int i = 0; i < len; ++i) v[i] = v[i] + a * d[i];
}

/* Optimizer.cpp:77-220 */
int orc_solve(const orc_problem* p, orc_trace* tr) {
  state_t st;
  setup(&st, p);
  const int N = st.N, len = orc_iterate_len(p);
  double* it = (double*)malloc(sizeof(double) * (size_t)len);
  double* dl = (double*)calloc((size_t)len, sizeof(double));
  double* daff = (double*)calloc((size_t)len, sizeof(double));
  double* trial = (double*)malloc(sizeof(double) * (size_t)len);
  double* K = (double*)malloc(sizeof(double) * (size_t)N * N);
  double* L = (double*)malloc(sizeof(double) * (size_t)N * N);
  double* D = (double*)malloc(sizeof(double) * (size_t)N);
  double* b = (double*)malloc(sizeof(double) * (size_t)N);
  if (tr->use_initial_iterate && tr->iterate) memcpy(it, tr->iterate, sizeof(double) * (size_t)len);
  else orc_initial_iterate(p, it);
  bind_iterate(&st, it, 0);
  bind_iterate(&st, dl, 1);

  const double tol = 1e-8;
  const int max_iter = 100;
  const double t0 = now_s();
  int iter = 0, converged = 0;
  tr->n_logged = 0;
  for (; iter < max_iter; ++iter) {
    const double f = objective(&st);
    residuals(&st, 0.0);
    double res, mu;
    res_and_mu(&st, &res, &mu);
    if (iter <= tr->cap_iters) {
      if (tr->f) tr->f[iter] = f;
      if (tr->res) tr->res[iter] = res;
      if (tr->mu) tr->mu[iter] = mu;
      tr->n_logged = iter + 1;
    }
    if (res < tol && mu < tol) { converged = 1; break; }
    if (tr->stop_after_cap && iter >= tr->cap_iters) break;

    assemble(&st, K);
    orc_ldlt(N, K, L, D);

    /* predictor (mu = 0); r_* are already those of residuals(st, 0) */
    aug_rhs(&st, b);
    if (iter < tr->cap_iters && tr->rhs_aff) memcpy(tr->rhs_aff + (size_t)iter * N, b, sizeof(double) * N);
    orc_solve_ldlt(N, L, D, b);
    if (iter < tr->cap_iters && tr->step_aff) memcpy(tr->step_aff + (size_t)iter * N, b, sizeof(double) * N);
    back_substitute(&st, b);
    const double a_aff = max_step(&st);
    memcpy(daff, dl, sizeof(double) * (size_t)len);

    /* mu after the full affine step (Optimizer.cpp:165-181) */
    memcpy(trial, it, sizeof(double) * (size_t)len);
    axpy_all(trial, a_aff, daff, len);
    bind_iterate(&st, trial, 0);
    const double mu_aff = mu_only(&st);
    bind_iterate(&st, it, 0);
    const double sigma = mu > 0.0 ? pow(mu_aff / mu, 3) : 0.0;
    const double mu_c = mu * sigma;

    /* corrector right-hand side (Optimizer.cpp:183-209) */
    residuals(&st, mu_c);
    {
      state_t da = st; /* views of the affine direction, same layout */
      bind_iterate(&da, daff, 1);
      for (int k = 0; k < st.ng; ++k) {
        rows_t* r = &st.g[k];
        const rows_t* a = &da.g[k];
        for (int j = 0; j < r->m; ++j) {
          if (r->lo) r->r_sl[j] = r->r_sl[j] + (a->dsl[j] * a->dlaml[j] + -(0.0 * 1.0));
          if (r->up) r->r_su[j] = r->r_su[j] + (a->dsu[j] * a->dlamu[j] + -(0.0 * 1.0));
        }
      }
      for (int i = 0; i < st.n; ++i) {
        if (st.ylo) st.r_y[i] = st.r_y[i] + (da.dy[i] * da.dlamy[i] + -(0.0 * 1.0));
        if (st.zup) st.r_z[i] = st.r_z[i] + (da.dz[i] * da.dlamz[i] + -(0.0 * 1.0));
      }
    }
    aug_rhs(&st, b);
    if (iter < tr->cap_iters && tr->rhs_cor) memcpy(tr->rhs_cor + (size_t)iter * N, b, sizeof(double) * N);
    orc_solve_ldlt(N, L, D, b);
    if (iter < tr->cap_iters && tr->step_cor) memcpy(tr->step_cor + (size_t)iter * N, b, sizeof(double) * N);
    back_substitute(&st, b);
    const double a = max_step(&st);
    if (iter < tr->cap_iters) {
      if (tr->alpha_aff) tr->alpha_aff[iter] = a_aff;
      if (tr->sigma) tr->sigma[iter] = sigma;
      if (tr->alpha) tr->alpha[iter] = a;
    }
    axpy_all(it, 0.995 * a, dl, len);
    /* slots of groups that do not exist carry Delta = 0 and stay untouched */
  }
  tr->seconds = now_s() - t0;
  tr->iterations = iter;
  tr->converged = converged;
  if (tr->iterate) memcpy(tr->iterate, it, sizeof(double) * (size_t)len);
  free(it); free(dl); free(daff); free(trial); free(K); free(L); free(D); free(b);
  free(st.work);
  return 0;
}
