// ipm-zoo_b200/csrc/ipmz_device.cuh -- device-side data layout shared by all kernels.
//
// Every kernel is batched: blockIdx.y selects a slot of the `active` list, which maps to a
// problem index; a single QP is a batch of one.  All per-problem arrays are contiguous
// slabs with a fixed stride, so a batch of independent QPs (cfg4) and one large QP (cfg2/3)
// run through the same code.
//
// HBM layout per problem (all FP64, all leading dimensions padded to a multiple of 4 so
// every row starts 32-byte aligned and 16-byte cp.async / double2 accesses are legal):
//   Q   [n  x ldq ]  ldq  = ns = pad4(n)        objective Hessian (row-major, symmetric)
//   M   [m  x ldm ]  ldm  = ns                  constraint rows, inequalities A then equalities C
//   MT  [n  x ldmt]  ldmt = ms = pad4(m)        transposed copy made once on the device, so
//                                               M^T*lambda and the condensed assembly read
//                                               contiguous rows
//   c, lx, ux [ns];  lo, up [ms]               bounds (eq rows: lo = up = d)
//   V, D, DA, R  "packs" of 5 n-vectors + 6 m-vectors: iterate, direction, affine
//                direction, shorthand residuals r_*        (slot order: enum NSlot/MSlot)
//   K   [N x ldk]    reduced matrix, factorized in place: strict lower = L, pivots in Dg
//
// Padding entries are zero-initialised and never written, so vectorised reads past n / m
// only ever multiply finite values by the zero padding of the matrices.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipmz {

enum NSlot { X = 0, LAMY = 1, LAMZ = 2, YS = 3, ZS = 4, N_NSLOTS = 5 };
enum MSlot { LAM = 0, SV = 1, LAML = 2, LAMU = 3, SL = 4, SU = 5, N_MSLOTS = 6 };

struct Shape {
  int n, m, mi;      // variables, constraint rows (mi inequalities first, then equalities)
  int ns, ms;        // padded vector lengths
  int ylo, zup;      // variable lower / upper bound slacks present
  int ilo, iup;      // inequality lower / upper side present (equality rows: both)
  int clamp_x;       // no g/h slacks in the system -> x also clamped to [l_x, u_x] (Optimizer.cpp:296)
  int ncomp;         // number of complementarity entries (denominator of mu)
  int hard_eq;       // EqualityHandling::None: equality rows carry lambda_C only (no t, v, w) -> indefinite KKT
  int reg_eq;        // EqualityHandling::Regularization: rows C x - d + delta p = 0, p in the SV slot, block -delta^2 I
  int pen_eq;        // EqualityHandling::PenaltyFunction*: rows C x - d - mu lambda = 0, multiplier only, block -mu I
  double delta_eq;
};

// FULL reduction (un-reduced Newton system, SymbolicOptimization.cpp:417-433): unknowns in the order
// the reference's block elimination removes them (complementarity slacks first, x and lambda last), so
// that the unpivoted LDL^T of the symmetrised matrix performs exactly that elimination and every
// pivot is nonzero:
//   dy dz dsl dsu | dlam_y dlam_z dlam_l dlam_u | ds | dx | dlam
// Variable-bound groups exist only when the bound side does (n entries each); the row-slack groups
// exist when any row has that side (m entries; rows without it are decoupled identity rows).
struct FullLayout {
  int oy, oz, osl, osu, oly, olz, oll, olu, os, ox, olam, N;
  int hasl, hasu;
};

static inline FullLayout full_layout(const Shape& s) {
  FullLayout f;
  const int me = s.m - s.mi;
  f.hasl = (s.m > 0) && (s.ilo || me > 0);
  f.hasu = (s.m > 0) && (s.iup || me > 0);
  int o = 0;
  f.oy = o;  o += s.ylo ? s.n : 0;
  f.oz = o;  o += s.zup ? s.n : 0;
  f.osl = o; o += f.hasl ? s.m : 0;
  f.osu = o; o += f.hasu ? s.m : 0;
  f.oly = o; o += s.ylo ? s.n : 0;
  f.olz = o; o += s.zup ? s.n : 0;
  f.oll = o; o += f.hasl ? s.m : 0;
  f.olu = o; o += f.hasu ? s.m : 0;
  f.os = o;  o += s.m;
  f.ox = o;  o += s.n;
  f.olam = o; o += s.m;
  f.N = o;
  return f;
}

// Per-problem scalars, device-resident across the whole solve.
struct Scal {
  double f, res, mu;
  double alpha_aff, mu_aff, sigma, mu_c, alpha;
  int iters;
  int done;       // 1: converged, 2: iteration cap
  int pad0, pad1;
};

struct View {
  Shape s;
  const double *Q, *M, *MT, *c, *lx, *ux, *lo, *up;
  int ldq, ldm, ldmt;
  size_t sQ, sM, sMT;      // per-problem strides in doubles
  double *V, *D, *DA, *R;  // packs
  size_t sp;               // pack stride = 5*ns + 6*ms
  double *Qx, *MTl;        // [ns]   Q x,  M^T lambda
  double *Mx;              // [ms]   M x (or M dx in the normal reduction)
  double *winv, *W;        // [ms]   (2,2) block of the augmented system without sign / its inverse
  double *rhs;             // [ns+ms] augmented right-hand side  b0 | b1
  double *sol;             // [Npad]  vector handed to the triangular solves
  double *tm, *tn;         // [ms], [ns] temporaries of the normal reduction
  double *out, *resid;     // [ns+ms] normal reduction: recovered step dx|dlam, augmented residual
  double *Qd;              // [ns]   Q dx during iterative refinement
  size_t ssol;
  double* K;               // reduced matrix / factor
  double* Dg;              // pivots
  size_t sK;
  int ldk, N, normal;      // N = n+m (augmented), n (normal) or FullLayout::N (full)
  int full;                // 1: un-reduced system in the FullLayout order
  int dual;                // 1: dual-Schur normal equations: K = augmented matrix, factorized in two stages (Hx, then S)
  FullLayout fl;
  Scal* sc;
  const int* active;       // slot -> problem index (nullptr: identity)
  double* partials;        // [slots][maxblk][8] reduction scratch
  int* counters;           // [slots]
  int maxblk;
  double tol, ftb, sigma_pow;
  int max_iter;
};

__device__ __forceinline__ int problem_of(const View& v) {
  return v.active ? v.active[blockIdx.y] : (int)blockIdx.y;
}
__device__ __forceinline__ double* nslot(double* pack, const Shape& s, int k) { return pack + (size_t)k * s.ns; }
__device__ __forceinline__ double* mslot(double* pack, const Shape& s, int k) {
  return pack + (size_t)N_NSLOTS * s.ns + (size_t)k * s.ms;
}
__device__ __forceinline__ const double* nslot(const double* pack, const Shape& s, int k) { return pack + (size_t)k * s.ns; }
__device__ __forceinline__ const double* mslot(const double* pack, const Shape& s, int k) {
  return pack + (size_t)N_NSLOTS * s.ns + (size_t)k * s.ms;
}

// Evaluation.cpp:267-271: 1/0 -> sqrt(DBL_MAX)
__device__ __forceinline__ double inv_guard(double v) {
  return v == 0.0 ? 1.3407807929942596e+154 : 1.0 / v;
}

static inline int pad4(int v) { return (v + 3) & ~3; }

}  // namespace ipmz
