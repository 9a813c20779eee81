// ipm-zoo_b200/csrc/full_system.cu -- the FULL (un-reduced) Newton system of the interior-point step.
//
// Reference: SymbolicOptimization::get_newton_system (SymbolicOptimization.cpp:417-433) derives the
// un-reduced system over every variable (x, lambda, s, the bound slacks and their multipliers:
// 5n + 6m unknowns with both-sided bounds); get_augmented_system (:451-463) then removes block rows
// from the bottom right by Gaussian elimination (:529-567).  The reference never evaluates the full
// system numerically; north_star asks for it as one of the three reductions.
//
// Here the system is symmetrised (each complementarity row  Lam dv + V dlam = -r  is divided by its
// slack v) and ordered as FullLayout (ipmz_device.cuh): complementarity slacks, their multipliers,
// the row slack s, then x and lambda.  In that order the unpivoted LDL^T of the factorization kernels
// performs the reference's block elimination numerically -- the pivots are Lam/V > 0, then -V/Lam < 0,
// then W = G^-1 Lam_g + H^-1 Lam_h > 0, and the trailing Schur complement is the augmented
// quasi-definite matrix -- so no pivoting is needed and the solved [dx; dlam] agrees with the
// reference's augmented Newton step to rounding (tests/test_gpu_parity.py).  All Delta's come out of
// one solve; there is no back-substitution pass.
//
//   row(dy)     : (lam_y/y) dy + dlam_y                     = -r_y / y
//   row(dlam_y) : dy - dx                                   = -r_lam_y
//   row(dz)     : (lam_z/z) dz + dlam_z                     = -r_z / z
//   row(dlam_z) : dz + dx                                   = -r_lam_z
//   row(dsl)    : (lam_l/sl) dsl + dlam_l                   = -r_sl / sl
//   row(dlam_l) : dsl - ds                                  = -r_lam_l
//   row(dsu)    : (lam_u/su) dsu + dlam_u                   = -r_su / su
//   row(dlam_u) : dsu + ds                                  = -r_lam_u
//   row(ds)     : -dlam_l + dlam_u - dlam                   = -r_s
//   row(dx)     : -dlam_y + dlam_z + Q dx + M^T dlam        = -r_x
//   row(dlam)   : -ds + M dx                                = -r_lam
// (residual definitions r_*: SymbolicOptimization.cpp:480-492, evaluated by k_residuals_rhs.)
#include "ipmz_device.cuh"
#include "ipmz_kernels.h"

namespace ipmz {

namespace {

constexpr int FS_ROWS = 8;
constexpr int FS_TPB = 256;

// One warp per row of K: a coalesced pass writes the whole row (zeros, or the dense Q / M / M^T
// blocks), then lane 0 drops the few +-1 / diagonal entries.  Bytes: N^2 written, n^2 + 2mn read.
__global__ void __launch_bounds__(32 * FS_ROWS) k_assemble_full(View v) {
  const int p = problem_of(v);
  const Shape& s = v.s;
  const FullLayout& f = v.fl;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * FS_ROWS + (threadIdx.x >> 5);
  if (r >= v.N) return;
  const double* V = v.V + (size_t)p * v.sp;
  double* Krow = v.K + (size_t)p * v.sK + (size_t)r * v.ldk;

  if (r >= f.ox && r < f.olam) {  // x row: Q | M^T
    const int i = r - f.ox;
    const double* q = v.Q + (size_t)p * v.sQ + (size_t)i * v.ldq;
    const double* mt = v.MT + (size_t)p * v.sMT + (size_t)i * v.ldmt;
    for (int c = lane; c < v.ldk; c += 32) {
      double val = 0.0;
      if (c >= f.ox && c < f.olam) val = q[c - f.ox];
      else if (c >= f.olam && c < f.N) val = mt[c - f.olam];
      Krow[c] = val;
    }
  } else if (r >= f.olam) {  // lambda row: M
    const int i = r - f.olam;
    const double* mr = v.M + (size_t)p * v.sM + (size_t)i * v.ldm;
    for (int c = lane; c < v.ldk; c += 32) Krow[c] = (c >= f.ox && c < f.olam) ? mr[c - f.ox] : 0.0;
  } else {
    for (int c = lane; c < v.ldk; c += 32) Krow[c] = 0.0;
  }
  __syncwarp();
  if (lane != 0) return;

  auto row_has = [&](int i, int lower) { return (i < s.mi) ? (lower ? s.ilo : s.iup) : 1; };
  if (s.ylo && r >= f.oy && r < f.oy + s.n) {
    const int i = r - f.oy;
    Krow[r] = inv_guard(nslot(V, s, YS)[i]) * nslot(V, s, LAMY)[i];
    Krow[f.oly + i] = 1.0;
  } else if (s.zup && r >= f.oz && r < f.oz + s.n) {
    const int i = r - f.oz;
    Krow[r] = inv_guard(nslot(V, s, ZS)[i]) * nslot(V, s, LAMZ)[i];
    Krow[f.olz + i] = 1.0;
  } else if (f.hasl && r >= f.osl && r < f.osl + s.m) {
    const int i = r - f.osl;
    if (row_has(i, 1)) {
      Krow[r] = inv_guard(mslot(V, s, SL)[i]) * mslot(V, s, LAML)[i];
      Krow[f.oll + i] = 1.0;
    } else {
      Krow[r] = 1.0;
    }
  } else if (f.hasu && r >= f.osu && r < f.osu + s.m) {
    const int i = r - f.osu;
    if (row_has(i, 0)) {
      Krow[r] = inv_guard(mslot(V, s, SU)[i]) * mslot(V, s, LAMU)[i];
      Krow[f.olu + i] = 1.0;
    } else {
      Krow[r] = 1.0;
    }
  } else if (s.ylo && r >= f.oly && r < f.oly + s.n) {
    const int i = r - f.oly;
    Krow[f.oy + i] = 1.0;
    Krow[f.ox + i] = -1.0;
  } else if (s.zup && r >= f.olz && r < f.olz + s.n) {
    const int i = r - f.olz;
    Krow[f.oz + i] = 1.0;
    Krow[f.ox + i] = 1.0;
  } else if (f.hasl && r >= f.oll && r < f.oll + s.m) {
    const int i = r - f.oll;
    if (row_has(i, 1)) {
      Krow[f.osl + i] = 1.0;
      Krow[f.os + i] = -1.0;
    } else {
      Krow[r] = 1.0;
    }
  } else if (f.hasu && r >= f.olu && r < f.olu + s.m) {
    const int i = r - f.olu;
    if (row_has(i, 0)) {
      Krow[f.osu + i] = 1.0;
      Krow[f.os + i] = 1.0;
    } else {
      Krow[r] = 1.0;
    }
  } else if (r >= f.os && r < f.ox) {
    const int i = r - f.os;
    if (row_has(i, 1)) Krow[f.oll + i] = -1.0;
    if (row_has(i, 0)) Krow[f.olu + i] = 1.0;
    Krow[f.olam + i] = -1.0;
  } else if (r >= f.ox && r < f.olam) {
    const int i = r - f.ox;
    if (s.ylo) Krow[f.oly + i] = -1.0;
    if (s.zup) Krow[f.olz + i] = 1.0;
  } else if (r >= f.olam) {
    Krow[f.os + (r - f.olam)] = -1.0;
  }
}

// Right-hand side of the symmetrised full system from the shorthand residuals R (predictor: mu = 0;
// corrector: complementarity rows hold v lam - sigma mu + dv_aff dlam_aff, Optimizer.cpp:188-209).
__global__ void __launch_bounds__(FS_TPB) k_full_rhs(View v) {
  const int p = problem_of(v);
  const Shape& s = v.s;
  const FullLayout& f = v.fl;
  const double* V = v.V + (size_t)p * v.sp;
  const double* R = v.R + (size_t)p * v.sp;
  double* sol = v.sol + (size_t)p * v.ssol;
  const int i = blockIdx.x * FS_TPB + threadIdx.x;
  if (i < s.n) {
    sol[f.ox + i] = -nslot(R, s, X)[i];
    if (s.ylo) {
      sol[f.oy + i] = -(inv_guard(nslot(V, s, YS)[i]) * nslot(R, s, YS)[i]);
      sol[f.oly + i] = -nslot(R, s, LAMY)[i];
    }
    if (s.zup) {
      sol[f.oz + i] = -(inv_guard(nslot(V, s, ZS)[i]) * nslot(R, s, ZS)[i]);
      sol[f.olz + i] = -nslot(R, s, LAMZ)[i];
    }
  }
  if (i < s.m) {
    const int lo = (i < s.mi) ? s.ilo : 1, up = (i < s.mi) ? s.iup : 1;
    sol[f.olam + i] = -mslot(R, s, LAM)[i];
    sol[f.os + i] = -mslot(R, s, SV)[i];
    if (f.hasl) {
      sol[f.osl + i] = lo ? -(inv_guard(mslot(V, s, SL)[i]) * mslot(R, s, SL)[i]) : 0.0;
      sol[f.oll + i] = lo ? -mslot(R, s, LAML)[i] : 0.0;
    }
    if (f.hasu) {
      sol[f.osu + i] = up ? -(inv_guard(mslot(V, s, SU)[i]) * mslot(R, s, SU)[i]) : 0.0;
      sol[f.olu + i] = up ? -mslot(R, s, LAMU)[i] : 0.0;
    }
  }
}

// Scatter the solved vector into the direction pack and reduce the single primal/dual step length
// over the non-negative variable kinds (Optimizer.cpp:270-342), as k_backsub_step does for the
// reduced systems.  MODE 0: affine direction DA / alpha_aff, MODE 1: final direction D / alpha.
template <int MODE>
__global__ void __launch_bounds__(FS_TPB) k_full_unpack(View v) {
  const int p = problem_of(v);
  const Shape& s = v.s;
  const FullLayout& f = v.fl;
  const double* V = v.V + (size_t)p * v.sp;
  double* D = (MODE == 0 ? v.DA : v.D) + (size_t)p * v.sp;
  const double* sol = v.sol + (size_t)p * v.ssol;
  const int i = blockIdx.x * FS_TPB + threadIdx.x;
  double a = 1.0;
  auto ratio = [&](double val, double d) {
    if (d < 0.0) a = fmin(a, -val / d);
  };
  if (i < s.n) {
    const double dx = sol[f.ox + i];
    nslot(D, s, X)[i] = dx;
    if (s.ylo) {
      const double dy = sol[f.oy + i], dly = sol[f.oly + i];
      nslot(D, s, YS)[i] = dy; nslot(D, s, LAMY)[i] = dly;
      ratio(nslot(V, s, YS)[i], dy); ratio(nslot(V, s, LAMY)[i], dly);
    }
    if (s.zup) {
      const double dz = sol[f.oz + i], dlz = sol[f.olz + i];
      nslot(D, s, ZS)[i] = dz; nslot(D, s, LAMZ)[i] = dlz;
      ratio(nslot(V, s, ZS)[i], dz); ratio(nslot(V, s, LAMZ)[i], dlz);
    }
    if (s.clamp_x) {
      const double x = nslot(V, s, X)[i];
      if (dx < 0.0) a = fmin(a, (v.lx[(size_t)p * s.ns + i] - x) / dx);
      if (dx > 0.0) a = fmin(a, (v.ux[(size_t)p * s.ns + i] - x) / dx);
    }
  }
  if (i < s.m) {
    const int lo = (i < s.mi) ? s.ilo : 1, up = (i < s.mi) ? s.iup : 1;
    mslot(D, s, LAM)[i] = sol[f.olam + i];
    mslot(D, s, SV)[i] = sol[f.os + i];
    if (lo) {
      const double dsl = sol[f.osl + i], dll = sol[f.oll + i];
      mslot(D, s, SL)[i] = dsl; mslot(D, s, LAML)[i] = dll;
      ratio(mslot(V, s, SL)[i], dsl); ratio(mslot(V, s, LAML)[i], dll);
    }
    if (up) {
      const double dsu = sol[f.osu + i], dlu = sol[f.olu + i];
      mslot(D, s, SU)[i] = dsu; mslot(D, s, LAMU)[i] = dlu;
      ratio(mslot(V, s, SU)[i], dsu); ratio(mslot(V, s, LAMU)[i], dlu);
    }
  }
  // block minimum by warp shuffles, then one atomic-free finish by the last block of the problem
  __shared__ double wmin[FS_TPB / 32];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
  if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = wmin[0];
    for (int k = 1; k < FS_TPB / 32; ++k) b = fmin(b, wmin[k]);
    double* part = v.partials + ((size_t)blockIdx.y * v.maxblk + blockIdx.x) * 8;
    part[0] = b;
    __threadfence();
    const int done = atomicAdd(&v.counters[blockIdx.y], 1);
    last = (done == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile double* part = v.partials + (size_t)blockIdx.y * v.maxblk * 8;
    double b = 1.0;
    for (unsigned k = 0; k < gridDim.x; ++k) b = fmin(b, part[(size_t)k * 8]);
    Scal& sc = v.sc[p];
    if (MODE == 0) sc.alpha_aff = b; else sc.alpha = b;
    v.counters[blockIdx.y] = 0;
  }
}

}  // namespace

void launch_assemble_full(cudaStream_t st, const View& v, int nslots) {
  dim3 grid((v.N + FS_ROWS - 1) / FS_ROWS, nslots);
  k_assemble_full<<<grid, 32 * FS_ROWS, 0, st>>>(v); count_launch();
}

static dim3 full_vec_grid(const View& v, int nslots) {
  const int len = v.s.ns > v.s.ms ? v.s.ns : v.s.ms;
  return dim3((len + FS_TPB - 1) / FS_TPB, nslots);
}

void launch_full_rhs(cudaStream_t st, const View& v, int nslots) {
  k_full_rhs<<<full_vec_grid(v, nslots), FS_TPB, 0, st>>>(v); count_launch();
}

void launch_full_unpack(cudaStream_t st, const View& v, int nslots, int mode) {
  if (mode == 0) k_full_unpack<0><<<full_vec_grid(v, nslots), FS_TPB, 0, st>>>(v);
  else k_full_unpack<1><<<full_vec_grid(v, nslots), FS_TPB, 0, st>>>(v);
  count_launch();
}

}  // namespace ipmz
