// ipm-zoo_b200/csrc/vector_bodies.cuh -- per-element bodies of the fused vector kernels of the Mehrotra
// predictor-corrector loop (reference: Optimizer.cpp:128-130, :147-157, :165-217, :240-342, :361-378).
// Each body handles element i of problem p and accumulates its reduction terms into `acc`; the grid-per-phase
// kernels of vector_kernels.cu and the persistent one-CTA-per-problem kernel of batch_fused.cu call the SAME
// bodies, so both paths evaluate the reference's formulas with identical roundings per element.
#pragma once
#include "ipmz_device.cuh"

namespace ipmz {

// The reference's initial point (EnvironmentBuilder.cpp:34-73): x, s at the mid-points of
// their bounds, every other variable 1.
__device__ __forceinline__ void initial_point_body(const View& v, int p, int i) {
  const Shape& s = v.s;
  double* V = v.V + (size_t)p * v.sp;
  if (i < s.n) {
    nslot(V, s, X)[i] = 0.5 * (v.lx[(size_t)p * s.ns + i] + v.ux[(size_t)p * s.ns + i]);
    nslot(V, s, LAMY)[i] = 1.0; nslot(V, s, LAMZ)[i] = 1.0;
    nslot(V, s, YS)[i] = 1.0; nslot(V, s, ZS)[i] = 1.0;
  }
  if (i < s.m && (s.hard_eq || s.reg_eq || s.pen_eq) && i >= s.mi) {  // EqualityHandling::None / Penalty*: only the multiplier exists
    mslot(V, s, LAM)[i] = 1.0;
    mslot(V, s, SV)[i] = s.reg_eq ? 1.0 : 0.0;  // Regularization: p = 1 (EnvironmentBuilder.cpp:56)
    mslot(V, s, LAML)[i] = 0.0; mslot(V, s, LAMU)[i] = 0.0;
    mslot(V, s, SL)[i] = 0.0; mslot(V, s, SU)[i] = 0.0;
  } else if (i < s.m) {
    const double mid = 0.5 * (v.lo[(size_t)p * s.ms + i] + v.up[(size_t)p * s.ms + i]);
    mslot(V, s, LAM)[i] = 1.0;
    mslot(V, s, SV)[i] = (i < s.mi) ? mid : 1.0;  // t (equality slack) starts at 1
    mslot(V, s, LAML)[i] = 1.0; mslot(V, s, LAMU)[i] = 1.0;
    mslot(V, s, SL)[i] = 1.0; mslot(V, s, SU)[i] = 1.0;
  }
  if (i == 0) {
    Scal& sc = v.sc[p];
    sc.iters = 0; sc.done = 0; sc.mu_c = 0.0; sc.alpha = 0.0; sc.alpha_aff = 0.0; sc.sigma = 0.0;
  }
}

// MODE 0 (start of an iteration, mu = 0): all shorthand residuals r_* (definitions:
// SymbolicOptimization.cpp:480-492), the (2,2) diagonal W / W^-1, the terms of the objective, residual norm and
// mean complementarity (Optimizer.cpp:128-130) and the predictor's augmented right-hand side.
// MODE 1 (corrector): complementarity rows become  v*lam - sigma*mu + dv_aff*dlam_aff
// (Optimizer.cpp:188-209) and the augmented right-hand side is rebuilt.
// acc: 0.5 x'Qx, c'x, sum r^2, sum |comp| (MODE 0 only; the caller zero-initialises).
template <int MODE>
__device__ __forceinline__ void residuals_rhs_body(const View& v, int p, int i, double (&acc)[4]) {
  const Shape& s = v.s;
  double* V = v.V + (size_t)p * v.sp;
  double* R = v.R + (size_t)p * v.sp;
  const double* DA = v.DA + (size_t)p * v.sp;
  double* rhs = v.rhs + (size_t)p * (s.ns + s.ms);
  const double mu = MODE == 0 ? 0.0 : v.sc[p].mu_c;
  if (i < s.n) {
    const double x = nslot(V, s, X)[i];
    double rx;
    if (MODE == 0) {
      const double qx = v.Qx[(size_t)p * s.ns + i];
      const double ci = v.c[(size_t)p * s.ns + i];
      rx = ci;
      if (s.zup) rx += nslot(V, s, LAMZ)[i];
      rx += qx;
      if (s.m > 0) rx += v.MTl[(size_t)p * s.ns + i];
      if (s.ylo) rx += -nslot(V, s, LAMY)[i];
      nslot(R, s, X)[i] = rx;
      acc[0] += (0.5 * x) * qx;
      acc[1] += ci * x;
      acc[2] += rx * rx;
    } else {
      rx = nslot(R, s, X)[i];
    }
    double tz = 0.0, ty = 0.0;
    if (s.ylo) {
      const double y = nslot(V, s, YS)[i], ly = nslot(V, s, LAMY)[i];
      double rly, ry;
      if (MODE == 0) {
        rly = (v.lx[(size_t)p * s.ns + i] + y) + -x;
        ry = y * ly + -(mu * 1.0);
        nslot(R, s, LAMY)[i] = rly;
        acc[2] += rly * rly + ry * ry;
        acc[3] += fabs(ry);
      } else {
        rly = nslot(R, s, LAMY)[i];
        ry = (y * ly + -(mu * 1.0)) + nslot(DA, s, YS)[i] * nslot(DA, s, LAMY)[i];
      }
      nslot(R, s, YS)[i] = ry;
      ty = inv_guard(y) * (ry + -(ly * rly));
    }
    if (s.zup) {
      const double z = nslot(V, s, ZS)[i], lz = nslot(V, s, LAMZ)[i];
      double rlz, rz;
      if (MODE == 0) {
        rlz = (x + z) + -v.ux[(size_t)p * s.ns + i];
        rz = z * lz + -(mu * 1.0);
        nslot(R, s, LAMZ)[i] = rlz;
        acc[2] += rlz * rlz + rz * rz;
        acc[3] += fabs(rz);
      } else {
        rlz = nslot(R, s, LAMZ)[i];
        rz = (z * lz + -(mu * 1.0)) + nslot(DA, s, ZS)[i] * nslot(DA, s, LAMZ)[i];
      }
      nslot(R, s, ZS)[i] = rz;
      tz = inv_guard(z) * (rz + -(lz * rlz));
    }
    double b;
    if (s.zup && s.ylo) b = (tz + -rx) + -ty;
    else if (s.zup) b = tz + -rx;
    else if (s.ylo) b = -(rx + ty);
    else b = -rx;
    rhs[i] = b;
  }

  if (i < s.m && s.hard_eq && i >= s.mi) {
    // EqualityHandling::None (SymbolicOptimization.cpp:137-140): r_lambda = C x - d, Newton row C dx = -r_lambda
    double rlam;
    if (MODE == 0) {
      rlam = v.Mx[(size_t)p * s.ms + i] + -v.lo[(size_t)p * s.ms + i];
      mslot(R, s, LAM)[i] = rlam;
      acc[2] += rlam * rlam;
      v.winv[(size_t)p * s.ms + i] = 0.0;  // the zero diagonal block of the indefinite KKT matrix
      v.W[(size_t)p * s.ms + i] = 0.0;
    } else {
      rlam = mslot(R, s, LAM)[i];
    }
    rhs[s.ns + i] = -rlam;
  } else if (i < s.m && s.reg_eq && i >= s.mi) {
    // EqualityHandling::Regularization (SymbolicOptimization.cpp:184-192): r_lambda = C x - d + delta p,
    // r_p = p + delta lambda; eliminating dp = -r_p - delta dlambda leaves C dx - delta^2 dlambda = -r_lambda + delta r_p
    double rlam, rp;
    if (MODE == 0) {
      const double pv = mslot(V, s, SV)[i], lam = mslot(V, s, LAM)[i];
      rlam = (v.Mx[(size_t)p * s.ms + i] + -v.lo[(size_t)p * s.ms + i]) + s.delta_eq * pv;
      rp = pv + s.delta_eq * lam;
      mslot(R, s, LAM)[i] = rlam;
      mslot(R, s, SV)[i] = rp;
      acc[2] += rlam * rlam + rp * rp;
      v.winv[(size_t)p * s.ms + i] = s.delta_eq * s.delta_eq;
      v.W[(size_t)p * s.ms + i] = 1.0 / (s.delta_eq * s.delta_eq);
    } else {
      rlam = mslot(R, s, LAM)[i];
      rp = mslot(R, s, SV)[i];
    }
    rhs[s.ns + i] = s.delta_eq * rp + -rlam;
  } else if (i < s.m && s.pen_eq && i >= s.mi) {
    // EqualityHandling::PenaltyFunction / PenaltyFunctionWithExtraDual (SymbolicOptimization.cpp:173-183; both reach
    // get_newton_system as the rows C dx - mu dlambda = d + mu lambda - C x): r_lambda = C x - d - mu lambda.
    // The reference's loop evaluates the matrix with the mu the previous iteration left behind (sigma mu; 1 before the
    // first iteration, EnvironmentBuilder.cpp:48), the predictor's r_lambda (and `res`) with mu = 0, the corrector's
    // with the new sigma mu (Optimizer.cpp:138-181: r_lambda holds mu but no e-vector, so it is simply re-evaluated).
    double rlam;
    if (MODE == 0) {
      const Scal& sc = v.sc[p];
      const double mu_env = sc.iters == 0 ? 1.0 : sc.mu_c;
      rlam = v.Mx[(size_t)p * s.ms + i] + -v.lo[(size_t)p * s.ms + i];
      mslot(R, s, LAM)[i] = rlam;
      acc[2] += rlam * rlam;
      v.winv[(size_t)p * s.ms + i] = mu_env;
      v.W[(size_t)p * s.ms + i] = inv_guard(mu_env);
    } else {
      // (v.Mx holds M dx by now in the normal reduction: start from the stored C x - d of this iteration)
      rlam = mslot(R, s, LAM)[i] + -(mu * mslot(V, s, LAM)[i]);
      mslot(R, s, LAM)[i] = rlam;
    }
    rhs[s.ns + i] = -rlam;
  } else if (i < s.m) {
    const int lo = (i < s.mi) ? s.ilo : 1, up = (i < s.mi) ? s.iup : 1;
    const double lam = mslot(V, s, LAM)[i], sv = mslot(V, s, SV)[i];
    double rlam, rsv;
    if (MODE == 0) {
      rlam = v.Mx[(size_t)p * s.ms + i] + -sv;
      const double ll = lo ? mslot(V, s, LAML)[i] : 0.0, lu = up ? mslot(V, s, LAMU)[i] : 0.0;
      if (lo && up) rsv = -((lam + ll) + -lu);
      else if (lo) rsv = -(lam + ll);
      else rsv = lu + -lam;
      mslot(R, s, LAM)[i] = rlam;
      mslot(R, s, SV)[i] = rsv;
      acc[2] += rlam * rlam + rsv * rsv;
    } else {
      rlam = mslot(R, s, LAM)[i];
      rsv = mslot(R, s, SV)[i];
    }
    double sl = 0, ll = 0, rll = 0, rsl = 0, su = 0, lu = 0, rlu = 0, rsu = 0;
    if (lo) {
      sl = mslot(V, s, SL)[i]; ll = mslot(V, s, LAML)[i];
      if (MODE == 0) {
        rll = (v.lo[(size_t)p * s.ms + i] + sl) + -sv;
        rsl = sl * ll + -(mu * 1.0);
        mslot(R, s, LAML)[i] = rll;
        acc[2] += rll * rll + rsl * rsl;
        acc[3] += fabs(rsl);
      } else {
        rll = mslot(R, s, LAML)[i];
        rsl = (sl * ll + -(mu * 1.0)) + mslot(DA, s, SL)[i] * mslot(DA, s, LAML)[i];
      }
      mslot(R, s, SL)[i] = rsl;
    }
    if (up) {
      su = mslot(V, s, SU)[i]; lu = mslot(V, s, LAMU)[i];
      if (MODE == 0) {
        rlu = (su + sv) + -v.up[(size_t)p * s.ms + i];
        rsu = su * lu + -(mu * 1.0);
        mslot(R, s, LAMU)[i] = rlu;
        acc[2] += rlu * rlu + rsu * rsu;
        acc[3] += fabs(rsu);
      } else {
        rlu = mslot(R, s, LAMU)[i];
        rsu = (su * lu + -(mu * 1.0)) + mslot(DA, s, SU)[i] * mslot(DA, s, LAMU)[i];
      }
      mslot(R, s, SU)[i] = rsu;
    }
    double wi, w, b;
    if (lo && up) {
      w = inv_guard(sl) * ll + inv_guard(su) * lu;
      wi = inv_guard(w);
      const double th = inv_guard(su) * (rsu + -(lu * rlu));
      const double tg = inv_guard(sl) * (rsl + -(ll * rll));
      b = wi * ((th + -rsv) + -tg) + -rlam;
    } else if (lo) {
      w = inv_guard(sl) * ll;
      wi = inv_guard(ll) * sl;
      b = -((rlam + inv_guard(ll) * (rsl + sl * rsv)) + -rll);
    } else {
      w = inv_guard(su) * lu;
      wi = inv_guard(lu) * su;
      b = (inv_guard(lu) * (rsu + -(su * rsv)) + -rlam) + -rlu;
    }
    if (MODE == 0) {
      v.winv[(size_t)p * s.ms + i] = wi;
      v.W[(size_t)p * s.ms + i] = w;
    }
    rhs[s.ns + i] = b;
  }
}

// Totals of residuals_rhs_body<0> -> objective, residual norm, mean complementarity, stopping test
// (Optimizer.cpp:128-133).
__device__ __forceinline__ void residuals_finish(const View& v, int p, const double* t) {
  const Shape& s = v.s;
  Scal& sc = v.sc[p];
  sc.f = t[0] + t[1];
  sc.res = sqrt(t[2]);
  sc.mu = s.ncomp ? t[3] / (double)s.ncomp : 0.0;
  if (sc.done == 0) {
    if (sc.res < v.tol && sc.mu < v.tol) sc.done = 1;
    else if (sc.iters >= v.max_iter) sc.done = 2;
  }
}

// Vector handed to the triangular solves.  Augmented: sol = [b0; b1].  Normal (primal
// condensed, eliminating dlam = W (M dx - b1)):  sol = b0 + M^T (W b1); stage 0 forms
// tm = W .* b1, stage 1 (after the M^T matvec into tn) forms sol.  `rvec` is the augmented
// right-hand side b0|b1 (the Newton rhs, or the residual during iterative refinement).
__device__ __forceinline__ void prepare_sol_body(const View& v, int p, int i, const double* rvec_all, int stage) {
  const Shape& s = v.s;
  const double* rvec = rvec_all + (size_t)p * (s.ns + s.ms);
  double* sol = v.sol + (size_t)p * v.ssol;
  if (!v.normal) {
    if (i < s.n) sol[i] = rvec[i];
    if (i < s.m) sol[s.n + i] = rvec[s.ns + i];
  } else if (stage == 0) {
    if (i < s.m) v.tm[(size_t)p * s.ms + i] = v.W[(size_t)p * s.ms + i] * rvec[s.ns + i];
  } else {
    if (i < s.n) sol[i] = rvec[i] + (s.m > 0 ? v.tn[(size_t)p * s.ns + i] : 0.0);
  }
}

// Normal reduction: out = (or +=) [dx ; W (M dx - b1)] from the condensed solve in `sol` and
// Mx = M dx.
__device__ __forceinline__ void recover_dual_body(const View& v, int p, int i, const double* rvec_all, int accumulate) {
  const Shape& s = v.s;
  const double* rvec = rvec_all + (size_t)p * (s.ns + s.ms);
  const double* sol = v.sol + (size_t)p * v.ssol;
  double* out = v.out + (size_t)p * (s.ns + s.ms);
  if (i < s.n) out[i] = accumulate ? out[i] + sol[i] : sol[i];
  if (i < s.m) {
    const double dl = v.W[(size_t)p * s.ms + i] * (v.Mx[(size_t)p * s.ms + i] - rvec[s.ns + i]);
    out[s.ns + i] = accumulate ? out[s.ns + i] + dl : dl;
  }
}

// Normal reduction, iterative refinement: residual of the AUGMENTED system for the recovered
// step,  resid = [b0 - (Hx dx + M^T dlam) ; b1 - (M dx - W^-1 dlam)], with Qd = Q dx,
// tn = M^T dlam, Mx = M dx already formed.
__device__ __forceinline__ void aug_residual_body(const View& v, int p, int i) {
  const Shape& s = v.s;
  const double* V = v.V + (size_t)p * v.sp;
  const double* rhs = v.rhs + (size_t)p * (s.ns + s.ms);
  const double* out = v.out + (size_t)p * (s.ns + s.ms);
  double* resid = v.resid + (size_t)p * (s.ns + s.ms);
  if (i < s.n) {
    double hd = 0.0;
    if (s.ylo) hd = hd + inv_guard(nslot(V, s, YS)[i]) * nslot(V, s, LAMY)[i];
    if (s.zup) hd = hd + inv_guard(nslot(V, s, ZS)[i]) * nslot(V, s, LAMZ)[i];
    const double kd = (v.Qd[(size_t)p * s.ns + i] + hd * out[i]) + (s.m > 0 ? v.tn[(size_t)p * s.ns + i] : 0.0);
    resid[i] = rhs[i] - kd;
  }
  if (i < s.m)
    resid[s.ns + i] = rhs[s.ns + i] - (v.Mx[(size_t)p * s.ms + i] - v.winv[(size_t)p * s.ms + i] * out[s.ns + i]);
}

// After the solve: split [dx; dlam], evaluate the eliminated variables' Delta definitions in
// the reference's reverse elimination order (Optimizer.cpp:361-378; formulas from
// SymbolicOptimization.cpp:499-527) and the terms of the single primal/dual step length
// (Optimizer.cpp:270-342).  MODE 0 writes the affine direction DA, MODE 1 the final direction D.
// amin: running minimum of the ratio test (the caller starts it at 1).
template <int MODE>
__device__ __forceinline__ void backsub_step_body(const View& v, int p, int i, double& amin) {
  const Shape& s = v.s;
  const double* V = v.V + (size_t)p * v.sp;
  const double* R = v.R + (size_t)p * v.sp;
  double* D = (MODE == 0 ? v.DA : v.D) + (size_t)p * v.sp;
  const double* sol = v.normal ? v.out + (size_t)p * (s.ns + s.ms) : v.sol + (size_t)p * v.ssol;
  auto ratio = [&](double val, double d) {
    if (d < 0.0) amin = fmin(amin, -val / d);
  };
  const double* Vn = V;
  const double* Rn = R;
  if (i < s.n) {
    const double dx = sol[i];
    nslot(D, s, X)[i] = dx;
    const double x = nslot(Vn, s, X)[i];
    if (s.ylo) {
      const double y = nslot(Vn, s, YS)[i], ly = nslot(Vn, s, LAMY)[i];
      const double ry = nslot(Rn, s, YS)[i], rly = nslot(Rn, s, LAMY)[i];
      const double dly = -((inv_guard(y) * ly) * ((dx + inv_guard(ly) * ry) + -rly));
      const double dy = -(inv_guard(ly) * (ry + y * dly));
      nslot(D, s, LAMY)[i] = dly; nslot(D, s, YS)[i] = dy;
      ratio(y, dy); ratio(ly, dly);
    }
    if (s.zup) {
      const double z = nslot(Vn, s, ZS)[i], lz = nslot(Vn, s, LAMZ)[i];
      const double rz = nslot(Rn, s, ZS)[i], rlz = nslot(Rn, s, LAMZ)[i];
      const double dlz = -((inv_guard(z) * lz) * ((inv_guard(lz) * rz + -rlz) + -dx));
      const double dz = -(inv_guard(lz) * (rz + z * dlz));
      nslot(D, s, LAMZ)[i] = dlz; nslot(D, s, ZS)[i] = dz;
      ratio(z, dz); ratio(lz, dlz);
    }
    if (s.clamp_x) {
      if (dx < 0.0) amin = fmin(amin, (v.lx[(size_t)p * s.ns + i] - x) / dx);
      if (dx > 0.0) amin = fmin(amin, (v.ux[(size_t)p * s.ns + i] - x) / dx);
    }
  }
  if (i < s.m && s.pen_eq && i >= s.mi) {
    mslot(D, s, LAM)[i] = v.normal ? sol[s.ns + i] : sol[s.n + i];  // free multiplier: no ratio test
  } else if (i < s.m && s.hard_eq && i >= s.mi) {
    mslot(D, s, LAM)[i] = sol[s.n + i];  // the multiplier is free: no ratio test
  } else if (i < s.m && s.reg_eq && i >= s.mi) {
    const double dlam = v.normal ? sol[s.ns + i] : sol[s.n + i];
    mslot(D, s, LAM)[i] = dlam;
    mslot(D, s, SV)[i] = -(mslot(Rn, s, SV)[i] + s.delta_eq * dlam);  // dp; both free: no ratio test
  } else if (i < s.m) {
    const int lo = (i < s.mi) ? s.ilo : 1, up = (i < s.mi) ? s.iup : 1;
    const double dlam = v.normal ? sol[s.ns + i] : sol[s.n + i];
    mslot(D, s, LAM)[i] = dlam;
    const double rsv = mslot(Rn, s, SV)[i];
    double sl = 0, ll = 0, rsl = 0, rll = 0, su = 0, lu = 0, rsu = 0, rlu = 0, tg = 0, th = 0;
    if (lo) {
      sl = mslot(Vn, s, SL)[i]; ll = mslot(Vn, s, LAML)[i];
      rsl = mslot(Rn, s, SL)[i]; rll = mslot(Rn, s, LAML)[i];
      tg = inv_guard(sl) * (rsl + -(ll * rll));
    }
    if (up) {
      su = mslot(Vn, s, SU)[i]; lu = mslot(Vn, s, LAMU)[i];
      rsu = mslot(Rn, s, SU)[i]; rlu = mslot(Rn, s, LAMU)[i];
      th = inv_guard(su) * (rsu + -(lu * rlu));
    }
    double dsv;
    if (lo && up) dsv = v.winv[(size_t)p * s.ms + i] * (((dlam + th) + -rsv) + -tg);
    else if (lo) dsv = (inv_guard(ll) * sl) * ((dlam + -rsv) + -tg);
    else dsv = (inv_guard(lu) * su) * ((dlam + th) + -rsv);
    mslot(D, s, SV)[i] = dsv;
    if (lo) {
      const double dll = -((inv_guard(sl) * ll) * ((dsv + inv_guard(ll) * rsl) + -rll));
      const double dsl = -(inv_guard(ll) * (rsl + sl * dll));
      mslot(D, s, LAML)[i] = dll; mslot(D, s, SL)[i] = dsl;
      ratio(sl, dsl); ratio(ll, dll);
    }
    if (up) {
      const double dlu = -((inv_guard(su) * lu) * ((inv_guard(lu) * rsu + -rlu) + -dsv));
      const double dsu = -(inv_guard(lu) * (rsu + su * dlu));
      mslot(D, s, LAMU)[i] = dlu; mslot(D, s, SU)[i] = dsu;
      ratio(su, dsu); ratio(lu, dlu);
    }
  }
}

// One element's |complementarity product| after the full affine step (Optimizer.cpp:165-181).
__device__ __forceinline__ void mu_affine_body(const View& v, int p, int i, double& sum) {
  const Shape& s = v.s;
  const double* V = v.V + (size_t)p * v.sp;
  const double* DA = v.DA + (size_t)p * v.sp;
  const double al = v.sc[p].alpha_aff;
  auto prod = [&](double a, double da, double b, double db) {
    sum += fabs((a + al * da) * (b + al * db) + -(0.0 * 1.0));
  };
  if (i < s.n) {
    if (s.ylo) prod(nslot(V, s, YS)[i], nslot(DA, s, YS)[i], nslot(V, s, LAMY)[i], nslot(DA, s, LAMY)[i]);
    if (s.zup) prod(nslot(V, s, ZS)[i], nslot(DA, s, ZS)[i], nslot(V, s, LAMZ)[i], nslot(DA, s, LAMZ)[i]);
  }
  if (i < s.m && !((s.hard_eq || s.reg_eq || s.pen_eq) && i >= s.mi)) {
    const int lo = (i < s.mi) ? s.ilo : 1, up = (i < s.mi) ? s.iup : 1;
    if (lo) prod(mslot(V, s, SL)[i], mslot(DA, s, SL)[i], mslot(V, s, LAML)[i], mslot(DA, s, LAML)[i]);
    if (up) prod(mslot(V, s, SU)[i], mslot(DA, s, SU)[i], mslot(V, s, LAMU)[i], mslot(DA, s, LAMU)[i]);
  }
}

// sigma = (mu_aff/mu)^3, mu_c = sigma*mu (Optimizer.cpp:178-180)
__device__ __forceinline__ void mu_affine_finish(const View& v, int p, double total) {
  const Shape& s = v.s;
  Scal& sc = v.sc[p];
  sc.mu_aff = s.ncomp ? total / (double)s.ncomp : 0.0;
  sc.sigma = sc.mu > 0.0 ? pow(sc.mu_aff / sc.mu, v.sigma_pow) : 0.0;
  sc.mu_c = sc.mu * sc.sigma;
}

}  // namespace ipmz
