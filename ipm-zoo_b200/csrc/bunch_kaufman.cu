// ipm-zoo_b200/csrc/bunch_kaufman.cu -- symmetric indefinite factorization with Bunch-Kaufman pivoting and
// its solve, on the device.
//
// Reference: LinearSolvers::symmetric_indefinite_factorization (LinearSolvers.cpp:76-207) and
// overwriting_solve_bunch_kaufman (:209-318): the unblocked LAPACK dsytf2 'L' scheme, alpha = (1+sqrt 17)/8,
// ipiv[k] >= 0 = 1x1 pivot interchanged with row ipiv[k], a negative pair = 2x2 pivot whose second row was
// interchanged with row -ipiv[k].  The reference never reaches it from Optimizer::solve (solve_indefinite_ is
// ASSERT(false), Optimizer.cpp:75); it is part of the public LinearSolvers header, and here it also carries the
// EqualityHandling::None (indefinite KKT) path of the solver.
//
// The factorization is BIT-EXACT against the reference, pivots included: pivot search and interchanges are
// index work, and every trailing element receives the reference's own sequence of individually rounded
// operations (a -= (rp * a_jk) * a_ik: a multiply, a multiply and a subtract, never fused -- the reference is
// compiled without FMA contraction for x86-64), so the algorithm can be laid out for the GPU freely as long as
// each element sees its updates in pivot order:
//   * the trailing matrix is kept FULLY symmetric in a scratch copy (element (a,b) and (b,a) are computed with
//     the same operands in the same order, so they stay bitwise equal); column k of the lower triangle is then
//     row k of the upper one and the pivot searches (column k, row/column imax) and the pivot column of the
//     rank-1 / rank-2 update are contiguous, coalesced row reads;
//   * rows are dealt cyclically to the CTAs of a team; one warp updates one row per trip, lanes along the row;
//   * team barriers only where another CTA could still read what one is about to write: one per pivot step while
//     the pivot is decided on column k alone and no interchange happens (the dominant-diagonal part of a KKT
//     matrix), three (search -> interchange -> update) otherwise; a single large matrix uses all SMs (cooperative
//     launch, grid barrier), a batch uses one CTA per matrix (__syncthreads);
//   * the finished column of L is mirrored into the dead upper row so that the solves read rows, not columns.
// HBM/L2-bound: the update reads and writes (n-k)^2 doubles per pivot (2 x the lower triangle), i.e.
// (2/3) n^3 * 8 B of traffic against L2 when the matrix fits (n <= ~3900 in 126 MB), HBM beyond.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/ipmz.h"
#include "ipmz_device.cuh"
#include "ipmz_kernels.h"

namespace cg = cooperative_groups;

namespace ipmz {
int ipmz_fail(int code, const std::string& msg);
int ipmz_ensure_device(int device);

namespace {

constexpr int BK_TPB = 512;
constexpr int BK_WARPS = BK_TPB / 32;

struct BkArgs {
  double* S;      // [count][n x ld] in: matrix (lower triangle significant); out: L / D below, L^T mirror above
  int ld;
  size_t sS;
  int n;
  int* ipiv;      // [count][sP]
  size_t sP;
  const int* active;
  double alpha;
  int mirror_input;  // 1: copy the lower triangle over the upper one first (input not known symmetric)
};

template <bool TEAM>
__device__ __forceinline__ void team_sync() {
  if (TEAM) cg::this_grid().sync();
  else __syncthreads();
}

// max |row[j]| over j in [j0, j1) except j == skip, and the FIRST index attaining it (the reference's
// strict `v > max` scan, LinearSolvers.cpp:86-101); NaNs never win, as in the reference.
__device__ void block_absmax(const double* __restrict__ row, int j0, int j1, int skip, double* out_val, int* out_arg,
                             double* sh_val, int* sh_arg) {
  double v = 0.0;
  int arg = 0x7fffffff;
  for (int j = j0 + (int)threadIdx.x; j < j1; j += BK_TPB) {
    if (j == skip) continue;
    const double t = fabs(row[j]);
    if (t > v) { v = t; arg = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ov > v || (ov == v && oa < arg)) { v = ov; arg = oa; }
  }
  __syncthreads();  // previous use of the scratch is over
  if ((threadIdx.x & 31) == 0) { sh_val[threadIdx.x >> 5] = v; sh_arg[threadIdx.x >> 5] = arg; }
  __syncthreads();
  v = sh_val[0]; arg = sh_arg[0];
#pragma unroll
  for (int w = 1; w < BK_WARPS; ++w) {
    const double ov = sh_val[w];
    const int oa = sh_arg[w];
    if (ov > v || (ov == v && oa < arg)) { v = ov; arg = oa; }
  }
  *out_val = v;
  *out_arg = (v == 0.0) ? 0 : arg;
}

// row[c] -= f(c, w0[c], w1[c]) for c in [c0, n) by one warp: scalar head up to a 16-byte boundary, then two
// columns per lane as double2 with four trips unrolled (eight independent 16-byte loads in flight per lane: the
// update is a pure streaming pass and needs the memory-level parallelism), scalar tail.  Every element still gets
// exactly one unfused subtract of the value f returns.
template <bool TWO, class F>
__device__ __forceinline__ void row_update(double* __restrict__ row, const double* __restrict__ w0,
                                           const double* __restrict__ w1, int c0, int n, int lane, F f) {
  if (c0 >= n) return;
  const int cs = (c0 + 1) & ~1;
  if (cs > c0 && lane == 0) row[c0] = __dsub_rn(row[c0], f(c0, w0[c0], TWO ? w1[c0] : 0.0));
  if (cs >= n) return;
  const int npairs = (n - cs) >> 1;
  double2* __restrict__ r2 = reinterpret_cast<double2*>(row + cs);
  const double2* __restrict__ a2 = reinterpret_cast<const double2*>(w0 + cs);
  const double2* __restrict__ b2 = reinterpret_cast<const double2*>(w1 + cs);
#pragma unroll 4
  for (int q = lane; q < npairs; q += 32) {
    double2 v = r2[q];
    const double2 a = a2[q];
    const double2 b = TWO ? b2[q] : make_double2(0.0, 0.0);
    const int c = cs + 2 * q;
    v.x = __dsub_rn(v.x, f(c, a.x, b.x));
    v.y = __dsub_rn(v.y, f(c + 1, a.y, b.y));
    r2[q] = v;
  }
  const int ct = cs + 2 * npairs;
  if (ct < n && lane == 0) row[ct] = __dsub_rn(row[ct], f(ct, w0[ct], TWO ? w1[ct] : 0.0));
}

template <bool TEAM>
__global__ void __launch_bounds__(BK_TPB, 1) k_bk_factor(BkArgs a) {
  __shared__ double sh_val[BK_WARPS];
  __shared__ int sh_arg[BK_WARPS];
  const int p = a.active ? a.active[blockIdx.y] : (int)blockIdx.y;
  double* __restrict__ S = a.S + (size_t)p * a.sS;
  int* __restrict__ ipiv = a.ipiv + (size_t)p * a.sP;
  const int n = a.n, ld = a.ld;
  const int team = gridDim.x, blk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto E = [&](int r, int c) -> double& { return S[(size_t)r * ld + c]; };
  // first row >= r owned by this CTA
  auto first_owned = [&](int r) { const int q = (r - blk + team - 1) / team; return blk + (q < 0 ? 0 : q) * team; };

  if (a.mirror_input) {
    for (int r = first_owned(0) + warp * team; r < n; r += BK_WARPS * team)
      for (int c = lane; c < r; c += 32) E(c, r) = E(r, c);
    team_sync<TEAM>();
  }

  int k = 0;
  int info = 0;         // LinearSolvers.cpp:103,111-117: only the FIRST zero column records kp = k, later ones keep kp = 0
  int pk = -1, pw = 0;  // previous step: first column and width of the L columns still to be mirrored
  while (k < n) {
    // ---- pivot search: every CTA evaluates it on the same (stable) data ----
    int width = 1, kp = k, imax;
    bool nothing = false, scanned_imax = false;
    double colmax;
    const double akk = fabs(E(k, k));
    block_absmax(S + (size_t)k * ld, k + 1, n, -1, &colmax, &imax, sh_val, sh_arg);
    if (akk == 0.0 && colmax == 0.0) {
      nothing = true;  // column k is zero (LinearSolvers.cpp:111-117)
      if (info == 0) { info = k; kp = k; } else kp = 0;
    } else if (!(akk >= __dmul_rn(a.alpha, colmax))) {
      double rowmax;
      int dummy;
      scanned_imax = true;
      block_absmax(S + (size_t)imax * ld, k, n, imax, &rowmax, &dummy, sh_val, sh_arg);
      if (__dmul_rn(akk, rowmax) >= __dmul_rn(__dmul_rn(a.alpha, colmax), colmax)) kp = k;
      else if (fabs(E(imax, imax)) >= __dmul_rn(a.alpha, rowmax)) kp = imax;
      else { kp = imax; width = 2; }
    }
    const int kk = k + width - 1;
    const bool swap = !nothing && kp != kk;
    // Barriers only where another CTA could still read what this one is about to write: the interchange moves rows
    // kk / kp and the update rewrites row imax, both of which other CTAs may still be scanning; a step that decided
    // on row k alone (the common case while the pivots are dominant) needs neither: row k is not written by the
    // update, and the mirror below writes rows < k only.
    if (swap || scanned_imax) team_sync<TEAM>();

    // ---- mirror the previous pivot's L column(s) into their (now dead) upper rows, and interchange ----
    if (pk >= 0) {
      for (int r = first_owned(pk + pw) + tid * team; r < n; r += BK_TPB * team) {
        E(pk, r) = E(r, pk);
        if (pw == 2) E(pk + 1, r) = E(r, pk + 1);
      }
    }
    if (swap) {
      // symmetric interchange of rows/columns kk and kp of the trailing block S[k:, k:]
      for (int r = first_owned(k) + tid * team; r < n; r += BK_TPB * team) {
        if (r == kk || r == kp) continue;
        const double t = E(r, kk); E(r, kk) = E(r, kp); E(r, kp) = t;
      }
      if (blk == kk % team) {
        for (int c = k + tid; c < n; c += BK_TPB) {
          if (c == kp) continue;  // the off-diagonal pair (kk,kp) / (kp,kk) maps onto itself
          if (c == kk) { const double t = E(kk, kk); E(kk, kk) = E(kp, kp); E(kp, kp) = t; continue; }
          const double t = E(kk, c); E(kk, c) = E(kp, c); E(kp, c) = t;
        }
      }
    }
    if (blk == 0 && tid == 0) {
      if (width == 1) ipiv[k] = kp;
      else { ipiv[k] = -kp; ipiv[k + 1] = -kp; }
    }
    if (swap) team_sync<TEAM>();

    // ---- trailing update: element (r, c) of S[k+width:, k+width:] with lo = min(r,c), hi = max(r,c) ----
    if (!nothing) {
      const double* __restrict__ w0 = S + (size_t)k * ld;  // row k = column k of the lower triangle (unscaled)
      if (width == 1) {
        const double rp = __ddiv_rn(1.0, w0[k]);
        for (int r = first_owned(k + 1) + warp * team; r < n; r += BK_WARPS * team) {
          double* __restrict__ row = S + (size_t)r * ld;
          const double wr = w0[r];
          const double sfr = __dmul_rn(rp, wr);  // = L(r,k)
          row_update<false>(row, w0, w0, k + 1, n, lane, [&](int c, double wc, double) {
            // lower (c <= r): A[r][c] -= (rp * A[c][k]) * A[r][k]; upper: the mirror of A[c][r]
            return (c <= r) ? __dmul_rn(__dmul_rn(rp, wc), wr) : __dmul_rn(sfr, wc);
          });
          if (lane == 0) row[k] = sfr;
        }
      } else {
        const double* __restrict__ w1 = S + (size_t)(k + 1) * ld;
        double d21 = w1[k];
        const double d11 = __ddiv_rn(w1[k + 1], d21);
        const double d22 = __ddiv_rn(w0[k], d21);
        const double t = __ddiv_rn(1.0, __dsub_rn(__dmul_rn(d11, d22), 1.0));
        d21 = __ddiv_rn(t, d21);
        for (int r = first_owned(k + 2) + warp * team; r < n; r += BK_WARPS * team) {
          double* __restrict__ row = S + (size_t)r * ld;
          const double e0r = w0[r], e1r = w1[r];
          const double wkr = __dmul_rn(d21, __dsub_rn(__dmul_rn(d11, e0r), e1r));
          const double wk1r = __dmul_rn(d21, __dsub_rn(__dmul_rn(d22, e1r), e0r));
          row_update<true>(row, w0, w1, k + 2, n, lane, [&](int c, double e0c, double e1c) {
            if (c <= r) {  // lo = c: A[r][c] -= A[r][k] * wk(c) + A[r][k+1] * wk1(c)
              const double wkc = __dmul_rn(d21, __dsub_rn(__dmul_rn(d11, e0c), e1c));
              const double wk1c = __dmul_rn(d21, __dsub_rn(__dmul_rn(d22, e1c), e0c));
              return __dadd_rn(__dmul_rn(e0r, wkc), __dmul_rn(e1r, wk1c));
            }
            // mirror of A[c][r]: A[c][k] * wk(r) + A[c][k+1] * wk1(r)
            return __dadd_rn(__dmul_rn(e0c, wkr), __dmul_rn(e1c, wk1r));
          });
          if (lane == 0) { row[k] = wkr; row[k + 1] = wk1r; }
        }
      }
      pk = k; pw = width;
    } else {
      pk = -1;
    }
    team_sync<TEAM>();
    k += width;
  }
  // mirror of the last pivot is empty (no rows below it) unless it was skipped earlier: nothing left to do
}

// x <- solution of (L D L^T with interchanges) x = b, LinearSolvers.cpp:209-318.  One CTA per system, x in shared
// memory; S holds L / D below the diagonal and L^T above it, so each pivot step reads one contiguous row.
// Parallel sums: agrees with the reference's sequential solve to rounding, not bitwise.
__global__ void __launch_bounds__(BK_TPB) k_bk_solve(BkArgs a, double* __restrict__ xall, size_t sx) {
  extern __shared__ double xs[];
  __shared__ double red[BK_WARPS];
  const int p = a.active ? a.active[blockIdx.y] : (int)blockIdx.y;
  const double* __restrict__ S = a.S + (size_t)p * a.sS;
  const int* __restrict__ ipiv = a.ipiv + (size_t)p * a.sP;
  double* __restrict__ x = xall + (size_t)p * sx;
  const int n = a.n, ld = a.ld, tid = threadIdx.x;
  for (int i = tid; i < n; i += BK_TPB) xs[i] = x[i];
  __syncthreads();
  int k = 0;
  while (k < n) {
    const double* __restrict__ r0 = S + (size_t)k * ld;
    if (ipiv[k] >= 0) {
      const int kp = ipiv[k];
      const double bk = xs[kp];  // value of b[k] after the interchange
      __syncthreads();
      if (tid == 0) { xs[kp] = xs[k]; xs[k] = bk / r0[k]; }
      const double mlt = -bk;
      __syncthreads();
      for (int i = k + 1 + tid; i < n; i += BK_TPB) xs[i] += r0[i] * mlt;
      k += 1;
    } else {
      const int kp = -ipiv[k];
      const double* __restrict__ r1 = S + (size_t)(k + 1) * ld;
      const double b0v = xs[k], b1v = xs[kp];
      __syncthreads();
      if (tid == 0) {
        xs[kp] = xs[k + 1];
        const double off = r1[k];
        const double a0 = r0[k] / off, a1 = r1[k + 1] / off;
        const double den = a0 * a1 - 1.0;
        const double c0 = b0v / off, c1 = b1v / off;
        xs[k] = (a1 * c0 - c1) / den;
        xs[k + 1] = (a0 * c1 - c0) / den;
      }
      const double m0 = -b0v, m1 = -b1v;
      __syncthreads();
      for (int i = k + 2 + tid; i < n; i += BK_TPB) xs[i] += r0[i] * m0 + r1[i] * m1;
      k += 2;
    }
    __syncthreads();
  }
  auto block_sum = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < BK_WARPS; ++w) s += red[w];
    return s;
  };
  k = n - 1;
  while (k >= 0) {
    if (ipiv[k] >= 0) {
      const double* __restrict__ r0 = S + (size_t)k * ld;
      double s = 0.0;
      for (int i = k + 1 + tid; i < n; i += BK_TPB) s += r0[i] * xs[i];
      s = block_sum(s);
      const int kp = ipiv[k];
      if (tid == 0) {
        const double v = xs[k] - s;
        xs[k] = xs[kp]; xs[kp] = v;  // kp == k: plain store of v
        if (kp == k) xs[k] = v;
      }
      k -= 1;
    } else {
      const double* __restrict__ r1 = S + (size_t)k * ld;        // column k
      const double* __restrict__ r0 = S + (size_t)(k - 1) * ld;  // column k-1
      double s1 = 0.0, s0 = 0.0;
      for (int i = k + 1 + tid; i < n; i += BK_TPB) { s1 += r1[i] * xs[i]; s0 += r0[i] * xs[i]; }
      s1 = block_sum(s1);
      s0 = block_sum(s0);
      const int kp = -ipiv[k];
      if (tid == 0) {
        xs[k - 1] -= s0;
        const double v = xs[k] - s1;
        xs[k] = xs[kp]; xs[kp] = v;
        if (kp == k) xs[k] = v;
      }
      k -= 2;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += BK_TPB) x[i] = xs[i];
}

// Same solve for n <= E * BK_TPB with the global-memory latency taken off the pivot chain: thread t owns the
// entries i = t + j * BK_TPB of x (j < E); the rows of L^T that the NEXT pivot step needs are prefetched into
// registers while the current step runs (the factor is final, so the row sequence is known from ipiv alone), the
// pivots, the 2x2 off-diagonals and ipiv sit in shared memory, and the interchange is folded into the update
// (the owner of x[kp] substitutes the incoming value), which leaves two block barriers per pivot step.
template <int E>
__global__ void __launch_bounds__(BK_TPB, 1) k_bk_solve_fast(BkArgs a, double* __restrict__ xall, size_t sx) {
  extern __shared__ double sm[];
  const int n = a.n, ld = a.ld, tid = threadIdx.x;
  double* xs = sm;            // [n]
  double* dg = sm + n;        // [n] S(k,k)
  double* od = sm + 2 * n;    // [n] S(k+1,k) (only read for 2x2 pivots)
  int* piv = reinterpret_cast<int*>(sm + 3 * n);
  __shared__ double red[2][BK_WARPS];
  const int p = a.active ? a.active[blockIdx.y] : (int)blockIdx.y;
  const double* __restrict__ S = a.S + (size_t)p * a.sS;
  const int* __restrict__ ipiv = a.ipiv + (size_t)p * a.sP;
  double* __restrict__ x = xall + (size_t)p * sx;
  for (int i = tid; i < n; i += BK_TPB) {
    xs[i] = x[i];
    dg[i] = S[(size_t)i * ld + i];
    od[i] = (i + 1 < n) ? S[(size_t)(i + 1) * ld + i] : 0.0;
    piv[i] = ipiv[i];
  }
  // entries of row r owned by this thread, columns > lim only (the rest of the row is not L^T of this pivot)
  auto fetch = [&](double (&dst)[E], int r, int lim) {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int i = tid + j * BK_TPB;
      dst[j] = (r >= 0 && r < n && i > lim && i < n) ? S[(size_t)r * ld + i] : 0.0;
    }
  };
  double c0[E], c1[E], n0[E], n1[E];
  __syncthreads();

  // ---- forward: L (with interchanges) and D ----
  int k = 0;
  fetch(c0, 0, (piv[0] >= 0) ? 0 : 1);
  fetch(c1, 1, 1);
  while (k < n) {
    const bool two = piv[k] < 0;
    const int w = two ? 2 : 1;
    const int kn = k + w;
    const bool ntwo = kn < n && piv[kn] < 0;
    fetch(n0, kn, ntwo ? kn + 1 : kn);  // in flight while this step runs
    fetch(n1, kn + 1, kn + 1);
    if (!two) {
      const int kp = piv[k];
      const double bk = xs[kp], bold = xs[k];
      __syncthreads();
      const double mlt = -bk;
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int i = tid + j * BK_TPB;
        if (i > k && i < n) xs[i] = ((i == kp) ? bold : xs[i]) + c0[j] * mlt;
      }
      if (tid == 0) xs[k] = bk / dg[k];
    } else {
      const int kp = -piv[k];
      const double b0v = xs[k], b1v = xs[kp], b1old = xs[k + 1];
      __syncthreads();
      const double m0 = -b0v, m1 = -b1v;
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int i = tid + j * BK_TPB;
        if (i > k + 1 && i < n) xs[i] = ((i == kp) ? b1old : xs[i]) + c0[j] * m0 + c1[j] * m1;
      }
      if (tid == 0) {
        const double off = od[k];
        const double a0 = dg[k] / off, a1 = dg[k + 1] / off;
        const double den = a0 * a1 - 1.0;
        const double q0 = b0v / off, q1 = b1v / off;
        xs[k] = (a1 * q0 - q1) / den;
        xs[k + 1] = (a0 * q1 - q0) / den;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < E; ++j) { c0[j] = n0[j]; c1[j] = n1[j]; }
    k = kn;
  }

  // ---- backward: L^T (with interchanges); the step at k is 2x2 when ipiv[k] < 0 (k = its second row) ----
  k = n - 1;
  {
    const bool two = piv[k] < 0;
    fetch(c0, k, k);                        // column k
    fetch(c1, two ? k - 1 : -1, k);         // column k-1 of a 2x2 pivot: entries below the block only
  }
  int phase = 0;
  while (k >= 0) {
    const bool two = piv[k] < 0;
    const int kn = k - (two ? 2 : 1);
    const bool ntwo = kn >= 0 && piv[kn] < 0;
    fetch(n0, kn, kn);
    fetch(n1, ntwo ? kn - 1 : -1, kn);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int i = tid + j * BK_TPB;
      if (i > k && i < n) { s0 += c0[j] * xs[i]; s1 += c1[j] * xs[i]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    double* rd = red[phase];  // two buffers: the next step's partials never overwrite sums still being read
    if ((tid & 31) == 0) { rd[tid >> 5] = s0; }
    __shared__ double red1[2][BK_WARPS];
    if ((tid & 31) == 0) { red1[phase][tid >> 5] = s1; }
    __syncthreads();
    if (tid == 0) {
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int wi = 0; wi < BK_WARPS; ++wi) { t0 += rd[wi]; t1 += red1[phase][wi]; }
      if (!two) {
        const int kp = piv[k];
        const double v = xs[k] - t0;
        xs[k] = xs[kp]; xs[kp] = v;
        if (kp == k) xs[k] = v;
      } else {
        const int kp = -piv[k];
        xs[k - 1] -= t1;
        const double v = xs[k] - t0;
        xs[k] = xs[kp]; xs[kp] = v;
        if (kp == k) xs[k] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < E; ++j) { c0[j] = n0[j]; c1[j] = n1[j]; }
    phase ^= 1;
    k = kn;
  }
  for (int i = tid; i < n; i += BK_TPB) x[i] = xs[i];
}

// lower triangle (strictly below the diagonal) -> upper triangle, for factors that arrive from the host
__global__ void k_bk_mirror(double* __restrict__ S, int ld, int n) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  for (int c = threadIdx.x & 31; c < r; c += 32) S[(size_t)c * ld + r] = S[(size_t)r * ld + c];
}

int g_bk_sms = 0;
bool g_bk_init = false;

}  // namespace

double bk_alpha() { return (1.0 + std::sqrt(17.0)) / 8.0; }

// Factor `nslots` matrices in place.  One matrix: all SMs as one team (cooperative launch) when it is large
// enough to feed them; a batch: one CTA per matrix.  Returns a cudaError_t.
int launch_bk_factor(cudaStream_t st, int nslots, const int* active, double* S, int ld, size_t sS, int n, int* ipiv,
                     size_t sP, int mirror_input) {
  if (n <= 0 || nslots <= 0) return 0;
  if (!g_bk_init) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_bk_sms, cudaDevAttrMultiProcessorCount, dev);
    g_bk_init = true;
  }
  BkArgs a{S, ld, sS, n, ipiv, sP, active, bk_alpha(), mirror_input};
  int team = 1;
  if (nslots == 1 && n >= 256) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bk_factor<true>, BK_TPB, 0);
    team = g_bk_sms * (per_sm > 0 ? 1 : 0);
    const int want = (n + BK_WARPS - 1) / BK_WARPS;  // at least one row per warp in the first steps
    if (team > want) team = want;
    if (team < 1) team = 1;
  }
  count_launch();
  if (team > 1) {
    void* args[] = {&a};
    return (int)cudaLaunchCooperativeKernel((void*)k_bk_factor<true>, dim3(team, 1), dim3(BK_TPB), args, 0, st);
  }
  k_bk_factor<false><<<dim3(1, nslots), BK_TPB, 0, st>>>(a);
  return (int)cudaGetLastError();
}

// Opt-in shared-memory sizes of the solve kernels; per device (called from factor_init, once for every device used).
constexpr size_t BK_SOLVE_MAX_SMEM = 200 * 1024;  // x alone: n <= 25600
int bk_init() {
  constexpr int FAST_E = 8;
  const size_t fast = (sizeof(double) * 3 + sizeof(int)) * (size_t)(FAST_E * BK_TPB);
  cudaError_t e = cudaFuncSetAttribute(k_bk_solve_fast<FAST_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_bk_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BK_SOLVE_MAX_SMEM);
  return (int)e;
}

int launch_bk_solve(cudaStream_t st, int nslots, const int* active, const double* S, int ld, size_t sS, int n,
                    const int* ipiv, size_t sP, double* x, size_t sx) {
  if (n <= 0 || nslots <= 0) return 0;
  BkArgs a{const_cast<double*>(S), ld, sS, n, const_cast<int*>(ipiv), sP, active, 0.0, 0};
  constexpr int FAST_E = 8;
  if (n <= FAST_E * BK_TPB) {  // latency-hidden variant: x, pivots and ipiv in shared memory, rows prefetched
    const size_t smem_fast = sizeof(double) * 3 * (size_t)n + sizeof(int) * (size_t)n;
    k_bk_solve_fast<FAST_E><<<dim3(1, nslots), BK_TPB, smem_fast, st>>>(a, x, sx);
    count_launch();
    return (int)cudaGetLastError();
  }
  const size_t smem = sizeof(double) * (size_t)n;
  if (smem > BK_SOLVE_MAX_SMEM) return (int)cudaErrorInvalidValue;
  k_bk_solve<<<dim3(1, nslots), BK_TPB, smem, st>>>(a, x, sx);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace ipmz

using namespace ipmz;

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ipmz_fail(IPMZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));   \
  } while (0)

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

extern "C" {

int ipmz_symmetric_indefinite_factorization(int n, const double* A, double* LD, int* ipiv) {
  if (n < 0 || (n > 0 && (!A || !LD || !ipiv))) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  if (n == 0) return IPMZ_OK;
  int rc;
  if ((rc = ipmz_ensure_device(0))) return rc;
  const int ld = pad4(n);
  DevBuf S, P;
  CUDA_TRY(cudaMalloc(&S.p, sizeof(double) * (size_t)n * ld));
  CUDA_TRY(cudaMalloc(&P.p, sizeof(int) * (size_t)n));
  CUDA_TRY(cudaMemset(S.p, 0, sizeof(double) * (size_t)n * ld));
  CUDA_TRY(cudaMemcpy2D(S.p, sizeof(double) * ld, A, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice));
  CUDA_TRY((cudaError_t)launch_bk_factor(nullptr, 1, nullptr, (double*)S.p, ld, (size_t)n * ld, n, (int*)P.p, (size_t)n, 1));
  CUDA_TRY(cudaDeviceSynchronize());
  // the reference returns a copy of the input whose lower triangle was overwritten: keep the caller's upper triangle
  std::vector<double> tmp((size_t)n * n);
  CUDA_TRY(cudaMemcpy2D(tmp.data(), sizeof(double) * n, S.p, sizeof(double) * ld, sizeof(double) * n, n,
                        cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(ipiv, P.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) LD[(size_t)i * n + j] = (j <= i) ? tmp[(size_t)i * n + j] : A[(size_t)i * n + j];
  return IPMZ_OK;
}

/* Device time of one factorization (CUDA events on the launch stream, `reps` factorizations of a pristine device
 * copy, restore copies outside the timed intervals): the roofline evidence of k_bk_factor in bench.py / tools. */
int ipmz_bk_factor_time(int n, const double* A, int reps, double* ms_per_factorization) {
  if (n <= 0 || !A || reps <= 0 || !ms_per_factorization) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  int rc;
  if ((rc = ipmz_ensure_device(0))) return rc;
  const int ld = pad4(n);
  const size_t bytes = sizeof(double) * (size_t)n * ld;
  DevBuf S, S0, P;
  CUDA_TRY(cudaMalloc(&S.p, bytes));
  CUDA_TRY(cudaMalloc(&S0.p, bytes));
  CUDA_TRY(cudaMalloc(&P.p, sizeof(int) * (size_t)n));
  CUDA_TRY(cudaMemset(S0.p, 0, bytes));
  CUDA_TRY(cudaMemcpy2D(S0.p, sizeof(double) * ld, A, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice));
  struct Events {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  } ev;
  CUDA_TRY(cudaEventCreate(&ev.e0));
  CUDA_TRY(cudaEventCreate(&ev.e1));
  cudaEvent_t e0 = ev.e0, e1 = ev.e1;
  double total = 0.0;
  for (int r = -1; r < reps; ++r) {  // r = -1: warm-up
    CUDA_TRY(cudaMemcpyAsync(S.p, S0.p, bytes, cudaMemcpyDeviceToDevice, nullptr));
    CUDA_TRY(cudaEventRecord(e0, nullptr));
    CUDA_TRY((cudaError_t)launch_bk_factor(nullptr, 1, nullptr, (double*)S.p, ld, (size_t)n * ld, n, (int*)P.p, (size_t)n, 1));
    CUDA_TRY(cudaEventRecord(e1, nullptr));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 0) total += ms;
  }
  *ms_per_factorization = total / reps;
  return IPMZ_OK;
}

int ipmz_overwriting_solve_bunch_kaufman(int n, const double* LD, const int* ipiv, double* b) {
  if (n < 0 || (n > 0 && (!LD || !ipiv || !b))) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  if (n == 0) return IPMZ_OK;
  if ((size_t)n * sizeof(double) > 200 * 1024) return ipmz_fail(IPMZ_ERR_ARG, "n too large for the one-CTA Bunch-Kaufman solve");
  int rc;
  if ((rc = ipmz_ensure_device(0))) return rc;
  const int ld = pad4(n);
  DevBuf S, P, X;
  CUDA_TRY(cudaMalloc(&S.p, sizeof(double) * (size_t)n * ld));
  CUDA_TRY(cudaMalloc(&P.p, sizeof(int) * (size_t)n));
  CUDA_TRY(cudaMalloc(&X.p, sizeof(double) * (size_t)ld));
  CUDA_TRY(cudaMemcpy2D(S.p, sizeof(double) * ld, LD, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(P.p, ipiv, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(X.p, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  k_bk_mirror<<<(n + 7) / 8, 256>>>((double*)S.p, ld, n);
  count_launch();
  CUDA_TRY((cudaError_t)launch_bk_solve(nullptr, 1, nullptr, (const double*)S.p, ld, (size_t)n * ld, n, (const int*)P.p,
                                        (size_t)n, (double*)X.p, (size_t)ld));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(b, X.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return IPMZ_OK;
}

}  // extern "C"
