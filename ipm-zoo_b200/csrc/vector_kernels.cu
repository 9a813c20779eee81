// ipm-zoo_b200/csrc/vector_kernels.cu -- fused per-iteration vector work of the Mehrotra
// predictor-corrector loop (reference: Optimizer.cpp:128-130, :147-157, :165-217, :240-342,
// :361-378).  All kernels are HBM-bound streaming passes over the packed iterate; the
// reductions (objective, residual norm, mean complementarity, step-length minimum) use warp
// shuffles inside a block and a deterministic last-block finish across blocks.
#include <cstdlib>

#include "ipmz_device.cuh"
#include "ipmz_kernels.h"
#include "vector_bodies.cuh"

namespace ipmz {

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-level reduce of K values (value k is a min if bit k of MINMASK is set, else a sum),
// then the last block of this problem's grid row combines the per-block partials in block
// order (deterministic) and calls fin(totals) from one thread.
template <int K, unsigned MINMASK, class Fin>
__device__ void reduce_finish(double (&v)[K], const View& vw, Fin fin) {
  __shared__ double sh[K][TPB / 32];
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double r = ((MINMASK >> k) & 1u) ? warp_min(v[k]) : warp_sum(v[k]);
    if (lane == 0) sh[k][warp] = r;
  }
  __syncthreads();
  const int slot = blockIdx.y;
  double* part = vw.partials + ((size_t)slot * vw.maxblk + blockIdx.x) * 8;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double r = sh[k][0];
      for (int w = 1; w < TPB / 32; ++w) r = ((MINMASK >> k) & 1u) ? fmin(r, sh[k][w]) : r + sh[k][w];
      part[k] = r;
    }
    __threadfence();
    const int t = atomicAdd(vw.counters + slot, 1);
    is_last = (t == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double tot[K];
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = ((MINMASK >> k) & 1u) ? 1.0e300 : 0.0;
    const volatile double* all = vw.partials + (size_t)slot * vw.maxblk * 8;
    for (unsigned b = 0; b < gridDim.x; ++b) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double r = all[(size_t)b * 8 + k];
        tot[k] = ((MINMASK >> k) & 1u) ? fmin(tot[k], r) : tot[k] + r;
      }
    }
    vw.counters[slot] = 0;
    fin(tot);
  }
}

}  // namespace

// y[r] = dot(A[r][0:cols], x) for one matrix per problem.  Used for Q x, M x, M^T lambda, M^T (W b1), M dx -- the
// only O(n^2)-byte movers outside the factorization (reference: Evaluation.cpp:35-41).  HBM-bound streaming:
// one warp per row with U = 8 independent 16-byte loads in flight per lane (the pass needs ~5 MB in flight per GPU at
// HBM latency), the matrix read with the streaming (evict-first) policy and x through the read-only path so the
// rows flowing through L1 do not evict the vector every warp re-reads, and CTAs small enough that every SM gets at
// least ~6 of them (a 4096-row matrix on 8-warp CTAs left 3.46 CTAs per SM: 59 % of the copy peak against 88 % for
// 8192 rows).  Several rows per warp were measured and dropped: fewer, fatter warps lose more than the x re-use wins
// (Q x at n = 8192: 1 row 5.8 TB/s, 2 rows 4.9, 4 rows 4.2; `tools/matvec_ab.py`).
// Long rows are split over 2 or 4 warps (a function of `cols` only): a warp that issues 8 loads and then waits for
// them keeps about half of them in flight on average, and 4096 such warps are too few to fill HBM (M x at cfg3 size:
// 3.9 TB/s with one warp per 64 KB row).
// The accumulation order of a row (chunks in order; inside a chunk two accumulators by trip parity, then the warp
// tree) depends on `cols` alone, never on the batch, so a QP gets bitwise the same products alone, in a batch or in
// a sub-batch.
template <int U, bool STREAM, int SPLIT>
__global__ void __launch_bounds__(TPB) k_matvec(const double* __restrict__ A, int lda, size_t sA, int rows,
                                                int cols, const double* __restrict__ x, size_t sx,
                                                double* __restrict__ y, size_t sy,
                                                const int* __restrict__ active) {
  static_assert(U % 2 == 0, "trip parity = unroll parity");
  constexpr int split = SPLIT;  // compile-time: the short-row instantiation carries no division or shared memory
  __shared__ double part_sum[SPLIT > 1 ? TPB / 32 : 1];
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // `split` warps share one long row (a function of cols only, see launch_matvec): each takes a contiguous chunk
  const int row = blockIdx.x * ((blockDim.x >> 5) / split) + wid / split;
  const int part = wid % split;
  const bool valid = row < rows;
  double sres = 0.0;
  if (valid) {
    const double2* a2 = reinterpret_cast<const double2*>(A + (size_t)p * sA + (size_t)row * lda);
    const double2* x2 = reinterpret_cast<const double2*>(x + (size_t)p * sx);
    const int c2 = (cols + 1) >> 1;
    const int chunk = SPLIT == 1 ? c2 : ((((c2 + split - 1) / split) + 63) & ~63);
    const int k0 = part * chunk, k1 = SPLIT == 1 ? c2 : min(c2, k0 + chunk);
    double acc0 = 0.0, acc1 = 0.0;
    for (int k = k0 + lane; k < k1; k += 32 * U) {
      double2 av[U], xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kk = k + 32 * u;
        const bool in = kk < k1;
        xv[u] = in ? (STREAM ? __ldg(x2 + kk) : x2[kk]) : make_double2(0.0, 0.0);
        av[u] = in ? (STREAM ? __ldcs(a2 + kk) : a2[kk]) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + 32 * u < k1) {
          double& t = (u & 1) ? acc1 : acc0;
          t = fma(av[u].x, xv[u].x, t);
          t = fma(av[u].y, xv[u].y, t);
        }
      }
    }
    sres = warp_sum(acc0 + acc1);
  }
  if (split == 1) {
    if (valid && lane == 0) y[(size_t)p * sy + row] = sres;
    return;
  }
  if (lane == 0) part_sum[wid] = sres;
  __syncthreads();
  if (valid && part == 0 && lane == 0) {
    double t = part_sum[wid];
    for (int q = 1; q < split; ++q) t += part_sum[wid + q];  // fixed order: chunk 0, 1, ...
    y[(size_t)p * sy + row] = t;
  }
}

// Short rows (cols <= 1024: the batched QPs, where a row is one or two trips of a warp): plain loop, two 16-byte
// loads in flight per lane, no predication.  Measured against the unrolled kernel above on 4096 QPs of n = 256:
// 200 us against 281 us per launch (6.4 TB/s, the copy peak) -- the whole row is in flight either way, and the
// unrolled kernel only adds dead slots and set-up.  Same accumulation order (trip parity), so the same bits.
__global__ void __launch_bounds__(TPB) k_matvec_short(const double* __restrict__ A, int lda, size_t sA, int rows,
                                                      int cols, const double* __restrict__ x, size_t sx,
                                                      double* __restrict__ y, size_t sy,
                                                      const int* __restrict__ active) {
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double2* a2 = reinterpret_cast<const double2*>(A + (size_t)p * sA + (size_t)row * lda);
  const double2* x2 = reinterpret_cast<const double2*>(x + (size_t)p * sx);
  const int c2 = (cols + 1) >> 1;
  double acc0 = 0.0, acc1 = 0.0;
  int k = lane;
  for (; k + 32 < c2; k += 64) {
    const double2 a = a2[k], b = a2[k + 32];
    const double2 u = x2[k], w = x2[k + 32];
    acc0 = fma(a.x, u.x, acc0); acc0 = fma(a.y, u.y, acc0);
    acc1 = fma(b.x, w.x, acc1); acc1 = fma(b.y, w.y, acc1);
  }
  for (; k < c2; k += 32) {
    const double2 a = a2[k];
    const double2 u = x2[k];
    acc0 = fma(a.x, u.x, acc0); acc0 = fma(a.y, u.y, acc0);
  }
  const double s = warp_sum(acc0 + acc1);
  if (lane == 0) y[(size_t)p * sy + row] = s;
}

// Tiled transpose M [m x n] -> MT [n x m], once per problem upload.
__global__ void k_transpose(const double* __restrict__ M, int ldm, size_t sM, double* __restrict__ MT,
                            int ldmt, size_t sMT, int m, int n) {
  __shared__ double tile[32][33];
  const int p = blockIdx.z;
  const double* src = M + (size_t)p * sM;
  double* dst = MT + (size_t)p * sMT;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < m && c < n) ? src[(size_t)r * ldm + c] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = c0 + i, c = r0 + threadIdx.x;  // row of MT = column of M
    if (r < n && c < m) dst[(size_t)r * ldmt + c] = tile[threadIdx.x][i];
  }
}

__global__ void k_initial_point(View v) {
  initial_point_body(v, problem_of(v), blockIdx.x * blockDim.x + threadIdx.x);
}

// Residuals, W, objective / residual norm / mean complementarity, stopping test and the augmented right-hand side
// (bodies: vector_bodies.cuh).  MODE 0: start of an iteration (mu = 0); MODE 1: corrector.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_residuals_rhs(View v) {
  const int p = problem_of(v);
  const int i = blockIdx.x * TPB + threadIdx.x;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // 0.5 x'Qx, c'x, sum r^2, sum |comp|
  residuals_rhs_body<MODE>(v, p, i, acc);
  if (MODE == 0) reduce_finish<4, 0u>(acc, v, [&](const double* t) { residuals_finish(v, p, t); });
}

__global__ void k_prepare_sol(View v, const double* __restrict__ rvec_all, int stage) {
  prepare_sol_body(v, problem_of(v), blockIdx.x * blockDim.x + threadIdx.x, rvec_all, stage);
}

__global__ void k_recover_dual(View v, const double* __restrict__ rvec_all, int accumulate) {
  recover_dual_body(v, problem_of(v), blockIdx.x * blockDim.x + threadIdx.x, rvec_all, accumulate);
}

__global__ void k_aug_residual(View v) {
  aug_residual_body(v, problem_of(v), blockIdx.x * blockDim.x + threadIdx.x);
}

// Dual-Schur normal equations (the reference's get_normal_equations, SymbolicOptimization.cpp:465-478): the vector
// steps between the three triangular solves of one Newton solve, rvec = augmented right-hand side b0|b1:
//   0: sol = b0                      -> Hx y = b0
//   1: lam = M y - b1                -> S dlam = M Hx^-1 b0 - b1
//   2: sol = b0 - M^T dlam           -> Hx dx = b0 - M^T dlam
//   3: sol[n ..] = dlam              (augmented layout [dx; dlam] for the back-substitution)
__global__ void k_dual_vec(View v, int stage, const double* __restrict__ rvec_all, double* __restrict__ lam_all) {
  const int p = problem_of(v);
  const Shape& s = v.s;
  const double* rvec = rvec_all + (size_t)p * (s.ns + s.ms);
  double* sol = v.sol + (size_t)p * v.ssol;
  double* lam = lam_all + (size_t)p * s.ms;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (stage == 0) {
    if (i < s.n) sol[i] = rvec[i];
  } else if (stage == 1) {
    if (i < s.m) lam[i] = v.Mx[(size_t)p * s.ms + i] - rvec[s.ns + i];
  } else if (stage == 2) {
    if (i < s.n) sol[i] = rvec[i] - v.tn[(size_t)p * s.ns + i];
  } else {
    if (i < s.m) sol[s.n + i] = lam[i];
  }
}

// Back-substitution + step length (bodies: vector_bodies.cuh).  MODE 0 writes the affine direction DA and
// alpha_aff, MODE 1 the final direction D and alpha.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_backsub_step(View v) {
  const int p = problem_of(v);
  const int i = blockIdx.x * TPB + threadIdx.x;
  double a[1] = {1.0};
  backsub_step_body<MODE>(v, p, i, a[0]);
  reduce_finish<1, 1u>(a, v, [&](const double* t) {
    Scal& sc = v.sc[p];
    const double al = fmin(1.0, t[0]);
    if (MODE == 0) sc.alpha_aff = al; else sc.alpha = al;
  });
}

// mu after the full affine step, sigma = (mu_aff/mu)^3, mu_c = sigma*mu (Optimizer.cpp:165-181).
__global__ void __launch_bounds__(TPB) k_mu_affine(View v) {
  const int p = problem_of(v);
  const int i = blockIdx.x * TPB + threadIdx.x;
  double acc[1] = {0.0};
  mu_affine_body(v, p, i, acc[0]);
  reduce_finish<1, 0u>(acc, v, [&](const double* t) { mu_affine_finish(v, p, t[0]); });
}

// v += fraction_to_boundary * alpha * dv over the whole pack (Optimizer.cpp:216-238).
__global__ void k_update(View v) {
  const int p = problem_of(v);
  double* V = v.V + (size_t)p * v.sp;
  const double* D = v.D + (size_t)p * v.sp;
  const double st = v.ftb * v.sc[p].alpha;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < v.sp) V[i] = V[i] + st * D[i];
  if (i == 0) v.sc[p].iters += 1;
}

// Compaction of the active problems that refine their condensed solves (mu < thr) -- one warp, ballot + prefix;
// the host counts the same predicate on the Scal records it has just read, so only the list stays on the device
// (an H2D copy here would queue behind the problem uploads of other sub-batches).
__global__ void k_refine_list(View v, int nact, double thr, int* __restrict__ out) {
  const int lane = threadIdx.x;
  int base_out = 0;
  for (int base = 0; base < nact; base += 32) {
    const int i = base + lane;
    const int p = i < nact ? (v.active ? v.active[i] : i) : 0;
    const bool need = i < nact && v.sc[p].mu < thr;
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (need) out[base_out + __popc(m & ((1u << lane) - 1u))] = p;
    base_out += __popc(m);
  }
}

// ---- host-side launchers --------------------------------------------------------------
void launch_matvec(cudaStream_t st, int nslots, const int* active, const double* A, int lda, size_t sA,
                   int rows, int cols, const double* x, size_t sx, double* y, size_t sy) {
  if (rows <= 0 || nslots <= 0) return;
  if (cols <= 1024) {
    k_matvec_short<<<dim3((rows + TPB / 32 - 1) / (TPB / 32), nslots), TPB, 0, st>>>(A, lda, sA, rows, cols, x, sx, y, sy,
                                                                                     active);
    count_launch();
    return;
  }
  const int split = cols >= 8192 ? 4 : cols >= 2048 ? 2 : 1;  // warps per row: by cols only (bitwise stable per shape)
  // warps per CTA: 8, or fewer when that would leave an SM with less than ~6 CTAs to balance
  const long long total = (long long)rows * nslots * split;
  int w = total >= 8LL * 148 * 6 ? 8 : total >= 4LL * 148 * 6 ? 4 : 2;
  if (w < split) w = split;
  const int rows_per_cta = w / split;
  const dim3 grid((rows + rows_per_cta - 1) / rows_per_cta, nslots);
#define MV(U_, S_, P_) k_matvec<U_, S_, P_><<<grid, 32 * w, 0, st>>>(A, lda, sA, rows, cols, x, sx, y, sy, active)
  if (split == 1) MV(8, true, 1);
  else if (split == 2) MV(8, true, 2);
  else MV(8, true, 4);
#undef MV
  count_launch();
}

void launch_transpose(cudaStream_t st, int count, const double* M, int ldm, size_t sM, double* MT, int ldmt,
                      size_t sMT, int m, int n) {
  if (m <= 0 || n <= 0 || count <= 0) return;
  dim3 grid((n + 31) / 32, (m + 31) / 32, count);
  k_transpose<<<grid, dim3(32, 8), 0, st>>>(M, ldm, sM, MT, ldmt, sMT, m, n); count_launch();
}

static dim3 vec_grid(const View& v, int nslots) {
  const int len = v.s.ns > v.s.ms ? v.s.ns : v.s.ms;
  return dim3((len + TPB - 1) / TPB, nslots);
}

void launch_initial_point(cudaStream_t st, const View& v, int nslots) {
  k_initial_point<<<vec_grid(v, nslots), TPB, 0, st>>>(v); count_launch();
}
void launch_residuals_rhs(cudaStream_t st, const View& v, int nslots, int mode) {
  if (mode == 0) k_residuals_rhs<0><<<vec_grid(v, nslots), TPB, 0, st>>>(v);
  else k_residuals_rhs<1><<<vec_grid(v, nslots), TPB, 0, st>>>(v);
  count_launch();
}
void launch_prepare_sol(cudaStream_t st, const View& v, int nslots, const double* rvec, int stage) {
  k_prepare_sol<<<vec_grid(v, nslots), TPB, 0, st>>>(v, rvec, stage); count_launch();
}
void launch_recover_dual(cudaStream_t st, const View& v, int nslots, const double* rvec, int accumulate) {
  k_recover_dual<<<vec_grid(v, nslots), TPB, 0, st>>>(v, rvec, accumulate); count_launch();
}
void launch_aug_residual(cudaStream_t st, const View& v, int nslots) {
  k_aug_residual<<<vec_grid(v, nslots), TPB, 0, st>>>(v); count_launch();
}
void launch_dual_vec(cudaStream_t st, const View& v, int nslots, int stage, const double* rvec, double* lam) {
  k_dual_vec<<<vec_grid(v, nslots), TPB, 0, st>>>(v, stage, rvec, lam); count_launch();
}
void launch_backsub_step(cudaStream_t st, const View& v, int nslots, int mode) {
  if (mode == 0) k_backsub_step<0><<<vec_grid(v, nslots), TPB, 0, st>>>(v);
  else k_backsub_step<1><<<vec_grid(v, nslots), TPB, 0, st>>>(v);
  count_launch();
}
void launch_mu_affine(cudaStream_t st, const View& v, int nslots) {
  k_mu_affine<<<vec_grid(v, nslots), TPB, 0, st>>>(v); count_launch();
}
void launch_refine_list(cudaStream_t st, const View& v, int nact, double thr, int* out) {
  k_refine_list<<<1, 32, 0, st>>>(v, nact, thr, out); count_launch();
}
void launch_update(cudaStream_t st, const View& v, int nslots) {
  dim3 grid((unsigned)((v.sp + TPB - 1) / TPB), nslots);
  k_update<<<grid, TPB, 0, st>>>(v); count_launch();
}

}  // namespace ipmz
