// ipm-zoo_b200/csrc/batch_fused.cu -- batches of small QPs (cfg4): the WHOLE Mehrotra predictor-corrector solve of
// one problem inside one persistent CTA, one launch per batch.
//
// Reference control flow per problem: Optimizer::solve_quasi_definite_ (Optimizer.cpp:77-220) with
// LinearSolvers::ldlt_decomposition / overwriting_solve_ldlt (LinearSolvers.cpp:14-74).  The multi-kernel batched
// schedule (solver.cu: ~110 launches per iteration, a host synchronisation per iteration to read the stopping test,
// every matrix pass re-streamed from HBM by a fresh grid) is replaced for systems that fit by:
//
//   * a persistent grid (2 CTAs per SM); each CTA pulls problem indices from a ticket counter and runs the complete
//     solve of that problem -- matvecs, residuals, stopping test, assembly, LDL^T, both Newton solves (with the
//     normal reduction's iterative refinement), centring, step length and update -- so there is no host round trip
//     and no wave quantisation: problems that converge early free their CTA for the next ticket;
//   * two co-resident CTAs per SM are always in different phases, so the latency-bound chains of one problem (the
//     one-warp 32 x 32 LDL^T of a diagonal block, the triangular sweeps, block reductions) overlap the DMMA-bound
//     phases of the other (condensed assembly M^T W M, trailing updates);
//   * the matrix K / its factor L of a problem is touched by one SM only and stays in L2 between the phases of an
//     iteration; Q, M and M^T are read-only and re-read from L2 while the problem is in flight.
//
// Phases of one iteration (all vector formulas are the bodies of vector_bodies.cuh, shared with the grid-per-phase
// kernels):
//   matvecs            Q x, M x, M^T lambda: one warp per row, four rows in flight per warp
//   residuals          r_*, W, objective, ||res||, mu, stopping test, predictor right-hand side
//   assembly           NORMAL: K = Q + Y^-1 L_y + Z^-1 L_z + M^T W M on the FP64 tensor pipe (64 x 64 tiles, accumulators
//                      start from Q, 16-wide k-slices of M^T through a 3-stage cp.async ring that runs across tile
//                      boundaries, the W scaling folded into the B fragment); AUGMENTED: a copy pass
//   factorization      right-looking LDL^T, 32-wide panels: panel in shared memory, diagonal block by one warp
//                      (warp_ldlt32), rows below on the tensor pipe with the 8 x 8 inverse blocks (panel_solve32),
//                      trailing matrix updated in L2 with both operands read from the shared-memory panel
//   solves             forward / pivots / backward with x in shared memory, 64-row blocks
//   back-substitution  eliminated Delta's, ratio test, centring parameter, corrector right-hand side, update
#include <cuda_runtime.h>
#include <stdlib.h>

#include "ipmz_device.cuh"
#include "ipmz_kernels.h"
#include "ldlt_device.cuh"
#include "vector_bodies.cuh"

namespace ipmz {

// debug builds (-DIPMZ_FUSED_CLOCKS): SM cycles per phase, summed over every CTA's thread 0
__device__ unsigned long long g_fused_clk[16];

namespace {

#ifdef IPMZ_FUSED_CLOCKS
#define FPH(i) do { if (threadIdx.x == 0) { const long long t__ = clock64(); atomicAdd(&g_fused_clk[i], (unsigned long long)(t__ - ph_last)); ph_last = t__; } } while (0)
#define FPH_DECL long long ph_last = clock64()
#define FSUB_BEGIN long long sub_t0__ = clock64()
#define FSUB_END(i) do { __syncthreads(); if (threadIdx.x == 0) atomicAdd(&g_fused_clk[i], (unsigned long long)(clock64() - sub_t0__)); } while (0)
#else
#define FPH(i) do {} while (0)
#define FPH_DECL do {} while (0)
#define FSUB_BEGIN do {} while (0)
#define FSUB_END(i) do {} while (0)
#endif

constexpr int FT = 256;          // threads per CTA
constexpr int FW = FT / 32;      // warps
constexpr int PP = 36;           // panel pitch: 36 mod 16 == 4 -> conflict-free DMMA fragment loads
constexpr int TS = 64;           // assembly tile
constexpr int FSTAGES = 3;
constexpr int STAGE_DOUBLES = 2 * TS * LDT;  // A-side + B-side rows of one 16-wide k-slice
constexpr int SB64 = 64, SP65 = 65;  // block of the triangular sweeps

struct FusedArgs {
  View v;
  int count;
  int* ticket;
  int refine_fixed;  // >= 0: that many refinement steps per condensed solve; -1: by the problem's mu (solver.cu policy)
  int smem_doubles;
  // streamed solve (ipmz_batch_solve_streamed): the kernel is launched BEFORE the problem data is uploaded; `ready` is a
  // device word the copy stream overwrites (4-byte H2D copy, stream-ordered after each chunk of problems) with the number
  // of problems whose data is resident.  A CTA holding ticket p waits until *ready > p, then builds the reference's
  // initial point and the transposed copy M^T itself (the upload path's k_initial_point / k_transpose could not be
  // scheduled while this persistent grid owns every SM).  nullptr: everything is resident already.
  const volatile int* ready;
  int* abort_flag;   // set when a wait on `ready` times out (the host returns an error instead of hanging the GPU)
  double* MT_w;      // writable alias of v.MT for the in-kernel transpose
  int dbg;  // experiments (IPMZ_FUSED_DBG): 1 = only the opening matvecs (30x), 2 = assembly (20x), 3 = + LDL^T, 4 = + one solve
};

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// CTA-wide reduction of K values in a fixed order (lanes by the shuffle tree, warps 0..7 in order): thread 0 gets the
// totals.  Ends with the values in red[k][0..FW); the caller synchronises before reusing `red`.
template <int K, unsigned MINMASK>
__device__ __forceinline__ void cta_reduce(double (&v)[K], double (*red)[FW], double* tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double r = ((MINMASK >> k) & 1u) ? wmin(v[k]) : wsum(v[k]);
    if (lane == 0) red[k][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double r = red[k][0];
      for (int w = 1; w < FW; ++w) r = ((MINMASK >> k) & 1u) ? fmin(r, red[k][w]) : r + red[k][w];
      tot[k] = r;
    }
  }
}

// y[r] = dot(A[r][0:cols], x), one warp per row, four rows in flight per warp (the rows are 1-2 KB: a single row per
// trip leaves the warp waiting on one L2 / HBM round trip).  Per row the accumulation order is that of
// k_matvec_short (two accumulators by trip parity, then the shuffle tree).
__device__ __forceinline__ void cta_matvec(const double* __restrict__ A, int lda, int rows, int cols, const double* x,
                                           double* y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c2 = (cols + 1) >> 1;
  const double2* x2 = reinterpret_cast<const double2*>(x);
  for (int r0 = warp * 4; r0 < rows; r0 += FW * 4) {
    double acc0[4] = {0.0, 0.0, 0.0, 0.0}, acc1[4] = {0.0, 0.0, 0.0, 0.0};
    const double2* a2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a2[u] = reinterpret_cast<const double2*>(A + (size_t)min(r0 + u, rows - 1) * lda);
    int k = lane;
    for (; k + 32 < c2; k += 64) {
      double2 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = a2[u][k]; b[u] = a2[u][k + 32]; }
      const double2 xa = x2[k], xb = x2[k + 32];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc0[u] = fma(a[u].x, xa.x, acc0[u]); acc0[u] = fma(a[u].y, xa.y, acc0[u]);
        acc1[u] = fma(b[u].x, xb.x, acc1[u]); acc1[u] = fma(b[u].y, xb.y, acc1[u]);
      }
    }
    for (; k < c2; k += 32) {
      const double2 xa = x2[k];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double2 a = a2[u][k];
        acc0[u] = fma(a.x, xa.x, acc0[u]); acc0[u] = fma(a.y, xa.y, acc0[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double s = wsum(acc0[u] + acc1[u]);
      if (lane == 0 && r0 + u < rows) y[r0 + u] = s;
    }
  }
}

// Matrix-vector product, TMA version (default; -DIPMZ_FUSED_NO_TMA_MV selects cta_matvec above): the matrix (rows contiguous, pitch lda) streams through
// two 32 KB shared-memory buffers by bulk copies (cp.async.bulk.shared.global, SASS UBLKCP): ONE instruction of one
// thread moves a whole chunk and completes on an mbarrier by transaction bytes.  Work split inside a chunk: S = 2^k
// lanes share one row (S >= cols / 16, so a lane's part of x is 8 double2 registers loaded once per call), FT / S rows
// per chunk; lane `seg` of a row takes the column pairs seg + S j (conflict-free reads, 8 independent 16-byte loads in
// flight per thread) and a row is finished by log2(S) shuffle steps.
constexpr int MV_NBUF = 2;
constexpr int MV_CHUNK = 4096;  // doubles per staging buffer
__shared__ __align__(8) unsigned long long g_mvbar[MV_NBUF];  // one mbarrier per staging buffer, initialised by the kernel
__shared__ unsigned g_mvph;                                    // their phase bits (uniform per CTA)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned da = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  // earlier generic-proxy accesses of the buffer (other phases use the same shared memory) before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;\n" ::"r"(ba), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(da), "l"(gsrc), "r"(bytes), "r"(ba) : "memory");
}
// bar: MV_NBUF mbarriers (count 1) initialised once per kernel; ph: their phase bits, carried by the caller
__device__ __noinline__ void cta_matvec_tma(const double* __restrict__ A, int lda, int rows, int cols, const double* x,
                                            double* y, double* sm) {
  const int tid = threadIdx.x;
  unsigned long long* bar = g_mvbar;
  const int c2 = (cols + 1) >> 1;
  int S = 4;
  while (8 * S < c2) S <<= 1;  // cols <= 512 -> S <= 32
  int R = FT / S;              // rows per chunk
  if (R * lda > MV_CHUNK) R = MV_CHUNK / lda;
  const int nch = (rows + R - 1) / R;
  const int rr = tid / S, seg = tid - rr * S;
  __syncthreads();  // the staging buffers are free and x is visible
  unsigned ph = g_mvph;
  if (tid == 0)
    for (int c = 0; c < MV_NBUF && c < nch; ++c)
      bulk_load(sm + c * MV_CHUNK, A + (size_t)c * R * lda, (unsigned)(min(R, rows - c * R) * lda) * 8u, bar + c);
  double2 xr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = seg + S * j;
    xr[j] = k < c2 ? reinterpret_cast<const double2*>(x)[k] : make_double2(0.0, 0.0);
  }
  for (int c = 0; c < nch; ++c) {
    const int bi = c % MV_NBUF;
    mbar_wait(bar + bi, (ph >> bi) & 1u);
    ph ^= 1u << bi;
    const int r0 = c * R, nr = min(R, rows - r0);
    {
      const double2* a2 = reinterpret_cast<const double2*>(sm + bi * MV_CHUNK + (size_t)(rr < nr ? rr : 0) * lda);
      double2 av[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = seg + S * j;
        av[j] = k < c2 ? a2[k] : make_double2(0.0, 0.0);
      }
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        acc0 = fma(av[j].x, xr[j].x, acc0); acc1 = fma(av[j].y, xr[j].y, acc1);
        acc2 = fma(av[j + 1].x, xr[j + 1].x, acc2); acc3 = fma(av[j + 1].y, xr[j + 1].y, acc3);
      }
      double t = (acc0 + acc1) + (acc2 + acc3);
      for (int o = S >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (seg == 0 && rr < nr) y[r0 + rr] = t;
    }
    __syncthreads();
    if (tid == 0 && c + MV_NBUF < nch)
      bulk_load(sm + bi * MV_CHUNK, A + (size_t)(c + MV_NBUF) * R * lda,
                (unsigned)(min(R, rows - (c + MV_NBUF) * R) * lda) * 8u, bar + bi);
  }
  if (tid == 0) g_mvph = ph;  // read by the next call after its opening barrier
}
#ifndef IPMZ_FUSED_NO_TMA_MV
#define MATVEC(A_, lda_, rows_, cols_, x_, y_) cta_matvec_tma(A_, lda_, rows_, cols_, x_, y_, sm)
#else
#define MATVEC(A_, lda_, rows_, cols_, x_, y_) cta_matvec(A_, lda_, rows_, cols_, x_, y_)
#endif

// ---- assembly ---------------------------------------------------------------------------------------------------
// NORMAL: K(lower) = Q + diag(hd) + MT diag(w) MT^T, hd = Y^-1 L_y + Z^-1 L_z.  sm: FSTAGES stages | w[ms] | hd[ns].
__device__ void assemble_normal(const View& v, int p, double* sm) {
  const Shape& s = v.s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps, warp tile 32 x 16
  double* wsm = sm + FSTAGES * STAGE_DOUBLES;
  double* hd = wsm + s.ms;
  const double* V = v.V + (size_t)p * v.sp;
  const double* Q = v.Q + (size_t)p * v.sQ;
  const double* MT = v.MT + (size_t)p * v.sMT;
  double* K = v.K + (size_t)p * v.sK;
  for (int i = tid; i < s.ms; i += FT) wsm[i] = i < s.m ? v.W[(size_t)p * s.ms + i] : 0.0;
  for (int i = tid; i < s.n; i += FT) {
    double h = 0.0;
    if (s.ylo) h = h + inv_guard(nslot(V, s, YS)[i]) * nslot(V, s, LAMY)[i];
    if (s.zup) h = h + inv_guard(nslot(V, s, ZS)[i]) * nslot(V, s, LAMZ)[i];
    hd[i] = h;
  }
  const int nt = (s.n + TS - 1) / TS;
  const int KT = (s.m + BK - 1) / BK;
  const int ntile = nt * (nt + 1) / 2;
  const int nsteps = ntile * KT;

  // load cursor (runs FSTAGES-1 steps ahead of the compute cursor, across tile boundaries)
  int l_ti = 0, l_tj = 0, l_kt = 0;
  auto load_next = [&](int stage) {
    double* As = sm + stage * STAGE_DOUBLES;
    const int kbase = l_kt * BK;
#pragma unroll
    for (int i = 0; i < 2 * TS * (BK / 2) / FT; ++i) {
      const int chunk = tid + i * FT;
      const int r = chunk >> 3, ck = (chunk & 7) * 2;
      const int gr = (r < TS ? l_ti * TS + r : l_tj * TS + (r - TS));
      const int k = kbase + ck;
      const bool ok = gr < s.n && k < s.ms;  // MT's padding columns [m, ms) are zero
      cp_async16(As + r * LDT + ck, MT + (size_t)(ok ? gr : 0) * v.ldmt + (ok ? k : 0), ok ? 16 : 0);
    }
    if (++l_kt == KT) { l_kt = 0; if (++l_tj > l_ti) { l_tj = 0; ++l_ti; } }
  };
  __syncthreads();  // wsm, hd visible; the stage buffers are free (previous phase done)
  int loaded = 0;
  for (; loaded < FSTAGES - 1; ++loaded) {
    if (loaded < nsteps) load_next(loaded);
    cp_async_commit();
  }
  int ti = 0, tj = 0;
  double acc[4][2][2];
  for (int step = 0, kt = 0; step < nsteps; ++step) {
    const int row0 = ti * TS + wm * 32, col0 = tj * TS + wn * 16;
    if (kt == 0) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int row = row0 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const int col = col0 + ni * 8 + 2 * q;
          double2 c = make_double2(0.0, 0.0);
          if (row < s.n && col < s.ns) c = *reinterpret_cast<const double2*>(Q + (size_t)row * v.ldq + col);
          if (row == col) c.x += hd[row < s.n ? row : 0];
          if (row == col + 1) c.y += hd[row < s.n ? row : 0];
          acc[mi][ni][0] = c.x; acc[mi][ni][1] = c.y;
        }
      }
    }
    cp_async_wait<FSTAGES - 2>();
    __syncthreads();
    if (loaded < nsteps) load_next(loaded % FSTAGES);
    cp_async_commit();
    ++loaded;
    const double* As = sm + (step % FSTAGES) * STAGE_DOUBLES;
    const double* Aw = As + (wm * 32 + g) * LDT + q;
    const double* Bw = As + (TS + wn * 16 + g) * LDT + q;
    const double* wk = wsm + kt * BK + q;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double af[4], bf[2];
      const double w = wk[kk * 4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) af[mi] = Aw[mi * 8 * LDT + kk * 4];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) bf[ni] = Bw[ni * 8 * LDT + kk * 4] * w;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
    }
    if (++kt == KT) {
      kt = 0;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int row = row0 + mi * 8 + g;
        if (row >= s.n) continue;
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const int col = col0 + ni * 8 + 2 * q;
          if (col > row) continue;
          double* dst = K + (size_t)row * v.ldk + col;
          if (col + 1 <= row) *reinterpret_cast<double2*>(dst) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          else dst[0] = acc[mi][ni][0];
        }
      }
      if (++tj > ti) { tj = 0; ++ti; }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// AUGMENTED: lower triangle of [[Q + diag(hd), .],[M, -W^-1]] (N = n + m), one warp per row.
__device__ void assemble_augmented(const View& v, int p) {
  const Shape& s = v.s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* V = v.V + (size_t)p * v.sp;
  double* K = v.K + (size_t)p * v.sK;
  for (int r = warp; r < v.N; r += FW) {
    double* Kr = K + (size_t)r * v.ldk;
    if (r < s.n) {
      const double* q = v.Q + (size_t)p * v.sQ + (size_t)r * v.ldq;
      for (int c = lane; c < r; c += 32) Kr[c] = q[c];
      if (lane == 0) {
        double dii = q[r];
        if (s.ylo) dii = dii + inv_guard(nslot(V, s, YS)[r]) * nslot(V, s, LAMY)[r];
        if (s.zup) dii = dii + inv_guard(nslot(V, s, ZS)[r]) * nslot(V, s, LAMZ)[r];
        Kr[r] = dii;
      }
    } else {
      const int j = r - s.n;
      const double* mr = v.M + (size_t)p * v.sM + (size_t)j * v.ldm;
      for (int c = lane; c < s.n; c += 32) Kr[c] = mr[c];
      for (int c = s.n + lane; c < r; c += 32) Kr[c] = 0.0;
      if (lane == 0) Kr[r] = -v.winv[(size_t)p * s.ms + j];
    }
  }
  __syncthreads();
}

// ---- factorization ----------------------------------------------------------------------------------------------
// In place on the lower triangle of K (N x N, leading dimension ld): strict lower = L, pivots -> Dg.
// sm: panel [round16(N) x PP] | dsm[32] | dinv[32] | colbuf[CBUF] | binv[INV_SUB].
__device__ void ldlt_panels(double* K, int ld, double* Dg, int N, double* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int rows_cap = (N + 15) & ~15;
  double* P = sm;
  double* dsm = P + (size_t)rows_cap * PP;
  double* dinv = dsm + 32;
  double* colbuf = dinv + 32;
  double* binv = colbuf + CBUF;
  for (int j0 = 0; j0 < N; j0 += SB) {
    const int jb = min(SB, N - j0);
    const int R = N - j0;                 // rows of the panel (diagonal block included)
    const int R16 = (R + 15) & ~15;
    // ---- panel -> shared memory (rows beyond R and columns beyond jb as zeros; the diagonal block's upper part too)
    for (int idx = tid; idx < R16 * (SB / 2); idx += FT) {
      const int r = idx >> 4, c = (idx & 15) * 2;
      int bytes = 0;
      if (r < R && c < jb && !(r < jb && c > r)) bytes = (jb - c >= 2) ? 16 : 8;
      cp_async16(P + r * PP + c, K + (size_t)(j0 + (bytes ? r : 0)) * ld + j0 + (bytes ? c : 0), bytes);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // ---- diagonal block by one warp
    if (warp == 0) warp_ldlt32<PP>(P, 0, jb, dsm, dinv, colbuf, binv, lane);
    __syncthreads();
    const int rem = R - jb;  // rows below the diagonal block (only when jb == 32)
    if (rem > 0) {
      panel_solve32<PP>(P + SB * PP, rem, P, dsm, binv, warp, lane, FW);
      __syncthreads();
    }
    // ---- L and the pivots back to global memory
    for (int idx = tid; idx < R * SB; idx += FT) {
      const int r = idx >> 5, c = idx & 31;
      if (c < jb && (r >= jb || c < r)) K[(size_t)(j0 + r) * ld + j0 + c] = P[r * PP + c];
    }
    if (tid < jb) Dg[j0 + tid] = dsm[tid];
    // ---- trailing update  C -= L_panel diag(d) L_panel^T  on the lower triangle, 16 x 16 tiles per warp
    if (rem > 0) {
      const double* T = P + SB * PP;
      double* C = K + (size_t)(j0 + SB) * ld + (j0 + SB);
      const int tm = (rem + 15) >> 4;
      const int ntask = tm * (tm + 1) / 2;
      for (int t = warp; t < ntask; t += FW) {
        int mi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while (mi * (mi + 1) / 2 > t) --mi;
        while ((mi + 1) * (mi + 2) / 2 <= t) ++mi;
        const int ni = t - mi * (mi + 1) / 2;
        const int ra = mi * 16 + g, rb = ni * 16 + g;
        double2 cv[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = mi * 16 + i * 8 + g, col = ni * 16 + j * 8 + 2 * q;
            cv[i][j] = make_double2(0.0, 0.0);
            if (row < rem && col <= row) {
              const double* src = C + (size_t)row * ld + col;
              if (col + 1 <= row) cv[i][j] = *reinterpret_cast<const double2*>(src);
              else cv[i][j].x = src[0];
            }
          }
        double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
        double af[8][2], bf[8][2];
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
          const double dk = dsm[k8 * 4 + q];
          af[k8][0] = T[ra * PP + k8 * 4 + q];
          af[k8][1] = T[(ra + 8) * PP + k8 * 4 + q];
          bf[k8][0] = T[rb * PP + k8 * 4 + q] * dk;
          bf[k8][1] = T[(rb + 8) * PP + k8 * 4 + q] * dk;
        }
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
          dmma884(acc[0][0], af[k8][0], bf[k8][0]);
          dmma884(acc[0][1], af[k8][0], bf[k8][1]);
          dmma884(acc[1][0], af[k8][1], bf[k8][0]);
          dmma884(acc[1][1], af[k8][1], bf[k8][1]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = mi * 16 + i * 8 + g, col = ni * 16 + j * 8 + 2 * q;
            if (row < rem && col <= row) {
              double* dst = C + (size_t)row * ld + col;
              if (col + 1 <= row) *reinterpret_cast<double2*>(dst) = make_double2(cv[i][j].x - acc[i][j][0], cv[i][j].y - acc[i][j][1]);
              else dst[0] = cv[i][j].x - acc[i][j][0];
            }
          }
      }
    }
    __syncthreads();  // panel buffer free, trailing matrix written
  }
}

// ---- solves -----------------------------------------------------------------------------------------------------
// x <- L^-1 x, x <- x / D, x <- L^-T x with the in-place factor; x (global, length N) is staged in shared memory.
// sm: sx[nblk * 64] | Ld[64 x 65] | part[4][64].
__device__ void ldlt_solve(const double* K, int ld, const double* Dg, int N, double* x, double* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (N + SB64 - 1) / SB64;
  double* sx = sm;
  double* Ld = sx + nblk * SB64;
  double* part = Ld + SB64 * SP65;
  for (int i = tid; i < nblk * SB64; i += FT) sx[i] = i < N ? x[i] : 0.0;
  __syncthreads();
  // forward
  for (int r = 0; r < nblk; ++r) {
    const int R0 = r * SB64, nr = min(SB64, N - R0);
    for (int idx = tid; idx < SB64 * SB64; idx += FT) {
      const int i = idx >> 6, c = idx & 63;
      Ld[i * SP65 + c] = (i < nr && c < i) ? K[(size_t)(R0 + i) * ld + R0 + c] : 0.0;
    }
    double sacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sacc[i] = 0.0;
    for (int c0 = 0; c0 < R0; c0 += 64) {
      const double xa = sx[c0 + lane], xb = sx[c0 + 32 + lane];
      double la[8], lb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 8 * warp + i;
        const double* Kr = K + (size_t)(R0 + (row < nr ? row : 0)) * ld + c0 + lane;
        la[i] = row < nr ? Kr[0] : 0.0;
        lb[i] = row < nr ? Kr[32] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sacc[i] = fma(lb[i], xb, fma(la[i], xa, sacc[i]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double t = wsum(sacc[i]);
      if (lane == 0) sx[R0 + 8 * warp + i] -= t;
    }
    __syncthreads();
    if (warp == 0) {
      double y0 = sx[R0 + lane], y1 = sx[R0 + lane + 32];
#pragma unroll 8
      for (int c = 0; c < SB64; ++c) {
        const double yc = __shfl_sync(0xffffffffu, c < 32 ? y0 : y1, c & 31);
        if (lane > c) y0 -= Ld[lane * SP65 + c] * yc;
        if (lane + 32 > c) y1 -= Ld[(lane + 32) * SP65 + c] * yc;
      }
      sx[R0 + lane] = y0;
      sx[R0 + lane + 32] = y1;
    }
    __syncthreads();
  }
  for (int i = tid; i < N; i += FT) sx[i] = sx[i] / Dg[i];
  __syncthreads();
  // backward
  const int c = tid & 63, grp = tid >> 6;
  for (int r = nblk - 1; r >= 0; --r) {
    const int R0 = r * SB64, nr = min(SB64, N - R0);
    for (int idx = tid; idx < SB64 * SB64; idx += FT) {
      const int i = idx >> 6, cc = idx & 63;
      Ld[i * SP65 + cc] = (i < nr && cc < i) ? K[(size_t)(R0 + i) * ld + R0 + cc] : 0.0;
    }
    double sacc = 0.0;
    if (c < nr) {
      double p8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) p8[u] = 0.0;
      int row = R0 + SB64 + grp;
      for (; row + 28 < N; row += 32) {
        double lv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) lv[u] = K[(size_t)(row + 4 * u) * ld + R0 + c];
#pragma unroll
        for (int u = 0; u < 8; ++u) p8[u] = fma(lv[u], sx[row + 4 * u], p8[u]);
      }
      for (; row < N; row += 4) p8[0] = fma(K[(size_t)row * ld + R0 + c], sx[row], p8[0]);
      sacc = ((p8[0] + p8[1]) + (p8[2] + p8[3])) + ((p8[4] + p8[5]) + (p8[6] + p8[7]));
    }
    part[grp * SB64 + c] = sacc;
    __syncthreads();
    if (tid < SB64) sx[R0 + tid] -= (part[tid] + part[SB64 + tid]) + (part[2 * SB64 + tid] + part[3 * SB64 + tid]);
    __syncthreads();
    if (warp == 0) {
      double x0 = sx[R0 + lane], x1 = sx[R0 + lane + 32];
#pragma unroll 8
      for (int i = SB64 - 1; i >= 0; --i) {
        const double xi = __shfl_sync(0xffffffffu, i < 32 ? x0 : x1, i & 31);
        if (lane < i) x0 -= Ld[i * SP65 + lane] * xi;
        if (lane + 32 < i) x1 -= Ld[i * SP65 + lane + 32] * xi;
      }
      sx[R0 + lane] = x0;
      sx[R0 + lane + 32] = x1;
    }
    __syncthreads();
  }
  for (int i = tid; i < N; i += FT) x[i] = sx[i];
  __syncthreads();
}

// ---- one Newton solve with the current factor ------------------------------------------------------------------
// Normal reduction: condensed solve of the augmented right-hand side rvec = b0|b1,
//   (Hx + M^T W M) dx = b0 + M^T W b1,   dlam = W (M dx - b1),   out (+)= [dx; dlam]     (solver.cu condensed_solve)
__device__ void condensed_solve(const View& v, int p, const double* rvec, int accumulate, double* sm) {
  const Shape& s = v.s;
  const int tid = threadIdx.x;
  const int len = max(s.n, s.m);
  double* sol = v.sol + (size_t)p * v.ssol;
  if (s.m > 0) {
    for (int i = tid; i < len; i += FT) prepare_sol_body(v, p, i, rvec, 0);
    __syncthreads();
    MATVEC(v.MT + (size_t)p * v.sMT, v.ldmt, s.n, s.m, v.tm + (size_t)p * s.ms, v.tn + (size_t)p * s.ns);
    __syncthreads();
  }
  for (int i = tid; i < len; i += FT) prepare_sol_body(v, p, i, rvec, 1);
  __syncthreads();
  {
    FSUB_BEGIN;
    ldlt_solve(v.K + (size_t)p * v.sK, v.ldk, v.Dg + (size_t)p * v.ldk, v.N, sol, sm);
    FSUB_END(8);
  }
  if (s.m > 0) {
    MATVEC(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, sol, v.Mx + (size_t)p * s.ms);
    __syncthreads();
  }
  for (int i = tid; i < len; i += FT) recover_dual_body(v, p, i, rvec, accumulate);
  __syncthreads();
}

template <int MODE>
__device__ void newton_direction(const View& v, int p, int nref, double (*red)[FW], double* sm) {
  const Shape& s = v.s;
  const int tid = threadIdx.x;
  const int len = max(s.n, s.m);
  const double* rhs = v.rhs + (size_t)p * (s.ns + s.ms);
  if (!v.normal) {
    for (int i = tid; i < len; i += FT) prepare_sol_body(v, p, i, v.rhs, 0);
    __syncthreads();
    ldlt_solve(v.K + (size_t)p * v.sK, v.ldk, v.Dg + (size_t)p * v.ldk, v.N, v.sol + (size_t)p * v.ssol, sm);
  } else {
    (void)rhs;
    condensed_solve(v, p, v.rhs, 0, sm);
    double* out = v.out + (size_t)p * (s.ns + s.ms);
    for (int r = 0; r < nref; ++r) {
      MATVEC(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, out, v.Qd + (size_t)p * s.ns);
      if (s.m > 0) {
        MATVEC(v.MT + (size_t)p * v.sMT, v.ldmt, s.n, s.m, out + s.ns, v.tn + (size_t)p * s.ns);
        // first refinement step: out's dx is still the vector the condensed solve just multiplied by M (Mx)
        if (r > 0) MATVEC(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, out, v.Mx + (size_t)p * s.ms);
      }
      __syncthreads();
      for (int i = tid; i < len; i += FT) aug_residual_body(v, p, i);
      __syncthreads();
      condensed_solve(v, p, v.resid, 1, sm);
    }
  }
  double a[1] = {1.0};
  for (int i = tid; i < len; i += FT) backsub_step_body<MODE>(v, p, i, a[0]);
  double tot[1];
  cta_reduce<1, 1u>(a, red, tot);
  if (tid == 0) {
    Scal& sc = v.sc[p];
    const double al = fmin(1.0, tot[0]);
    if (MODE == 0) sc.alpha_aff = al; else sc.alpha = al;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FT, 2) k_ipm_batch(FusedArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[4][FW];
  __shared__ int s_p;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MV_NBUF; ++i) mbar_init(g_mvbar + i, 1);
    g_mvph = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const View& v = a.v;
  const Shape& s = v.s;
  const int tid = threadIdx.x;
  const int len = max(s.n, s.m);
  for (;;) {
    if (tid == 0) s_p = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int p = s_p;
    __syncthreads();
    if (p >= a.count) break;
    Scal& sc = v.sc[p];
    if (a.ready) {
      __shared__ int s_ok;
      if (tid == 0) {
        const long long t0 = clock64();
        int ok = 1;
        while (*a.ready <= p) {
          __nanosleep(2000);
          if (clock64() - t0 > (8ll << 30)) { ok = 0; break; }  // ~4 s: the upload never arrived
        }
        if (!ok) atomicExch(a.abort_flag, 1);
        s_ok = ok;
        __threadfence();
      }
      __syncthreads();
      if (!s_ok) {
        if (tid == 0) { sc.iters = 0; sc.done = 3; }
        __syncthreads();
        continue;
      }
      // initial point (EnvironmentBuilder.cpp:34-73) and M^T, as ipmz_batch_upload's kernels would have left them
      for (int i = tid; i < len; i += FT) initial_point_body(v, p, i);
      if (s.m > 0) {
        const double* M = v.M + (size_t)p * v.sM;
        double* MT = a.MT_w + (size_t)p * v.sMT;
        const int lane = tid & 31, warp = tid >> 5;
        double* tile = sm + warp * (32 * 33);
        const int tr = (s.m + 31) >> 5, tc = (s.n + 31) >> 5;
        for (int t = warp; t < tr * tc; t += FW) {
          const int r0 = (t / tc) * 32, c0 = (t % tc) * 32;
          for (int i = 0; i < 32; ++i)
            tile[i * 33 + lane] = (r0 + i < s.m && c0 + lane < s.n) ? M[(size_t)(r0 + i) * v.ldm + c0 + lane] : 0.0;
          __syncwarp();
          for (int i = 0; i < 32; ++i)
            if (c0 + i < s.n && r0 + lane < s.m) MT[(size_t)(c0 + i) * v.ldmt + r0 + lane] = tile[lane * 33 + i];
          __syncwarp();
        }
      }
      __syncthreads();
    }
    if (tid == 0) {  // a fresh solve restarts the counters and keeps the iterate (warm start, as ipmz_solve)
      sc.iters = 0; sc.done = 0; sc.mu_c = 0.0; sc.alpha = 0.0; sc.alpha_aff = 0.0; sc.sigma = 0.0;
    }
    double* V = v.V + (size_t)p * v.sp;
#ifdef IPMZ_FUSED_DBG_MODES
    if (a.dbg == 1) {
      for (int rep = 0; rep < 30; ++rep) {
        MATVEC(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, V, v.Qx + (size_t)p * s.ns);
        MATVEC(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, V, v.Mx + (size_t)p * s.ms);
        MATVEC(v.MT + (size_t)p * v.sMT, v.ldmt, s.n, s.m, V + (size_t)N_NSLOTS * s.ns, v.MTl + (size_t)p * s.ns);
      }
      __syncthreads();
      continue;
    }
    if (a.dbg >= 2) {
      MATVEC(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, V, v.Qx + (size_t)p * s.ns);
      MATVEC(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, V, v.Mx + (size_t)p * s.ms);
      __syncthreads();
      {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < len; i += FT) residuals_rhs_body<0>(v, p, i, acc);
        __syncthreads();
      }
      for (int rep = 0; rep < 20; ++rep) {
        assemble_normal(v, p, sm);
        if (a.dbg >= 3) ldlt_panels(v.K + (size_t)p * v.sK, v.ldk, v.Dg + (size_t)p * v.ldk, v.N, sm);
        if (a.dbg >= 4) ldlt_solve(v.K + (size_t)p * v.sK, v.ldk, v.Dg + (size_t)p * v.ldk, v.N, v.sol + (size_t)p * v.ssol, sm);
      }
      __syncthreads();
      continue;
    }
#endif
    FPH_DECL;
    for (;;) {
      // ---- Q x, M x, M^T lambda
      MATVEC(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, V, v.Qx + (size_t)p * s.ns);
      if (s.m > 0) {
        MATVEC(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, V, v.Mx + (size_t)p * s.ms);
        MATVEC(v.MT + (size_t)p * v.sMT, v.ldmt, s.n, s.m, V + (size_t)N_NSLOTS * s.ns, v.MTl + (size_t)p * s.ns);
      }
      __syncthreads();
      FPH(0);
      // ---- residuals, W, objective / res / mu, stopping test, predictor right-hand side
      {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < len; i += FT) residuals_rhs_body<0>(v, p, i, acc);
        double tot[4];
        cta_reduce<4, 0u>(acc, red, tot);
        if (tid == 0) residuals_finish(v, p, tot);
        __syncthreads();
      }
      FPH(1);
      if (sc.done != 0) break;
      int nref = 0;
      if (v.normal) nref = a.refine_fixed >= 0 ? a.refine_fixed : ((s.reg_eq || sc.mu < 1e-3) ? 1 : 0);
      // ---- assembly + factorization
      if (v.normal && s.m > 0) assemble_normal(v, p, sm);
      else assemble_augmented(v, p);
      FPH(2);
      ldlt_panels(v.K + (size_t)p * v.sK, v.ldk, v.Dg + (size_t)p * v.ldk, v.N, sm);
      FPH(3);
      // ---- predictor
      newton_direction<0>(v, p, nref, red, sm);
      FPH(4);
      {
        double m1[1] = {0.0};
        for (int i = tid; i < len; i += FT) mu_affine_body(v, p, i, m1[0]);
        double tot[1];
        cta_reduce<1, 0u>(m1, red, tot);
        if (tid == 0) mu_affine_finish(v, p, tot[0]);
        __syncthreads();
      }
      // ---- corrector
      {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < len; i += FT) residuals_rhs_body<1>(v, p, i, acc);
        __syncthreads();
      }
      FPH(5);
      newton_direction<1>(v, p, nref, red, sm);
      FPH(6);
      // ---- v += 0.995 alpha dv (Optimizer.cpp:216-238)
      {
        const double st = v.ftb * sc.alpha;
        const double* D = v.D + (size_t)p * v.sp;
        for (size_t i = tid; i < v.sp; i += FT) V[i] = V[i] + st * D[i];
        __syncthreads();
        if (tid == 0) sc.iters += 1;
      }
      FPH(7);
    }
    __syncthreads();
  }
}

int fused_smem_doubles(const View& v) {
  const Shape& s = v.s;
  const int rows_cap = (v.N + 15) & ~15;
  const int ldlt = rows_cap * PP + 32 + 32 + CBUF + INV_SUB;
  const int syrk = FSTAGES * STAGE_DOUBLES + s.ms + s.ns;
  const int nblk = (v.N + SB64 - 1) / SB64;
  const int solve = nblk * SB64 + SB64 * SP65 + 4 * SB64;
  int m = ldlt > syrk ? ldlt : syrk;
  if (solve > m) m = solve;
  if (MV_NBUF * MV_CHUNK > m) m = MV_NBUF * MV_CHUNK;
  if (FW * 32 * 33 > m) m = FW * 32 * 33;  // one 32 x 33 transpose tile per warp (streamed mode)
  return m;
}

}  // namespace

// Does the persistent one-CTA-per-problem kernel cover this workspace?  AUGMENTED or NORMAL reduction, LDL^T
// (no Bunch-Kaufman rows), and a panel that fits the shared memory of one CTA.
bool fused_batch_applicable(const View& v) {
  if (v.full || v.s.hard_eq) return false;
  if (v.N < 1 || v.N > 512) return false;  // cta_matvec_tma keeps a lane's part of x in 8 double2 registers
  return (size_t)fused_smem_doubles(v) * sizeof(double) <= 200 * 1024;
}

// debug: read and clear the per-phase cycle counters (zeros unless built with -DIPMZ_FUSED_CLOCKS)
int fused_read_clocks(unsigned long long* out16) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_fused_clk, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) return (int)e;
  unsigned long long z[16] = {0};
  return (int)cudaMemcpyToSymbol(g_fused_clk, z, sizeof(z));
}

int fused_batch_init() {
  return (int)cudaFuncSetAttribute(k_ipm_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

// One launch: every problem of the batch from its current iterate to convergence.  `ticket` is a device int the
// launcher resets on the stream.  Returns a cudaError_t.
int launch_ipm_batch(cudaStream_t st, const View& v, int count, int refine_fixed, int* ticket, const int* ready,
                     int* abort_flag) {
  FusedArgs a;
  a.ready = ready;
  a.abort_flag = abort_flag;
  a.MT_w = const_cast<double*>(v.MT);
  a.v = v;
  a.v.active = nullptr;
  a.count = count;
  a.ticket = ticket;
  a.refine_fixed = refine_fixed;
  a.smem_doubles = fused_smem_doubles(v);
  a.dbg = getenv("IPMZ_FUSED_DBG") ? atoi(getenv("IPMZ_FUSED_DBG")) : 0;
  const size_t smem = (size_t)a.smem_doubles * sizeof(double);
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, nsm = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ipm_batch, FT, smem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) per_sm = 1;
  if (const char* e = getenv("IPMZ_FUSED_CTAS_PER_SM")) {  // experiments: fewer problems in flight = smaller L2 footprint
    const int want = atoi(e);
    if (want >= 1 && want < per_sm) per_sm = want;
  }
  int grid = nsm * per_sm;
  if (grid > count) grid = count;
  k_ipm_batch<<<grid, FT, smem, st>>>(a);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace ipmz
