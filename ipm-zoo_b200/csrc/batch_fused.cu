// ipm-zoo_b200/csrc/batch_fused.cu -- batches of small QPs (cfg4): the WHOLE Mehrotra predictor-corrector solve of
// every problem of a batch inside one persistent kernel, one launch per batch.
//
// Reference control flow per problem: Optimizer::solve_quasi_definite_ (Optimizer.cpp:77-220) with
// LinearSolvers::ldlt_decomposition / overwriting_solve_ldlt (LinearSolvers.cpp:14-74).  The multi-kernel batched
// schedule (solver.cu: ~110 launches per iteration, a host synchronisation per iteration to read the stopping test,
// every matrix pass re-streamed from HBM by a fresh grid) is replaced for systems that fit by:
//
//   * a persistent grid (2 CTAs per SM); a CTA takes a problem from a device-side queue and runs one complete Mehrotra
//     iteration of it -- matvecs, residuals, stopping test, assembly, LDL^T, both Newton solves (with the normal
//     reduction's iterative refinement), centring, step length and update -- then hands the problem back, so there is
//     no host round trip and no wave quantisation: the batch advances as one front (fewest iterations first), problems
//     that converge leave the queue (see "work distribution" below);
//   * two co-resident CTAs per SM are in different phases most of the time, so the latency-bound chains of one problem
//     (the one-warp 32 x 32 LDL^T of a diagonal block, the triangular sweeps, block reductions) overlap the DMMA-bound
//     phases of the other (condensed assembly M^T W M, trailing updates);
//   * the matrix K / its factor L of a problem is stored tile-major (64 x 64 tiles, each contiguous) and touched by one
//     SM per iteration; Q and M are read-only; the transposed copy M^T the grid-per-phase kernels use is never read.
//
// Phases of one iteration (all vector formulas are the bodies of vector_bodies.cuh, shared with the grid-per-phase
// kernels):
//   matvecs            Q x, M x, M^T lambda: rows staged by bulk copies (cp.async.bulk + mbarrier), 16 lanes per row;
//                      M^T v accumulated by columns from the same staged rows (M x and M^T lambda are one pass)
//   residuals          r_*, W, objective, ||res||, mu, stopping test, predictor right-hand side
//   assembly           NORMAL: K = Q + Y^-1 L_y + Z^-1 L_z + M^T W M on the FP64 tensor pipe (64 x 64 tiles, accumulators
//                      start from Q, 16-row slices of M through a 3-stage cp.async ring that runs across tile
//                      boundaries, fragments read down the slice columns, the W scaling folded into the B fragment);
//                      AUGMENTED: a copy pass
//   factorization      right-looking LDL^T, 32-wide panels: panel in shared memory, diagonal block by one warp
//                      (warp_ldlt32), rows below on the tensor pipe with the 8 x 8 inverse blocks (panel_solve32),
//                      trailing matrix updated in L2 with both operands read from the shared-memory panel
//   solves             forward / pivots / backward over the factor's 64 x 64 tiles (one bulk copy each, 2-stage
//                      mbarrier ring), x in shared memory, diagonal tiles by 4-column substitution blocks of one warp
//   back-substitution  eliminated Delta's, ratio test, centring parameter, corrector right-hand side, update
#include <cuda_runtime.h>
#include <algorithm>
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
#ifndef IPMZ_FUSED_CTAS
#define IPMZ_FUSED_CTAS 2        // resident CTAs per SM the kernel is compiled for (register budget, panel pitch)
#endif
#if IPMZ_FUSED_CTAS >= 3
constexpr int PP = 34;           // three CTAs per SM leave 75 KB each: the 256 x 32 panel fits with pitch 34 (fragment loads 2-way)
#else
constexpr int PP = 36;           // panel pitch: 36 mod 16 == 4 -> conflict-free DMMA fragment loads
#endif
constexpr int TS = 64;           // assembly tile
#ifndef IPMZ_FSTAGES
#define IPMZ_FSTAGES 3
#endif
constexpr int FSTAGES = IPMZ_FSTAGES;
constexpr int PA = TS + 4;      // pitch of a k-slice row (68 mod 16 == 4: conflict-free DMMA fragment loads down a column)
constexpr int STAGE_DOUBLES = 2 * BK * PA;  // A-side + B-side columns of one 16-row slice of M
constexpr int SB64 = 64;  // block of the triangular sweeps
// The reduced matrix / its factor live TILE-MAJOR in global memory in the fused path: the lower triangle as 64 x 64 tiles
// (tile (tr, tc), tc <= tr, at index tr (tr + 1) / 2 + tc), each tile stored with the pitch it has in shared memory.  A
// triangular sweep then fetches a tile with ONE bulk copy (35 KB contiguous; 64 row copies of 512 B each were bound by
// the copy engine's per-instruction cost: 4500 cycles per tile), and the footprint of K shrinks from N ldk to about
// 0.55 N^2 doubles.  Rows / columns past N and the pitch padding are never written and stay zero from the allocation.
constexpr int TP = 70;               // tile pitch: rows 16-byte aligned; 70 * 8 B = 140 banks: the row-per-lane 16-byte loads of
                                     // the diagonal chains are conflict-free (pitch 68: 8-way), the 4-lanes-per-row products 2-way
constexpr int TILE_DOUBLES = SB64 * TP;
__host__ __device__ __forceinline__ size_t ktile_base(int tr, int tc) { return (size_t)(tr * (tr + 1) / 2 + tc) * TILE_DOUBLES; }
__device__ __forceinline__ size_t kidx(int row, int col) {
  return ktile_base(row >> 6, col >> 6) + (size_t)((row & 63) * TP + (col & 63));
}

struct FusedArgs {
  View v;
  int count;
  int* ticket;       // control words (see "work distribution"): [0] next fresh problem (ticket), [3] finished, bucket heads / tails
  int* queue;        // slots of the per-iteration-count buckets (nullptr: one CTA keeps a problem to the end)
  int refine_fixed;  // >= 0: that many refinement steps per condensed solve; -1: by the problem's mu (solver.cu policy)
  int smem_doubles;
  // streamed solve (ipmz_batch_solve_streamed): the kernel is launched BEFORE the problem data is uploaded; `ready` is a
  // device word the copy stream overwrites (4-byte H2D copy, stream-ordered after each chunk of problems) with the number
  // of problems whose data is resident.  A CTA holding ticket p waits until *ready > p, then builds the reference's
  // initial point itself (the upload path's k_initial_point could not be
  // scheduled while this persistent grid owns every SM).  nullptr: everything is resident already.
  const volatile int* ready;
  int* abort_flag;   // set when a wait on `ready` times out (the host returns an error instead of hanging the GPU)
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

// Matrix-vector product: the matrix (rows contiguous, pitch lda) streams through
// two 32 KB shared-memory buffers by bulk copies (cp.async.bulk.shared.global, SASS UBLKCP): ONE instruction of one
// thread moves a whole chunk and completes on an mbarrier by transaction bytes.  Work split inside a chunk: S = 2^k
// lanes share one row (S >= cols / 16, so a lane's part of x is 8 double2 registers loaded once per call), FT / S rows
// per chunk; lane `seg` of a row takes the column pairs seg + S j (conflict-free reads, 8 independent 16-byte loads in
// flight per thread) and a row is finished by log2(S) shuffle steps.
constexpr int MV_NBUF = 2;
constexpr int MV_CHUNK = 4096;  // doubles per staging buffer
__shared__ __align__(8) unsigned long long g_mvbar[MV_NBUF];  // one mbarrier per staging buffer, initialised by the kernel
__shared__ unsigned g_mvph;                                    // their phase bits (uniform per CTA)
// L2 eviction priorities (createpolicy): the working set of the problems in flight (Q, M and K of 296 problems: 370 MB)
// is three times the L2.  M is re-read 7 to 13 times per iteration, Q 1.5 to 3.5 times: M's lines are kept
// (evict_last), Q's are marked evict_first so that streaming Q does not push M and the factor out.
// IPMZ_L2_HINTS=0 (compile time) builds without the hints for A/B runs.
#ifndef IPMZ_L2_HINTS
#define IPMZ_L2_HINTS 0
#endif
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
enum { L2_NORMAL = 0, L2_KEEP = 1, L2_STREAM = 2 };
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar,
                                          int hint = L2_NORMAL) {
  const unsigned da = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  // earlier generic-proxy accesses of the buffer (other phases use the same shared memory) before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;\n" ::"r"(ba), "r"(bytes) : "memory");
#if IPMZ_L2_HINTS
  if (hint != L2_NORMAL) {
    const unsigned long long pol = hint == L2_KEEP ? l2_policy_evict_last() : l2_policy_evict_first();
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(da), "l"(gsrc), "r"(bytes), "r"(ba), "l"(pol) : "memory");
    return;
  }
#endif
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(da), "l"(gsrc), "r"(bytes), "r"(ba) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(void* smem_dst, const void* gsrc, int src_bytes, unsigned long long pol) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
#if IPMZ_L2_HINTS
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(sa), "l"(gsrc), "r"(src_bytes), "l"(pol));
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(src_bytes));
#endif
}
__device__ __forceinline__ double2 ld_stream2(const double* p) {  // 16-byte load, L2 evict_first
#if IPMZ_L2_HINTS
  double2 r;
  const unsigned long long pol = l2_policy_evict_first();
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;\n" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
  return r;
#else
  return *reinterpret_cast<const double2*>(p);
#endif
}
// bar: MV_NBUF mbarriers (count 1) initialised once per kernel; ph: their phase bits, carried by the caller.
// ROWDOT: y[r] = dot(A[r][0:cols], x) (above).  COLACC: yt[j] = sum_r A[r][j] vt[r] from the SAME staged chunk -- thread j
// owns column j (and j + FT), four accumulators by r mod 4, conflict-free reads -- so M^T v needs no transposed copy of
// M and the opening M x / M^T lambda of an iteration are ONE pass over M.
template <bool ROWDOT, bool COLACC>
__device__ __noinline__ void cta_matvec_tma(const double* __restrict__ A, int lda, int rows, int cols, const double* x,
                                            double* y, const double* vt, double* yt, double* sm, int hint) {
  const int tid = threadIdx.x;
  unsigned long long* bar = g_mvbar;
  const int c2 = (cols + 1) >> 1;
  int S = 4;
  while (8 * S < c2) S <<= 1;  // cols <= 512 -> S <= 32
  int R = ROWDOT ? FT / S : MV_CHUNK / lda;  // rows per chunk
  if (R * lda > MV_CHUNK) R = MV_CHUNK / lda;
  if (COLACC) R &= ~3;  // chunks start at multiples of four rows (16-byte reads of the staged vector); R >= 8
  const int nch = (rows + R - 1) / R;
  const int rr = tid / S, seg = tid - rr * S;
  const bool exact = c2 == 8 * S;  // every lane's eight column pairs exist (cfg4: 256 columns, 16 lanes per row)
  double* svt = sm + MV_NBUF * MV_CHUNK;  // COLACC: vt staged in shared memory (broadcast reads in the inner loop)
  __syncthreads();  // the staging buffers are free and x is visible
  unsigned ph = g_mvph;
  if (tid == 0)
    for (int c = 0; c < MV_NBUF && c < nch; ++c)
      bulk_load(sm + c * MV_CHUNK, A + (size_t)c * R * lda, (unsigned)(min(R, rows - c * R) * lda) * 8u, bar + c, hint);
  double2 xr[8];
  if (ROWDOT) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = seg + S * j;
      xr[j] = (exact || k < c2) ? reinterpret_cast<const double2*>(x)[k] : make_double2(0.0, 0.0);
    }
  }
  if (COLACC) {
    for (int i = tid; i < ((rows + 3) & ~3); i += FT) svt[i] = i < rows ? vt[i] : 0.0;
    __syncthreads();
  }
  double ca[4] = {0.0, 0.0, 0.0, 0.0}, cb[4] = {0.0, 0.0, 0.0, 0.0};
  const bool col_a = tid < cols, col_b = tid + FT < cols;
  for (int c = 0; c < nch; ++c) {
    const int bi = c % MV_NBUF;
    mbar_wait(bar + bi, (ph >> bi) & 1u);  // (one polling warp + a barrier instead: no gain, 68.5 vs 68.1 ms)
    ph ^= 1u << bi;
    const int r0 = c * R, nr = min(R, rows - r0);
    const double* buf = sm + bi * MV_CHUNK;
    if (ROWDOT) {
      const double2* a2 = reinterpret_cast<const double2*>(buf + (size_t)(rr < nr ? rr : 0) * lda) + seg;
      double2 av[8];
      if (exact) {  // uniform branch: no zero-initialisation / predicates on the common path
#pragma unroll
        for (int j = 0; j < 8; ++j) av[j] = a2[S * j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) av[j] = seg + S * j < c2 ? a2[S * j] : make_double2(0.0, 0.0);
      }
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        acc0 = fma(av[j].x, xr[j].x, acc0); acc1 = fma(av[j].y, xr[j].y, acc1);
        acc2 = fma(av[j + 1].x, xr[j + 1].x, acc2); acc3 = fma(av[j + 1].y, xr[j + 1].y, acc3);
      }
      double t = (acc0 + acc1) + (acc2 + acc3);
      if (S >= 32) t += __shfl_xor_sync(0xffffffffu, t, 16);
      if (S >= 16) t += __shfl_xor_sync(0xffffffffu, t, 8);
      if (S >= 8) t += __shfl_xor_sync(0xffffffffu, t, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      if (seg == 0 && rr < nr) y[r0 + rr] = t;
    }
    if (COLACC && col_a) {
      // accumulator u takes the rows r0 + r + u of every group of four (static indexing: a dynamically indexed
      // accumulator array would live in local memory); rows past `rows` are zeros in svt and stale (finite or not,
      // times zero is avoided by the bound) data in the buffer
      const double* sa = buf + tid;
      const double* sv = svt + r0;
      const int n4 = nr & ~3;
      for (int r = 0; r < n4; r += 4) {
        const double2 v01 = *reinterpret_cast<const double2*>(sv + r), v23 = *reinterpret_cast<const double2*>(sv + r + 2);
        ca[0] = fma(sa[(size_t)r * lda], v01.x, ca[0]);
        ca[1] = fma(sa[(size_t)(r + 1) * lda], v01.y, ca[1]);
        ca[2] = fma(sa[(size_t)(r + 2) * lda], v23.x, ca[2]);
        ca[3] = fma(sa[(size_t)(r + 3) * lda], v23.y, ca[3]);
        if (col_b) {
          cb[0] = fma(sa[(size_t)r * lda + FT], v01.x, cb[0]);
          cb[1] = fma(sa[(size_t)(r + 1) * lda + FT], v01.y, cb[1]);
          cb[2] = fma(sa[(size_t)(r + 2) * lda + FT], v23.x, cb[2]);
          cb[3] = fma(sa[(size_t)(r + 3) * lda + FT], v23.y, cb[3]);
        }
      }
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (n4 + u < nr) {
          ca[u] = fma(sa[(size_t)(n4 + u) * lda], sv[n4 + u], ca[u]);
          if (col_b) cb[u] = fma(sa[(size_t)(n4 + u) * lda + FT], sv[n4 + u], cb[u]);
        }
    }
    __syncthreads();
    if (tid == 0 && c + MV_NBUF < nch)
      bulk_load(sm + bi * MV_CHUNK, A + (size_t)(c + MV_NBUF) * R * lda,
                (unsigned)(min(R, rows - (c + MV_NBUF) * R) * lda) * 8u, bar + bi, hint);
  }
  if (COLACC) {
    if (col_a) yt[tid] = (ca[0] + ca[1]) + (ca[2] + ca[3]);
    if (col_b) yt[tid + FT] = (cb[0] + cb[1]) + (cb[2] + cb[3]);
  }
  if (tid == 0) g_mvph = ph;  // read by the next call after its opening barrier
}
#define MATVEC(A_, lda_, rows_, cols_, x_, y_) cta_matvec_tma<true, false>(A_, lda_, rows_, cols_, x_, y_, nullptr, nullptr, sm, L2_KEEP)
// the same for Q (streamed: evict_first)
#define MATVEC_Q(A_, lda_, rows_, cols_, x_, y_) cta_matvec_tma<true, false>(A_, lda_, rows_, cols_, x_, y_, nullptr, nullptr, sm, L2_STREAM)
// yt = A^T vt (A stored by rows)
#define MATVEC_T(A_, lda_, rows_, cols_, vt_, yt_) cta_matvec_tma<false, true>(A_, lda_, rows_, cols_, nullptr, nullptr, vt_, yt_, sm, L2_KEEP)
// y = A x and yt = A^T vt in one pass over A
#define MATVEC_BOTH(A_, lda_, rows_, cols_, x_, y_, vt_, yt_) cta_matvec_tma<true, true>(A_, lda_, rows_, cols_, x_, y_, vt_, yt_, sm, L2_KEEP)

// ---- assembly ---------------------------------------------------------------------------------------------------
// NORMAL: K(lower) = Q + diag(hd) + M^T diag(w) M, hd = Y^-1 L_y + Z^-1 L_z.  sm: FSTAGES stages | w[ms] | hd[ns].
// The operands are 16-row slices of M itself (rows = constraints = the contraction index): a stage holds the 64 columns
// of the tile's row block and the 64 columns of its column block, [k][i] with pitch 68, and the DMMA fragments are read
// down the columns -- no transposed copy of M exists in the fused path.
__device__ void assemble_normal(const View& v, int p, double* sm) {
  const Shape& s = v.s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps, warp tile 32 x 16
  double* wsm = sm + FSTAGES * STAGE_DOUBLES;
  double* hd = wsm + s.ms;
  const double* V = v.V + (size_t)p * v.sp;
  const double* Q = v.Q + (size_t)p * v.sQ;
  const double* M = v.M + (size_t)p * v.sM;
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

  // load cursor (runs FSTAGES-1 steps ahead of the compute cursor, across tile boundaries).  A thread's four 16-byte
  // chunks of a stage: rows lrow and lrow + 8 of the slice, A side (tile row block) and B side (tile column block).
  // The source addresses are two running pointers that advance by 16 rows of M per step and are rebuilt ten times per
  // assembly (the address arithmetic per chunk used to be 38 % of this function's instructions).
  int l_ti = 0, l_tj = 0, l_kt = 0;
  const int lrow = tid >> 5, lck = (tid & 31) * 2;
  const int ldm = v.ldm, m = s.m, ns = s.ns, n = s.n;
  const unsigned sm_u = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(lrow * PA + lck) * 8u;
  const size_t step8 = (size_t)8 * ldm, step16 = (size_t)16 * ldm;
  const double* pa = M;
  const double* pb = M;
  bool cola = false, colb = false;
  auto tile_ptrs = [&]() {  // M's padding columns [n, ns) are zero, columns past ns do not exist
    const int ca = l_ti * TS + lck, cb = l_tj * TS + lck;
    cola = ca < ns; colb = cb < ns;
    pa = M + (size_t)lrow * ldm + (cola ? ca : 0);
    pb = M + (size_t)lrow * ldm + (colb ? cb : 0);
  };
  tile_ptrs();
  auto load_next = [&](int stage) {
    const unsigned dst = sm_u + (unsigned)(stage * STAGE_DOUBLES) * 8u;
    const int k0 = l_kt * BK + lrow;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool okk = k0 + 8 * h < m;
      const bool oka = okk && cola, okb = okk && colb;
      const double* sa = oka ? pa + (h ? step8 : 0) : M;
      const double* sb = okb ? pb + (h ? step8 : 0) : M;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst + (unsigned)(8 * h * PA) * 8u), "l"(sa), "r"(oka ? 16 : 0));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst + (unsigned)((BK + 8 * h) * PA) * 8u), "l"(sb), "r"(okb ? 16 : 0));
    }
    pa += step16; pb += step16;
    if (++l_kt == KT) {
      l_kt = 0;
      if (++l_tj > l_ti) { l_tj = 0; ++l_ti; }
      tile_ptrs();
    }
  };
  __syncthreads();  // wsm, hd visible; the stage buffers are free (previous phase done)
  int loaded = 0, lstage = 0;  // lstage = loaded % FSTAGES, kept incrementally
  for (; loaded < FSTAGES - 1; ++loaded) {
    if (loaded < nsteps) load_next(lstage);
    if (++lstage == FSTAGES) lstage = 0;
    cp_async_commit();
  }
  int stage = 0;
  for (int ti = 0; ti < nt; ++ti)
    for (int tj = 0; tj <= ti; ++tj) {
      const int row0 = ti * TS + wm * 32 + g, col0 = tj * TS + wn * 16 + 2 * q;
      double acc[4][2][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int row = row0 + mi * 8;
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const int col = col0 + ni * 8;
          double2 c = make_double2(0.0, 0.0);
          if (row < n && col < ns) c = *reinterpret_cast<const double2*>(Q + (size_t)row * v.ldq + col);
          if (row == col) c.x += hd[row < n ? row : 0];
          if (row == col + 1) c.y += hd[row < n ? row : 0];
          acc[mi][ni][0] = c.x; acc[mi][ni][1] = c.y;
        }
      }
#pragma unroll 1
      for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<FSTAGES - 2>();
        __syncthreads();
        if (loaded < nsteps) load_next(lstage);
        if (++lstage == FSTAGES) lstage = 0;
        cp_async_commit();
        ++loaded;
        const double* As = sm + stage * STAGE_DOUBLES;
        if (++stage == FSTAGES) stage = 0;
        const double* Aw = As + q * PA + wm * 32 + g;
        const double* Bw = As + (BK + q) * PA + wn * 16 + g;
        const double* wk = wsm + kt * BK + q;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
          double af[4], bf[2];
          const double w = wk[kk * 4];
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) af[mi] = Aw[kk * 4 * PA + mi * 8];
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) bf[ni] = Bw[kk * 4 * PA + ni * 8] * w;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
        }
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int row = row0 + mi * 8;
        if (row >= n) continue;
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const int col = col0 + ni * 8;
          if (col > row) continue;
          double* dst = K + kidx(row, col);
          if (col + 1 <= row) *reinterpret_cast<double2*>(dst) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          else dst[0] = acc[mi][ni][0];
        }
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
    if (r < s.n) {
      const double* q = v.Q + (size_t)p * v.sQ + (size_t)r * v.ldq;
      for (int c = lane; c < r; c += 32) K[kidx(r, c)] = q[c];
      if (lane == 0) {
        double dii = q[r];
        if (s.ylo) dii = dii + inv_guard(nslot(V, s, YS)[r]) * nslot(V, s, LAMY)[r];
        if (s.zup) dii = dii + inv_guard(nslot(V, s, ZS)[r]) * nslot(V, s, LAMZ)[r];
        K[kidx(r, r)] = dii;
      }
    } else {
      const int j = r - s.n;
      const double* mr = v.M + (size_t)p * v.sM + (size_t)j * v.ldm;
      for (int c = lane; c < s.n; c += 32) K[kidx(r, c)] = mr[c];
      for (int c = s.n + lane; c < r; c += 32) K[kidx(r, c)] = 0.0;
      if (lane == 0) K[kidx(r, r)] = -v.winv[(size_t)p * s.ms + j];
    }
  }
  __syncthreads();
}

// ---- factorization ----------------------------------------------------------------------------------------------
// In place on the lower triangle of K (N x N, tile-major: kidx): strict lower = L, pivots -> Dg.
// sm: panel [round16(N) x PP] | dsm[32] | dinv[32] | colbuf[CBUF] | binv[INV_SUB].
__device__ void ldlt_panels(double* K, double* Dg, int N, double* sm) {
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
    // ---- panel -> shared memory (rows beyond R and columns beyond jb as zeros; the diagonal block's upper part too).
    // A thread keeps its chunk column and walks down the rows 16 at a time: the tile-major row address is one
    // kidx() per row visited, the diagonal block's triangle test only touches the first two trips.
    {
      const int c = (tid & 15) * 2;
      for (int r = tid >> 4; r < R16; r += FT / 16) {
        int bytes = 0;
        if (r < R && c < jb && !(r < jb && c > r)) bytes = (jb - c >= 2) ? 16 : 8;
        cp_async16(P + r * PP + c, K + kidx(j0 + (bytes ? r : 0), j0) + (bytes ? c : 0), bytes);
      }
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
    // ---- L and the pivots back to global memory: the strict lower triangle of the diagonal block element-wise, the
    // rows below it as 16-byte stores (jb == 32 whenever rows below exist)
    for (int idx = tid; idx < jb * SB; idx += FT) {
      const int r = idx >> 5, c = idx & 31;
      if (c < r) K[kidx(j0 + r, j0) + c] = P[r * PP + c];
    }
    {
      const int c = (tid & 15) * 2;
      for (int r = jb + (tid >> 4); r < R; r += FT / 16)
        *reinterpret_cast<double2*>(K + kidx(j0 + r, j0) + c) = *reinterpret_cast<const double2*>(P + r * PP + c);
    }
    if (tid < jb) Dg[j0 + tid] = dsm[tid];
    // ---- trailing update  C -= L_panel diag(d) L_panel^T  on the lower triangle, 16 x 16 tiles per warp
    if (rem > 0) {
      const double* T = P + SB * PP;
      const int t0 = j0 + SB;  // first row / column of the trailing matrix
      const int tm = (rem + 15) >> 4;
      const int ntask = tm * (tm + 1) / 2;
      // the C values of a warp's NEXT tile are requested before the DMMAs of the current one (the trailing matrix
      // lives in L2: a tile loaded on demand would expose one L2 round trip per 32 DMMAs)
      // tiles are numbered row by row over the lower triangle; a warp visits t = warp, warp + FW, ...: the (mi, ni)
      // of the next one follows from the current by walking FW positions (no sqrt decode per tile)
      auto tile_step = [&](int& mi, int& ni, int by) {
        ni += by;
        while (ni > mi) { ni -= mi + 1; ++mi; }
      };
      // a 16 x 16 tile of the trailing matrix (t0 is a multiple of 32) lies inside one 64 x 64 storage tile: one tile
      // base per task, constant per-lane offsets
      const int lane_off = g * TP + 2 * q;
      auto tile_ptr = [&](int mi, int ni) -> size_t { return kidx(t0 + mi * 16, t0 + ni * 16) + lane_off; };
      auto load_c = [&](int mi, int ni, double2 (&cv)[2][2]) {
        const double* base = K + tile_ptr(mi, ni);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = mi * 16 + i * 8 + g, col = ni * 16 + j * 8 + 2 * q;
            cv[i][j] = make_double2(0.0, 0.0);
            if (row < rem && col <= row) {
              const double* src = base + i * 8 * TP + j * 8;
              if (col + 1 <= row) cv[i][j] = *reinterpret_cast<const double2*>(src);
              else cv[i][j].x = src[0];
            }
          }
      };
      double2 cvn[2][2];
      int mi_n = 0, ni_n = 0;
      tile_step(mi_n, ni_n, warp);
      if (warp < ntask) load_c(mi_n, ni_n, cvn);
      for (int t = warp; t < ntask; t += FW) {
        const int mi = mi_n, ni = ni_n;
        // accumulators start from C and take the products with the sign folded into the B fragment (the reference's own
        // order: sum -= L[j][k] * L[i][k] * D[k], LinearSolvers.cpp:32-34); k in two halves of 16 keeps the fragment
        // registers at 16 doubles (a full set of 32 next to the prefetched tile spilled the accumulators)
        double acc[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) { acc[i][j][0] = cvn[i][j].x; acc[i][j][1] = cvn[i][j].y; }
        if (t + FW < ntask) { tile_step(mi_n, ni_n, FW); load_c(mi_n, ni_n, cvn); }
        const int ra = mi * 16 + g, rb = ni * 16 + g;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          double af[4][2], bf[4][2];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int kk = (kh * 4 + k4) * 4 + q;
            const double dk = -dsm[kk];
            af[k4][0] = T[ra * PP + kk];
            af[k4][1] = T[(ra + 8) * PP + kk];
            bf[k4][0] = T[rb * PP + kk] * dk;
            bf[k4][1] = T[(rb + 8) * PP + kk] * dk;
          }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            dmma884(acc[0][0], af[k4][0], bf[k4][0]);
            dmma884(acc[0][1], af[k4][0], bf[k4][1]);
            dmma884(acc[1][0], af[k4][1], bf[k4][0]);
            dmma884(acc[1][1], af[k4][1], bf[k4][1]);
          }
        }
        double* obase = K + tile_ptr(mi, ni);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = mi * 16 + i * 8 + g, col = ni * 16 + j * 8 + 2 * q;
            if (row < rem && col <= row) {
              double* dst = obase + i * 8 * TP + j * 8;
              if (col + 1 <= row) *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
              else dst[0] = acc[i][j][0];
            }
          }
      }
    }
    __syncthreads();  // panel buffer free, trailing matrix written
  }
}

// ---- solves -----------------------------------------------------------------------------------------------------
// x <- L^-1 x, x <- x / D, x <- L^-T x with the in-place factor; x (global, length N) is staged in shared memory.
// The factor is consumed as 64 x 64 tiles in ONE fixed sequence -- forward (0,0) (1,0) (1,1) (2,0) ... (b,b), then the
// same tiles in reverse order for the transposed sweep -- through a ring of SOLVE_STAGES shared-memory stages filled by
// one bulk copy per tile (tile-major factor), one tile ahead of the arithmetic across block-row and sweep boundaries (a
// third stage costs L1: 91.0 vs 86.4 ms; the first version loaded every diagonal block and every off-diagonal strip on
// demand: 16 exposed L2 / HBM latencies per solve).  Off-diagonal tiles are 64 x 64 matrix-vector products over all 256 threads (4 lanes per row /
// column, shuffle reduction); a diagonal tile is the substitution chain of one warp, four columns per step.
// sm: sx[nblk * 64] | SOLVE_STAGES tiles [64 x TP].
#ifndef IPMZ_SOLVE_STAGES
#define IPMZ_SOLVE_STAGES 2
#endif
constexpr int SOLVE_STAGES = IPMZ_SOLVE_STAGES;
// position in the tile sequence: forward (0,0) (1,0) (1,1) (2,0) ..., the last tile twice, then the same way back
struct TileCursor {
  int r = 0, c = 0;
  __device__ __forceinline__ void advance(int k, int ntile) {  // from step k to step k + 1
    if (k + 1 < ntile) { if (c < r) ++c; else { ++r; c = 0; } }
    else if (k + 1 > ntile) { if (c > 0) --c; else { --r; c = r; } }
  }
};

__shared__ __align__(8) unsigned long long g_svbar[4];  // one mbarrier per ring stage (SOLVE_STAGES <= 4)
__shared__ unsigned g_svph;                             // their phase bits

// Tile (r, c) of the factor -> stage: ONE bulk copy (cp.async.bulk, 35 KB contiguous in the tile-major layout) that
// completes on the stage's mbarrier by transaction bytes.
__device__ __forceinline__ void solve_tile_issue(double* stage, unsigned long long* bar, const double* K, int r, int c) {
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  const unsigned sa = (unsigned)__cvta_generic_to_shared(stage);
  constexpr unsigned bytes = TILE_DOUBLES * 8u;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;\n" ::"r"(ba), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(sa), "l"(K + ktile_base(r, c)), "r"(bytes), "r"(ba) : "memory");
}

struct TriFwd { double l10, l20, l21, l30, l31, l32; double2 a01, a23, b01, b23; };
__device__ __forceinline__ void tri64_forward_load(TriFwd& k, const double* T, int b, const double* r0, const double* r1) {
  const int c0 = 4 * b;
  const double* d = T + c0 * TP + c0;
  k.l10 = d[TP]; k.l20 = d[2 * TP]; k.l21 = d[2 * TP + 1];
  k.l30 = d[3 * TP]; k.l31 = d[3 * TP + 1]; k.l32 = d[3 * TP + 2];
  k.a01 = *reinterpret_cast<const double2*>(r0 + c0); k.a23 = *reinterpret_cast<const double2*>(r0 + c0 + 2);
  k.b01 = *reinterpret_cast<const double2*>(r1 + c0); k.b23 = *reinterpret_cast<const double2*>(r1 + c0 + 2);
}
template <bool HI>
__device__ __forceinline__ void tri64_forward_block(const TriFwd& k, int b, double& y0, double& y1, int lane) {
  const int c0 = 4 * b;
  const double src = HI ? y1 : y0;
  const double v0 = __shfl_sync(0xffffffffu, src, (c0 + 0) & 31);
  const double v1 = __shfl_sync(0xffffffffu, src, (c0 + 1) & 31);
  const double v2 = __shfl_sync(0xffffffffu, src, (c0 + 2) & 31);
  const double v3 = __shfl_sync(0xffffffffu, src, (c0 + 3) & 31);
  const double x0 = v0;
  const double x1 = fma(-k.l10, x0, v1);
  const double x2 = fma(-k.l21, x1, fma(-k.l20, x0, v2));
  const double x3 = fma(-k.l32, x2, fma(-k.l31, x1, fma(-k.l30, x0, v3)));
  const double xs = (lane & 3) == 0 ? x0 : (lane & 3) == 1 ? x1 : (lane & 3) == 2 ? x2 : x3;  // c0 is a multiple of 4
  if (!HI) {
    const double u0 = fma(-k.a23.y, x3, fma(-k.a23.x, x2, fma(-k.a01.y, x1, fma(-k.a01.x, x0, y0))));
    y0 = lane >= c0 + 4 ? u0 : (lane >= c0 ? xs : y0);
    y1 = fma(-k.b23.y, x3, fma(-k.b23.x, x2, fma(-k.b01.y, x1, fma(-k.b01.x, x0, y1))));
  } else {
    const int rr = lane + 32;
    const double u1 = fma(-k.b23.y, x3, fma(-k.b23.x, x2, fma(-k.b01.y, x1, fma(-k.b01.x, x0, y1))));
    y1 = rr >= c0 + 4 ? u1 : (rr >= c0 ? xs : y1);
  }
}
// Rolled loops on purpose (one warp's dependent chain, run once per tile: straight-line code of 16 blocks is fetched
// cold from the instruction cache every time); the coefficients of block b + 1 are loaded before block b's chain, and
// the in-block rows are selected, not branched (tools/tri_probe.cu: 2500 cycles per tile against 3500 for the
// per-column chain, bitwise the same values; pitch 70 makes the row-per-lane 16-byte loads conflict-free).
__device__ __forceinline__ void tri64_forward(const double* T, double* y, int lane) {
  double y0 = y[lane], y1 = y[lane + 32];
  const double* r0 = T + lane * TP;
  const double* r1 = T + (lane + 32) * TP;
  TriFwd ka, kb;
  tri64_forward_load(ka, T, 0, r0, r1);
#pragma unroll 1
  for (int b = 0; b < 8; b += 2) {
    tri64_forward_load(kb, T, b + 1, r0, r1);
    tri64_forward_block<false>(ka, b, y0, y1, lane);
    tri64_forward_load(ka, T, b + 2, r0, r1);
    tri64_forward_block<false>(kb, b + 1, y0, y1, lane);
  }
#pragma unroll 1
  for (int b = 8; b < 16; b += 2) {
    tri64_forward_load(kb, T, b + 1, r0, r1);
    tri64_forward_block<true>(ka, b, y0, y1, lane);
    tri64_forward_load(ka, T, b + 2 < 16 ? b + 2 : 15, r0, r1);
    tri64_forward_block<true>(kb, b + 1, y0, y1, lane);
  }
  y[lane] = y0;
  y[lane + 32] = y1;
}

// the transposed (unit upper triangular) solve with the same tile: rows 63 .. 0, four at a time
struct TriBwd { double l10, l20, l21, l30, l31, l32, a0, a1, a2, a3, b0, b1, b2, b3; };
__device__ __forceinline__ void tri64_backward_load(TriBwd& k, const double* T, int b, int lane) {
  const int c0 = 4 * b;
  const double* d = T + c0 * TP + c0;
  k.l10 = d[TP]; k.l20 = d[2 * TP]; k.l21 = d[2 * TP + 1];
  k.l30 = d[3 * TP]; k.l31 = d[3 * TP + 1]; k.l32 = d[3 * TP + 2];
  const double* c = T + c0 * TP + lane;
  k.a0 = c[0]; k.a1 = c[TP]; k.a2 = c[2 * TP]; k.a3 = c[3 * TP];
  k.b0 = c[32]; k.b1 = c[TP + 32]; k.b2 = c[2 * TP + 32]; k.b3 = c[3 * TP + 32];  // used for c0 > 32 only
}
template <bool HI>
__device__ __forceinline__ void tri64_backward_block(const TriBwd& k, int b, double& y0, double& y1, int lane) {
  const int c0 = 4 * b;
  const double src = HI ? y1 : y0;
  const double v0 = __shfl_sync(0xffffffffu, src, (c0 + 0) & 31);
  const double v1 = __shfl_sync(0xffffffffu, src, (c0 + 1) & 31);
  const double v2 = __shfl_sync(0xffffffffu, src, (c0 + 2) & 31);
  const double v3 = __shfl_sync(0xffffffffu, src, (c0 + 3) & 31);
  const double x3 = v3;
  const double x2 = fma(-k.l32, x3, v2);
  const double x1 = fma(-k.l21, x2, fma(-k.l31, x3, v1));
  const double x0 = fma(-k.l10, x1, fma(-k.l20, x2, fma(-k.l30, x3, v0)));
  const double xs = (lane & 3) == 0 ? x0 : (lane & 3) == 1 ? x1 : (lane & 3) == 2 ? x2 : x3;
  if (HI) {
    const int rr = lane + 32;
    const double u1 = fma(-k.b0, x0, fma(-k.b1, x1, fma(-k.b2, x2, fma(-k.b3, x3, y1))));
    y1 = rr < c0 ? u1 : (rr < c0 + 4 ? xs : y1);
    y0 = fma(-k.a0, x0, fma(-k.a1, x1, fma(-k.a2, x2, fma(-k.a3, x3, y0))));
  } else {
    const double u0 = fma(-k.a0, x0, fma(-k.a1, x1, fma(-k.a2, x2, fma(-k.a3, x3, y0))));
    y0 = lane < c0 ? u0 : (lane < c0 + 4 ? xs : y0);
  }
}
__device__ __forceinline__ void tri64_backward(const double* T, double* y, int lane) {
  double y0 = y[lane], y1 = y[lane + 32];
  TriBwd ka, kb;
  tri64_backward_load(ka, T, 15, lane);
#pragma unroll 1
  for (int b = 15; b >= 8; b -= 2) {
    tri64_backward_load(kb, T, b - 1, lane);
    tri64_backward_block<true>(ka, b, y0, y1, lane);
    tri64_backward_load(ka, T, b - 2, lane);
    tri64_backward_block<true>(kb, b - 1, y0, y1, lane);
  }
#pragma unroll 1
  for (int b = 7; b >= 0; b -= 2) {
    tri64_backward_load(kb, T, b - 1, lane);
    tri64_backward_block<false>(ka, b, y0, y1, lane);
    tri64_backward_load(ka, T, b - 2 >= 0 ? b - 2 : 0, lane);
    tri64_backward_block<false>(kb, b - 1, y0, y1, lane);
  }
  y[lane] = y0;
  y[lane + 32] = y1;
}

__device__ void ldlt_solve(const double* K, const double* Dg, int N, double* x, double* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (N + SB64 - 1) / SB64;
  const int ntile = nblk * (nblk + 1) / 2, nsteps = 2 * ntile;
  double* sx = sm;
  double* ring = sx + nblk * SB64;
  TileCursor lc, cc;  // load cursor (SOLVE_STAGES - 1 steps ahead) and compute cursor
  int lk = 0;
  __syncthreads();  // the previous phase is done with this shared memory; g_svph of the previous solve is visible
  unsigned ph = g_svph;
  auto issue = [&]() {  // thread 0: tile of step lk -> its stage
    if (lk < nsteps) {
      solve_tile_issue(ring + (lk % SOLVE_STAGES) * TILE_DOUBLES, g_svbar + lk % SOLVE_STAGES, K, lc.r, lc.c);
      lc.advance(lk, ntile);
    }
    ++lk;
  };
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < SOLVE_STAGES - 1; ++k) issue();
  }
  for (int i = tid; i < nblk * SB64; i += FT) sx[i] = i < N ? x[i] : 0.0;
  const int ti = tid >> 2, tp = tid & 3;  // off-diagonal tiles: 4 lanes per row (forward) / column (backward)
  for (int k = 0; k < nsteps; ++k) {
#ifdef IPMZ_FUSED_CLOCKS
    const long long tk0 = clock64();
#endif
    // warp 0 alone polls the stage's mbarrier; the other warps sleep in the hardware barrier (eight warps spinning on
    // try_wait would take issue slots from the co-resident CTA)
    if (warp == 0) mbar_wait(g_svbar + k % SOLVE_STAGES, (ph >> (k % SOLVE_STAGES)) & 1u);  // all lanes: they read the tile next
    ph ^= 1u << (k % SOLVE_STAGES);
    __syncthreads();  // tile k has landed; step k - 1 is finished, its stage is free
    if (tid == 0) issue();
#ifdef IPMZ_FUSED_CLOCKS
    const long long tk1 = clock64();
#endif
    if (k == ntile) {  // between the sweeps: the pivots
      for (int i = tid; i < N; i += FT) sx[i] = sx[i] / Dg[i];
      __syncthreads();
    }
    const double* T = ring + (k % SOLVE_STAGES) * TILE_DOUBLES;
    const int r = cc.r, c = cc.c;
    cc.advance(k, ntile);
    const int R0 = r * SB64, C0 = c * SB64;
    if (k < ntile) {
      if (c < r) {  // y_r -= L(r,c) x_c
        const double* Tr = T + ti * TP + tp;
        const double* xc = sx + C0 + tp;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          s0 = fma(Tr[4 * j], xc[4 * j], s0);
          s1 = fma(Tr[4 * j + 4], xc[4 * j + 4], s1);
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (tp == 0) sx[R0 + ti] -= s;
      } else if (warp == 0) {  // unit lower triangular 64 x 64
        tri64_forward(T, sx + R0, lane);
      }
    } else {
      if (c < r) {  // x_c -= L(r,c)^T x_r
        const double* Tc = T + tp * TP + ti;
        const double* xr = sx + R0 + tp;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          s0 = fma(Tc[4 * j * TP], xr[4 * j], s0);
          s1 = fma(Tc[(4 * j + 4) * TP], xr[4 * j + 4], s1);
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (tp == 0) sx[C0 + ti] -= s;
      } else if (warp == 0) {  // unit upper triangular (the transposed tile)
        tri64_backward(T, sx + R0, lane);
      }
    }
#ifdef IPMZ_FUSED_CLOCKS
    if (tid == 0) {
      const long long tk2 = clock64();
      atomicAdd(&g_fused_clk[9], (unsigned long long)(tk1 - tk0));
      atomicAdd(&g_fused_clk[c == r ? 10 : 11], (unsigned long long)(tk2 - tk1));
      atomicAdd(&g_fused_clk[c == r ? 12 : 13], 1ull);
    }
#endif
  }
  __syncthreads();
  if (tid == 0) g_svph = ph;
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
    MATVEC_T(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, v.tm + (size_t)p * s.ms, v.tn + (size_t)p * s.ns);
    __syncthreads();
  }
  for (int i = tid; i < len; i += FT) prepare_sol_body(v, p, i, rvec, 1);
  __syncthreads();
  {
    FSUB_BEGIN;
    ldlt_solve(v.K + (size_t)p * v.sK, v.Dg + (size_t)p * v.ldk, v.N, sol, sm);
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
    ldlt_solve(v.K + (size_t)p * v.sK, v.Dg + (size_t)p * v.ldk, v.N, v.sol + (size_t)p * v.ssol, sm);
  } else {
    (void)rhs;
    condensed_solve(v, p, v.rhs, 0, sm);
    double* out = v.out + (size_t)p * (s.ns + s.ms);
    for (int r = 0; r < nref; ++r) {
      MATVEC_Q(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, out, v.Qd + (size_t)p * s.ns);
      if (s.m > 0) {
        MATVEC_T(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, out + s.ns, v.tn + (size_t)p * s.ns);
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

// ---- work distribution ----------------------------------------------------------------------------------------
// Queue mode (default): the unit of work is ONE ITERATION of one problem.  All of a problem's state between iterations
// lives in global memory, so any CTA can continue it.  A CTA takes a fresh problem if one is resident (ticket < *ready),
// otherwise the waiting problem with the FEWEST iterations done (one FIFO bucket per iteration count, lowest non-empty
// bucket first), runs one Mehrotra iteration on it and appends it to the next bucket unless the stopping test fired.
// Least-iterations-first is critical-path-first for chains of similar length: the batch advances as one front, a share
// of 512 problems keeps all 296 resident CTAs busy to the end (problem-granular tickets: 1.73 waves), and in a streamed
// solve late arrivals catch up with the front instead of queueing behind it, so everything finishes together shortly
// after the last upload chunk.  Publication: every thread's stores -> __syncthreads -> thread 0: __threadfence, then the
// queue slot; acquisition: thread 0 reads the slot, __threadfence (drops stale L1 lines of a problem this SM saw an
// iteration ago), __syncthreads.
// Control words: [0] next fresh problem, [3] problems finished, [4] highest bucket used, [8 + 2 b] head and
// [8 + 2 b + 1] tail of bucket b (b = iterations done, 1 .. max_iter); slots of bucket b at queue + (b - 1) * count.
__device__ __forceinline__ int vload(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void vstore(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// thread 0 only.  Returns the problem (fresh = 1: first touch), -1 when every problem is finished, -2 on the watchdog.
// (scalar arguments: passing the kernel's argument struct by reference would move the whole View into local memory)
__device__ __noinline__ int acquire_work(int* ctl, const int* queue, int count, const volatile int* ready, int& fresh) {
  const long long t0 = clock64();
  for (;;) {
    const int n = vload(ctl + 0);
    int avail = count;
    if (ready) { avail = *ready; if (avail > count) avail = count; }
    if (n < avail) {
      if (atomicCAS(ctl + 0, n, n + 1) == n) { fresh = 1; return n; }
      continue;
    }
    const int maxb = vload(ctl + 4);
    bool lost = false;
    for (int b = 1; b <= maxb; ++b) {
      int* hb = ctl + 8 + 2 * b;
      const int h = vload(hb), t = vload(hb + 1);
      if (h < t) {
        if (atomicCAS(hb, h, h + 1) == h) {  // slot h is ours; its writer bumped the tail first, the value follows
          const int* slot = queue + (size_t)(b - 1) * count + h;
          int p;
          while ((p = vload(slot)) < 0) {}
          fresh = 0;
          return p;
        }
        lost = true;  // another CTA took it: look again from the lowest bucket
        break;
      }
    }
    if (lost) continue;
    if (vload(ctl + 3) >= count) return -1;
    __nanosleep(500);
    if (clock64() - t0 > (8ll << 30)) return -2;  // ~4 s without work: an upload that never arrived
  }
}
// thread 0 only, after __syncthreads + __threadfence: problem p has `iters` iterations done and needs another one
__device__ __forceinline__ void release_work(int* ctl, int* queue, int count, int p, int iters) {
  const int t = atomicAdd(ctl + 8 + 2 * iters + 1, 1);
  vstore(queue + (size_t)(iters - 1) * count + t, p);
  atomicMax(ctl + 4, iters);
}

__global__ void __launch_bounds__(FT, IPMZ_FUSED_CTAS) k_ipm_batch(FusedArgs a) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[4][FW];
  __shared__ int s_p;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MV_NBUF; ++i) mbar_init(g_mvbar + i, 1);
    for (int i = 0; i < SOLVE_STAGES; ++i) mbar_init(g_svbar + i, 1);
    g_mvph = 0;
    g_svph = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const View& v = a.v;
  const Shape& s = v.s;
  const int tid = threadIdx.x;
  const int len = max(s.n, s.m);
  const bool fifo = a.queue != nullptr;
  __shared__ int s_fresh;
  for (;;) {
    if (tid == 0) {
      int fr = 1, q;
      if (fifo) {
        q = acquire_work(a.ticket, a.queue, a.count, a.ready, fr);
        __threadfence();
      } else {
        q = atomicAdd(a.ticket, 1);
        if (q >= a.count) q = -1;
      }
      s_p = q;
      s_fresh = fr;
    }
    __syncthreads();
    const int p = s_p;
    const int fresh = s_fresh;
    __syncthreads();
    if (p < 0) {
      if (p == -2 && tid == 0 && a.abort_flag) atomicExch(a.abort_flag, 1);
      break;
    }
    Scal& sc = v.sc[p];
    if (fresh && a.ready) {
      if (!fifo) {  // problem-granular tickets: wait for this problem's data
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
      }
      // initial point (EnvironmentBuilder.cpp:34-73), as ipmz_batch_upload's kernel would have left it (the fused path
      // never reads the transposed copy M^T the upload path builds for the grid-per-phase kernels)
      for (int i = tid; i < len; i += FT) initial_point_body(v, p, i);
      __syncthreads();
    }
    if (fresh && tid == 0) {  // a fresh solve restarts the counters and keeps the iterate (warm start, as ipmz_solve)
      sc.iters = 0; sc.done = 0; sc.mu_c = 0.0; sc.alpha = 0.0; sc.alpha_aff = 0.0; sc.sigma = 0.0;
    }
    double* V = v.V + (size_t)p * v.sp;
#ifdef IPMZ_FUSED_DBG_MODES
    if (a.dbg == 1) {
      for (int rep = 0; rep < 30; ++rep) {
        MATVEC_Q(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, V, v.Qx + (size_t)p * s.ns);
        MATVEC_BOTH(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, V, v.Mx + (size_t)p * s.ms, V + (size_t)N_NSLOTS * s.ns,
                    v.MTl + (size_t)p * s.ns);
      }
      __syncthreads();
      continue;
    }
    if (a.dbg >= 2) {
      MATVEC_Q(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, V, v.Qx + (size_t)p * s.ns);
      MATVEC(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, V, v.Mx + (size_t)p * s.ms);
      __syncthreads();
      {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < len; i += FT) residuals_rhs_body<0>(v, p, i, acc);
        __syncthreads();
      }
      for (int rep = 0; rep < 20; ++rep) {
        assemble_normal(v, p, sm);
        if (a.dbg >= 3) ldlt_panels(v.K + (size_t)p * v.sK, v.Dg + (size_t)p * v.ldk, v.N, sm);
        if (a.dbg >= 4) ldlt_solve(v.K + (size_t)p * v.sK, v.Dg + (size_t)p * v.ldk, v.N, v.sol + (size_t)p * v.ssol, sm);
      }
      __syncthreads();
      continue;
    }
#endif
    FPH_DECL;
    for (;;) {
      // ---- Q x, M x, M^T lambda
      MATVEC_Q(v.Q + (size_t)p * v.sQ, v.ldq, s.n, s.n, V, v.Qx + (size_t)p * s.ns);
      if (s.m > 0) {
        MATVEC_BOTH(v.M + (size_t)p * v.sM, v.ldm, s.m, s.n, V, v.Mx + (size_t)p * s.ms, V + (size_t)N_NSLOTS * s.ns,
                    v.MTl + (size_t)p * s.ns);
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
      ldlt_panels(v.K + (size_t)p * v.sK, v.Dg + (size_t)p * v.ldk, v.N, sm);
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
      if (fifo) break;  // one iteration per acquisition: the problem goes back to the queue
    }
    __syncthreads();
    if (fifo && tid == 0) {
      __threadfence();  // the iteration's stores (all threads, ordered by the barrier above) before the hand-over
      if (sc.done != 0) atomicAdd(a.ticket + 3, 1);
      else release_work(a.ticket, a.queue, a.count, p, sc.iters);
    }
  }
}

int fused_smem_doubles(const View& v) {
  const Shape& s = v.s;
  const int rows_cap = (v.N + 15) & ~15;
  const int ldlt = rows_cap * PP + 32 + 32 + CBUF + INV_SUB;
  const int syrk = FSTAGES * STAGE_DOUBLES + s.ms + s.ns;
  const int nblk = (v.N + SB64 - 1) / SB64;
  const int solve = nblk * SB64 + SOLVE_STAGES * TILE_DOUBLES;
  int m = ldlt > syrk ? ldlt : syrk;
  if (solve > m) m = solve;
  if (MV_NBUF * MV_CHUNK + 512 > m) m = MV_NBUF * MV_CHUNK + 512;  // staging buffers + the staged vector of A^T v
  return m;
}

}  // namespace

// doubles of one problem's reduced matrix in the fused path's tile-major layout (the workspace allocates at least this)
size_t fused_k_doubles(int N) {
  const int nblk = (N + SB64 - 1) / SB64;
  return (size_t)(nblk * (nblk + 1) / 2) * TILE_DOUBLES;
}

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
int fused_ctl_words(int max_iter) { return 8 + 2 * (std::max(1, max_iter) + 1); }
size_t fused_work_words(int count, int max_iter) {
  return (size_t)fused_ctl_words(max_iter) + (size_t)count * std::max(1, max_iter);
}

int launch_ipm_batch(cudaStream_t st, const View& v, int count, int refine_fixed, int* ticket, const int* ready,
                     int* abort_flag, size_t work_words) {
  FusedArgs a;
  a.ready = ready;
  a.abort_flag = abort_flag;
  a.v = v;
  a.v.active = nullptr;
  a.count = count;
  a.ticket = ticket;
  a.refine_fixed = refine_fixed;
  a.smem_doubles = fused_smem_doubles(v);
  a.dbg = getenv("IPMZ_FUSED_DBG") ? atoi(getenv("IPMZ_FUSED_DBG")) : 0;
  const size_t smem = (size_t)a.smem_doubles * sizeof(double);
  // `ticket`: control words, then the bucket slots (fused_work_words).  IPMZ_FUSED_QUEUE=0 / debug modes / a buffer
  // that is too small: problem-granular tickets.
  static const int use_queue = getenv("IPMZ_FUSED_QUEUE") ? atoi(getenv("IPMZ_FUSED_QUEUE")) : 1;
  const int ctlw = fused_ctl_words(v.max_iter);
  const bool fifo = use_queue && a.dbg == 0 && work_words >= fused_work_words(count, v.max_iter);
  a.queue = fifo ? ticket + ctlw : nullptr;
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(int) * (fifo ? ctlw : 8), st);
  if (e != cudaSuccess) return (int)e;
  if (fifo) {
    e = cudaMemsetAsync(a.queue, 0xFF, sizeof(int) * (size_t)count * std::max(1, v.max_iter), st);  // -1: slot not written yet
    if (e != cudaSuccess) return (int)e;
  }
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
