// ipm-zoo_b200/csrc/factor.cu -- blocked right-looking FP64 LDL^T (quasi-definite augmented
// system) / root-free Cholesky (normal equations) of the reduced Newton matrix, in place in HBM.
//
// Reference: LinearSolvers::ldlt_decomposition (LinearSolvers.cpp:14-42), a row-oriented
// scalar triple loop, no pivoting, `D[i] == 0 -> 1e-8` (:28).  Here:
//   k_diag_ldlt   one CTA factors the NB x NB diagonal block in shared memory; its 32 x 32
//                 sub-blocks are factored by one warp with the columns exchanged by shuffles.
//   k_trsm_panel  the rows below the block: X (D L_kk^T) = A, 64 rows per CTA staged in
//                 shared memory.
//   k_syrk_ldl    trailing update C -= P diag(d) P^T on the FP64 tensor pipe
//                 (mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4; tcgen05 has no FP64 kind), operands
//                 staged with a 3-deep cp.async pipeline; the only dense contraction of the path.
// The same k_syrk_ldl forms the condensed matrix Hx + M^T W M of the normal reduction.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ipmz_device.cuh"
#include "ipmz_kernels.h"
#include "ldlt_device.cuh"

namespace ipmz {

__device__ long long g_phase_clk[16];

namespace {

#ifdef IPMZ_PHASE_CLOCKS
#define PHASE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_phase_clk[i] = clock64(); } while (0)
#else
#define PHASE(i) do {} while (0)
#endif

__global__ void __launch_bounds__(256) k_diag_ldlt(const double* src, double* dst,
                                                   int ld, size_t sK, double* __restrict__ Dg, size_t sD,
                                                   double* __restrict__ Ginv, size_t sInv,
                                                   int k0, int nb, const int* __restrict__ active) {
  extern __shared__ double sm[];
  double* S = sm;
  double* dsm = sm + NB * SP;
  double* dinv = dsm + NB;
  double* colbuf = dinv + NB;          // exchange buffers of warp_ldlt32
  double* binv = colbuf + CBUF;      // 4 sub-blocks x 4 blocks x INV_BLK
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const double* A = src + (size_t)p * sK + (size_t)k0 * ld + k0;
  double* O = dst + (size_t)p * sK + (size_t)k0 * ld + k0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  PHASE(0);
  // zero-fill so that partial sub-blocks / tiles never read uninitialised shared memory
  for (int i = tid; i < NB * SP; i += 256) S[i] = 0.0;
  __syncthreads();
  async_block_load<true>(S, A, ld, nb, nb, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  PHASE(1);

  for (int j0 = 0; j0 < nb; j0 += SB) {
    const int jb = min(SB, nb - j0);
    if (j0 == 0) PHASE(2);
    if (warp == 0) warp_ldlt32(S, j0, jb, dsm, dinv, colbuf, binv + (j0 / SB) * INV_SUB, lane);
    __syncthreads();
    if (j0 == 0) PHASE(3);
    const int base = j0 + jb, rem = nb - base;
    if (rem > 0) {
      panel_solve32(S + base * SP + j0, rem, S + j0 * SP + j0, dsm + j0, binv + (j0 / SB) * INV_SUB, warp, lane, 8);
      __syncthreads();
      if (j0 == 0) PHASE(4);
      smem_update<true>(S + base * SP + base, S + base * SP + j0, S + base * SP + j0, dsm + j0, rem, rem, jb,
                        warp, lane, 8);
      __syncthreads();
      if (j0 == 0) PHASE(5);
    }
  }
  PHASE(6);

  for (int r = warp; r < nb; r += 8) {
    for (int c = lane; c < r; c += 32) O[(size_t)r * ld + c] = S[r * SP + c];
    if (lane == 0) O[(size_t)r * ld + r] = dsm[r];
  }
  for (int t = tid; t < nb; t += 256) Dg[(size_t)p * sD + k0 + t] = dsm[t];
  {
    double* gi = Ginv + (size_t)p * sInv + (size_t)(k0 / 8) * INV_BLK;
    const int nblk = (nb + SB - 1) / SB * 4;
    for (int t = tid; t < nblk * INV_BLK; t += 256) gi[t] = binv[t];
  }
  PHASE(7);
}

__global__ void __launch_bounds__(256) k_trsm_panel(const double* src, double* dst,
                                                    int ld, size_t sK, const double* __restrict__ Dg, size_t sD,
                                                    const double* __restrict__ Ginv, size_t sInv,
                                                    double* __restrict__ Wp, size_t sW,
                                                    int k0, int nb, int N, const int* __restrict__ active) {
  extern __shared__ double sm[];
  double* S = sm;                  // L_kk (strict lower)
  double* T = sm + NB * SP;        // this CTA's rows of the panel
  double* dsm = T + RB * SP;
  double* binv = dsm + NB;
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = k0 + nb + blockIdx.x * RB;
  const int nr = min(RB, N - r0);
  const double* Lkk = dst + (size_t)p * sK + (size_t)k0 * ld + k0;
  const double* Ain = src + (size_t)p * sK + (size_t)r0 * ld + k0;
  double* Aout = dst + (size_t)p * sK + (size_t)r0 * ld + k0;

  if (nb < NB || nr < RB) {
    for (int i = tid; i < (NB + RB) * SP; i += 256) sm[i] = 0.0;
    __syncthreads();
  }
  async_block_load<true>(S, Lkk, ld, nb, nb, tid);
  async_block_load<false>(T, Ain, ld, nr, nb, tid);
  cp_async_commit();
  for (int t = tid; t < nb; t += 256) dsm[t] = Dg[(size_t)p * sD + k0 + t];
  {
    const double* gi = Ginv + (size_t)p * sInv + (size_t)(k0 / 8) * INV_BLK;
    const int nblk = (nb + SB - 1) / SB * 4;
    for (int t = tid; t < nblk * INV_BLK; t += 256) binv[t] = gi[t];
  }
  cp_async_wait<0>();
  __syncthreads();

  for (int c0 = 0; c0 < nb; c0 += SB) {
    const int cb = min(SB, nb - c0);
    panel_solve32(T + c0, nr, S + c0 * SP + c0, dsm + c0, binv + (c0 / SB) * INV_SUB, warp, lane, 8);
    __syncthreads();
    const int base = c0 + cb, rem = nb - base;
    if (rem > 0) {
      smem_update<false>(T + base, T + c0, S + base * SP + c0, dsm + c0, nr, rem, cb, warp, lane, 8);
      __syncthreads();
    }
  }

  // L into the factor, W = L D (the pre-scaled B operand of the trailing updates) into the panel buffer
  double* Wout = Wp + (size_t)p * sW + (size_t)r0 * WLD + (k0 % WLD);
  for (int r = warp; r < nr; r += 8)
    for (int c = lane; c < nb; c += 32) {
      const double l = T[r * SP + c];
      Aout[(size_t)r * ld + c] = l;
      Wout[(size_t)r * WLD + c] = l * dsm[c];
    }
}

// ------------------------------------------------------------------------------------------
// DMMA SYRK:  Cout(lower) = Cin + sign * P diag(d) P^T
//
// CTA tile BM x BN (BM = 2 BN), 8 warps as WM x WN, each warp (BM/WM) x (BN/WN) built from
// m8n8k4 DMMAs.  Two CTAs are resident per SM (<= 128 registers, ~92 KB shared memory each), so
// while one CTA waits for its C tile or for a cp.async stage the other keeps the FP64 tensor
// pipe busy.  Operands: 3-stage cp.async pipeline of BK = 16 wide k-slices, rows padded to 20
// doubles so fragment loads are bank-conflict free.
constexpr int BM = 128, BN = 64, STAGES = 3, WARPS_M = 4, WARPS_N = 2;
constexpr int MI = BM / (WARPS_M * 8), NI = BN / (WARPS_N * 8);
constexpr int TILE_RATIO = BM / BN;
constexpr size_t SYRK_SMEM = (size_t)(STAGES * (BM + BN) * LDT) * sizeof(double);
static_assert(((BM + BN) * (BK / 2)) % 256 == 0, "stage loads must divide evenly over 256 threads");

struct SyrkArgs {
  const double* Cin;
  double* Cout;
  int ldc;
  size_t sC;
  const double* PA;  // rows x kdim, A operand (rows of the tile)
  int lda;
  size_t sA;
  const double* PB;  // rows x kdim, B operand = P diag(d), pre-scaled (columns of the tile)
  int ldb;
  size_t sB;
  int rows, kdim;
  double sign;
  const int* active;
  int tn;      // number of BN-wide tile columns
  int mode;    // 1: only the first `fcols` BM-wide block columns; otherwise: all but the first `fcols`
  int fcols;
  int ntiles;  // tiles of this launch; CTAs are persistent and stride over them
};

__device__ __forceinline__ bool syrk_tile(const SyrkArgs& a, int t, int& ti, int& tj) {
  if (a.mode == 1) {  // only the first a.fcols BM-wide block columns (look-ahead / panel-internal update)
    const int w = TILE_RATIO * a.fcols;
    ti = t / w;
    tj = t - ti * w;
    if (tj > TILE_RATIO * ti + (TILE_RATIO - 1)) return false;  // above the diagonal
  } else {  // all tiles on/below the diagonal except the first a.fcols block columns
    int u = (int)((sqrt(8.0 * (double)t / TILE_RATIO + 1.0) - 1.0) * 0.5);
    while (TILE_RATIO * u * (u + 1) / 2 > t) --u;
    while (TILE_RATIO * (u + 1) * (u + 2) / 2 <= t) ++u;
    ti = u + a.fcols;
    tj = t - TILE_RATIO * u * (u + 1) / 2 + TILE_RATIO * a.fcols;
  }
  return tj < a.tn;
}

__global__ void __launch_bounds__(256, 2) k_syrk_ldl(SyrkArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = As + STAGES * BM * LDT;

  const int p = a.active ? a.active[blockIdx.y] : blockIdx.y;
  const double* PA = a.PA + (size_t)p * a.sA;
  const double* PB = a.PB + (size_t)p * a.sB;
  const double* Cin = a.Cin + (size_t)p * a.sC;
  double* Cout = a.Cout + (size_t)p * a.sC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int g = lane >> 2, q = lane & 3;
  const int KT = (a.kdim + BK - 1) / BK;

  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    int ti, tj;
    if (!syrk_tile(a, t, ti, tj)) continue;
    const int row0 = ti * BM, col0 = tj * BN;
    const int wrow = row0 + wm * (MI * 8), wcol = col0 + wn * (NI * 8);

    auto load_stage = [&](int stage, int kt) {
      const int kbase = kt * BK;
      double* Asd = As + stage * BM * LDT;
      double* Bsd = Bs + stage * BN * LDT;
#pragma unroll
      for (int i = 0; i < (BM + BN) * (BK / 2) / 256; ++i) {
        const int chunk = tid + i * 256;
        const int r = chunk >> 3, ck = (chunk & 7) * 2;
        const int k = kbase + ck;
        const bool isA = r < BM;
        const int gr = isA ? row0 + r : col0 + (r - BM);
        const bool ok = (gr < a.rows) && (k < a.kdim);
        const double* base = isA ? PA : PB;
        const int ld = isA ? a.lda : a.ldb;
        const double* srcp = base + (size_t)(ok ? gr : 0) * ld + (ok ? k : 0);
        double* dstp = isA ? Asd + r * LDT + ck : Bsd + (r - BM) * LDT + ck;
        cp_async16(dstp, srcp, ok ? 16 : 0);
      }
    };

    __syncthreads();  // the previous tile's last stage is no longer read by any warp
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < KT) load_stage(s, s);
      cp_async_commit();
    }

    // Accumulators start from sign*C (all loads independent and in flight while the cp.async
    // prologue lands), so the epilogue is store-only:  Cout = sign * (sign*Cin + PA PB^T).
    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const int row = wrow + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int col = wcol + ni * 8 + 2 * q;
        double2 cv = make_double2(0.0, 0.0);
        if (row < a.rows && col <= row) {
          const size_t off = (size_t)row * a.ldc + col;
          if (col + 1 <= row) cv = *reinterpret_cast<const double2*>(Cin + off);
          else cv.x = Cin[off];
        }
        acc[mi][ni][0] = a.sign * cv.x;
        acc[mi][ni][1] = a.sign * cv.y;
      }
    }
    {
      // warm L2 with the C tile this CTA visits next, so its accumulator loads do not wait on HBM
      const int tn2 = t + gridDim.x;
      int ti2, tj2;
      if (tn2 < a.ntiles && syrk_tile(a, tn2, ti2, tj2)) {
        const int r = ti2 * BM + (tid >> 1);
        const int c = tj2 * BN + (tid & 1) * 32;
        if (r < a.rows && c <= r) {
          const double* pf = Cin + (size_t)r * a.ldc + c;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 16));
        }
      }
    }

    for (int kt = 0; kt < KT; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        const int nk = kt + STAGES - 1;
        if (nk < KT) load_stage(nk % STAGES, nk);
        cp_async_commit();
      }
      const int stage = kt % STAGES;
      const double* Aw = As + stage * BM * LDT + (wm * (MI * 8) + g) * LDT + q;
      const double* Bw = Bs + stage * BN * LDT + (wn * (NI * 8) + g) * LDT + q;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double af[MI], bf[NI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) af[mi] = Aw[mi * 8 * LDT + kk * 4];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bf[ni] = Bw[ni * 8 * LDT + kk * 4];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
      }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const int row = wrow + mi * 8 + g;
      if (row >= a.rows) continue;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int col = wcol + ni * 8 + 2 * q;
        if (col > row) continue;  // strictly upper part of a diagonal-crossing tile
        const size_t off = (size_t)row * a.ldc + col;
        if (col + 1 <= row) {
          double2 o;
          o.x = a.sign * acc[mi][ni][0];
          o.y = a.sign * acc[mi][ni][1];
          *reinterpret_cast<double2*>(Cout + off) = o;
        } else {
          Cout[off] = a.sign * acc[mi][ni][0];
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// Warp-specialised persistent variant of the same update for one large matrix:
// one CTA per SM (≈205 KB shared memory), 128 x 128 tiles, 16 consumer warps (4 x 4, 32 x 32
// each) + 1 producer warp.  The producer streams 16-wide k-slices of both operands through a
// 5-stage ring with cp.async; completion is signalled on per-stage "full" mbarriers
// (cp.async.mbarrier.arrive) and consumers release slots on "empty" mbarriers, so there is no
// CTA-wide barrier in the loop and the producer runs ahead across tile boundaries (the next
// tile's operands arrive while the consumers store C).  CTAs stride over the tile list; the
// launch uses fewer CTAs than SMs when the look-ahead schedule reserves SMs for the panel
// kernels of the side stream.
constexpr int WS_BM = 128, WS_BN = 128, WS_STAGES = 5, WS_CONSUMERS = 16, WS_PRODUCERS = 4;
constexpr int WS_THREADS = (WS_CONSUMERS + WS_PRODUCERS) * 32;
constexpr int WS_MI = 4, WS_NI = 4;
constexpr int WS_STAGE_DOUBLES = (WS_BM + WS_BN) * LDT;
constexpr size_t WS_SMEM = (size_t)WS_STAGES * WS_STAGE_DOUBLES * sizeof(double) + 2 * WS_STAGES * sizeof(unsigned long long);

__device__ __forceinline__ bool ws_tile(const SyrkArgs& a, int t, int& ti, int& tj) {
  if (a.mode == 1) {
    ti = t / a.fcols;
    tj = t - ti * a.fcols;
    if (tj > ti) return false;
  } else {
    int u = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (u * (u + 1) / 2 > t) --u;
    while ((u + 1) * (u + 2) / 2 <= t) ++u;
    ti = u + a.fcols;
    tj = t - u * (u + 1) / 2 + a.fcols;
  }
  return true;
}

__global__ void __launch_bounds__(WS_THREADS, 1) k_syrk_ws(SyrkArgs a) {
  extern __shared__ __align__(16) double smem[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + WS_STAGES * WS_STAGE_DOUBLES);
  unsigned long long* empty = full + WS_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* PA = a.PA;
  const double* PB = a.PB;
  const int KT = (a.kdim + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < WS_STAGES; ++s) {
      mbar_init(full + s, 32 * WS_PRODUCERS);  // one cp.async-completion arrive per producer lane
      mbar_init(empty + s, WS_CONSUMERS);  // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= WS_CONSUMERS) {
    // ---------------- producers (one warp per SM sub-partition) ----------------
    const int pl = tid - WS_CONSUMERS * 32;  // 0 .. 32*WS_PRODUCERS-1
    unsigned it = 0;
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
      int ti, tj;
      if (!ws_tile(a, t, ti, tj)) continue;
      const int row0 = ti * WS_BM, col0 = tj * WS_BN;
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const unsigned slot = it % WS_STAGES;
        mbar_wait(empty + slot, ((it / WS_STAGES) & 1u) ^ 1u);
        double* Sd = smem + slot * WS_STAGE_DOUBLES;
        const int kbase = kt * BK;
#pragma unroll 8
        for (int i = 0; i < (WS_BM + WS_BN) * (BK / 2) / (32 * WS_PRODUCERS); ++i) {
          const int chunk = pl + i * 32 * WS_PRODUCERS;
          const int r = chunk >> 3, ck = (chunk & 7) * 2;
          const int k = kbase + ck;
          const bool isA = r < WS_BM;
          const int gr = isA ? row0 + r : col0 + (r - WS_BM);
          const bool ok = (gr < a.rows) && (k < a.kdim);
          const double* srcp = (isA ? PA + (size_t)(ok ? gr : 0) * a.lda : PB + (size_t)(ok ? gr : 0) * a.ldb) + (ok ? k : 0);
          cp_async16(Sd + r * LDT + ck, srcp, ok ? 16 : 0);
        }
        mbar_arrive_cp_async(full + slot);
      }
    }
    cp_async_wait<0>();
    return;
  }

  // ---------------- consumers ----------------
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, q = lane & 3;
  const double* Cin = a.Cin;
  double* Cout = a.Cout;
  unsigned it = 0;
  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    int ti, tj;
    if (!ws_tile(a, t, ti, tj)) continue;
    const int row0 = ti * WS_BM, col0 = tj * WS_BN;
    const int wrow = row0 + wm * 32, wcol = col0 + wn * 32;
    double acc[WS_MI][WS_NI][2];
#pragma unroll
    for (int mi = 0; mi < WS_MI; ++mi) {
      const int row = wrow + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < WS_NI; ++ni) {
        const int col = wcol + ni * 8 + 2 * q;
        double2 cv = make_double2(0.0, 0.0);
        if (row < a.rows && col <= row) {
          const size_t off = (size_t)row * a.ldc + col;
          if (col + 1 <= row) cv = *reinterpret_cast<const double2*>(Cin + off);
          else cv.x = Cin[off];
        }
        acc[mi][ni][0] = a.sign * cv.x;
        acc[mi][ni][1] = a.sign * cv.y;
      }
    }
    {
      // warm L2 with the C tile this CTA visits next
      const int t2 = t + gridDim.x;
      int ti2, tj2;
      if (t2 < a.ntiles && ws_tile(a, t2, ti2, tj2)) {
        const int r = ti2 * WS_BM + (tid >> 2);
        const int c = tj2 * WS_BN + (tid & 3) * 32;
        if (r < a.rows && c <= r) {
          const double* pf = Cin + (size_t)r * a.ldc + c;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 16));
        }
      }
    }
    for (int kt = 0; kt < KT; ++kt, ++it) {
      const unsigned slot = it % WS_STAGES;
      mbar_wait(full + slot, (it / WS_STAGES) & 1u);
      const double* Aw = smem + slot * WS_STAGE_DOUBLES + (wm * 32 + g) * LDT + q;
      const double* Bw = smem + slot * WS_STAGE_DOUBLES + WS_BM * LDT + (wn * 32 + g) * LDT + q;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double af[WS_MI], bf[WS_NI];
#pragma unroll
        for (int mi = 0; mi < WS_MI; ++mi) af[mi] = Aw[mi * 8 * LDT + kk * 4];
#pragma unroll
        for (int ni = 0; ni < WS_NI; ++ni) bf[ni] = Bw[ni * 8 * LDT + kk * 4];
#pragma unroll
        for (int mi = 0; mi < WS_MI; ++mi)
#pragma unroll
          for (int ni = 0; ni < WS_NI; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
#pragma unroll
    for (int mi = 0; mi < WS_MI; ++mi) {
      const int row = wrow + mi * 8 + g;
      if (row >= a.rows) continue;
#pragma unroll
      for (int ni = 0; ni < WS_NI; ++ni) {
        const int col = wcol + ni * 8 + 2 * q;
        if (col > row) continue;
        const size_t off = (size_t)row * a.ldc + col;
        if (col + 1 <= row) {
          double2 o;
          o.x = a.sign * acc[mi][ni][0];
          o.y = a.sign * acc[mi][ni][1];
          *reinterpret_cast<double2*>(Cout + off) = o;
        } else {
          Cout[off] = a.sign * acc[mi][ni][0];
        }
      }
    }
  }
}

}  // namespace

static int g_num_sms = 0;
static int g_use_ws = 1;      // IPMZ_SYRK_WS=0 selects the 2-CTA/SM kernel everywhere (A/B testing)
static int g_ws_reserve = 16;  // SMs left to the side stream while the main stream updates

int factor_init() {
  cudaError_t e;
  if (const char* s = getenv("IPMZ_SYRK_WS")) g_use_ws = atoi(s);
  if (const char* s = getenv("IPMZ_WS_RESERVE")) g_ws_reserve = atoi(s);
  e = cudaFuncSetAttribute(k_diag_ldlt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_syrk_ldl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_syrk_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
  if (e != cudaSuccess) return (int)e;
  const int te = trsv_init();
  if (te != 0) return te;
  const int be = bk_init();
  if (be != 0) return be;
  const int fe = fused_batch_init();
  if (fe != 0) return fe;
  return dataflow_init();
}


static void launch_syrk_mode(cudaStream_t st, int nslots, const int* active, const double* Cin, double* Cout,
                             int ldc, size_t sC, const double* PA, int lda, size_t sA, const double* PB, int ldb,
                             size_t sB, int rows, int kdim, double sign, int mode, int fcols, bool side_stream) {
  if (rows <= 0 || kdim <= 0 || nslots <= 0) return;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (nslots == 1 && rows >= 1024 && g_use_ws) {
    // warp-specialised persistent kernel, 128 x 128 tiles (fcols counts 128-wide block columns in both)
    const int T = (rows + WS_BM - 1) / WS_BM;
    int tiles;
    if (mode == 1) tiles = T * fcols;
    else { const int U = T - fcols; tiles = U > 0 ? U * (U + 1) / 2 : 0; }
    if (tiles <= 0) return;
    SyrkArgs a{Cin, Cout, ldc, sC, PA, lda, sA, PB, ldb, sB, rows, kdim, sign, active, 0, mode, fcols, tiles};
    // SMs are left free only when a side stream runs panel kernels next to this launch (look-ahead schedule); the
    // condensed assembly owns the GPU: 2080 tiles at n = 8192 are 16 rounds on 132 CTAs, 15 on 148
    int ctas = g_num_sms - (side_stream ? g_ws_reserve : 0);
    if (ctas < 1) ctas = 1;
    if (ctas > tiles) ctas = tiles;
    k_syrk_ws<<<ctas, WS_THREADS, WS_SMEM, st>>>(a); count_launch();
    return;
  }
  const int T = (rows + BM - 1) / BM;
  int tiles;
  if (mode == 1) {
    tiles = T * TILE_RATIO * fcols;
  } else {
    const int U = T - fcols;
    tiles = U > 0 ? TILE_RATIO * U * (U + 1) / 2 : 0;
  }
  if (tiles <= 0) return;
  SyrkArgs a{Cin, Cout, ldc, sC, PA, lda, sA, PB, ldb, sB, rows, kdim, sign, active, (rows + BN - 1) / BN,
             mode, fcols, tiles};
  // One CTA per tile: with the look-ahead schedule the high-priority panel kernels of the side
  // stream get SM slots as tiles retire (persistent CTAs would hold every slot until the end).
  dim3 grid(tiles, nslots);
  k_syrk_ldl<<<grid, 256, SYRK_SMEM, st>>>(a); count_launch();
}

void launch_syrk_ldl(cudaStream_t st, int nslots, const int* active, const double* Cin, double* Cout, int ldc,
                     size_t sC, const double* PA, int lda, size_t sA, const double* PB, int ldb, size_t sB, int rows,
                     int kdim, double sign) {
  launch_syrk_mode(st, nslots, active, Cin, Cout, ldc, sC, PA, lda, sA, PB, ldb, sB, rows, kdim, sign, 0, 0, false);
}

// Optional per-launch instrumentation (bench roofline): events around every kernel.
struct LaunchHooks {
  struct Rec { cudaEvent_t a, b; int kind; };
  std::vector<Rec>* recs = nullptr;
  double* flops_syrk = nullptr;
  int* n_syrk = nullptr;
  void begin(cudaStream_t st, int kind) const {
    if (!recs) return;
    Rec r; r.kind = kind;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    recs->push_back(r);
  }
  void end(cudaStream_t st) const {
    if (recs) cudaEventRecord(recs->back().b, st);
  }
};

// The scaled-panel buffer is double-buffered by outer (2NB-wide) block: with look-ahead the next
// double panel is written while the trailing update still reads the current one.
static double* wpanel_of(const FactorPlan& fp, int k0) {
  return fp.wpanel + (size_t)((k0 / WLD) & 1) * (size_t)fp.N * WLD;
}

// diag + panel solve of the NB-wide block column at k0 (reads `in`, writes dst)
static void launch_panel(cudaStream_t st, const FactorPlan& fp, const double* in, double* dst, double* Dg, int k0,
                         const LaunchHooks& hk) {
  const int ncols = fp.ncols > 0 ? fp.ncols : fp.N;
  const int nb = ncols - k0 < NB ? ncols - k0 : NB;
  hk.begin(st, 0);
  k_diag_ldlt<<<dim3(1, fp.nslots), 256, DIAG_SMEM, st>>>(in, dst, fp.ld, fp.sK, Dg, fp.sD, fp.inv, fp.sInv, k0, nb,
                                                          fp.active);
  count_launch();
  hk.end(st);
  const int rem = fp.N - k0 - nb;
  if (rem > 0) {
    hk.begin(st, 1);
    k_trsm_panel<<<dim3((rem + RB - 1) / RB, fp.nslots), 256, TRSM_SMEM, st>>>(in, dst, fp.ld, fp.sK, Dg, fp.sD,
                                                                               fp.inv, fp.sInv, wpanel_of(fp, k0), fp.sW, k0,
                                                                               nb, fp.N, fp.active);
    count_launch();
    hk.end(st);
  }
}

// trailing update of the matrix that starts at row/column c0 with the panel columns [p0, p0+kdim)
static void launch_trailing(cudaStream_t st, const FactorPlan& fp, const double* in, double* dst, const double* Dg,
                            int c0, int p0, int kdim, int mode, int fcols, const LaunchHooks& hk) {
  const int rem = fp.N - c0;
  if (rem <= 0) return;
  const size_t off = (size_t)c0 * fp.ld + c0;
  hk.begin(st, 2);
  launch_syrk_mode(st, fp.nslots, fp.active, in + off, dst + off, fp.ld, fp.sK, dst + (size_t)c0 * fp.ld + p0,
                   fp.ld, fp.sK, wpanel_of(fp, p0) + (size_t)c0 * WLD + (p0 % WLD), WLD, fp.sW, rem, kdim, -1.0, mode,
                   fcols, fp.la != nullptr);
  hk.end(st);
  if (hk.flops_syrk) {
    // algorithmic flops: lower triangle of the updated region, 2 flops per multiply-add
    const double r = (double)rem, w = (double)(fcols * BM < rem ? fcols * BM : rem);
    const double elems = mode == 1 ? (r * w - w * (w - 1) / 2) : (r - w) * (r - w + 1) / 2;
    *hk.flops_syrk += (double)fp.nslots * 2.0 * elems * (double)kdim;
    if (hk.n_syrk) *hk.n_syrk += 1;
  }
}

// Two NB-wide panels (columns [k0, k0+2NB)) factored back to back: panel A, its update of
// block column B only, panel B.  After this the trailing matrix can be updated with K = 2NB.
static void launch_double_panel(cudaStream_t st, const FactorPlan& fp, const double* in, double* dst, double* Dg,
                                int k0, const LaunchHooks& hk) {
  launch_panel(st, fp, in, dst, Dg, k0, hk);
  const int k1 = k0 + NB;
  if (k1 >= fp.N) return;
  launch_trailing(st, fp, in, dst, Dg, k1, k0, NB, 1, 1, hk);
  launch_panel(st, fp, dst, dst, Dg, k1, hk);
}

static void ldlt_schedule(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg,
                          const LaunchHooks& hk, bool allow_overlap) {
  const LookAhead* la = fp.la;
  if (fp.ncols > 0 && fp.ncols < fp.N) {
    // partial elimination (dual-Schur normal equations): the first ncols columns, one panel at a time; the last panel
    // may be narrower than NB, the trailing update starts right after it
    for (int k0 = 0; k0 < fp.ncols; k0 += NB) {
      const int nb = fp.ncols - k0 < NB ? fp.ncols - k0 : NB;
      const double* in = (k0 == 0) ? src : dst;
      launch_panel(st, fp, in, dst, Dg, k0, hk);
      launch_trailing(st, fp, in, dst, Dg, k0 + nb, k0, nb, 0, 0, hk);
    }
    return;
  }
  const bool two_level = fp.N > 8 * NB;
  const bool overlap = allow_overlap && two_level && la && la->side;
  if (!two_level) {
    // one panel at a time (small matrices and batches)
    for (int k0 = 0; k0 < fp.N; k0 += NB) {
      const double* in = (k0 == 0) ? src : dst;
      launch_panel(st, fp, in, dst, Dg, k0, hk);
      launch_trailing(st, fp, in, dst, Dg, k0 + NB, k0, NB, 0, 0, hk);
    }
    return;
  }
  // Two-level blocking with look-ahead: the trailing matrix is updated with K = 2NB per pass
  // (half the C traffic and half the per-tile prologue of K = NB); the next double panel is
  // factored on the high-priority side stream while the main stream updates the rest.
  const int OB = 2 * NB;
  launch_double_panel(st, fp, src, dst, Dg, 0, hk);
  for (int k0 = 0; k0 + OB < fp.N; k0 += OB) {
    const double* in = (k0 == 0) ? src : dst;
    const int c0 = k0 + OB;
    launch_trailing(st, fp, in, dst, Dg, c0, k0, OB, 1, 2, hk);  // next double block column first
    cudaStream_t ps = st;
    if (overlap) {
      cudaEventRecord(la->e_col, st);
      cudaStreamWaitEvent(la->side, la->e_col, 0);
      ps = la->side;
    }
    launch_double_panel(ps, fp, dst, dst, Dg, c0, hk);
    if (overlap) cudaEventRecord(la->e_panel, la->side);
    launch_trailing(st, fp, in, dst, Dg, c0, k0, OB, 0, 2, hk);  // the rest, concurrently
    if (overlap) cudaStreamWaitEvent(st, la->e_panel, 0);
  }
}

__global__ void k_negate_block(const double* __restrict__ src, int lds, size_t sS, double* __restrict__ dst, int ldd,
                               size_t sD, int m, const int* __restrict__ active) {
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= m) return;
  const double* s = src + (size_t)p * sS + (size_t)r * lds;
  double* d = dst + (size_t)p * sD + (size_t)r * ldd;
  for (int c = lane; c <= r; c += 32) d[c] = -s[c];
}

void launch_negate_block(cudaStream_t st, int nslots, const int* active, const double* src, int lds, size_t sS,
                         double* dst, int ldd, size_t sD, int m) {
  if (m <= 0 || nslots <= 0) return;
  k_negate_block<<<dim3((m + 7) / 8, nslots), 256, 0, st>>>(src, lds, sS, dst, ldd, sD, m, active);
  count_launch();
}

void launch_ldlt(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg) {
  if (fp.df && fp.nslots == 1 && !fp.active && !(fp.ncols > 0 && fp.ncols < fp.N)) {
    launch_ldlt_dataflow(st, *fp.df, src, dst, Dg, fp.inv);
    return;
  }
  ldlt_schedule(st, fp, src, dst, Dg, LaunchHooks{}, true);
}

int lookahead_create(LookAhead* la) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaError_t e = cudaStreamCreateWithPriority(&la->side, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&la->e_col, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&la->e_panel, cudaEventDisableTiming);
  return (int)e;
}

void lookahead_destroy(LookAhead* la) {
  if (la->e_col) cudaEventDestroy(la->e_col);
  if (la->e_panel) cudaEventDestroy(la->e_panel);
  if (la->side) cudaStreamDestroy(la->side);
  la->side = nullptr; la->e_col = nullptr; la->e_panel = nullptr;
}

// Debug: the overlapped schedule with events around every launch; out rows = (kind, start_ms, end_ms)
// relative to the first launch.  Event records perturb the overlap slightly.
int launch_ldlt_timeline(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg,
                         double* out, int cap, int* nrec) {
  std::vector<LaunchHooks::Rec> recs;
  LaunchHooks hk;
  hk.recs = &recs;
  cudaEvent_t base;
  cudaEventCreate(&base);
  cudaEventRecord(base, st);
  ldlt_schedule(st, fp, src, dst, Dg, hk, true);
  cudaError_t e = cudaStreamSynchronize(st);
  if (fp.la && fp.la->side && e == cudaSuccess) e = cudaStreamSynchronize(fp.la->side);
  int n = 0;
  for (auto& r : recs) {
    float t0 = 0.f, t1 = 0.f;
    if (e == cudaSuccess) { cudaEventElapsedTime(&t0, base, r.a); cudaEventElapsedTime(&t1, base, r.b); }
    if (n < cap) { out[3 * n] = r.kind; out[3 * n + 1] = t0; out[3 * n + 2] = t1; ++n; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  cudaEventDestroy(base);
  *nrec = n;
  return (int)e;
}

int launch_ldlt_profiled(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg,
                         double ms[3], double* flops_syrk, int* n_syrk) {
  if (fp.df && fp.nslots == 1 && !fp.active) {
    // the whole factorization is one kernel: its duration against the algorithmic N^3/3 flops
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    launch_ldlt_dataflow(st, *fp.df, src, dst, Dg, fp.inv);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaEventSynchronize(e1);
    float t = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ms[2] += t;
    *flops_syrk += (double)fp.N * fp.N * fp.N / 3.0;
    if (n_syrk) *n_syrk += 1;
    return (int)e;
  }
  std::vector<LaunchHooks::Rec> recs;
  LaunchHooks hk;
  hk.recs = &recs;
  hk.flops_syrk = flops_syrk;
  hk.n_syrk = n_syrk;
  ldlt_schedule(st, fp, src, dst, Dg, hk, false);  // same schedule, serialised on one stream
  cudaError_t e = cudaStreamSynchronize(st);
  for (auto& r : recs) {
    float t = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.kind] += t;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  return (int)e;
}

namespace {
__global__ void k_dmma_probe(double* out, int iters) {
  double c[8][2];
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

int read_phase_clocks(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, g_phase_clk, sizeof(long long) * 16);
}

int fp64_peak_probe(cudaStream_t st, double* tflops) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(double) * (size_t)sms * 2 * 256);
  if (e != cudaSuccess) return (int)e;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(a, st);
    k_dmma_probe<<<sms * 2, 256, 0, st>>>(out, iters);
    cudaEventRecord(b, st);
    e = cudaEventSynchronize(b);
    if (e != cudaSuccess) break;
    float t; cudaEventElapsedTime(&t, a, b);
    if (r > 0 && t < best) best = t;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(out);
  if (e != cudaSuccess) return (int)e;
  *tflops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * 8.0 * (double)sms * 2 / ((double)best * 1e-3) * 1e-12;
  return 0;
}

}  // namespace ipmz
