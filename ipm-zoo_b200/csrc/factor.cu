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

#include <vector>

#include "ipmz_device.cuh"
#include "ipmz_kernels.h"

namespace ipmz {

namespace {

constexpr int NB = 128;  // panel width
constexpr int SB = 32;   // sub-block factored by one warp
constexpr int SP = 132;  // shared-memory pitch: 132 mod 16 == 4 -> DMMA fragment loads (row = lane/4,
                         // k = lane%4) of a half-warp hit 16 distinct 8-byte banks
constexpr int RB = 64;   // panel rows per CTA in k_trsm_panel

constexpr size_t DIAG_SMEM = (size_t)(NB * SP + 2 * NB) * sizeof(double);
constexpr size_t TRSM_SMEM = (size_t)((NB + RB) * SP + 2 * NB) * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Asynchronous copy of a [rows x cols] block (global, leading dimension ld) into shared memory
// (pitch SP); with LOWER only columns <= row are fetched.  16-byte cp.async chunks, all in
// flight at once (a plain load/store loop would serialise on the L2 latency).
template <bool LOWER>
__device__ __forceinline__ void async_block_load(double* S, const double* A, int ld, int rows, int cols, int tid) {
  const int half = (cols + 1) >> 1;
  for (int idx = tid; idx < rows * half; idx += 256) {
    const int r = idx / half, c = (idx - r * half) * 2;
    if (LOWER && c > r) continue;
    cp_async16(S + r * SP + c, A + (size_t)r * ld + c, (cols - c >= 2) ? 16 : 8);
  }
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// In shared memory, on the FP64 tensor pipe:  C[rows x cols] -= A[rows x kb] diag(d) B[cols x kb]^T.
// Each warp takes 16 x 16 output micro-tiles (4 DMMA accumulators).  LOWER: C is the lower
// triangle of a square block (A and B index the same rows) and only tiles on/below the
// diagonal are touched.
template <bool LOWER>
__device__ __forceinline__ void smem_update(double* C, const double* A, const double* B, const double* d,
                                            int rows, int cols, int kb, int warp, int lane, int nwarps) {
  const int g = lane >> 2, q = lane & 3;
  const int tm = (rows + 15) >> 4, tn = (cols + 15) >> 4;
  for (int t = warp; t < tm * tn; t += nwarps) {
    const int mi = t / tn, ni = t - mi * tn;
    if (LOWER && ni > mi) continue;
    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    const int ra = mi * 16 + g, rb = ni * 16 + g;
    for (int k = 0; k < kb; k += 4) {
      const int kk = k + q;
      const bool kok = kk < kb;
      const double dk = kok ? d[kk] : 0.0;
      double af[2], bf[2];
      af[0] = (kok && ra < rows) ? A[ra * SP + kk] : 0.0;
      af[1] = (kok && ra + 8 < rows) ? A[(ra + 8) * SP + kk] : 0.0;
      bf[0] = (kok && rb < cols) ? B[rb * SP + kk] * dk : 0.0;
      bf[1] = (kok && rb + 8 < cols) ? B[(rb + 8) * SP + kk] * dk : 0.0;
      dmma884(acc[0][0], af[0], bf[0]);
      dmma884(acc[0][1], af[0], bf[1]);
      dmma884(acc[1][0], af[1], bf[0]);
      dmma884(acc[1][1], af[1], bf[1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = mi * 16 + i * 8 + g, col = ni * 16 + j * 8 + 2 * q;
        if (row < rows) {
          if (col < cols && (!LOWER || col <= row)) C[row * SP + col] -= acc[i][j][0];
          if (col + 1 < cols && (!LOWER || col + 1 <= row)) C[row * SP + col + 1] -= acc[i][j][1];
        }
      }
  }
}

// One thread owns one row of X (D L^T) = A restricted to a 32-wide column block whose unit-lower
// factor Lb (pitch SP) and pivots are in shared memory: right-looking substitution, so the 31-c
// updates of step c are independent FMAs.
__device__ __forceinline__ void row_solve32(double* row, const double* Lb, const double* dinv, int cb) {
  double v[SB];
#pragma unroll
  for (int c = 0; c < SB; ++c) v[c] = c < cb ? row[c] : 0.0;
#pragma unroll
  for (int c = 0; c < SB; ++c) {
    if (c < cb) {
      const double w = v[c];  // = L_rc * d_c
#pragma unroll
      for (int c2 = 0; c2 < SB; ++c2)
        if (c2 > c && c2 < cb) v[c2] -= w * Lb[c2 * SP + c];
      v[c] = w * dinv[c];
    }
  }
#pragma unroll
  for (int c = 0; c < SB; ++c)
    if (c < cb) row[c] = v[c];
}

__global__ void __launch_bounds__(256) k_diag_ldlt(const double* src, double* dst,
                                                   int ld, size_t sK, double* __restrict__ Dg, size_t sD,
                                                   int k0, int nb, const int* __restrict__ active) {
  extern __shared__ double sm[];
  double* S = sm;
  double* dsm = sm + NB * SP;
  double* dinv = dsm + NB;
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const double* A = src + (size_t)p * sK + (size_t)k0 * ld + k0;
  double* O = dst + (size_t)p * sK + (size_t)k0 * ld + k0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  async_block_load<true>(S, A, ld, nb, nb, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  for (int j0 = 0; j0 < nb; j0 += SB) {
    const int jb = min(SB, nb - j0);
    if (warp == 0) {
      // lane r holds row r of the sub-block; column c is exchanged with shuffles
      double a[SB];
#pragma unroll
      for (int c = 0; c < SB; ++c) a[c] = (lane < jb && c <= lane) ? S[(j0 + lane) * SP + j0 + c] : 0.0;
#pragma unroll
      for (int c = 0; c < SB; ++c) {
        if (c < jb) {
          const double acol = a[c];
          double d = __shfl_sync(0xffffffffu, acol, c);
          if (d == 0.0) d = 1e-8;  // LinearSolvers.cpp:28
          const double rinv = __drcp_rn(d);
          const double l = acol * rinv;
#pragma unroll
          for (int c2 = c + 1; c2 < SB; ++c2) {
            const double o = __shfl_sync(0xffffffffu, acol, c2);
            a[c2] -= l * o;
          }
          if (lane == c) { dsm[j0 + c] = d; dinv[j0 + c] = rinv; }
          if (lane > c) a[c] = l;
        }
      }
#pragma unroll
      for (int c = 0; c < SB; ++c)
        if (lane < jb && c < lane) S[(j0 + lane) * SP + j0 + c] = a[c];
    }
    __syncthreads();
    const int base = j0 + jb, rem = nb - base;
    if (tid < rem) row_solve32(S + (base + tid) * SP + j0, S + j0 * SP + j0, dinv + j0, jb);
    __syncthreads();
    if (rem > 0)
      smem_update<true>(S + base * SP + base, S + base * SP + j0, S + base * SP + j0, dsm + j0, rem, rem, jb,
                        warp, lane, 8);
    __syncthreads();
  }

  for (int r = warp; r < nb; r += 8) {
    for (int c = lane; c < r; c += 32) O[(size_t)r * ld + c] = S[r * SP + c];
    if (lane == 0) O[(size_t)r * ld + r] = dsm[r];
  }
  for (int t = tid; t < nb; t += 256) Dg[(size_t)p * sD + k0 + t] = dsm[t];
}

__global__ void __launch_bounds__(256) k_trsm_panel(const double* src, double* dst,
                                                    int ld, size_t sK, const double* __restrict__ Dg, size_t sD,
                                                    int k0, int nb, int N, const int* __restrict__ active) {
  extern __shared__ double sm[];
  double* S = sm;                  // L_kk (strict lower)
  double* T = sm + NB * SP;        // this CTA's rows of the panel
  double* dsm = T + RB * SP;
  double* dinv = dsm + NB;
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = k0 + nb + blockIdx.x * RB;
  const int nr = min(RB, N - r0);
  const double* Lkk = dst + (size_t)p * sK + (size_t)k0 * ld + k0;
  const double* Ain = src + (size_t)p * sK + (size_t)r0 * ld + k0;
  double* Aout = dst + (size_t)p * sK + (size_t)r0 * ld + k0;

  async_block_load<true>(S, Lkk, ld, nb, nb, tid);
  async_block_load<false>(T, Ain, ld, nr, nb, tid);
  cp_async_commit();
  for (int t = tid; t < nb; t += 256) {
    const double d = Dg[(size_t)p * sD + k0 + t];
    dsm[t] = d;
    dinv[t] = __drcp_rn(d);
  }
  cp_async_wait<0>();
  __syncthreads();

  for (int c0 = 0; c0 < nb; c0 += SB) {
    const int cb = min(SB, nb - c0);
    if (tid < nr) row_solve32(T + tid * SP + c0, S + c0 * SP + c0, dinv + c0, cb);
    __syncthreads();
    const int base = c0 + cb, rem = nb - base;
    if (rem > 0)
      smem_update<false>(T + base, T + c0, S + base * SP + c0, dsm + c0, nr, rem, cb, warp, lane, 8);
    __syncthreads();
  }

  for (int r = warp; r < nr; r += 8)
    for (int c = lane; c < nb; c += 32) Aout[(size_t)r * ld + c] = T[r * SP + c];
}

// ------------------------------------------------------------------------------------------
// DMMA SYRK:  Cout(lower) = Cin + sign * P diag(d) P^T
//
// CTA tile BM x BN (BM = 2 BN), 8 warps as WM x WN, each warp (BM/WM) x (BN/WN) built from
// m8n8k4 DMMAs.  Two CTAs are resident per SM (<= 128 registers, ~92 KB shared memory each), so
// while one CTA waits for its C tile or for a cp.async stage the other keeps the FP64 tensor
// pipe busy.  Operands: 3-stage cp.async pipeline of BK = 16 wide k-slices, rows padded to 20
// doubles so fragment loads are bank-conflict free.
constexpr int BM = 128, BN = 64, BK = 16, STAGES = 3, WARPS_M = 4, WARPS_N = 2;
constexpr int MI = BM / (WARPS_M * 8), NI = BN / (WARPS_N * 8);
constexpr int LDT = BK + 4;
constexpr int TILE_RATIO = BM / BN;
constexpr size_t SYRK_SMEM = (size_t)(STAGES * (BM + BN) * LDT + STAGES * BK) * sizeof(double);
static_assert(((BM + BN) * (BK / 2)) % 256 == 0, "stage loads must divide evenly over 256 threads");

struct SyrkArgs {
  const double* Cin;
  double* Cout;
  int ldc;
  size_t sC;
  const double* P;
  int ldp;
  size_t sP;
  const double* d;
  size_t sd;
  int rows, kdim;
  double sign;
  const int* active;
  int tn;        // number of BN-wide tile columns
  int tj_limit;  // only tile columns < tj_limit are updated (look-ahead split); <= 0: all
  int tj_first;  // first tile column handled by this launch
};

__global__ void __launch_bounds__(256, 2) k_syrk_ldl(SyrkArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = As + STAGES * BM * LDT;
  double* ds = Bs + STAGES * BN * LDT;

  const int p = a.active ? a.active[blockIdx.y] : blockIdx.y;
  // linear index over the tiles on/below the diagonal: row block ti holds TILE_RATIO*(ti+1)
  // column tiles (the last row block is clipped to tn)
  const int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)t / TILE_RATIO + 1.0) - 1.0) * 0.5);
  while (TILE_RATIO * ti * (ti + 1) / 2 > t) --ti;
  while (TILE_RATIO * (ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const int tj = t - TILE_RATIO * ti * (ti + 1) / 2;
  if (tj >= a.tn) return;
  if (tj < a.tj_first || (a.tj_limit > 0 && tj >= a.tj_limit)) return;
  const int row0 = ti * BM, col0 = tj * BN;

  const double* P = a.P + (size_t)p * a.sP;
  const double* dv = a.d + (size_t)p * a.sd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int g = lane >> 2, q = lane & 3;
  const int wrow = row0 + wm * (MI * 8), wcol = col0 + wn * (NI * 8);

  const int KT = (a.kdim + BK - 1) / BK;
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
      const double* srcp = P + (size_t)(ok ? gr : 0) * a.ldp + (ok ? k : 0);
      double* dstp = isA ? Asd + r * LDT + ck : Bsd + (r - BM) * LDT + ck;
      cp_async16(dstp, srcp, ok ? 16 : 0);
    }
    if (tid < 8) {
      const int k = kbase + tid * 2;
      const bool ok = k < a.kdim;
      cp_async16(ds + stage * BK + tid * 2, dv + (ok ? k : 0), ok ? 16 : 0);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  // Accumulators start from sign*C (all loads independent and in flight while the cp.async
  // prologue lands), so the epilogue is store-only:  Cout = sign * (sign*Cin + P d P^T).
  const double* Cin = a.Cin + (size_t)p * a.sC;
  double* Cout = a.Cout + (size_t)p * a.sC;
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
    const double* dw = ds + stage * BK + q;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double af[MI], bf[NI];
      const double dk = dw[kk * 4];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) af[mi] = Aw[mi * 8 * LDT + kk * 4];
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) bf[ni] = Bw[ni * 8 * LDT + kk * 4] * dk;
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

}  // namespace

int factor_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_diag_ldlt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_syrk_ldl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYRK_SMEM);
  return (int)e;
}

void launch_syrk_ldl(cudaStream_t st, int nslots, const int* active, const double* Cin, double* Cout, int ldc,
                     size_t sC, const double* P, int ldp, size_t sP, const double* d, size_t sd, int rows,
                     int kdim, double sign) {
  if (rows <= 0 || kdim <= 0 || nslots <= 0) return;
  const int T = (rows + BM - 1) / BM;
  SyrkArgs a{Cin, Cout, ldc, sC, P, ldp, sP, d, sd, rows, kdim, sign, active, (rows + BN - 1) / BN, 0, 0};
  dim3 grid(TILE_RATIO * T * (T + 1) / 2, nslots);
  k_syrk_ldl<<<grid, 256, SYRK_SMEM, st>>>(a); count_launch();
}

void launch_ldlt(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg) {
  for (int k0 = 0; k0 < fp.N; k0 += NB) {
    const int nb = fp.N - k0 < NB ? fp.N - k0 : NB;
    const double* in = (k0 == 0) ? src : dst;
    k_diag_ldlt<<<dim3(1, fp.nslots), 256, DIAG_SMEM, st>>>(in, dst, fp.ld, fp.sK, Dg, fp.sD, k0, nb, fp.active); count_launch();
    const int rem = fp.N - k0 - nb;
    if (rem > 0) {
      k_trsm_panel<<<dim3((rem + RB - 1) / RB, fp.nslots), 256, TRSM_SMEM, st>>>(in, dst, fp.ld, fp.sK, Dg, fp.sD,
                                                                                 k0, nb, fp.N, fp.active); count_launch();
      const size_t off = (size_t)(k0 + nb) * fp.ld + (k0 + nb);
      launch_syrk_ldl(st, fp.nslots, fp.active, in + off, dst + off, fp.ld, fp.sK,
                      dst + (size_t)(k0 + nb) * fp.ld + k0, fp.ld, fp.sK, Dg + k0, fp.sD, rem, nb, -1.0);
    }
  }
}


int launch_ldlt_profiled(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg,
                         double ms[3], double* flops_syrk, int* n_syrk) {
  struct Rec { cudaEvent_t a, b; int kind; };
  std::vector<Rec> recs;
  auto begin = [&](int kind) {
    Rec r; r.kind = kind;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    recs.push_back(r);
  };
  auto end = [&]() { cudaEventRecord(recs.back().b, st); };
  for (int k0 = 0; k0 < fp.N; k0 += NB) {
    const int nb = fp.N - k0 < NB ? fp.N - k0 : NB;
    const double* in = (k0 == 0) ? src : dst;
    begin(0);
    k_diag_ldlt<<<dim3(1, fp.nslots), 256, DIAG_SMEM, st>>>(in, dst, fp.ld, fp.sK, Dg, fp.sD, k0, nb, fp.active);
    count_launch();
    end();
    const int rem = fp.N - k0 - nb;
    if (rem > 0) {
      begin(1);
      k_trsm_panel<<<dim3((rem + RB - 1) / RB, fp.nslots), 256, TRSM_SMEM, st>>>(in, dst, fp.ld, fp.sK, Dg, fp.sD,
                                                                                 k0, nb, fp.N, fp.active);
      count_launch();
      end();
      const size_t off = (size_t)(k0 + nb) * fp.ld + (k0 + nb);
      begin(2);
      launch_syrk_ldl(st, fp.nslots, fp.active, in + off, dst + off, fp.ld, fp.sK,
                      dst + (size_t)(k0 + nb) * fp.ld + k0, fp.ld, fp.sK, Dg + k0, fp.sD, rem, nb, -1.0);
      end();
      if (flops_syrk) *flops_syrk += (double)fp.nslots * (double)rem * (double)rem * (double)nb;
      if (n_syrk) *n_syrk += 1;
    }
  }
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
