// ipm-zoo_b200/csrc/ldlt_device.cuh -- device building blocks shared by the multi-kernel LDL^T
// schedule (factor.cu) and the persistent dataflow factorization (dataflow.cu): cp.async and
// mbarrier wrappers, the m8n8k4 FP64 MMA, the one-warp 32 x 32 LDL^T, the tensor-pipe panel
// solve with 8 x 8 inverse blocks and the in-shared-memory rank-k update.
#pragma once
#include "ipmz_device.cuh"

namespace ipmz {
namespace {

constexpr int NB = 128;  // panel width
constexpr int SB = 32;   // sub-block factored by one warp
constexpr int SP = 132;  // shared-memory pitch: 132 mod 16 == 4 -> DMMA fragment loads (row = lane/4,
                         // k = lane%4) of a half-warp hit 16 distinct 8-byte banks
constexpr int RB = 64;   // panel rows per CTA in k_trsm_panel
constexpr int WLD = 256; // leading dimension of the scaled-panel buffer W = L D (two NB-wide panels)

constexpr size_t DIAG_SMEM = (size_t)(NB * SP + 2 * NB + 8 * 32 + 16 * 96) * sizeof(double);
constexpr size_t TRSM_SMEM = (size_t)((NB + RB) * SP + NB + 16 * 96) * sizeof(double);

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
template <bool LOWER, int NT = 256>
__device__ __forceinline__ void async_block_load(double* S, const double* A, int ld, int rows, int cols, int tid) {
  if (cols == NB) {  // full-width block: 64 chunks per row, index arithmetic by shifts
    for (int idx = tid; idx < rows * (NB / 2); idx += NT) {
      const int r = idx >> 6, c = (idx & 63) * 2;
      if (LOWER && c > r) continue;
      cp_async16(S + r * SP + c, A + (size_t)r * ld + c, 16);
    }
    return;
  }
  const int half = (cols + 1) >> 1;
  for (int idx = tid; idx < rows * half; idx += NT) {
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

// 1/d to <= 1 ulp: MUFU.RCP64H seed + two Newton steps (~50 cycles on the dependent chain of the
// pivots; the IEEE division sequence is ~70).
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// In shared memory, on the FP64 tensor pipe:  C[rows x cols] -= A[rows x kb] diag(d) B[cols x kb]^T.
// Each warp takes 16 x 16 output micro-tiles (4 DMMA accumulators).  LOWER: C is the lower
// triangle of a square block (A and B index the same rows) and only tiles on/below the
// diagonal are touched.
template <bool LOWER, int P = SP>
__device__ __forceinline__ void smem_update(double* C, const double* A, const double* B, const double* d,
                                            int rows, int cols, int kb, int warp, int lane, int nwarps) {
  const int g = lane >> 2, q = lane & 3;
  const int tm = (rows + 15) >> 4, tn = (cols + 15) >> 4;
  for (int t = warp; t < tm * tn; t += nwarps) {
    const int mi = t / tn, ni = t - mi * tn;
    if (LOWER && ni > mi) continue;
    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    const int ra = mi * 16 + g, rb = ni * 16 + g;
    if (kb == SB && mi * 16 + 16 <= rows && ni * 16 + 16 <= cols) {  // warp-uniform: mma.sync needs all lanes
      // full interior tile: all fragment loads of the 8 k-steps in flight, then 32 DMMAs
      double af[8][2], bf[8][2];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const double dk = d[s * 4 + q];
        af[s][0] = A[ra * P + s * 4 + q];
        af[s][1] = A[(ra + 8) * P + s * 4 + q];
        bf[s][0] = B[rb * P + s * 4 + q] * dk;
        bf[s][1] = B[(rb + 8) * P + s * 4 + q] * dk;
      }
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        dmma884(acc[0][0], af[s][0], bf[s][0]);
        dmma884(acc[0][1], af[s][0], bf[s][1]);
        dmma884(acc[1][0], af[s][1], bf[s][0]);
        dmma884(acc[1][1], af[s][1], bf[s][1]);
      }
    } else
    for (int k = 0; k < kb; k += 4) {
      const int kk = k + q;
      const bool kok = kk < kb;
      const double dk = kok ? d[kk] : 0.0;
      double af[2], bf[2];
      af[0] = (kok && ra < rows) ? A[ra * P + kk] : 0.0;
      af[1] = (kok && ra + 8 < rows) ? A[(ra + 8) * P + kk] : 0.0;
      bf[0] = (kok && rb < cols) ? B[rb * P + kk] * dk : 0.0;
      bf[1] = (kok && rb + 8 < cols) ? B[(rb + 8) * P + kk] * dk : 0.0;
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
          if (col < cols && (!LOWER || col <= row)) C[row * P + col] -= acc[i][j][0];
          if (col + 1 < cols && (!LOWER || col + 1 <= row)) C[row * P + col + 1] -= acc[i][j][1];
        }
      }
  }
}

constexpr int IP = 12;               // pitch of an 8 x 8 inverse block (conflict-free B fragments)
constexpr int INV_BLK = 8 * IP;      // doubles per inverse block
constexpr int INV_SUB = 4 * INV_BLK; // per 32-wide sub-block

// Unpivoted LDL^T of one 32 x 32 diagonal sub-block by ONE warp: lane r keeps row r in registers.
// Columns are eliminated FOUR at a time: the lanes exchange their four current column entries
// through shared memory (one store, broadcast loads), every lane factors the 4 x 4 diagonal block
// redundantly (the dependent chain of the pivots: 4 reciprocals), solves its own row against it
// and, after one more exchange of the unscaled entries w = l d, applies the rank-4 update to its
// row.  One exchange round trip per 4 pivots instead of per pivot.  Then the 8 x 8 blocks
// Binv_b = D_b^-1 L_bb^-1 (b = 0..3) are formed so that the rows below are solved on the tensor
// pipe:  X_b = R_b Binv_b^T  with  R_b = A_b - sum_{l<b} X_l D_l L_bl^T.
constexpr int CBUF = 8 * SB;  // two 32 x 4 exchange buffers
__device__ __forceinline__ double pivot_of(double d) { return d == 0.0 ? 1e-8 : d; }  // LinearSolvers.cpp:28
template <int P = SP>
__device__ __forceinline__ void warp_ldlt32(double* S, int j0, int jb, double* dsm, double* dinv, double* colbuf,
                                            double* binv, int lane) {
  double a[SB];
#pragma unroll
  for (int c = 0; c < SB; ++c) a[c] = (lane < jb && c <= lane) ? S[(j0 + lane) * P + j0 + c] : 0.0;
  double* xbuf = colbuf;            // [32][4] current entries of the four columns
  double* wbuf = colbuf + 4 * SB;   // [32][4] the same after elimination inside the block (w = l d)
#pragma unroll
  for (int cb = 0; cb < SB / 4; ++cb) {
    constexpr int dummy = 0; (void)dummy;
    const int c0 = 4 * cb;
    if (c0 < jb) {
      *reinterpret_cast<double2*>(xbuf + 4 * lane) = make_double2(a[c0], a[c0 + 1]);
      *reinterpret_cast<double2*>(xbuf + 4 * lane + 2) = make_double2(a[c0 + 2], a[c0 + 3]);
      __syncwarp();
      // 4 x 4 diagonal block (rows c0..c0+3), lower part, unscaled
      const double p00 = xbuf[4 * c0];
      const double2 p1 = *reinterpret_cast<const double2*>(xbuf + 4 * (c0 + 1));      // p10 p11
      const double2 p2a = *reinterpret_cast<const double2*>(xbuf + 4 * (c0 + 2));     // p20 p21
      const double p22 = xbuf[4 * (c0 + 2) + 2];
      const double2 p3a = *reinterpret_cast<const double2*>(xbuf + 4 * (c0 + 3));     // p30 p31
      const double2 p3b = *reinterpret_cast<const double2*>(xbuf + 4 * (c0 + 3) + 2); // p32 p33
      const double d0 = pivot_of(p00), r0 = fast_rcp(d0);
      const double l10 = p1.x * r0, l20 = p2a.x * r0, l30 = p3a.x * r0;
      const double d1 = pivot_of(fma(-l10, p1.x, p1.y)), r1 = fast_rcp(d1);
      const double t21 = fma(-l20, p1.x, p2a.y), t31 = fma(-l30, p1.x, p3a.y);
      const double l21 = t21 * r1, l31 = t31 * r1;
      const double d2 = pivot_of(fma(-l21, t21, fma(-l20, p2a.x, p22))), r2 = fast_rcp(d2);
      const double t32 = fma(-l31, t21, fma(-l30, p2a.x, p3b.x));
      const double l32 = t32 * r2;
      const double d3 = pivot_of(fma(-l32, t32, fma(-l31, t31, fma(-l30, p3a.x, p3b.y)))), r3 = fast_rcp(d3);
      // this lane's row against the block: l_k = w_k / d_k, w_k = x_k - sum_{m<k} l_m (w of row c0+k)_m
      const double w0 = a[c0], li0 = w0 * r0;
      const double w1 = fma(-li0, p1.x, a[c0 + 1]), li1 = w1 * r1;
      const double w2 = fma(-li1, t21, fma(-li0, p2a.x, a[c0 + 2])), li2 = w2 * r2;
      const double w3 = fma(-li2, t32, fma(-li1, t31, fma(-li0, p3a.x, a[c0 + 3]))), li3 = w3 * r3;
      *reinterpret_cast<double2*>(wbuf + 4 * lane) = make_double2(w0, w1);
      *reinterpret_cast<double2*>(wbuf + 4 * lane + 2) = make_double2(w2, w3);
      if (lane == c0) {
        *reinterpret_cast<double2*>(dsm + j0 + c0) = make_double2(d0, d1);
        *reinterpret_cast<double2*>(dsm + j0 + c0 + 2) = make_double2(d2, d3);
        *reinterpret_cast<double2*>(dinv + j0 + c0) = make_double2(r0, r1);
        *reinterpret_cast<double2*>(dinv + j0 + c0 + 2) = make_double2(r2, r3);
      }
      if (lane > c0) a[c0] = li0;
      if (lane > c0 + 1) a[c0 + 1] = li1;
      if (lane > c0 + 2) a[c0 + 2] = li2;
      if (lane > c0 + 3) a[c0 + 3] = li3;
      __syncwarp();
      // rank-4 update of this lane's row: a[j] -= sum_k l_k w_jk  (the next block's columns first)
#pragma unroll
      for (int j = c0 + 4; j < SB; ++j) {
        const double2 wa = *reinterpret_cast<const double2*>(wbuf + 4 * j);
        const double2 wb = *reinterpret_cast<const double2*>(wbuf + 4 * j + 2);
        a[j] = fma(-li3, wb.y, fma(-li2, wb.x, fma(-li1, wa.y, fma(-li0, wa.x, a[j]))));
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SB; ++c)
    if (lane < jb && c < lane) S[(j0 + lane) * P + j0 + c] = a[c];
  __syncwarp();
  // lane (b, j): column j of L_bb^-1 by forward substitution, scaled by D^-1
  {
    const int bb = lane >> 3, j = lane & 7;
    const double* Lb = S + (j0 + 8 * bb) * P + j0 + 8 * bb;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < i) s += ((8 * bb + i < jb && k >= j) ? Lb[i * P + k] : 0.0) * x[k];
      if (i > j) x[i] = -s;
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const double dn = (8 * bb + n < jb) ? dinv[j0 + 8 * bb + n] : 0.0;
      binv[bb * INV_BLK + n * IP + j] = (n >= j) ? dn * x[n] : 0.0;
    }
  }
}

// Rows [0, nrows) of T (shared, pitch SP), columns [c0, c0+32):  X (D L^T) = A on the tensor
// pipe.  L = the unit-lower 32 x 32 block at Lb (pitch SP), d its pivots, binv its four
// D^-1 L^-1 blocks.  Each warp owns 16-row tiles and runs the four 8-column stages on them
// without block-level synchronisation (rows are independent).
template <int P = SP>
__device__ __forceinline__ void panel_solve32(double* T, int nrows, const double* Lb, const double* d,
                                              const double* binv, int warp, int lane, int nwarps) {
  const int g = lane >> 2, q = lane & 3;
  for (int mt = warp; mt * 16 < nrows; mt += nwarps) {
    double* T0 = T + (mt * 16 + g) * P;
    double* T1 = T0 + 8 * P;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double acc0[2], acc1[2];
      acc0[0] = T0[8 * b + 2 * q]; acc0[1] = T0[8 * b + 2 * q + 1];
      acc1[0] = T1[8 * b + 2 * q]; acc1[1] = T1[8 * b + 2 * q + 1];
      if (b > 0) {
        double af0[6], af1[6], bf[6];
#pragma unroll
        for (int s = 0; s < 2 * b; ++s) {
          const int k = 4 * s + q;
          af0[s] = T0[k];
          af1[s] = T1[k];
          bf[s] = -(Lb[(8 * b + g) * P + k] * d[k]);
        }
#pragma unroll
        for (int s = 0; s < 2 * b; ++s) {
          dmma884(acc0, af0[s], bf[s]);
          dmma884(acc1, af1[s], bf[s]);
        }
        T0[8 * b + 2 * q] = acc0[0]; T0[8 * b + 2 * q + 1] = acc0[1];
        T1[8 * b + 2 * q] = acc1[0]; T1[8 * b + 2 * q + 1] = acc1[1];
        __syncwarp();
      }
      double x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
      const double r00 = T0[8 * b + q], r01 = T0[8 * b + 4 + q];
      const double r10 = T1[8 * b + q], r11 = T1[8 * b + 4 + q];
      const double i0 = binv[b * INV_BLK + g * IP + q], i1 = binv[b * INV_BLK + g * IP + 4 + q];
      dmma884(x0, r00, i0);
      dmma884(x1, r10, i0);
      dmma884(x0, r01, i1);
      dmma884(x1, r11, i1);
      __syncwarp();
      T0[8 * b + 2 * q] = x0[0]; T0[8 * b + 2 * q + 1] = x0[1];
      T1[8 * b + 2 * q] = x1[0]; T1[8 * b + 2 * q + 1] = x1[1];
      __syncwarp();
    }
  }
}


constexpr int BK = 16;        // k-slice of the DMMA update kernels
constexpr int LDT = BK + 4;   // shared-memory pitch of a k-slice row (conflict-free fragment loads)

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared.b64 st, [%0];\n}\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned long long* bar) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WS_WAIT:\n"
      "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
      "@p bra WS_DONE;\n"
      "bra WS_WAIT;\n"
      "WS_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}


// Bounded wait: gives up after ~2^26 polls (seconds) and raises *sticky, so that a transfer that never
// completes (a faulting bulk copy, a protocol bug) ends the kernel with an error instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned parity, int* sticky) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
  }
  if (sticky) atomicExch(sticky, 1);
}

}  // namespace
}  // namespace ipmz
