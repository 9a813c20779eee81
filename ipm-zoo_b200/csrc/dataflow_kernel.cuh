// ipm-zoo_b200/csrc/dataflow_kernel.cuh -- the persistent dataflow LDL^T kernel (see dataflow.cu for the
// design).  Included by exactly two translation units: dataflow.cu (DF_TMA = 0, cp.async operand loads) and
// dataflow_tma.cu (DF_TMA = 1, TMA operand loads); each gets its own kernel in its anonymous namespace.
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "ipmz_kernels.h"
#include "ldlt_device.cuh"
#include "ldlt_schedule.hpp"

namespace ipmz {

namespace {

#ifndef DF_TMA
#define DF_TMA 0
#endif
// DF_TMA = 1: the UPD operands are moved by TMA (cp.async.bulk.tensor.2d, one elected producer thread, full
// barriers completed by transaction bytes, 128-byte swizzled k-slices: no row padding, 6 stages);
// DF_TMA = 0: 16-byte cp.async by the four producer warps, rows padded to 20 doubles, 5 stages.
constexpr int DF_STAGES = DF_TMA ? 6 : 5;
constexpr int DF_CONSUMERS = 8, DF_PRODUCERS = 4;
constexpr int DF_MI = 8, DF_NI = 4;  // consumer warp tile = 64 x 32 (2 x 4 warps)
constexpr int DF_CONSUMER_REGS = 232, DF_PRODUCER_REGS = 40;
constexpr int DF_LAG = 2;  // k-slices consumer group 1 trails group 0 (<= DF_STAGES - 2)  // setmaxnreg after the role split
constexpr int DF_CTHREADS = DF_CONSUMERS * 32;
constexpr int DF_PTHREADS = DF_PRODUCERS * 32;
constexpr int DF_THREADS = DF_CTHREADS + DF_PTHREADS;
constexpr int DF_OP_DOUBLES = DF_TILE * (DF_TMA ? BK : LDT);  // one operand slice (128 rows)
constexpr int DF_STAGE_DOUBLES = 2 * DF_OP_DOUBLES;
constexpr int DF_RING_DOUBLES = DF_STAGES * DF_STAGE_DOUBLES;
constexpr int DF_BULK_DOUBLES = NB * SP + RB * SP + 2 * NB + CBUF + 16 * 96;
constexpr int DF_DATA_DOUBLES = DF_RING_DOUBLES > DF_BULK_DOUBLES ? DF_RING_DOUBLES : DF_BULK_DOUBLES;
constexpr size_t DF_SMEM = (size_t)DF_DATA_DOUBLES * sizeof(double) + 256;
static_assert(DF_TILE == NB && DF_HALF == RB, "tile grid of the schedule = panel geometry of the kernels");
static_assert(DF_SMEM <= 232448, "shared memory per CTA");

struct DfArgs {
  const double* src;
  double* dst;
  double* W;
  double* Dg;
  double* Ginv;
  const int4* tasks;
  int* ticket;
  int* abort;
  int* sticky;  // never reset: set together with abort, read by the host after a solve
  int* rdy;
  int* cnt;
  long long* tlog;
  int N, ld, nt, ntasks;
};

__device__ __forceinline__ void bar_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void csync() { bar_named(2, DF_CTHREADS); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
__device__ __forceinline__ int smid() {
  int s;
  asm volatile("mov.u32 %0, %%smid;\n" : "=r"(s));
  return s;
}

__device__ __forceinline__ int need_of(const DfArgs& a, int i) { return (a.N - i * DF_TILE) > DF_HALF ? 2 : 1; }

// Executed by the first producer warp: wait until every input of `tk` has been published.
// One lane per flag; a watchdog turns a scheduler/protocol bug into an error instead of a hang.
__device__ __forceinline__ void wait_deps(const DfArgs& a, const int4 tk, int lane) {
  const int type = tk.x & 0xff, i = tk.y, j = tk.z;
  const int* flag = nullptr;
  int want = 0;
  if (type == DF_DIAG) {
    if (lane == 0) { flag = a.cnt + (size_t)i * a.nt + j; want = j; }
  } else if (type == DF_TRSM) {
    if (lane == 0) { flag = a.cnt + (size_t)i * a.nt + j; want = j; }
    if (lane == 1) { flag = a.rdy + (size_t)j * a.nt + j; want = 1; }
  } else {
    const int k0 = tk.w & 0xffff, k1 = tk.w >> 16;
    if (lane == 0) { flag = a.cnt + (size_t)i * a.nt + j; want = k0; }
    const int l = lane - 1;
    if (l >= 0 && (l >> 1) < k1 - k0) {
      const int k = k0 + (l >> 1), r = (l & 1) ? j : i;
      flag = a.rdy + (size_t)k * a.nt + r;
      want = need_of(a, r);
    }
  }
  long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    const bool ok = flag == nullptr || ld_acquire(flag) >= want;
    if (__all_sync(0xffffffffu, ok)) break;
    __nanosleep(64);
    if ((spin & 1023u) == 1023u) {
      if (t0 == 0) t0 = gtimer();
      const bool bail = ld_acquire(a.abort) != 0 || gtimer() - t0 > 4000000000LL;
      if (__any_sync(0xffffffffu, bail)) {
        if (lane == 0) { atomicExch(a.abort, 1); atomicExch(a.sticky, 1); }
        break;
      }
    }
  }
  __syncwarp();  // lane 0 publishes the task: order it after every lane's acquire
}

// ---- DIAG(k): LDL^T of the diagonal tile, 16 consumer warps (body of k_diag_ldlt, factor.cu) ----
// preloaded: the updated tile is already in S (fused DIAGU task: written from the accumulators)
__device__ __forceinline__ void df_diag_task(const DfArgs& a, double* sm, int k, int tid, long long* ph, bool preloaded) {
  double* S = sm;
  double* dsm = sm + NB * SP + RB * SP;
  double* dinv = dsm + NB;
  double* colbuf = dinv + NB;
  double* binv = colbuf + CBUF;
  const int lane = tid & 31, warp = tid >> 5;
  const int k0 = k * NB;
  const int nb = min(NB, a.N - k0);
  const double* A = (k == 0 ? a.src : a.dst) + (size_t)k0 * a.ld + k0;
  double* O = a.dst + (size_t)k0 * a.ld + k0;

  const long long c_begin = ph ? clock64() : 0;
  long long c_ldlt = 0;
  if (!preloaded) {
    if (nb < NB) {  // ragged tile: partial sub-blocks must not read uninitialised shared memory
      for (int i = tid; i < NB * SP; i += DF_CTHREADS) S[i] = 0.0;
      csync();
    }
    async_block_load<true, DF_CTHREADS>(S, A, a.ld, nb, nb, tid);
    cp_async_commit();
    cp_async_wait<0>();
    csync();
  }
  const long long c_loaded = ph ? clock64() : 0;
  // Look-ahead inside the tile: after the 32 x 32 sub-block j0 is factored, all warps first bring the NEXT
  // diagonal sub-block up to date (its 32 rows of the panel solve + its own update); then warp 0 factors it
  // while warps 1-7 solve and update the rest of the trailing part of step j0.  The dependent chain of the
  // tile is 4 one-warp factorizations + 3 short critical parts instead of 4 full solve/update rounds.
  for (int j0 = 0; j0 < nb; j0 += SB) {
    const int jb = min(SB, nb - j0);
    if (warp == 0) {
      const long long c0 = ph ? clock64() : 0;
      warp_ldlt32(S, j0, jb, dsm, dinv, colbuf, binv + (j0 / SB) * INV_SUB, lane);
      if (ph) c_ldlt += clock64() - c0;
    }
    csync();  // L32 / inverse blocks of j0 are there, and every update of step j0 - 32 has landed
    const int base = j0 + jb, rem = nb - base;
    if (rem > 0) {
      const double* Lb = S + j0 * SP + j0;
      const double* bi = binv + (j0 / SB) * INV_SUB;
      const int crit = min(SB, rem);  // rows / columns of the next diagonal sub-block
      panel_solve32(S + base * SP + j0, crit, Lb, dsm + j0, bi, warp, lane, DF_CONSUMERS);
      csync();
      smem_update<true>(S + base * SP + base, S + base * SP + j0, S + base * SP + j0, dsm + j0, crit, crit, jb, warp,
                        lane, DF_CONSUMERS);
      csync();
      if (warp > 0 && rem > crit) {
        // the rest of step j0 on warps 1-7 (warp 0 is already factoring the next sub-block)
        const int rest = rem - crit, b2 = base + crit;
        panel_solve32(S + b2 * SP + j0, rest, Lb, dsm + j0, bi, warp - 1, lane, DF_CONSUMERS - 1);
        bar_named(3, DF_CTHREADS - 32);
        // rectangular part: rows >= b2, the columns of the next sub-block
        smem_update<false>(S + b2 * SP + base, S + b2 * SP + j0, S + base * SP + j0, dsm + j0, rest, crit, jb, warp - 1,
                           lane, DF_CONSUMERS - 1);
        // lower triangle from b2 on
        smem_update<true>(S + b2 * SP + b2, S + b2 * SP + j0, S + b2 * SP + j0, dsm + j0, rest, rest, jb, warp - 1, lane,
                          DF_CONSUMERS - 1);
      }
    }
  }
  csync();
  const long long c_factored = ph ? clock64() : 0;
  // strict lower part = L, diagonal = pivots; 16-byte stores (row bases and even columns are aligned)
#pragma unroll 4
  for (int r = warp; r < nb; r += DF_CONSUMERS) {
    const double dr = dsm[r];
#pragma unroll
    for (int h = 0; h < NB / 64; ++h) {
      const int c = 2 * lane + 64 * h;
      if (c > r) continue;
      const double2 v = *reinterpret_cast<const double2*>(S + r * SP + c);
      double* o = O + (size_t)r * a.ld + c;
      if (c + 1 < r) *reinterpret_cast<double2*>(o) = v;
      else if (c + 1 == r) *reinterpret_cast<double2*>(o) = make_double2(v.x, dr);
      else *o = dr;  // c == r
    }
  }
  for (int t = tid; t < nb; t += DF_CTHREADS) a.Dg[k0 + t] = dsm[t];
  if (ph && tid == 0) {
    ph[0] = c_loaded - c_begin;       // load
    ph[1] = c_ldlt;                   // the four one-warp 32 x 32 factorizations
    ph[2] = c_factored - c_loaded;    // whole elimination loop
    ph[3] = clock64() - c_factored;   // store (issue)
  }
  {
    double* gi = a.Ginv + (size_t)(k0 / 8) * INV_BLK;
    const int nblk = (nb + SB - 1) / SB * 4;
    for (int t = tid; t < nblk * INV_BLK; t += DF_CTHREADS) gi[t] = binv[t];
  }
}

// ---- TRSM(i,k,h): 64 rows of the panel below the diagonal tile,  X (D L_kk^T) = A ----
// Register-resident: every warp keeps its 8 x 128 row block as 16 DMMA accumulator fragments and
// sweeps the 8-column blocks left to right with no block-level synchronisation:
//   X_b = R_b Binv_b^T                      (Binv_b = D_b^-1 L_bb^-1, the 8 x 8 inverse blocks of DIAG)
//   R_b' -= (X_b D_b) L_b'b^T   for b' > b   (independent accumulators: the tensor pipe stays full)
// Only L_kk sits in shared memory (B fragments); accumulator -> A-fragment conversion is a
// shuffle inside each quad.  FP64 work 64 x 128^2 / 2 FMA = 8.2k cycles on one SM's DMMA pipe.
__device__ __forceinline__ double quad_afrag(const double (&c)[2], int src, int odd, unsigned) {
  const double v0 = __shfl_sync(0xffffffffu, c[0], src);
  const double v1 = __shfl_sync(0xffffffffu, c[1], src);
  return odd ? v1 : v0;
}
__device__ __forceinline__ void df_trsm_task(const DfArgs& a, double* sm, int i, int k, int h, int tid, long long* ph) {
  double* S = sm;
  double* dsm = sm + NB * SP + RB * SP;
  double* binv = dsm + 2 * NB + CBUF;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int k0 = k * NB;
  const int r0 = i * DF_TILE + h * RB;
  const double* Lkk = a.dst + (size_t)k0 * a.ld + k0;
  const double* Ain = (k == 0 ? a.src : a.dst) + (size_t)r0 * a.ld + k0;
  double* Aout = a.dst + (size_t)r0 * a.ld + k0;
  double* Wout = a.W + (size_t)r0 * a.ld + k0;

  const long long c_begin = ph ? clock64() : 0;
  // this lane's part of the warp's 8 rows, in accumulator layout (row g, columns 8b + 2q, +1): the loads
  // are issued first, their latency hides behind the cp.async issue loop of L_kk
  const int row = r0 + 8 * warp + g;
  const bool row_ok = row < a.N;
  double acc[16][2];
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    double2 v = make_double2(0.0, 0.0);
    if (row_ok) v = __ldcg(reinterpret_cast<const double2*>(Ain + (size_t)(8 * warp + g) * a.ld + 8 * b + 2 * q));
    acc[b][0] = v.x;
    acc[b][1] = v.y;
  }
  const long long c_issued = ph ? clock64() : 0;
  async_block_load<true, DF_CTHREADS>(S, Lkk, a.ld, NB, NB, tid);
  cp_async_commit();
  for (int t = tid; t < NB; t += DF_CTHREADS) dsm[t] = __ldcg(a.Dg + k0 + t);
  {
    const double* gi = a.Ginv + (size_t)(k0 / 8) * INV_BLK;
    for (int t = tid; t < 16 * INV_BLK; t += DF_CTHREADS) binv[t] = __ldcg(gi + t);
  }
  const long long c_acc = ph ? clock64() : 0;
  cp_async_wait<0>();
  csync();
  const long long c_loaded = ph ? clock64() : 0;

  const int src0 = (lane & ~3) | (q >> 1), src1 = src0 + 2, odd = q & 1;
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    // X_b = R_b Binv_b^T
    const double r_0 = quad_afrag(acc[b], src0, odd, 0), r_1 = quad_afrag(acc[b], src1, odd, 0);
    const double i_0 = binv[b * INV_BLK + g * IP + q], i_1 = binv[b * INV_BLK + g * IP + 4 + q];
    double x[2] = {0.0, 0.0};
    dmma884(x, r_0, i_0);
    dmma884(x, r_1, i_1);
    acc[b][0] = x[0];
    acc[b][1] = x[1];
    if (b < 15) {
      // A fragments of -(X_b D_b), then the independent updates of every later block
      const double a_0 = -(quad_afrag(x, src0, odd, 0) * dsm[8 * b + q]);
      const double a_1 = -(quad_afrag(x, src1, odd, 0) * dsm[8 * b + 4 + q]);
      const double* Sb = S + g * SP + 8 * b + q;
#pragma unroll
      for (int b2 = b + 1; b2 < 16; ++b2) {
        const double f_0 = Sb[(8 * b2) * SP], f_1 = Sb[(8 * b2) * SP + 4];
        dmma884(acc[b2], a_0, f_0);
        dmma884(acc[b2], a_1, f_1);
      }
    }
  }
  const long long c_solved = ph ? clock64() : 0;
  if (ph && tid == 0) {
    ph[0] = c_issued - c_begin;   // L_kk cp.async issue loop
    ph[1] = c_acc - c_issued;     // accumulator loads issued/landed
    ph[2] = c_loaded - c_acc;     // wait for L_kk + barrier
    ph[3] = c_solved - c_loaded;  // solve
  }
  if (row_ok) {
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      const int c = 8 * b + 2 * q;
      *reinterpret_cast<double2*>(Aout + (size_t)(8 * warp + g) * a.ld + c) = make_double2(acc[b][0], acc[b][1]);
      *reinterpret_cast<double2*>(Wout + (size_t)(8 * warp + g) * a.ld + c) =
          make_double2(acc[b][0] * dsm[c], acc[b][1] * dsm[c + 1]);
    }
  }
}


#if DF_TMA
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
// 16 x 128 box (columns c0.., rows c1..) of a row-major FP64 matrix into shared memory, 128-byte swizzle
__device__ __forceinline__ void tma_load_2d(double* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(d),
      "l"(map), "r"(c0), "r"(c1), "r"(b)
      : "memory");
}
#endif

#if DF_TMA
__global__ void __launch_bounds__(DF_THREADS, 1) k_ldlt_dataflow(DfArgs a, const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) double smem[];
#else
__global__ void __launch_bounds__(DF_THREADS, 1) k_ldlt_dataflow(DfArgs a) {
  extern __shared__ __align__(16) double smem[];
#endif
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + DF_DATA_DOUBLES);
  unsigned long long* empty = full + DF_STAGES;
  unsigned long long* tq_full = empty + DF_STAGES;
  unsigned long long* tq_empty = tq_full + 2;
  int4* tq = reinterpret_cast<int4*>(tq_empty + 2);     // 2 published tasks
  int4* ptask = tq + 2;                                 // 2 producer-side broadcast slots
  int* tq_ticket = reinterpret_cast<int*>(ptask + 2);   // ticket numbers of the published tasks
  int* ptask_ticket = tq_ticket + 2;
  volatile int* prog0 = ptask_ticket + 2;  // k-slices consumed so far by consumer group 0
  int* upd_done = ptask_ticket + 3;        // [2] consumer warps that finished the UPD task in queue slot 0 / 1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < DF_STAGES; ++s) {
      mbar_init(full + s, DF_TMA ? 1 : DF_PTHREADS);  // TMA: the issuing thread + transaction bytes; else one
                                                      // cp.async-completion arrive per producer thread
      mbar_init(empty + s, DF_CONSUMERS); // one arrive per consumer warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tq_full + s, 1);
      mbar_init(tq_empty + s, DF_CONSUMERS);
      upd_done[s] = 0;
    }
    *prog0 = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= DF_CONSUMERS) {
    // ======================= producers =======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(DF_PRODUCER_REGS));
    const int pl = tid - DF_CTHREADS;
    const int pwarp = warp - DF_CONSUMERS;
    unsigned it = 0, pq = 0;
    int4 tk;
    int tnum;
    auto fetch = [&]() {
      const int b = pq & 1;
      if (pwarp == 0) {
        int t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        int4 x = make_int4(DF_DONE, 0, 0, 0);
        if (t < a.ntasks) {
          x = __ldg(a.tasks + t);
          wait_deps(a, x, lane);
        }
        if (lane == 0) { ptask[b] = x; ptask_ticket[b] = t; }
      }
      bar_named(1, DF_PTHREADS);
      tk = ptask[b];
      tnum = ptask_ticket[b];
    };
    fetch();
    for (;;) {
      const int slot = pq & 1;
      if (pl == 0) {
        mbar_wait_bounded(tq_empty + slot, ((pq >> 1) & 1u) ^ 1u, a.sticky);
        tq[slot] = tk;
        tq_ticket[slot] = tnum;
        mbar_arrive(tq_full + slot);
      }
      ++pq;
      const int type = tk.x & 0xff;
      if (type == DF_DONE) break;
      if (type == DF_UPD || type == DF_DIAGU) {
        const int i = tk.y, j = tk.z, k0 = tk.w & 0xffff, k1 = tk.w >> 16;
        const int KT = (k1 - k0) * (DF_TILE / BK);
        const double* PA = a.dst + (size_t)i * DF_TILE * a.ld + (size_t)k0 * DF_TILE;
        const double* PB = a.W + (size_t)j * DF_TILE * a.ld + (size_t)k0 * DF_TILE;
        const int rowsA = a.N - i * DF_TILE, rowsB = a.N - j * DF_TILE;
#if DF_TMA
        if (pl == 0) {
          for (int kt = 0; kt < KT; ++kt, ++it) {
            const unsigned s = it % DF_STAGES;
            mbar_wait_bounded(empty + s, ((it / DF_STAGES) & 1u) ^ 1u, a.sticky);
            double* Sd = smem + s * DF_STAGE_DOUBLES;
            const int kcol = k0 * DF_TILE + kt * BK;
            mbar_arrive_expect_tx(full + s, (unsigned)(DF_STAGE_DOUBLES * sizeof(double)));
            tma_load_2d(Sd, &mapA, kcol, i * DF_TILE, full + s);                 // rows of L (zero beyond N)
            tma_load_2d(Sd + DF_OP_DOUBLES, &mapB, kcol, j * DF_TILE, full + s);  // rows of W = L D
          }
        } else {
          it += KT;
        }
        (void)PA; (void)PB; (void)rowsA; (void)rowsB;
#else
        for (int kt = 0; kt < KT; ++kt, ++it) {
          const unsigned s = it % DF_STAGES;
          mbar_wait_bounded(empty + s, ((it / DF_STAGES) & 1u) ^ 1u, a.sticky);
          double* Sd = smem + s * DF_STAGE_DOUBLES;
          const int kbase = kt * BK;
#pragma unroll 8
          for (int c = 0; c < 2 * DF_TILE * (BK / 2) / DF_PTHREADS; ++c) {
            const int chunk = pl + c * DF_PTHREADS;
            const int r = chunk >> 3, ck = (chunk & 7) * 2;
            const bool isA = r < DF_TILE;
            const int rr = isA ? r : r - DF_TILE;
            const bool ok = rr < (isA ? rowsA : rowsB);
            const double* srcp = (isA ? PA : PB) + (size_t)(ok ? rr : 0) * a.ld + kbase + ck;
            cp_async16(Sd + r * LDT + ck, srcp, ok ? 16 : 0);
          }
          mbar_arrive_cp_async(full + s);
        }
#endif
        if (type == DF_DIAGU) {  // the factorization part re-uses the ring memory: park like a bulk task
          __syncthreads();
          fetch();
          __syncthreads();
        } else {
          fetch();
        }
      } else {
        __syncthreads();  // consumers own the ring memory for the bulk task
        fetch();
        __syncthreads();
      }
    }
    cp_async_wait<0>();
    return;
  }

  // ======================= consumers =======================
  // the registers the producers gave back: 232 per consumer thread -- the 64 x 32 DMMA accumulator
  // tile (128 registers) and the one-warp 32 x 32 LDL^T of the DIAG task (row in registers) both fit
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(DF_CONSUMER_REGS));
  const int wm = warp & 1, wn = warp >> 1;
  const int g = lane >> 2, q = lane & 3;
  // The two halves of the consumer tile (warps 0-3: columns 0..63, warps 4-7: columns 64..127) run
  // DF_LAG k-slices apart, and every SM sub-partition hosts one warp of each half.  One warp alone can
  // keep its sub-partition's DMMA pipe full, so while one half loads its C tile, waits on a stage or
  // stores its result, the other half's warp has the pipe: tile transitions cost no tensor time.
  // There is therefore no CTA-wide barrier on the UPD path (completion is counted per warp).
  const int grp = warp >> 2;
  unsigned it = 0, cq = 0;
  for (;;) {
    const int slot = cq & 1;
    mbar_wait_bounded(tq_full + slot, (cq >> 1) & 1u, a.sticky);
    const int4 tk = tq[slot];
    const int tnum = tq_ticket[slot];
    ++cq;
    const int type = tk.x & 0xff;
    if (type == DF_DONE) break;
    long long t_start = 0;
    if (a.tlog && tid == 0) t_start = gtimer();
    const int i = tk.y, j = tk.z;
    if (type == DF_UPD || type == DF_DIAGU) {
      const int k0 = tk.w & 0xffff, k1 = tk.w >> 16;
      const int KT = (k1 - k0) * (DF_TILE / BK);
      const bool diag = i == j;
      const int row0 = i * DF_TILE, col0 = j * DF_TILE;
      const int wrow = row0 + wm * 64, wcol = col0 + wn * 32;
      const bool live = !diag || wn <= 2 * wm + 1;  // warp tiles strictly above the diagonal are skipped
      const double* Cin = k0 == 0 ? a.src : a.dst;
      if (grp == 1 && type == DF_UPD) {  // trail group 0 (the producers can always run DF_STAGES slices ahead of
                                         // this group); not in the fused chain task, whose latency matters
        while ((int)(*prog0 - it) < DF_LAG) __nanosleep(200);
      }
      double acc[DF_MI][DF_NI][2];
#pragma unroll
      for (int mi = 0; mi < DF_MI; ++mi) {
        const int row = wrow + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < DF_NI; ++ni) {
          const int col = wcol + ni * 8 + 2 * q;
          double2 cv = make_double2(0.0, 0.0);
          if (live && row < a.N && (!diag || col <= row)) {
            const double* p = Cin + (size_t)row * a.ld + col;
            if (!diag || col + 1 <= row) cv = __ldcg(reinterpret_cast<const double2*>(p));
            else cv.x = __ldcg(p);
          }
          acc[mi][ni][0] = -cv.x;
          acc[mi][ni][1] = -cv.y;
        }
      }
      long long c_t0 = 0, c_first = 0, c_loop = 0, c_stored = 0;
      if (a.tlog) c_t0 = clock64();
      for (int kt = 0; kt < KT; ++kt, ++it) {
        const unsigned s = it % DF_STAGES;
        mbar_wait_bounded(full + s, (it / DF_STAGES) & 1u, a.sticky);
        if (a.tlog && kt == 1) c_first = clock64();
        if (live) {
#if DF_TMA
          // 128-byte swizzle: the 16-byte chunk c of row r sits at chunk position c ^ (r & 7); r & 7 == g here
          const double* Aw = smem + s * DF_STAGE_DOUBLES + (wm * 64 + g) * BK + (q & 1);
          const double* Bw = smem + s * DF_STAGE_DOUBLES + DF_OP_DOUBLES + (wn * 32 + g) * BK + (q & 1);
#else
          const double* Aw = smem + s * DF_STAGE_DOUBLES + (wm * 64 + g) * LDT + q;
          const double* Bw = smem + s * DF_STAGE_DOUBLES + DF_TILE * LDT + (wn * 32 + g) * LDT + q;
#endif
#pragma unroll
          for (int kk = 0; kk < BK / 4; ++kk) {
            double af[DF_MI], bf[DF_NI];
#if DF_TMA
            const int sw = ((2 * kk + (q >> 1)) ^ g) * 2;
#pragma unroll
            for (int mi = 0; mi < DF_MI; ++mi) af[mi] = Aw[mi * 8 * BK + sw];
#pragma unroll
            for (int ni = 0; ni < DF_NI; ++ni) bf[ni] = Bw[ni * 8 * BK + sw];
#else
#pragma unroll
            for (int mi = 0; mi < DF_MI; ++mi) af[mi] = Aw[mi * 8 * LDT + kk * 4];
#pragma unroll
            for (int ni = 0; ni < DF_NI; ++ni) bf[ni] = Bw[ni * 8 * LDT + kk * 4];
#endif
#pragma unroll
            for (int mi = 0; mi < DF_MI; ++mi)
#pragma unroll
              for (int ni = 0; ni < DF_NI; ++ni) dmma884(acc[mi][ni], af[mi], bf[ni]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (tid == 0) *prog0 = (int)(it + 1);
      }
      if (a.tlog) c_loop = clock64();
      if (type == DF_DIAGU) {
        // fused: the updated diagonal tile goes from the accumulators straight into the shared-memory
        // tile that DIAG factors (every slice has been consumed once all threads pass the barrier)
        __syncthreads();
        if (live) {
#pragma unroll
          for (int mi = 0; mi < DF_MI; ++mi) {
            const int lr = wm * 64 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < DF_NI; ++ni) {
              const int lc = wn * 32 + ni * 8 + 2 * q;
              if (lc <= lr) smem[lr * SP + lc] = -acc[mi][ni][0];
              if (lc + 1 <= lr) smem[lr * SP + lc + 1] = -acc[mi][ni][1];
            }
          }
        }
        csync();
      } else {
      if (live) {
#pragma unroll
        for (int mi = 0; mi < DF_MI; ++mi) {
          const int row = wrow + mi * 8 + g;
          if (row >= a.N) continue;
#pragma unroll
          for (int ni = 0; ni < DF_NI; ++ni) {
            const int col = wcol + ni * 8 + 2 * q;
            if (diag && col > row) continue;
            double* p = a.dst + (size_t)row * a.ld + col;
            if (!diag || col + 1 <= row) *reinterpret_cast<double2*>(p) = make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
            else *p = -acc[mi][ni][0];
          }
        }
      }
      if (a.tlog) c_stored = clock64();
      // every thread orders its own stores, the last of the eight warps publishes the tile
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        const int old = atomicAdd(upd_done + slot, 1);
        if (old == DF_CONSUMERS - 1) {
          upd_done[slot] = 0;
          st_release(a.cnt + (size_t)i * a.nt + j, k1);
          if (a.tlog) a.tlog[(size_t)tnum * 8 + 1] = gtimer();
        }
      }
      if (a.tlog && tid == 0) {
        long long* ph = a.tlog + (size_t)tnum * 8 + 4;
        ph[0] = c_first - c_t0;          // C tile in registers + first k-slice done
        ph[1] = c_loop - c_first;        // remaining k-slices
        ph[2] = c_stored - c_loop;       // C stores issued
        ph[3] = clock64() - c_stored;    // fence + completion count
      }
      }  // UPD epilogue
    }
    if (type != DF_UPD) {
      // bulk tasks on the consumer warps (DIAGU arrives here with its tile already in shared memory)
      if (type != DF_DIAGU) __syncthreads();  // producers have published and stopped touching the ring
      long long* ph = (a.tlog && type != DF_DIAGU) ? a.tlog + (size_t)tnum * 8 + 4 : nullptr;
      if (type == DF_TRSM) df_trsm_task(a, smem, i, j, (tk.x >> 8) & 0xff, tid, ph);
      else df_diag_task(a, smem, j, tid, ph, type == DF_DIAGU);
#if DF_TMA
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes of the tile storage before TMA refills it
#endif
      csync();
      if (tid == 0) {
        __threadfence();
        if (type == DF_TRSM) {
          red_release_add(a.rdy + (size_t)j * a.nt + i, 1);
        } else {
          if (type == DF_DIAGU) st_release(a.cnt + (size_t)j * a.nt + j, j);
          st_release(a.rdy + (size_t)j * a.nt + j, 1);
        }
      }
      __syncthreads();
    }
    if (a.tlog && tid == 0) {
      long long* rec = a.tlog + (size_t)tnum * 8;
      rec[0] = t_start;
      if (type != DF_UPD) rec[1] = gtimer();
      rec[2] = smid();
      rec[3] = tk.x | ((long long)tk.w << 32);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tq_empty + slot);
  }
}


// One launch of the dataflow kernel of this translation unit (dataflow.cu: cp.async build, dataflow_tma.cu: TMA build)
static int df_kernel_init() {
  return (int)cudaFuncSetAttribute(k_ldlt_dataflow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM);
}

#if DF_TMA
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// rows x ld row-major FP64, box = 16 columns x 128 rows, 128-byte swizzle, zeros out of bounds
static bool make_map(CUtensorMap* m, const double* base, int rows, int ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)DF_TILE};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#endif

#if DF_TMA
// Same kernel with the two UPD operands taken from arbitrary row-major matrices: C(i,j) -= A(i, k-range) B(j, k-range)^T.
// Used for the condensed assembly M^T W M (A = MT, B = -MT diag(W)) as a list of UPD tasks with preset flags.
static int df_kernel_launch_operands(cudaStream_t st, const DfArgs& a, int ctas, const double* A, int rowsA, int ldA,
                                     const double* B, int rowsB, int ldB) {
  CUtensorMap mapA, mapB;
  if (!make_map(&mapA, A, rowsA, ldA) || !make_map(&mapB, B, rowsB, ldB)) return (int)cudaErrorNotSupported;
  k_ldlt_dataflow<<<ctas, DF_THREADS, DF_SMEM, st>>>(a, mapA, mapB);
  return 0;
}
#endif

static int df_kernel_launch(cudaStream_t st, const DfArgs& a, int ctas, const double* W) {
#if DF_TMA
  CUtensorMap mapA, mapB;
  if (!make_map(&mapA, a.dst, a.N, a.ld) || !make_map(&mapB, W, a.N, a.ld)) return (int)cudaErrorNotSupported;
  k_ldlt_dataflow<<<ctas, DF_THREADS, DF_SMEM, st>>>(a, mapA, mapB);
#else
  (void)W;
  k_ldlt_dataflow<<<ctas, DF_THREADS, DF_SMEM, st>>>(a);
#endif
  return 0;
}

}  // namespace

}  // namespace ipmz
