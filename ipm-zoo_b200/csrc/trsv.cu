// ipm-zoo_b200/csrc/trsv.cu -- forward / diagonal / backward solves with the in-place LDL^T
// factor (reference: LinearSolvers::overwriting_solve_ldlt, LinearSolvers.cpp:44-74).
//
// HBM-bound: each sweep reads the strict lower triangle once (N^2/2 doubles).  One launch per
// sweep: block-row r of L is owned by one CTA which streams the tiles L[r][j] as the x_j blocks
// they multiply are published by the CTAs of earlier block rows (flag per block, release /
// acquire through __threadfence).  CTAs take their block row from an atomic ticket, so a CTA
// only ever waits on CTAs that started before it -- no dependence on dispatch order.
#include "ipmz_device.cuh"
#include "ipmz_kernels.h"

namespace ipmz {

namespace {

constexpr int TB = 64;  // block size of the solves
constexpr int TP = TB + 1;

struct TrsvArgs {
  const double* K;
  const double* Dg;
  double* x;
  int ld, N, nblk;
  size_t sK, sD, sx;
  int* flags;
  int* ticket;
  int epoch, cap_blocks, total;
  const int* active;
};

__device__ __forceinline__ void wait_flag(const int* flag, int epoch) {
  if (threadIdx.x == 0) {
    while (*reinterpret_cast<const volatile int*>(flag) != epoch) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ int take_ticket(const TrsvArgs& a) {
  __shared__ int tk;
  if (threadIdx.x == 0) tk = atomicAdd(a.ticket, 1);
  __syncthreads();
  return tk;
}

__device__ __forceinline__ void release_ticket(const TrsvArgs& a, int tk) {
  // the CTA holding the last ticket re-arms the counter for the next launch
  if (threadIdx.x == 0 && tk == a.total - 1) *a.ticket = 0;
}

// L y = b, unit lower
__global__ void __launch_bounds__(256) k_trsv_forward(TrsvArgs a) {
  __shared__ double Ld[TB * TP];
  __shared__ double xj[TB];
  __shared__ double acc[TB];
  const int tk = take_ticket(a);
  const int slot = tk / a.nblk, r = tk - slot * a.nblk;
  const int p = a.active ? a.active[slot] : slot;
  const double* K = a.K + (size_t)p * a.sK;
  double* x = a.x + (size_t)p * a.sx;
  int* flags = a.flags + (size_t)slot * a.cap_blocks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R0 = r * TB, nr = min(TB, a.N - R0);

  // diagonal block and right-hand side do not depend on other CTAs
  for (int idx = tid; idx < TB * TB; idx += 256) {
    const int i = idx / TB, c = idx - i * TB;
    Ld[i * TP + c] = (i < nr && c < i) ? K[(size_t)(R0 + i) * a.ld + R0 + c] : 0.0;
  }
  if (tid < TB) acc[tid] = tid < nr ? x[R0 + tid] : 0.0;
  __syncthreads();

  for (int j = 0; j < r; ++j) {
    // rows of this warp: warp*8 .. +7; prefetch the tile before waiting on x_j
    double l0[8], l1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = warp * 8 + i;
      const double* Lr = K + (size_t)(R0 + row) * a.ld + (size_t)j * TB;
      l0[i] = row < nr ? Lr[lane] : 0.0;
      l1[i] = row < nr ? Lr[lane + 32] : 0.0;
    }
    wait_flag(flags + j, a.epoch);
    if (tid < TB) xj[tid] = __ldcg(x + (size_t)j * TB + tid);
    __syncthreads();
    const double x0 = xj[lane], x1 = xj[lane + 32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double s = l0[i] * x0 + l1[i] * x1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) acc[warp * 8 + i] -= s;
    }
    __syncthreads();
  }

  if (warp == 0) {
    double y0 = acc[lane], y1 = acc[lane + 32];
#pragma unroll 8
    for (int c = 0; c < TB; ++c) {
      const double yc = __shfl_sync(0xffffffffu, c < 32 ? y0 : y1, c & 31);
      if (lane > c) y0 -= Ld[lane * TP + c] * yc;
      if (lane + 32 > c) y1 -= Ld[(lane + 32) * TP + c] * yc;
    }
    if (lane < nr) x[R0 + lane] = y0;
    if (lane + 32 < nr) x[R0 + lane + 32] = y1;
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile int*>(flags + r) = a.epoch;
  }
  release_ticket(a, tk);
}

// L^T x = D^-1 y
__global__ void __launch_bounds__(256) k_trsv_backward(TrsvArgs a) {
  __shared__ double Ld[TB * TP];
  __shared__ double xj[TB];
  __shared__ double part[4][TB];
  __shared__ double acc[TB];
  const int tk = take_ticket(a);
  const int slot = tk / a.nblk, r = a.nblk - 1 - (tk - slot * a.nblk);
  const int p = a.active ? a.active[slot] : slot;
  const double* K = a.K + (size_t)p * a.sK;
  const double* Dg = a.Dg + (size_t)p * a.sD;
  double* x = a.x + (size_t)p * a.sx;
  int* flags = a.flags + (size_t)slot * a.cap_blocks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R0 = r * TB, nr = min(TB, a.N - R0);

  for (int idx = tid; idx < TB * TB; idx += 256) {
    const int i = idx / TB, c = idx - i * TB;
    Ld[i * TP + c] = (i < nr && c < i) ? K[(size_t)(R0 + i) * a.ld + R0 + c] : 0.0;
  }
  if (tid < TB) acc[tid] = tid < nr ? x[R0 + tid] / Dg[R0 + tid] : 0.0;
  __syncthreads();

  const int c = tid & 63, grp = tid >> 6;  // column of this block row's columns, row group
  for (int j = a.nblk - 1; j > r; --j) {
    const int J0 = j * TB, nj = min(TB, a.N - J0);
    double lv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int row = grp * 16 + i;
      lv[i] = (row < nj && c < nr) ? K[(size_t)(J0 + row) * a.ld + R0 + c] : 0.0;
    }
    wait_flag(flags + j, a.epoch);
    if (tid < TB) xj[tid] = tid < nj ? __ldcg(x + J0 + tid) : 0.0;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += lv[i] * xj[grp * 16 + i];
    part[grp][c] = s;
    __syncthreads();
    if (tid < TB) acc[tid] -= (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    __syncthreads();
  }

  if (warp == 0) {
    double x0 = acc[lane], x1 = acc[lane + 32];
#pragma unroll 8
    for (int i = TB - 1; i >= 0; --i) {
      const double xi = __shfl_sync(0xffffffffu, i < 32 ? x0 : x1, i & 31);
      if (lane < i) x0 -= Ld[i * TP + lane] * xi;
      if (lane + 32 < i) x1 -= Ld[i * TP + lane + 32] * xi;
    }
    if (lane < nr) x[R0 + lane] = x0;
    if (lane + 32 < nr) x[R0 + lane + 32] = x1;
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile int*>(flags + r) = a.epoch;
  }
  release_ticket(a, tk);
}

}  // namespace

void launch_ldlt_solve(cudaStream_t st, const FactorPlan& fp, const double* K, const double* Dg, double* x,
                       size_t sx, TrsvWork& w) {
  if (fp.N <= 0 || fp.nslots <= 0) return;
  const int nblk = (fp.N + TB - 1) / TB;
  TrsvArgs a;
  a.K = K; a.Dg = Dg; a.x = x; a.ld = fp.ld; a.N = fp.N; a.nblk = nblk;
  a.sK = fp.sK; a.sD = fp.sD; a.sx = sx;
  a.flags = w.flags; a.ticket = w.ticket; a.cap_blocks = w.cap_blocks;
  a.total = nblk * fp.nslots; a.active = fp.active;
  a.epoch = ++w.epoch;
  k_trsv_forward<<<a.total, 256, 0, st>>>(a); count_launch();
  a.epoch = ++w.epoch;
  k_trsv_backward<<<a.total, 256, 0, st>>>(a); count_launch();
}

}  // namespace ipmz
