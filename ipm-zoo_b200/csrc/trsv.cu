// ipm-zoo_b200/csrc/trsv.cu -- forward / diagonal / backward solves with the in-place LDL^T
// factor (reference: LinearSolvers::overwriting_solve_ldlt, LinearSolvers.cpp:44-74).
//
// HBM-bound: each sweep reads the strict lower triangle once (N^2/2 doubles).  One launch per
// sweep: block-row r of L is owned by one CTA which streams the tiles L[r][j] as the x_j blocks
// they multiply are published by the CTAs of earlier block rows (flag per block, release /
// acquire through __threadfence).  CTAs take their block row from an atomic ticket, so a CTA
// only ever waits on CTAs that started before it -- no dependence on dispatch order.
#include <stdlib.h>

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


// ------------------------------------------------------------------------------------------
// Small systems (N <= TRSV_SMALL_MAX; the batched QPs of cfg4): ONE CTA per problem runs the
// forward sweep, the pivot scaling and the backward sweep back to back with x in shared memory --
// no inter-CTA flags, one launch per solve, and the factor is read a second time while it is still
// in L2.  Block rows of 64: the part left of (below) the diagonal block is a coalesced
// matrix-vector product over the whole CTA, the 64 x 64 diagonal block is solved by one warp.
constexpr int TRSV_SMALL_MAX = 512;

__global__ void __launch_bounds__(256, 4) k_trsv_small(TrsvArgs a) {
  extern __shared__ double sx[];  // x, padded to a multiple of 64
  __shared__ double Ld[TB * TP];
  __shared__ double part[4][TB];
  const int slot = blockIdx.x;
  const int p = a.active ? a.active[slot] : slot;
  const double* K = a.K + (size_t)p * a.sK;
  const double* Dg = a.Dg + (size_t)p * a.sD;
  double* x = a.x + (size_t)p * a.sx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = a.nblk, N = a.N;
  for (int i = tid; i < nblk * TB; i += 256) sx[i] = i < N ? x[i] : 0.0;
  __syncthreads();

  // ---- forward: L y = b ----
  for (int r = 0; r < nblk; ++r) {
    const int R0 = r * TB, nr = min(TB, N - R0);
    for (int idx = tid; idx < TB * TB; idx += 256) {
      const int i = idx / TB, c = idx - i * TB;
      Ld[i * TP + c] = (i < nr && c < i) ? K[(size_t)(R0 + i) * a.ld + R0 + c] : 0.0;
    }
    // rows 8*warp .. +7 of the block: dot products with the finished part of y
    double s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.0;
    // two 32-column chunks per trip: 16 independent loads in flight per lane (the loop is latency bound)
    for (int c0 = 0; c0 < R0; c0 += 64) {
      const double xa = sx[c0 + lane], xb = sx[c0 + 32 + lane];
      double la[8], lb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 8 * warp + i;
        const double* Kr = K + (size_t)(R0 + row) * a.ld + c0 + lane;
        la[i] = row < nr ? Kr[0] : 0.0;
        lb[i] = row < nr ? Kr[32] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = fma(lb[i], xb, fma(la[i], xa, s[i]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double v = s[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sx[R0 + 8 * warp + i] -= v;
    }
    __syncthreads();
    if (warp == 0) {
      double y0 = sx[R0 + lane], y1 = sx[R0 + lane + 32];
#pragma unroll 8
      for (int c = 0; c < TB; ++c) {
        const double yc = __shfl_sync(0xffffffffu, c < 32 ? y0 : y1, c & 31);
        if (lane > c) y0 -= Ld[lane * TP + c] * yc;
        if (lane + 32 > c) y1 -= Ld[(lane + 32) * TP + c] * yc;
      }
      sx[R0 + lane] = y0;
      sx[R0 + lane + 32] = y1;
    }
    __syncthreads();
  }
  // ---- pivots ----
  for (int i = tid; i < N; i += 256) sx[i] = sx[i] / Dg[i];
  __syncthreads();
  // ---- backward: L^T x = y ----
  const int c = tid & 63, grp = tid >> 6;
  for (int r = nblk - 1; r >= 0; --r) {
    const int R0 = r * TB, nr = min(TB, N - R0);
    for (int idx = tid; idx < TB * TB; idx += 256) {
      const int i = idx / TB, cc = idx - i * TB;
      Ld[i * TP + cc] = (i < nr && cc < i) ? K[(size_t)(R0 + i) * a.ld + R0 + cc] : 0.0;
    }
    // column R0 + c against the finished rows below the block, rows split over the 4 thread groups
    double sacc = 0.0;
    if (c < nr) {
      // 8 independent loads / partial sums per trip (latency bound otherwise)
      double p8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) p8[u] = 0.0;
      int row = R0 + TB + grp;
      for (; row + 28 < N; row += 32) {
        double lv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) lv[u] = K[(size_t)(row + 4 * u) * a.ld + R0 + c];
#pragma unroll
        for (int u = 0; u < 8; ++u) p8[u] = fma(lv[u], sx[row + 4 * u], p8[u]);
      }
      for (; row < N; row += 4) p8[0] = fma(K[(size_t)row * a.ld + R0 + c], sx[row], p8[0]);
      sacc = ((p8[0] + p8[1]) + (p8[2] + p8[3])) + ((p8[4] + p8[5]) + (p8[6] + p8[7]));
    }
    part[grp][c] = sacc;
    __syncthreads();
    if (tid < TB) sx[R0 + tid] -= (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    __syncthreads();
    if (warp == 0) {
      double x0 = sx[R0 + lane], x1 = sx[R0 + lane + 32];
#pragma unroll 8
      for (int i = TB - 1; i >= 0; --i) {
        const double xi = __shfl_sync(0xffffffffu, i < 32 ? x0 : x1, i & 31);
        if (lane < i) x0 -= Ld[i * TP + lane] * xi;
        if (lane + 32 < i) x1 -= Ld[i * TP + lane + 32] * xi;
      }
      sx[R0 + lane] = x0;
      sx[R0 + lane + 32] = x1;
    }
    __syncthreads();
  }
  for (int i = tid; i < N; i += 256) x[i] = sx[i];
}

// ------------------------------------------------------------------------------------------
// Streaming solves for one large factor (single matrix, N >= the dataflow threshold):
//   forward   y_r = L_rr^-1   (b_r       - sum_{j<r} L_rj   y_j)
//   backward  x_r = L_rr^-T   (y_r / D_r - sum_{j>r} L_jr^T x_j)
// One CTA per 128-row block row (ticket order = dependency order).  Its off-diagonal tiles
// stream through a 9-stage cp.async ring of 16 x 128 slices whose addresses do not depend on
// other CTAs, so HBM latency is hidden and the whole last tile is resident when x_{r-1} arrives.
// The diagonal tile is prepared while the CTA waits: H = blockdiag(L_bb^-1) L_rr with the 8 x 8
// inverse blocks of the factorization (the same blocks its panel solves use; no larger explicit
// inverse -- unpivoted LDL^T of a KKT matrix has |L_ij| >> 1 and a 128 x 128 inverse loses all
// accuracy), which leaves ONE 8 x 8 product per block step on the dependent chain.  The backward
// sweep runs the same code on the reversed transpose of the tile (again unit lower triangular).
// Blocks are exchanged through a self-validating buffer (every entry starts as an all-ones
// NaN pattern and is polled until it changes: one L2 round trip per hop, no fence, no flag).
constexpr int SV_TB = 128, SV_SR = 16, SV_SPT = SV_TB / SV_SR, SV_STAGES = SV_SPT + 1, SV_THREADS = 256;
constexpr int SV_STAGE_DOUBLES = SV_SR * SV_TB;
constexpr int SV_NBLK8 = SV_TB / 8;
constexpr int SV_H_DOUBLES = 64 * (SV_NBLK8 * (SV_NBLK8 + 1) / 2);  // block-row packed lower triangle
constexpr int SV_SMEM_DOUBLES = SV_STAGES * SV_STAGE_DOUBLES + SV_H_DOUBLES + SV_TB + 8 * SV_TB;
constexpr size_t SV_SMEM = (size_t)SV_SMEM_DOUBLES * sizeof(double);
static_assert(SV_SMEM <= 232448 - 64, "shared memory per CTA");
constexpr unsigned long long SV_NOT_YET = ~0ull;
constexpr int SV_IP = 12, SV_INV_BLK = 8 * SV_IP;  // layout of the 8 x 8 inverse blocks (ldlt_device.cuh)

struct SvArgs {
  const double* K;
  const double* Dg;
  const double* Ginv;
  double* x;      // in: right-hand side, out: solution
  double* xl;     // [2][nblk*128] exchange buffers: forward blocks y_r, backward blocks x_r
  int* ticket;
  int* sticky;    // set when a poll hits its watchdog (never reset; read by the host after a solve)
  int ld, N, nblk;
  long long* tlog;  // debug: [2][nblk][16] stamps (nullptr: off)
};

__device__ __forceinline__ long long sv_now() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

__device__ __forceinline__ void sv_cp16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(src_bytes));
}

__device__ __forceinline__ double sv_poll(const double* p, int* sticky) {
  unsigned long long v;
  long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    if (v != SV_NOT_YET) break;
    if ((spin & 4095u) == 4095u) {  // watchdog: a protocol bug must not hang the device
      const long long t = sv_now();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) {
        if (sticky) atomicExch(sticky, 1);  // the host reports the solve as failed
        break;
      }
    }
  }
  return __longlong_as_double((long long)v);
}
// one non-blocking look (so that the looks at several chunks overlap); sv_poll_from continues from it
__device__ __forceinline__ unsigned long long sv_peek(const double* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double sv_poll_from(unsigned long long v, const double* p, int* sticky) {
  return v != SV_NOT_YET ? __longlong_as_double((long long)v) : sv_poll(p, sticky);
}
__device__ __forceinline__ void sv_publish(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(__double_as_longlong(v)) : "memory");
}

// sums v[0..15] over the warp; afterwards v[0] on lane l is the total of entry
// 8*bit4(l) + 4*bit3(l) + 2*bit2(l) + bit1(l)
__device__ __forceinline__ void warp_reduce16(double (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = up ? v[i] : v[i + half];
      const double keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Packed lower-triangular storage of the prepared diagonal tile, COLUMN-major: column q (8-block
// l = q / 8) holds rows 8l..127 contiguously, so the solve's access pattern (lane = row, all
// lanes the same column) is bank-conflict free.
__device__ __forceinline__ int sv_hc(int p, int q) {
  const int l = q >> 3;
  return 1024 * l - 32 * l * (l - 1) + (q & 7) * (SV_TB - 8 * l) + p - 8 * l;
}

// One block row of one sweep.  FWD: y_r = L_rr^-1 (b_r - sum_{j<r} L_rj y_j), published only.
// !FWD: x_r = L_rr^-T (y_r / D_r - sum_{j>r} L_jr^T x_j), published and written to a.x.
template <bool FWD>
__device__ __forceinline__ void sv_block_row(const SvArgs& a, const int r, double* svm, const int tid) {
  double* ring = svm;
  double* Hs = ring + SV_STAGES * SV_STAGE_DOUBLES;  // prepared diagonal tile
  double* ts = Hs + SV_H_DOUBLES;                    // 128: right-hand side of the diagonal step
  double* part = ts + SV_TB;                         // 8 x 128 (backward: per-warp column sums; then 4 x 32 finished x)
  const int lane = tid & 31, warp = tid >> 5;
  const int R0 = r * SV_TB, nr = min(SV_TB, a.N - R0);
  const int ntiles = FWD ? r : a.nblk - 1 - r;
  const int total = SV_SPT * ntiles;
  double* fwd = a.xl;
  double* bwd = a.xl + (size_t)a.nblk * SV_TB;
  double* pub = FWD ? fwd : bwd;
  long long* tlog = a.tlog ? a.tlog + ((size_t)(FWD ? 0 : a.nblk) + r) * 16 : nullptr;
// stamps are SM cycle counters; slot 7 holds one globaltimer reading taken together with stamp 6 so
// that the per-SM counters can be put on a common time axis afterwards.  The __syncwarp
// re-converges warp 0: a diverged warp takes the slow path of every later shuffle.
#define SV_STAMP(k) do { if (tlog) { if (tid == 0) tlog[(k)] = clock64(); __syncwarp(); } } while (0)
  SV_STAMP(0);

  auto issue = [&](int st) {
    const int t = st / SV_SPT, sl = st - t * SV_SPT;
    double* dst = ring + (st % SV_STAGES) * SV_STAGE_DOUBLES;
    const int j = FWD ? t : a.nblk - 1 - t;
    const int row0 = (FWD ? R0 : j * SV_TB) + sl * SV_SR;
    const double* base = a.K + (size_t)row0 * a.ld + (FWD ? j * SV_TB : R0);
    const int row_lim = a.N - row0;
#pragma unroll
    for (int c = 0; c < SV_SR * SV_TB / 2 / SV_THREADS; ++c) {
      const int idx = tid + c * SV_THREADS;
      const int rr = idx >> 6, cc = (idx & 63) * 2;
      const bool ok = rr < row_lim;
      sv_cp16(dst + rr * SV_TB + cc, base + (size_t)(ok ? rr : 0) * a.ld + cc, ok ? 16 : 0);
    }
  };
#pragma unroll
  for (int st = 0; st < SV_STAGES - 1; ++st) {
    if (st < total) issue(st);
    asm volatile("cp.async.commit_group;\n" ::);
  }

  // ---- diagonal tile, prepared while the first tiles are in flight ----
  // local index p: forward p = row; backward p = 127 - row (reversed transpose, again unit lower)
  // strictly-lower 8 x 8 blocks: coalesced reads of the source rows, scattered into the packed layout
  for (int sr = 8 + warp; sr < SV_TB; sr += SV_THREADS / 32) {
    const int ncol = 8 * (sr >> 3);
    const double* src = a.K + (size_t)(R0 + sr) * a.ld + R0;
#pragma unroll
    for (int c0 = 0; c0 < SV_TB - 8; c0 += 32) {
      const int sc = c0 + lane;
      if (sc < ncol) {
        const double v = sr < nr ? src[sc] : 0.0;
        const int p = FWD ? sr : SV_TB - 1 - sc, q = FWD ? sc : SV_TB - 1 - sr;
        Hs[sv_hc(p, q)] = v;
      }
    }
  }
  // diagonal blocks: L_bb^-1 = D_b (D_b^-1 L_bb^-1), from the 8 x 8 inverse blocks of the factorization
  for (int e = tid; e < SV_NBLK8 * 64; e += SV_THREADS) {
    const int b = e >> 6, i = (e >> 3) & 7, c = e & 7;
    double v = 0.0;
    if (c <= i) {
      const int gb = FWD ? b : SV_NBLK8 - 1 - b, gi = FWD ? i : 7 - c, gc = FWD ? c : 7 - i;
      if (8 * gb + gi < nr)
        v = a.Dg[R0 + 8 * gb + gi] * a.Ginv[(size_t)(R0 / 8 + gb) * SV_INV_BLK + gi * SV_IP + gc];
    }
    Hs[sv_hc(8 * b + i, 8 * b + c)] = v;
  }
  __syncthreads();
  // H_{b,l} = L_bb^-1 L_{b,l}, in place: 8 threads (one per row) per block
  {
    const int grp = tid >> 3, i = tid & 7;
    constexpr int NPAIR = SV_NBLK8 * (SV_NBLK8 - 1) / 2;
    for (int pr0 = 0; pr0 < NPAIR; pr0 += SV_THREADS / 8) {
      const int pr = pr0 + grp;
      const bool on = pr < NPAIR;
      int b = 1, l = 0;
      if (on) {
        while (b * (b + 1) / 2 <= pr) ++b;
        l = pr - b * (b - 1) / 2;
      }
      double h[8];
      if (on) {
#pragma unroll
        for (int c = 0; c < 8; ++c) h[c] = 0.0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const double li = Hs[sv_hc(8 * b + i, 8 * b + m)];
#pragma unroll
          for (int c = 0; c < 8; ++c) h[c] = fma(li, Hs[sv_hc(8 * b + m, 8 * l + c)], h[c]);
        }
      }
      __syncwarp();
      if (on) {
#pragma unroll
        for (int c = 0; c < 8; ++c) Hs[sv_hc(8 * b + i, 8 * l + c)] = h[c];
      }
      __syncwarp();
    }
  }

  // Accumulators of the tile products.  Blocks of x arrive in four 32-entry chunks (the producer
  // publishes each as soon as it is final), and a tile is consumed chunk by chunk:
  //   forward   out[row] += sum_c S[row][c] x[c]:   chunk q = columns 32q..32q+31, lane = column within the
  //             chunk, the warp owns rows 2*warp, 2*warp+1 of every 16-row slice -> acc[2*slice + i]
  //   backward  out[c] += sum_row S[row][c] x[row]: chunk q = rows 32q..32q+31 = slices 2q, 2q+1, lane owns
  //             columns 4*lane..+3 -> acc[0..3]; lane e < 16 fetches x[16*(e/2) + 2*warp + e%2] for its warp
  double acc[16];
  // address this lane fetches of chunk q of the block at x_blk (backward: only lanes 4q..4q+3)
  auto chunk_src = [&](const double* x_blk, int q) -> const double* {
    return FWD ? x_blk + 32 * q + lane : x_blk + 16 * (lane >> 1) + 2 * warp + (lane & 1);
  };
  auto chunk = [&](const double* x_blk, int q, unsigned long long peeked, int slot0, bool real) {
    if (FWD) {
      const double xc = real ? sv_poll_from(peeked, chunk_src(x_blk, q), a.sticky) : 0.0;
#pragma unroll
      for (int sl = 0; sl < SV_SPT; ++sl) {
        int slot = slot0 + sl;
        slot = slot >= SV_STAGES ? slot - SV_STAGES : slot;
        const double* S = ring + slot * SV_STAGE_DOUBLES + (2 * warp) * SV_TB + 32 * q + lane;
        acc[2 * sl] = fma(S[0], xc, acc[2 * sl]);
        acc[2 * sl + 1] = fma(S[SV_TB], xc, acc[2 * sl + 1]);
      }
    } else {
      double xm = 0.0;
      if (real && (lane >> 2) == q) xm = sv_poll_from(peeked, chunk_src(x_blk, q), a.sticky);
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int sl = 2 * q + h;
        int slot = slot0 + sl;
        slot = slot >= SV_STAGES ? slot - SV_STAGES : slot;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double xr = __shfl_sync(0xffffffffu, xm, 2 * sl + i);
          const double* rowp = ring + slot * SV_STAGE_DOUBLES + (2 * warp + i) * SV_TB + 4 * lane;
          const double2 l01 = *reinterpret_cast<const double2*>(rowp);
          const double2 l23 = *reinterpret_cast<const double2*>(rowp + 2);
          acc[0] = fma(l01.x, xr, acc[0]); acc[1] = fma(l01.y, xr, acc[1]);
          acc[2] = fma(l23.x, xr, acc[2]); acc[3] = fma(l23.y, xr, acc[3]);
        }
      }
    }
  };

  // The critical section below (last tile, right-hand side, diagonal step, publish) is straight-line
  // code that each CTA executes once: run cold it is dominated by instruction-cache misses.  So every
  // CTA first runs it once on whatever is in shared memory with all side effects masked (pass 0)
  // while it would be waiting for its inputs anyway.
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const bool real = pass != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    // right-hand side entry of this thread's row, fetched before the tiles so that its latency is hidden
    // (forward: b; backward: the forward result of this block row over the pivot)
    double rhs_val = 0.0;
    if (real) {
      if (FWD) {
        const int e = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        const int row = SV_SR * (e >> 1) + 2 * warp + (e & 1);
        if (row < nr) rhs_val = a.x[R0 + row];
      } else if (tid < nr) {
        rhs_val = sv_poll(fwd + R0 + tid, a.sticky) / a.Dg[R0 + tid];
      }
    }
    if (real) {
      SV_STAMP(1);
      for (int t = 0; t + 1 < ntiles; ++t) {  // streaming tiles: their x blocks were published long ago
        const int j = FWD ? t : a.nblk - 1 - t;
        const double* x_blk = pub + (size_t)j * SV_TB;
        double xq[4];
        if (FWD) {
#pragma unroll
          for (int q = 0; q < 4; ++q) xq[q] = sv_poll(x_blk + 32 * q + lane, a.sticky);
        } else {
          xq[0] = lane < 16 ? sv_poll(x_blk + 16 * (lane >> 1) + 2 * warp + (lane & 1), a.sticky) : 0.0;
        }
#pragma unroll
        for (int sl = 0; sl < SV_SPT; ++sl) {
          const int st = SV_SPT * t + sl;
          asm volatile("cp.async.wait_group %0;\n" ::"n"(SV_STAGES - 2));
          __syncthreads();
          if (st + SV_STAGES - 1 < total) issue(st + SV_STAGES - 1);
          asm volatile("cp.async.commit_group;\n" ::);
          const double* S = ring + (st % SV_STAGES) * SV_STAGE_DOUBLES + (2 * warp) * SV_TB;
          if (FWD) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[2 * sl] = fma(S[32 * q + lane], xq[q], acc[2 * sl]);
              acc[2 * sl + 1] = fma(S[SV_TB + 32 * q + lane], xq[q], acc[2 * sl + 1]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double xr = __shfl_sync(0xffffffffu, xq[0], 2 * sl + i);
              const double2 l01 = *reinterpret_cast<const double2*>(S + i * SV_TB + 4 * lane);
              const double2 l23 = *reinterpret_cast<const double2*>(S + i * SV_TB + 4 * lane + 2);
              acc[0] = fma(l01.x, xr, acc[0]); acc[1] = fma(l01.y, xr, acc[1]);
              acc[2] = fma(l23.x, xr, acc[2]); acc[3] = fma(l23.y, xr, acc[3]);
            }
          }
        }
      }
    }
    if (ntiles > 0) {
      // last tile: nothing left to issue and all of it has landed -- one wait, one barrier, then chunk by chunk
      const int t = ntiles - 1;
      const int j = FWD ? t : a.nblk - 1 - t;
      if (real) {
        SV_STAMP(2);
        asm volatile("cp.async.wait_group 0;\n" ::);
      }
      __syncthreads();
      const int slot0 = (SV_SPT * t) % SV_STAGES;
      const double* x_blk = pub + (size_t)j * SV_TB;
      // chunks in the order the producer publishes them (its warp 0 first: forward rows 0..31,
      // backward rows 127..96); one look at all four first, so that their round trips overlap
      unsigned long long pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        pk[q] = (real && (FWD || (lane >> 2) == q)) ? sv_peek(chunk_src(x_blk, q)) : SV_NOT_YET;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = FWD ? k : 3 - k;
        chunk(x_blk, q, pk[q], slot0, real);
        if (real && k == 0) SV_STAMP(8);
        if (real && k == 1) SV_STAMP(9);
        if (real && k == 2) SV_STAMP(3);
      }
      if (real) SV_STAMP(4);
    }

    // ---- right-hand side of the diagonal step, in local (possibly reversed) order ----
    if (FWD) {
      warp_reduce16(acc, lane);
      const int e = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      const int row = SV_SR * (e >> 1) + 2 * warp + (e & 1);
      if ((lane & 1) == 0) ts[row] = row < nr ? rhs_val - acc[0] : 0.0;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) part[warp * SV_TB + 4 * lane + q] = acc[q];
      __syncthreads();
      if (tid < SV_TB) {
        double s2 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s2 += part[w * SV_TB + tid];
        ts[SV_TB - 1 - tid] = tid < nr ? rhs_val - s2 : 0.0;
      }
    }
    __syncthreads();
    if (real) SV_STAMP(5);

    // ---- diagonal step: v = blockdiag(L_bb^-1) t, then v_i -= H[i][8b..8b+7] x_b block by block.
    // Warp W (rows 32W..32W+31) finalises its four blocks with shuffles only and publishes them at
    // once; the 32 finished values cross to the later warps through shared memory (4 barriers, not 16).
    if (tid < SV_TB) {
      const int p = tid, pb = p >> 3, wq = p >> 5, lb = pb & 3;
      const int row = FWD ? p : SV_TB - 1 - p;
      double v = 0.0;
      {
        const double* tb = ts + 8 * pb;
#pragma unroll
        for (int c = 0; c < 8; ++c) v = fma(Hs[sv_hc(p, 8 * pb + c)], tb[c], v);
      }
      double* xw = part;  // the column sums of the backward sweep are consumed by now
      for (int W = 0; W < 4; ++W) {
        if (wq == W) {
          double h[24];
#pragma unroll
          for (int c = 0; c < 24; ++c) h[c] = (c >> 3) < lb ? Hs[sv_hc(p, 32 * W + c)] : 0.0;
#pragma unroll
          for (int sb = 0; sb < 3; ++sb) {
            // block sb of this warp is final: broadcast its 8 values through shared memory
            if (lb == sb) xw[32 * W + lane] = v;
            __syncwarp();
            const double2 b01 = *reinterpret_cast<const double2*>(xw + 32 * W + 8 * sb);
            const double2 b23 = *reinterpret_cast<const double2*>(xw + 32 * W + 8 * sb + 2);
            const double2 b45 = *reinterpret_cast<const double2*>(xw + 32 * W + 8 * sb + 4);
            const double2 b67 = *reinterpret_cast<const double2*>(xw + 32 * W + 8 * sb + 6);
            double s0 = h[8 * sb] * b01.x, s1 = h[8 * sb + 4] * b45.x;
            s0 = fma(h[8 * sb + 1], b01.y, s0); s1 = fma(h[8 * sb + 5], b45.y, s1);
            s0 = fma(h[8 * sb + 2], b23.x, s0); s1 = fma(h[8 * sb + 6], b67.x, s1);
            s0 = fma(h[8 * sb + 3], b23.y, s0); s1 = fma(h[8 * sb + 7], b67.y, s1);
            if (lb > sb) v -= s0 + s1;
          }
          if (real) {
            sv_publish(pub + R0 + row, row < nr ? v : 0.0);
            if (!FWD && row < nr) a.x[R0 + row] = v;
          }
          if (lb == 3) xw[32 * W + lane] = v;
        }
        double h[32];
        if (wq > W) {
#pragma unroll
          for (int c = 0; c < 32; ++c) h[c] = Hs[sv_hc(p, 32 * W + c)];
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        if (real) SV_STAMP(10 + W);
        if (wq > W) {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          const double* xv = xw + 32 * W;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            s0 = fma(h[c], xv[c], s0);
            s1 = fma(h[c + 8], xv[c + 8], s1);
            s2 = fma(h[c + 16], xv[c + 16], s2);
            s3 = fma(h[c + 24], xv[c + 24], s3);
          }
          v -= (s0 + s1) + (s2 + s3);
        }
      }
    }
    __syncthreads();  // pass 0: the scratch is re-used by pass 1
  }  // pass
  SV_STAMP(6);
  if (tlog && tid == 0) tlog[7] = sv_now();
#undef SV_STAMP
}

// Both sweeps of one solve in ONE launch: tickets 0..nblk-1 are the forward block rows in order,
// tickets nblk..2nblk-1 the backward block rows from the last to the first.  A ticket only ever
// waits on lower tickets, so any number of CTAs may be resident.  The backward CTAs prepare their
// diagonal tiles and prefetch their first tiles while the forward sweep is still running.
__global__ void __launch_bounds__(SV_THREADS, 1) k_trsv_fused(SvArgs a) {
  extern __shared__ __align__(16) double svm[];
  __shared__ int s_tk;
  const int tid = threadIdx.x;
  if (tid == 0) s_tk = atomicAdd(a.ticket, 1);
  __syncthreads();
  const int order = s_tk;
  if (order < a.nblk) sv_block_row<true>(a, order, svm, tid);
  else sv_block_row<false>(a, 2 * a.nblk - 1 - order, svm, tid);
  if (tid == 0 && order == 2 * a.nblk - 1) *a.ticket = 0;  // last ticket re-arms the counter for the next launch
}

}  // namespace

int trsv_init() {
  return (int)cudaFuncSetAttribute(k_trsv_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SV_SMEM);
}

static long long* g_sv_log = nullptr;  // debug: device buffer [2][nblk][16]
void trsv_set_debug_log(long long* dev) { g_sv_log = dev; }

void launch_ldlt_solve(cudaStream_t st, const FactorPlan& fp, const double* K, const double* Dg, double* x,
                       size_t sx, TrsvWork& w) {
  if (fp.N <= 0 || fp.nslots <= 0) return;
  static const int use_stream = getenv("IPMZ_TRSV_STREAM") ? atoi(getenv("IPMZ_TRSV_STREAM")) : 1;
  if (use_stream && fp.df && fp.nslots == 1 && !fp.active) {
    const DataflowPlan& p = *fp.df;
    SvArgs v;
    v.K = K; v.Dg = Dg; v.Ginv = fp.inv; v.x = x; v.xl = p.xl; v.ticket = p.solve_ticket; v.sticky = p.solve_ticket + 1;
    v.ld = fp.ld; v.N = fp.N; v.nblk = p.nt; v.tlog = g_sv_log;
    cudaMemsetAsync(p.xl, 0xff, sizeof(double) * 2 * (size_t)p.nt * SV_TB, st);  // every entry "not yet"
    k_trsv_fused<<<2 * p.nt, SV_THREADS, SV_SMEM, st>>>(v); count_launch();
    return;
  }
  const int nblk = (fp.N + TB - 1) / TB;
  if (fp.N <= TRSV_SMALL_MAX) {
    TrsvArgs a{};
    a.K = K; a.Dg = Dg; a.x = x; a.ld = fp.ld; a.N = fp.N; a.nblk = nblk;
    a.sK = fp.sK; a.sD = fp.sD; a.sx = sx; a.active = fp.active;
    k_trsv_small<<<fp.nslots, 256, sizeof(double) * (size_t)nblk * TB, st>>>(a); count_launch();
    return;
  }
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
