// tools/tri_probe.cu -- cycles of the 64 x 64 triangular tile solves of the fused batch kernel, one warp alone on an SM.
#include <cuda_runtime.h>
#include <cstdio>
#ifndef TPV
#define TPV 70
#endif
constexpr int TP = TPV;
#define FWD_ONLY
__device__ __forceinline__ void chain_old(const double* T, double* y, int lane) {
  double y0 = y[lane], y1 = y[lane + 32];
#pragma unroll 8
  for (int cc = 0; cc < 64; ++cc) {
    const double yc = __shfl_sync(0xffffffffu, cc < 32 ? y0 : y1, cc & 31);
    if (lane > cc) y0 -= T[lane * TP + cc] * yc;
    if (lane + 32 > cc) y1 -= T[(lane + 32) * TP + cc] * yc;
  }
  y[lane] = y0; y[lane + 32] = y1;
}
struct TriCoef { double l10, l20, l21, l30, l31, l32; double2 a01, a23, b01, b23; };
__device__ __forceinline__ void tri_load(TriCoef& k, const double* T, int b, const double* r0, const double* r1) {
  const int c0 = 4 * b;
  const double* d = T + c0 * TP + c0;
  k.l10 = d[TP]; k.l20 = d[2 * TP]; k.l21 = d[2 * TP + 1];
  k.l30 = d[3 * TP]; k.l31 = d[3 * TP + 1]; k.l32 = d[3 * TP + 2];
  k.a01 = *reinterpret_cast<const double2*>(r0 + c0); k.a23 = *reinterpret_cast<const double2*>(r0 + c0 + 2);
  k.b01 = *reinterpret_cast<const double2*>(r1 + c0); k.b23 = *reinterpret_cast<const double2*>(r1 + c0 + 2);
}
template <bool HI>
__device__ __forceinline__ void blk(const TriCoef& k, int b, double& y0, double& y1, int lane) {
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
  const double xs = (lane & 3) == 0 ? x0 : (lane & 3) == 1 ? x1 : (lane & 3) == 2 ? x2 : x3;
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
__device__ __forceinline__ void chain_new(const double* T, double* y, int lane) {
  double y0 = y[lane], y1 = y[lane + 32];
  const double* r0 = T + lane * TP; const double* r1 = T + (lane + 32) * TP;
  TriCoef ka, kb;
  tri_load(ka, T, 0, r0, r1);
#pragma unroll 1
  for (int b = 0; b < 8; b += 2) {
    tri_load(kb, T, b + 1, r0, r1);
    blk<false>(ka, b, y0, y1, lane);
    tri_load(ka, T, b + 2, r0, r1);
    blk<false>(kb, b + 1, y0, y1, lane);
  }
#pragma unroll 1
  for (int b = 8; b < 16; b += 2) {
    tri_load(kb, T, b + 1, r0, r1);
    blk<true>(ka, b, y0, y1, lane);
    tri_load(ka, T, b + 2 < 16 ? b + 2 : 15, r0, r1);
    blk<true>(kb, b + 1, y0, y1, lane);
  }
  y[lane] = y0; y[lane + 32] = y1;
}
__global__ void probe(long long* clk, double* out, int reps) {
  __shared__ double T[64 * TP];
  __shared__ double y[64];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * TP; i += blockDim.x) T[i] = 1e-3 * ((i * 7) % 13);
  if (threadIdx.x < 64) y[threadIdx.x] = 1.0 + threadIdx.x;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  __shared__ double ya[64], yb[64];
  ya[lane] = y[lane]; ya[lane + 32] = y[lane + 32]; yb[lane] = y[lane]; yb[lane + 32] = y[lane + 32];
  __syncwarp();
  chain_old(T, ya, lane); chain_new(T, yb, lane);
  __syncwarp();
  int diff = (ya[lane] != yb[lane]) + (ya[lane + 32] != yb[lane + 32]);
  diff = __reduce_add_sync(0xffffffffu, diff);
  if (lane == 0) clk[4] = diff;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) { chain_old(T, y, lane); __syncwarp(); }
  long long t1 = clock64();
  for (int r = 0; r < reps; ++r) { chain_new(T, y, lane); __syncwarp(); }
  long long t2 = clock64();
  // shuffle chain, double
  double a = y[lane];
  for (int r = 0; r < 1024; ++r) a = __shfl_sync(0xffffffffu, a, (r + 1) & 31) + 1e-9;
  long long t3 = clock64();
  // barrier-free LDS chain
  int idx = lane;
  for (int r = 0; r < 1024; ++r) { double v = T[idx]; idx = (int)(v * 1e-30) + ((idx + 33) & 1023); }
  long long t4 = clock64();
  if (lane == 0) { clk[0] = (t1 - t0) / reps; clk[1] = (t2 - t1) / reps; clk[2] = (t3 - t2) / 1024; clk[3] = (t4 - t3) / 1024; }
  out[lane] = y[lane] + a + idx;
}
int main() {
  long long* clk; double* out;
  cudaMalloc(&clk, 64); cudaMalloc(&out, 512);
  probe<<<1, 256>>>(clk, out, 200);
  long long h[5];
  cudaMemcpy(h, clk, 40, cudaMemcpyDeviceToHost);
  printf("per 64x64 tile: per-column chain %lld cycles, 4-column blocks %lld cycles; shfl(double)+dadd %lld; lds chain %lld\n", h[0], h[1], h[2], h[3]);
  printf("pitch %d, entries that differ between the two chains: %lld; %s\n", TP, h[4], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
