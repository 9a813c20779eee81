// tools/lat_probe.cu -- dependent-chain latencies (cycles) of the FP64 instructions the
// latency-bound panel / triangular-solve kernels are built from, on one warp of one SM.
#include <cuda_runtime.h>
#include <cstdio>
#define N 2048
__global__ void probe(double* out, long long* clk, double seed) {
  __shared__ double sm[256];
  sm[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  double a = seed, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  int k = 0;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = fma(a, b, c);
  t1 = clock64(); clk[k++] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = a * b;
  t1 = clock64(); clk[k++] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = a + c;
  t1 = clock64(); clk[k++] = t1 - t0;
  // DMMA dependent chain (accumulator)
  double acc[2] = {a, a};
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[0]), "+d"(acc[1]) : "d"(b), "d"(c));
  t1 = clock64(); clk[k++] = t1 - t0;
  a += acc[0] + acc[1];
  // SHFL (double) chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (i + 1) & 31);
  t1 = clock64(); clk[k++] = t1 - t0;
  // LDS.64 pointer-chase-like chain (address depends on previous value)
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double v = sm[idx & 255]; idx = (int)(v * 0.0) + ((idx + 1) & 255); a += 0; }
  t1 = clock64(); clk[k++] = t1 - t0;
  // 1.0/d division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) a = 1.0 / (a + 1.5);
  t1 = clock64(); clk[k++] = t1 - t0;
  // __drcp_rn chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) a = __drcp_rn(a + 1.5);
  t1 = clock64(); clk[k++] = t1 - t0;
  // rcp.approx + 2 Newton
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    double d = a + 1.5, r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0); r = fma(r, e, r);
    e = fma(-d, r, 1.0); r = fma(r, e, r);
    a = r;
  }
  t1 = clock64(); clk[k++] = t1 - t0;
  // independent DFMA throughput, one warp (8 chains)
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = a + j;
  t0 = clock64();
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, c);
  }
  t1 = clock64(); clk[k++] = t1 - t0;
#pragma unroll
  for (int j = 0; j < 8; ++j) a += x[j];
  // independent DMMA throughput, one warp (8 accumulators)
  double m[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) { m[j][0] = a; m[j][1] = j; }
  t0 = clock64();
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(m[j][0]), "+d"(m[j][1]) : "d"(b), "d"(c));
  }
  t1 = clock64(); clk[k++] = t1 - t0;
#pragma unroll
  for (int j = 0; j < 8; ++j) a += m[j][0] + m[j][1];
  out[threadIdx.x] = a + idx;
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 8 * 256); cudaMalloc(&clk, 8 * 32);
  probe<<<1, 32>>>(out, clk, 1.0);
  probe<<<1, 32>>>(out, clk, 1.0);
  cudaDeviceSynchronize();
  long long h[32]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA chain", "DMUL chain", "DADD chain", "DMMA acc chain", "SHFL.f64 chain", "LDS.64 chain", "1.0/d chain", "__drcp_rn chain", "rcp.approx+2NR chain", "DFMA 8-indep (per instr)", "DMMA 8-indep (per instr)"};
  for (int i = 0; i < 11; ++i) printf("%-28s %.1f cycles/op\n", names[i], (double)h[i] / (i < 9 ? N : (i == 9 ? N * 8 : N)));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
