// tools/fp64_probe.cu -- measures the FP64 ceilings this path is judged against on the box:
// DFMA issue rate, DMMA (mma.sync f64) issue rate for every shape ptxas accepts on sm_100a,
// cuBLAS DGEMM / DSYRK and cuSOLVER DPOTRF / cuBLAS DTRSV as same-box yardsticks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_probe.cu -lcublas -lcusolver -o tools/fp64_probe
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
  const double x = 1.0000001, y = 1e-9;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma884_kernel(double* out, int iters) {
  double c[NACC][2];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma1688_kernel(double* out, int iters) {
  double c[NACC][4];
  double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-3 + i;
  b[0] = 1.0 + threadIdx.x * 1e-6; b[1] = 0.5;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma16816_kernel(double* out, int iters) {
  double c[NACC][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-6 * i;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs %d clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  const int sms = prop.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int bps : {1, 2}) {
      const int threads = warps * 32, blocks = sms * bps;
      if (warps * bps > 64) continue;
      double ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
      double fl = 2.0 * 16 * iters * (double)threads * blocks;
      printf("DFMA warps/blk %2d blk/SM %d : %.2f TFLOP/s\n", warps, bps, fl / ms * 1e-9);
      ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters); });
      fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)warps * blocks;
      printf("DMMA m8n8k4  x8acc warps/blk %2d blk/SM %d : %.2f TFLOP/s\n", warps, bps, fl / ms * 1e-9);
      ms = time_ms([&] { dmma1688_kernel<8><<<blocks, threads>>>(out, iters); });
      fl = 2.0 * 16 * 8 * 8 * 8 * iters * (double)warps * blocks;
      printf("DMMA m16n8k8 x8acc warps/blk %2d blk/SM %d : %.2f TFLOP/s\n", warps, bps, fl / ms * 1e-9);
      ms = time_ms([&] { dmma16816_kernel<8><<<blocks, threads>>>(out, iters / 2); });
      fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 2) * (double)warps * blocks;
      printf("DMMA m16n8k16 x8acc warps/blk %2d blk/SM %d : %.2f TFLOP/s\n", warps, bps, fl / ms * 1e-9);
    }
  }
  {
    double ms = time_ms([&] { dmma1688_kernel<2><<<sms, 256>>>(out, iters); });
    printf("DMMA m16n8k8 x2acc 8 warps: %.2f TFLOP/s\n", 2.0 * 16 * 8 * 8 * 2 * iters * 8.0 * sms / ms * 1e-9);
    ms = time_ms([&] { dmma1688_kernel<4><<<sms, 256>>>(out, iters); });
    printf("DMMA m16n8k8 x4acc 8 warps: %.2f TFLOP/s\n", 2.0 * 16 * 8 * 8 * 4 * iters * 8.0 * sms / ms * 1e-9);
    ms = time_ms([&] { dmma1688_kernel<16><<<sms, 256>>>(out, iters); });
    printf("DMMA m16n8k8 x16acc 8 warps: %.2f TFLOP/s\n", 2.0 * 16 * 8 * 8 * 16 * iters * 8.0 * sms / ms * 1e-9);
  }
  // library yardsticks
  cublasHandle_t cb; cublasCreate(&cb);
  cusolverDnHandle_t cs; cusolverDnCreate(&cs);
  for (int n : {2048, 4096, 8192}) {
    double *A, *B, *Cm;
    size_t bytes = sizeof(double) * (size_t)n * n;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&Cm, bytes));
    std::vector<double> h((size_t)n * n);
    for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
    const double one = 1.0, zero = 0.0, mone = -1.0;
    double ms = time_ms([&] { cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, Cm, n); });
    printf("cuBLAS DGEMM NT n=%d : %.3f ms %.2f TFLOP/s\n", n, ms, 2.0 * n * n * n / ms * 1e-9);
    ms = time_ms([&] { cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, Cm, n); });
    printf("cuBLAS DGEMM TN n=%d : %.3f ms %.2f TFLOP/s\n", n, ms, 2.0 * n * n * n / ms * 1e-9);
    ms = time_ms([&] { cublasDsyrk(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &mone, A, n, &one, Cm, n); });
    printf("cuBLAS DSYRK n=k=%d : %.3f ms %.2f TFLOP/s\n", n, ms, 1.0 * n * n * n / ms * 1e-9);
    ms = time_ms([&] { cublasDsyrk(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, 256, &mone, A, n, &one, Cm, n); });
    printf("cuBLAS DSYRK n=%d k=256 : %.3f ms %.2f TFLOP/s\n", n, ms, 1.0 * n * n * 256 / ms * 1e-9);
    // SPD matrix for potrf: Cm = A*A^T + n*I
    cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, A, n, &zero, Cm, n);
    std::vector<double> hc((size_t)n * n);
    CK(cudaMemcpy(hc.data(), Cm, bytes, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) hc[(size_t)i * n + i] += n;
    int lwork = 0; cusolverDnDpotrf_bufferSize(cs, CUBLAS_FILL_MODE_LOWER, n, Cm, n, &lwork);
    double* work; int* info; CK(cudaMalloc(&work, sizeof(double) * lwork)); CK(cudaMalloc(&info, sizeof(int)));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      CK(cudaMemcpy(Cm, hc.data(), bytes, cudaMemcpyHostToDevice));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, n, Cm, n, work, lwork, info);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t;
    }
    int hinfo; CK(cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost));
    printf("cuSOLVER DPOTRF n=%d : %.3f ms %.2f TFLOP/s (info %d)\n", n, best, (double)n * n * n / 3.0 / best * 1e-9, hinfo);
    double* x; CK(cudaMalloc(&x, sizeof(double) * n)); CK(cudaMemcpy(x, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    ms = time_ms([&] { cublasDtrsv(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, Cm, n, x, 1); });
    printf("cuBLAS DTRSV n=%d : %.3f ms (%.1f GB/s on n^2/2 doubles)\n", n, ms, 4.0 * n * n / ms * 1e-6);
    cudaFree(A); cudaFree(B); cudaFree(Cm); cudaFree(work); cudaFree(info); cudaFree(x);
  }
  return 0;
}
