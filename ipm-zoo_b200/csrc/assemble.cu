// ipm-zoo_b200/csrc/assemble.cu -- forms the reduced Newton matrix in HBM.
//
// Reference: Optimizer.cpp:387-391 -> Evaluation::evaluate_matrix (Evaluation.cpp:53-77) ->
// concatenate_matrices_ (Optimizer.cpp:441-501): every symbolic block of the augmented LHS is
// densified into temporaries and copied.  Here one coalesced streaming pass writes each row of
// K straight from Q / M / M^T and the current iterate (bytes: n^2 + 2mn read, N^2 written).
// For the normal reduction this pass writes Hx = Q + Y^-1 L_y + Z^-1 L_z and the condensed
// term M^T W M is accumulated by the DMMA kernel k_syrk_ldl (factor.cu).
#include "ipmz_device.cuh"
#include "ipmz_kernels.h"

namespace ipmz {

__global__ void __launch_bounds__(256) k_assemble(View v) {
  const int p = problem_of(v);
  const Shape& s = v.s;
  const int r = blockIdx.x;
  double* V = v.V + (size_t)p * v.sp;
  double* Krow = v.K + (size_t)p * v.sK + (size_t)r * v.ldk;
  if (r < s.n) {
    double diag = 0.0;
    const double* q = v.Q + (size_t)p * v.sQ + (size_t)r * v.ldq;
    double dii = q[r];
    if (s.ylo) dii = dii + inv_guard(nslot(V, s, YS)[r]) * nslot(V, s, LAMY)[r];
    if (s.zup) dii = dii + inv_guard(nslot(V, s, ZS)[r]) * nslot(V, s, LAMZ)[r];
    diag = dii;
    for (int c = threadIdx.x; c < s.n; c += blockDim.x) Krow[c] = (c == r) ? diag : q[c];
    if (!v.normal) {
      const double* mt = v.MT + (size_t)p * v.sMT + (size_t)r * v.ldmt;
      for (int c = threadIdx.x; c < s.m; c += blockDim.x) Krow[s.n + c] = mt[c];
    }
  } else {
    const int j = r - s.n;
    const double* mr = v.M + (size_t)p * v.sM + (size_t)j * v.ldm;
    for (int c = threadIdx.x; c < s.n; c += blockDim.x) Krow[c] = mr[c];
    const double wi = v.winv[(size_t)p * s.ms + j];
    for (int c = threadIdx.x; c < s.m; c += blockDim.x) Krow[s.n + c] = (c == j) ? -wi : 0.0;
  }
  for (int c = v.N + threadIdx.x; c < v.ldk; c += blockDim.x) Krow[c] = 0.0;
}

__global__ void __launch_bounds__(256) k_scale_cols(const double* __restrict__ in, double* __restrict__ out, int ld,
                                                    size_t sM, int cols, const double* __restrict__ d, size_t sd,
                                                    const int* __restrict__ active) {
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const double* src = in + (size_t)p * sM + (size_t)blockIdx.x * ld;
  double* dst = out + (size_t)p * sM + (size_t)blockIdx.x * ld;
  const double* dv = d + (size_t)p * sd;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dst[c] = src[c] * dv[c];
}

void launch_scale_cols(cudaStream_t st, int nslots, const int* active, const double* in, double* out, int ld,
                       size_t sM, int rows, int cols, const double* d, size_t sd) {
  if (rows <= 0 || cols <= 0 || nslots <= 0) return;
  k_scale_cols<<<dim3(rows, nslots), 256, 0, st>>>(in, out, ld, sM, cols, d, sd, active); count_launch();
}

void launch_assemble(cudaStream_t st, const View& v, int nslots) {
  dim3 grid(v.N, nslots);
  k_assemble<<<grid, 256, 0, st>>>(v); count_launch();
}

}  // namespace ipmz
