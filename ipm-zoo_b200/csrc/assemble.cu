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

// One warp per row of K, 8 rows per CTA, 16-byte accesses (every leading dimension is a multiple
// of 4 doubles and every row base 32-byte aligned; the padding of Q / M / MT is zero).
constexpr int AS_ROWS = 8;

__device__ __forceinline__ void copy_row2(double* __restrict__ dst, const double* __restrict__ src, int len2, int lane) {
  for (int c = lane; c < len2; c += 32)
    reinterpret_cast<double2*>(dst)[c] = reinterpret_cast<const double2*>(src)[c];
}

__global__ void __launch_bounds__(32 * AS_ROWS) k_assemble(View v) {
  const int p = problem_of(v);
  const Shape& s = v.s;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * AS_ROWS + (threadIdx.x >> 5);
  if (r >= v.N) return;
  double* V = v.V + (size_t)p * v.sp;
  double* Krow = v.K + (size_t)p * v.sK + (size_t)r * v.ldk;
  const bool even_n = (s.n & 1) == 0;  // the (1,2) / (2,2) blocks start at column n: 16-byte aligned only then
  if (r < s.n) {
    const double* q = v.Q + (size_t)p * v.sQ + (size_t)r * v.ldq;
    copy_row2(Krow, q, (v.normal ? v.ldk : s.n) >> 1, lane);  // normal: N = n and ldk = ldq = pad4(n), padding copies zeros
    if (!v.normal && (s.n & 1)) { if (lane == 0) Krow[s.n - 1] = q[s.n - 1]; }
    __syncwarp();
    if (lane == 0) {
      double dii = q[r];
      if (s.ylo) dii = dii + inv_guard(nslot(V, s, YS)[r]) * nslot(V, s, LAMY)[r];
      if (s.zup) dii = dii + inv_guard(nslot(V, s, ZS)[r]) * nslot(V, s, LAMZ)[r];
      Krow[r] = dii;
    }
    if (!v.normal) {
      const double* mt = v.MT + (size_t)p * v.sMT + (size_t)r * v.ldmt;
      if (even_n) copy_row2(Krow + s.n, mt, (s.m + 1) >> 1, lane);  // odd m: also copies one zero of MT's padding into K's
      else for (int c = lane; c < s.m; c += 32) Krow[s.n + c] = mt[c];
      for (int c = v.N + lane; c < v.ldk; c += 32) Krow[c] = 0.0;
    }
  } else {
    const int j = r - s.n;
    const double* mr = v.M + (size_t)p * v.sM + (size_t)j * v.ldm;
    copy_row2(Krow, mr, s.n >> 1, lane);
    if (s.n & 1) { if (lane == 0) Krow[s.n - 1] = mr[s.n - 1]; }
    const double wi = v.winv[(size_t)p * s.ms + j];
    for (int c = lane; c < s.m; c += 32) Krow[s.n + c] = (c == j) ? -wi : 0.0;
    for (int c = v.N + lane; c < v.ldk; c += 32) Krow[c] = 0.0;
  }
}

__global__ void __launch_bounds__(32 * AS_ROWS) k_scale_cols(const double* __restrict__ in, double* __restrict__ out,
                                                             int ld, size_t sM, int rows, int cols,
                                                             const double* __restrict__ d, size_t sd,
                                                             const int* __restrict__ active, double sign) {
  const int p = active ? active[blockIdx.y] : blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * AS_ROWS + (threadIdx.x >> 5);
  if (r >= rows) return;
  const double2* src = reinterpret_cast<const double2*>(in + (size_t)p * sM + (size_t)r * ld);
  double2* dst = reinterpret_cast<double2*>(out + (size_t)p * sM + (size_t)r * ld);
  const double2* dv = reinterpret_cast<const double2*>(d + (size_t)p * sd);
  // ld and the stride of d are multiples of 4 and the padding of `in` is zero: whole double2's
  for (int c = lane; c < (cols + 1) >> 1; c += 32) {
    const double2 a = src[c], w = dv[c];
    dst[c] = make_double2(sign * (a.x * w.x), 2 * c + 1 < cols ? sign * (a.y * w.y) : 0.0);
  }
}

void launch_scale_cols(cudaStream_t st, int nslots, const int* active, const double* in, double* out, int ld,
                       size_t sM, int rows, int cols, const double* d, size_t sd, double sign) {
  if (rows <= 0 || cols <= 0 || nslots <= 0) return;
  k_scale_cols<<<dim3((rows + AS_ROWS - 1) / AS_ROWS, nslots), 32 * AS_ROWS, 0, st>>>(in, out, ld, sM, rows, cols, d, sd,
                                                                                     active, sign);
  count_launch();
}

void launch_assemble(cudaStream_t st, const View& v, int nslots) {
  dim3 grid((v.N + AS_ROWS - 1) / AS_ROWS, nslots);
  k_assemble<<<grid, 32 * AS_ROWS, 0, st>>>(v); count_launch();
}

}  // namespace ipmz
