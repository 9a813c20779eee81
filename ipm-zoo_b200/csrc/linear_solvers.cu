// ipm-zoo_b200/csrc/linear_solvers.cu -- C ABI mirror of NumericalOptimization::LinearSolvers
// (include/NumericalOptimization/LinearSolvers.h:11-17) on host buffers, and the
// device-resident factor+solve object that bench.py times against the FP64 roofline.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/ipmz.h"
#include "ipmz_device.cuh"
#include "ipmz_kernels.h"

namespace ipmz {
int ipmz_fail(int code, const std::string& msg);
int ipmz_ensure_device(int device);
}  // namespace ipmz
using namespace ipmz;

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ipmz_fail(IPMZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));   \
  } while (0)

struct ipmz_factor_s {
  int device = 0, n = 0, ld = 0;
  double *A = nullptr, *L = nullptr, *Dg = nullptr, *b = nullptr, *x = nullptr, *inv = nullptr, *wpanel = nullptr;
  TrsvWork tw{};
  LookAhead la{};
  DataflowPlan* df = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

static FactorPlan plan(const ipmz_factor_s* h) {
  FactorPlan fp;
  fp.N = h->n; fp.ld = h->ld; fp.sK = (size_t)h->n * h->ld; fp.sD = (size_t)h->ld; fp.nslots = 1; fp.active = nullptr;
  fp.inv = h->inv; fp.sInv = factor_inv_stride(h->ld);
  fp.wpanel = h->wpanel; fp.sW = factor_wpanel_stride(h->n);
  fp.la = &h->la;
  fp.df = h->df;
  return fp;
}

extern "C" {

int ipmz_factor_create(int n, int device, ipmz_factor_handle* out) {
  if (!out || n <= 0) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  int rc;
  if ((rc = ipmz_ensure_device(device))) return rc;
  ipmz_factor_s* h = new ipmz_factor_s();
  h->device = device; h->n = n; h->ld = pad4(n);
  const size_t bytes = sizeof(double) * (size_t)n * h->ld;
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc(&h->A, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->L, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->Dg, sizeof(double) * h->ld);
  if (e == cudaSuccess) e = cudaMalloc(&h->b, sizeof(double) * h->ld);
  if (e == cudaSuccess) e = cudaMalloc(&h->x, sizeof(double) * h->ld);
  if (e == cudaSuccess) e = cudaMalloc(&h->inv, sizeof(double) * factor_inv_stride(h->ld));
  if (e == cudaSuccess) e = cudaMalloc(&h->wpanel, sizeof(double) * factor_wpanel_stride(h->n));
  h->tw.cap_blocks = (n + 63) / 64;
  if (e == cudaSuccess) e = cudaMalloc(&h->tw.flags, sizeof(int) * h->tw.cap_blocks);
  if (e == cudaSuccess) e = cudaMalloc(&h->tw.ticket, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(h->A, 0, bytes);
  if (e == cudaSuccess) e = cudaMemset(h->L, 0, bytes);
  if (e == cudaSuccess) e = cudaMemset(h->b, 0, sizeof(double) * h->ld);
  if (e == cudaSuccess) e = cudaMemset(h->x, 0, sizeof(double) * h->ld);
  if (e == cudaSuccess) e = cudaMemset(h->tw.flags, 0, sizeof(int) * h->tw.cap_blocks);
  if (e == cudaSuccess) e = cudaMemset(h->tw.ticket, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = (cudaError_t)lookahead_create(&h->la);
  if (e == cudaSuccess && dataflow_min_n() > 0 && n >= dataflow_min_n())
    e = (cudaError_t)dataflow_plan_create(&h->df, n, h->ld);
  if (e == cudaSuccess) e = cudaEventCreate(&h->e0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->e1);
  if (e != cudaSuccess) {
    ipmz_factor_destroy(h);
    return ipmz_fail(IPMZ_ERR_ALLOC, std::string("factor_create: ") + cudaGetErrorString(e));
  }
  *out = h;
  return IPMZ_OK;
}

int ipmz_factor_destroy(ipmz_factor_handle h) {
  if (!h) return IPMZ_OK;
  cudaSetDevice(h->device);
  cudaFree(h->A); cudaFree(h->L); cudaFree(h->Dg); cudaFree(h->b); cudaFree(h->x); cudaFree(h->inv); cudaFree(h->wpanel);
  cudaFree(h->tw.flags); cudaFree(h->tw.ticket);
  lookahead_destroy(&h->la);
  dataflow_plan_destroy(h->df);
  if (h->e0) cudaEventDestroy(h->e0);
  if (h->e1) cudaEventDestroy(h->e1);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return IPMZ_OK;
}

int ipmz_factor_set_matrix(ipmz_factor_handle h, const double* A_host) {
  if (!h || !A_host) return ipmz_fail(IPMZ_ERR_ARG, "null argument");
  int rc;
  if ((rc = ipmz_ensure_device(h->device))) return rc;
  CUDA_TRY(cudaMemcpy2D(h->A, sizeof(double) * h->ld, A_host, sizeof(double) * h->n, sizeof(double) * h->n, h->n,
                        cudaMemcpyHostToDevice));
  return IPMZ_OK;
}

int ipmz_factor_set_rhs(ipmz_factor_handle h, const double* b_host) {
  if (!h || !b_host) return ipmz_fail(IPMZ_ERR_ARG, "null argument");
  int rc;
  if ((rc = ipmz_ensure_device(h->device))) return rc;
  CUDA_TRY(cudaMemcpy(h->b, b_host, sizeof(double) * h->n, cudaMemcpyHostToDevice));
  return IPMZ_OK;
}

int ipmz_factor_run(ipmz_factor_handle h, int reps, int nrhs, double* ms_total) {
  if (!h || reps <= 0 || nrhs < 0) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  int rc;
  if ((rc = ipmz_ensure_device(h->device))) return rc;
  const FactorPlan fp = plan(h);
  CUDA_TRY(cudaEventRecord(h->e0, h->st));
  for (int r = 0; r < reps; ++r) {
    launch_ldlt(h->st, fp, h->A, h->L, h->Dg);
    for (int k = 0; k < nrhs; ++k) {
      CUDA_TRY(cudaMemcpyAsync(h->x, h->b, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->st));
      launch_ldlt_solve(h->st, fp, h->L, h->Dg, h->x, (size_t)h->ld, h->tw);
    }
  }
  CUDA_TRY(cudaEventRecord(h->e1, h->st));
  CUDA_TRY(cudaEventSynchronize(h->e1));
  CUDA_TRY(cudaGetLastError());
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->e0, h->e1));
  if (ms_total) *ms_total = ms;
  if (h->df) {
    int aborted = 0;
    CUDA_TRY((cudaError_t)dataflow_abort_flag(h->st, *h->df, &aborted));
    if (aborted) return ipmz_fail(IPMZ_ERR_CUDA, "dataflow factorization / streaming solve: a dependency wait timed out");
  }
  return IPMZ_OK;
}

// One profiled factorization: ms[0..2] = device time of the diagonal-block, panel and
// trailing-update kernels (CUDA events around each launch on the library stream),
// flops_syrk = algorithmic flops of the trailing updates, n_syrk = their launch count.
int ipmz_factor_profile(ipmz_factor_handle h, double* ms3, double* flops_syrk, int* n_syrk) {
  if (!h || !ms3) return ipmz_fail(IPMZ_ERR_ARG, "null argument");
  int rc;
  if ((rc = ipmz_ensure_device(h->device))) return rc;
  ms3[0] = ms3[1] = ms3[2] = 0.0;
  if (flops_syrk) *flops_syrk = 0.0;
  if (n_syrk) *n_syrk = 0;
  FactorPlan fp = plan(h);
  fp.la = nullptr;
  const int e = launch_ldlt_profiled(h->st, fp, h->A, h->L, h->Dg, ms3, flops_syrk, n_syrk);
  if (e != 0) return ipmz_fail(IPMZ_ERR_CUDA, std::string("factor_profile: ") + cudaGetErrorString((cudaError_t)e));
  return IPMZ_OK;
}

int ipmz_fp64_peak_probe(int device, double* tflops) {
  if (!tflops) return ipmz_fail(IPMZ_ERR_ARG, "null argument");
  int rc;
  if ((rc = ipmz_ensure_device(device))) return rc;
  const int e = fp64_peak_probe(0, tflops);
  if (e != 0) return ipmz_fail(IPMZ_ERR_CUDA, std::string("fp64_peak_probe: ") + cudaGetErrorString((cudaError_t)e));
  return IPMZ_OK;
}

int ipmz_factor_get_solution(ipmz_factor_handle h, double* x_host) {
  if (!h || !x_host) return ipmz_fail(IPMZ_ERR_ARG, "null argument");
  int rc;
  if ((rc = ipmz_ensure_device(h->device))) return rc;
  CUDA_TRY(cudaMemcpy(x_host, h->x, sizeof(double) * h->n, cudaMemcpyDeviceToHost));
  return IPMZ_OK;
}

// L in the reference's format: unit diagonal, zeros above (LinearSolvers.cpp:18,38).
int ipmz_factor_get_ld(ipmz_factor_handle h, double* L_host, double* D_host) {
  if (!h) return ipmz_fail(IPMZ_ERR_ARG, "null argument");
  int rc;
  if ((rc = ipmz_ensure_device(h->device))) return rc;
  const int n = h->n;
  if (L_host) {
    CUDA_TRY(cudaMemcpy2D(L_host, sizeof(double) * n, h->L, sizeof(double) * h->ld, sizeof(double) * n, n,
                          cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
      L_host[(size_t)i * n + i] = 1.0;
      for (int j = i + 1; j < n; ++j) L_host[(size_t)i * n + j] = 0.0;
    }
  }
  if (D_host) CUDA_TRY(cudaMemcpy(D_host, h->Dg, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return IPMZ_OK;
}

int ipmz_ldlt_decomposition(int n, const double* A, double* L, double* D) {
  if (n < 0 || (n > 0 && (!A || !L || !D))) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  if (n == 0) return IPMZ_OK;
  ipmz_factor_handle h = nullptr;
  int rc = ipmz_factor_create(n, 0, &h);
  if (rc) return rc;
  rc = ipmz_factor_set_matrix(h, A);
  double ms;
  if (!rc) rc = ipmz_factor_run(h, 1, 0, &ms);
  if (!rc) rc = ipmz_factor_get_ld(h, L, D);
  ipmz_factor_destroy(h);
  return rc;
}

int ipmz_overwriting_solve_ldlt(int n, const double* L, const double* D, double* b) {
  if (n < 0 || (n > 0 && (!L || !D || !b))) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  if (n == 0) return IPMZ_OK;  // LinearSolvers.cpp:46-48
  ipmz_factor_handle h = nullptr;
  int rc = ipmz_factor_create(n, 0, &h);
  if (rc) return rc;
  cudaError_t e = cudaMemcpy2D(h->L, sizeof(double) * h->ld, L, sizeof(double) * n, sizeof(double) * n, n,
                               cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->Dg, D, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->x, b, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    FactorPlan fp = plan(h);
    fp.df = nullptr;  // L and D come from the caller: the streaming solve needs the 8 x 8 inverse blocks that
                      // only a factorization on this handle produces -> block-row sweeps on L, D alone
    launch_ldlt_solve(h->st, fp, h->L, h->Dg, h->x, (size_t)h->ld, h->tw);
    e = cudaStreamSynchronize(h->st);
  }
  if (e == cudaSuccess) e = cudaMemcpy(b, h->x, sizeof(double) * n, cudaMemcpyDeviceToHost);
  ipmz_factor_destroy(h);
  if (e != cudaSuccess) return ipmz_fail(IPMZ_ERR_CUDA, std::string("solve_ldlt: ") + cudaGetErrorString(e));
  return IPMZ_OK;
}

int ipmz_factor_info(ipmz_factor_handle h, int* dataflow, int* ntasks, double* simulated_us) {
  if (!h) return ipmz_fail(IPMZ_ERR_ARG, "null argument");
  if (dataflow) *dataflow = h->df ? 1 : 0;
  if (ntasks) *ntasks = dataflow_plan_ntasks(h->df);
  if (simulated_us) *simulated_us = dataflow_plan_sim_us(h->df);
  return IPMZ_OK;
}

int ipmz_schedule_check(int n, int workers, int* counts3, double* makespan_us, double* work_us) {
  if (n <= 0) return ipmz_fail(IPMZ_ERR_ARG, "bad argument");
  return dataflow_schedule_check(n, workers, counts3, makespan_us, work_us) ? IPMZ_OK
                                                                             : ipmz_fail(IPMZ_ERR_ARG, "invalid schedule");
}

int ipmz_assembly_schedule_check(int n, int m, int* ntasks) {
  return dataflow_assembly_schedule_check(n, m, ntasks) ? IPMZ_OK : ipmz_fail(IPMZ_ERR_ARG, "invalid assembly task list");
}

int ipmz_debug_factor_timeline(ipmz_factor_handle h, double* out, int cap, int* nrec) {
  if (!h) return 1;
  if (ipmz_ensure_device(h->device)) return 2;
  const FactorPlan fp = plan(h);
  return launch_ldlt_timeline(h->st, fp, h->A, h->L, h->Dg, out, cap, nrec);
}

int ipmz_debug_phase_clocks(long long* out16) { return read_phase_clocks(out16); }

// Dataflow factorization with its per-task log (tools/df_tasklog.py): 4 x int64 per ticket
// (start ns, end ns, SM id, task words); *sim_us = makespan of the host-simulated schedule.
int ipmz_debug_factor_tasklog(ipmz_factor_handle h, long long* out, int cap_tasks, int* ntasks, double* sim_us) {
  if (!h || !h->df) return 1;
  if (ipmz_ensure_device(h->device)) return 2;
  if (sim_us) *sim_us = dataflow_plan_sim_us(h->df);
  const FactorPlan fp = plan(h);
  int rc = launch_ldlt_dataflow_logged(h->st, *h->df, h->A, h->L, h->Dg, fp.inv, out, cap_tasks, ntasks);
  if (rc) return 100 + rc;
  int ab = 0;
  rc = dataflow_abort_flag(h->st, *h->df, &ab);
  return rc ? 100 + rc : (ab ? 3 : 0);
}
// One solve with the per-block-row timestamps of the streaming sweeps: out[2][nblk][16].
int ipmz_debug_trsv_log(ipmz_factor_handle h, long long* out, int nblk) {
  if (!h || !h->df) return 1;
  if (ipmz_ensure_device(h->device)) return 2;
  long long* dev = nullptr;
  const size_t bytes = sizeof(long long) * 2 * (size_t)nblk * 16;
  if (cudaMalloc(&dev, bytes) != cudaSuccess) return 3;
  cudaMemset(dev, 0, bytes);
  const FactorPlan fp = plan(h);
  cudaMemcpyAsync(h->x, h->b, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->st);
  trsv_set_debug_log(dev);
  launch_ldlt_solve(h->st, fp, h->L, h->Dg, h->x, (size_t)h->ld, h->tw);
  trsv_set_debug_log(nullptr);
  cudaError_t e = cudaStreamSynchronize(h->st);
  if (e == cudaSuccess) e = cudaMemcpy(out, dev, bytes, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  return e == cudaSuccess ? 0 : 100 + (int)e;
}
int ipmz_debug_factor_ntasks(ipmz_factor_handle h) { return (h && h->df) ? dataflow_plan_ntasks(h->df) : 0; }

void* ipmz_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void ipmz_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
