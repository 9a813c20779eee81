// ipm-zoo_b200/csrc/solver.cu -- host driver of the device-resident Mehrotra
// predictor-corrector loop and the C ABI of include/ipmz.h.
//
// Reference control flow: Optimizer::solve_quasi_definite_ (Optimizer.cpp:77-220).  The
// iterate, slacks, duals, residuals and directions stay in HBM for the whole solve; per
// iteration the host reads back one small Scal record per problem (f, res, mu, done) to run
// the reference's stopping test and to maintain the list of still-active problems of a batch.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ipmz.h"
#include "ipmz_device.cuh"
#include "ipmz_kernels.h"

namespace ipmz {

std::atomic<unsigned long long> g_launch_count{0};
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int ipmz_fail(int code, const std::string& msg) { return fail(code, msg); }

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(IPMZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
  } while (0)

static bool g_factor_init_done[64] = {false};

static int ensure_device(int device) {
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0)
    return fail(IPMZ_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") +
                                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
  if (device < 0 || device >= cnt) return fail(IPMZ_ERR_ARG, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  if (device < 64 && !g_factor_init_done[device]) {
    const int rc = factor_init();
    if (rc != 0) return fail(IPMZ_ERR_CUDA, std::string("factor_init: ") + cudaGetErrorString((cudaError_t)rc));
    // freed handle buffers stay in the device's default pool (up to 16 GB) for the next handle
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long keep = 16ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g_factor_init_done[device] = true;
  }
  return IPMZ_OK;
}

int ipmz_ensure_device(int device) { return ensure_device(device); }

template <class T>
static int dalloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) return fail(IPMZ_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemset(*p, 0, count * sizeof(T));
  if (e != cudaSuccess) return fail(IPMZ_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
  return IPMZ_OK;
}

// Page-locked host mirror of the per-problem Scal records: the per-iteration readback is then a true asynchronous
// D2H copy on the stream's own queue (a pageable destination is staged by the driver and can queue behind the
// large H2D problem uploads of other sub-batches, which serialises upload and solve).
struct PinnedScal {
  Scal* p = nullptr;
  size_t n = 0;
  ~PinnedScal() { if (p) cudaFreeHost(p); }
  bool resize(size_t count) {
    if (p) cudaFreeHost(p);
    p = nullptr; n = 0;
    if (cudaHostAlloc((void**)&p, sizeof(Scal) * (count ? count : 1), cudaHostAllocDefault) != cudaSuccess) return false;
    std::memset(p, 0, sizeof(Scal) * (count ? count : 1));
    n = count;
    return true;
  }
  Scal* data() { return p; }
  Scal& operator[](size_t i) { return p[i]; }
  const Scal& operator[](size_t i) const { return p[i]; }
};

// Everything one batch of `count` equally-shaped QPs needs on the device.
struct Workspace {
  int device = 0, count = 0;
  ipmz_options opt{};
  View v{};
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<void*> allocs;
  double *Q = nullptr, *M = nullptr, *MT = nullptr, *c = nullptr, *lx = nullptr, *ux = nullptr, *lo = nullptr,
         *up = nullptr;
  int* active_dev = nullptr;
  double* inv = nullptr;
  double* wpanel = nullptr;
  double* MTW = nullptr;  // MT diag(W), B operand of the condensed assembly
  int* ipiv = nullptr;    // [count][ldk] Bunch-Kaufman pivots (EqualityHandling::None: indefinite KKT)
  int launch_error = 0;   // first cudaError_t a Bunch-Kaufman launcher returned (surfaced by run_ipm / newton_step)
  TrsvWork tw{};
  LookAhead la{};
  DataflowPlan* df = nullptr;  // single large QP: persistent dataflow LDL^T
  // dual-Schur normal equations: S = W^-1 + M Hx^-1 M^T as its own m x m matrix, its pivots and factorization scratch
  double *S = nullptr, *Dg2 = nullptr, *inv2 = nullptr, *wpanel2 = nullptr;
  int ldS = 0;
  AssemblyPlan* asmp = nullptr;  // single large QP, NORMAL: condensed assembly as UPD tasks of the dataflow kernel
  int Naug = 0;
  bool fused = false;         // batch handles: the persistent one-CTA-per-problem kernel (batch_fused.cu)
  int* fused_ticket = nullptr;  // control words + iteration queue of the fused batch kernel
  size_t fused_queue_cap = 0;
  // streamed solve: copy stream, the device word it overwrites with the number of resident problems, abort flag
  cudaStream_t cps = nullptr;
  cudaEvent_t ev_arm = nullptr;
  int* ready_dev = nullptr;
  int* abort_dev = nullptr;
  int* ready_host = nullptr;  // pinned: cumulative problem counts per chunk (the source of the 4-byte copies)
  int refine = 0;  // iterative-refinement steps of the normal reduction
  int refine_extra = 0;   // +1 on the late iterations of a single QP (run_ipm), see there
  bool refine_auto = false;
  // problems of the current iteration that take refinement steps (default policy: those with mu < 1e-3)
  int* refine_dev = nullptr;  // [count] compacted on the device by k_refine_list
  int nref = 0;               // how many (counted on the host from the same Scal records)
  bool refine_all = true;     // every active problem refines: the kernels run on the active list itself
  // host mirrors
  PinnedScal sc_host;
  std::vector<int> active_host;
  // trace of the last solve (single-problem handles)
  std::vector<double> tr_f, tr_res, tr_mu, tr_alpha_aff, tr_sigma, tr_alpha;
  double* steps_dev = nullptr;  // [max_iter][2][Naug] when record_steps
  double* Rlast = nullptr;      // single QP: the shorthand residuals r_* of the last Newton iteration (corrector values)
  int tr_iters = 0;
  bool iterate_set = false;

  ~Workspace() {
    cudaSetDevice(device);
    lookahead_destroy(&la);
    dataflow_plan_destroy(df);
    dataflow_assembly_plan_destroy(asmp);
    if (pooled && st) {
      cudaStreamSynchronize(st);
      if (cps) cudaStreamSynchronize(cps);
      for (void* p : allocs) cudaFreeAsync(p, st);  // back to the device's pool: the next handle of this shape reuses it
      cudaStreamSynchronize(st);
    } else {
      for (void* p : allocs) cudaFree(p);
    }
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_arm) cudaEventDestroy(ev_arm);
    if (ready_host) cudaFreeHost(ready_host);
    if (cps) cudaStreamDestroy(cps);
    if (st) cudaStreamDestroy(st);
  }
  // Device buffers of a handle can come from the device's default memory pool, stream-ordered on the handle's stream
  // (cudaMallocAsync; release threshold raised in ensure_device): creating a handle of a shape that was used before
  // costs no cudaMalloc.  Opt-in (IPMZ_POOL_ALLOC=1); default: plain cudaMalloc.
  bool pooled = false;
  template <class T>
  int alloc(T** p, size_t n) {
    if (pooled && st) {
      *p = nullptr;
      if (n == 0) n = 1;
      cudaError_t e = cudaMallocAsync((void**)p, n * sizeof(T), st);
      if (e != cudaSuccess) return fail(IPMZ_ERR_ALLOC, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
      allocs.push_back(*p);
      e = cudaMemsetAsync(*p, 0, n * sizeof(T), st);
      if (e != cudaSuccess) return fail(IPMZ_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
      return IPMZ_OK;
    }
    const int rc = dalloc(p, n);
    if (rc == IPMZ_OK) allocs.push_back(*p);
    return rc;
  }
  // like alloc, with the zero-fill ORDERED ON THE HANDLE'S STREAM: for the buffers the upload writes next on that
  // stream (the legacy-stream memset of alloc() is not ordered against a non-blocking stream)
  template <class T>
  int alloc_st(T** p, size_t n) {
    if (pooled || !st) return alloc(p, n);
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) return fail(IPMZ_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    allocs.push_back(*p);
    e = cudaMemsetAsync(*p, 0, n * sizeof(T), st);
    if (e != cudaSuccess) return fail(IPMZ_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
    return IPMZ_OK;
  }
};

static int check_problem(const ipmz_problem* p) {
  if (!p) return fail(IPMZ_ERR_ARG, "null problem");
  if (p->n <= 0) return fail(IPMZ_ERR_ARG, "n must be positive");
  if (p->m_ineq < 0 || p->m_eq < 0) return fail(IPMZ_ERR_ARG, "negative row count");
  if (!p->Q || !p->c || !p->l_x || !p->u_x) return fail(IPMZ_ERR_ARG, "Q, c, l_x, u_x are required");
  const bool ineq = p->ineq_bounds != IPMZ_BOUNDS_NONE && p->m_ineq > 0;
  const bool eq = p->equalities && p->m_eq > 0;
  if (ineq && (!p->A || !p->l_A || !p->u_A)) return fail(IPMZ_ERR_ARG, "A, l_A, u_A are required");
  if (eq && (!p->C || !p->d)) return fail(IPMZ_ERR_ARG, "C, d are required");
  if (p->ineq_bounds < 0 || p->ineq_bounds > 3 || p->var_bounds < 0 || p->var_bounds > 3)
    return fail(IPMZ_ERR_ARG, "bounds selector out of range");
  return IPMZ_OK;
}

static void fill_shape(Shape& s, const ipmz_problem* p) {
  const bool ineq = p->ineq_bounds != IPMZ_BOUNDS_NONE && p->m_ineq > 0;
  const bool eq = p->equalities && p->m_eq > 0;
  s.n = p->n;
  s.mi = ineq ? p->m_ineq : 0;
  s.m = s.mi + (eq ? p->m_eq : 0);
  s.ns = pad4(s.n);
  s.ms = pad4(s.m);
  s.ylo = (p->var_bounds == IPMZ_BOUNDS_LOWER || p->var_bounds == IPMZ_BOUNDS_BOTH);
  s.zup = (p->var_bounds == IPMZ_BOUNDS_UPPER || p->var_bounds == IPMZ_BOUNDS_BOTH);
  s.ilo = ineq && (p->ineq_bounds == IPMZ_BOUNDS_LOWER || p->ineq_bounds == IPMZ_BOUNDS_BOTH);
  s.iup = ineq && (p->ineq_bounds == IPMZ_BOUNDS_UPPER || p->ineq_bounds == IPMZ_BOUNDS_BOTH);
  s.clamp_x = ineq ? 0 : 1;
  s.hard_eq = (eq && p->equalities == IPMZ_EQ_NONE) ? 1 : 0;
  s.reg_eq = (eq && p->equalities == IPMZ_EQ_REGULARIZATION) ? 1 : 0;
  s.pen_eq = (eq && p->equalities == IPMZ_EQ_PENALTY) ? 1 : 0;
  s.delta_eq = 1e-4;  // overwritten from the options by create_workspace
  s.ncomp = (s.ilo + s.iup) * s.mi + ((s.hard_eq || s.reg_eq || s.pen_eq) ? 0 : 2 * (s.m - s.mi)) + (s.ylo + s.zup) * s.n;
}

// host [rows x cols] dense (count blocks back to back) -> device pitched rows
static int upload_pitched(double* dst, int ld, const double* src, int cols, size_t rows, cudaStream_t st) {
  if (rows == 0 || cols == 0) return IPMZ_OK;
  CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), src, (size_t)cols * sizeof(double),
                             (size_t)cols * sizeof(double), rows, cudaMemcpyHostToDevice, st));
  return IPMZ_OK;
}

// the reference's bound checks (EnvironmentBuilder.cpp:10-17) over problems [q0, q1)
static int check_bounds(const Workspace& w, const ipmz_problem* p, int q0, int q1) {
  const Shape& s = w.v.s;
  for (size_t i = (size_t)q0 * s.n; i < (size_t)q1 * s.n; ++i)
    if (!(p->l_x[i] < p->u_x[i])) return fail(IPMZ_ERR_BOUNDS, "l_x < u_x violated");
  for (size_t i = (size_t)q0 * s.mi; i < (size_t)q1 * s.mi; ++i)
    if (!(p->l_A[i] <= p->u_A[i])) return fail(IPMZ_ERR_BOUNDS, "l_A <= u_A violated");
  return IPMZ_OK;
}

// H2D of the data of problems [q0, q1) on stream `st` (host arrays hold all `count` problems back to back)
static int upload_range(Workspace& w, const ipmz_problem* p, int q0, int q1, cudaStream_t st) {
  const Shape& s = w.v.s;
  const int me = s.m - s.mi;
  const size_t C = (size_t)(q1 - q0), o = (size_t)q0;
  if (q1 <= q0) return IPMZ_OK;
  int rc;
  if ((rc = upload_pitched(w.Q + o * w.v.sQ, w.v.ldq, p->Q + o * s.n * s.n, s.n, C * s.n, st))) return rc;
  if ((rc = upload_pitched(w.c + o * s.ns, s.ns, p->c + o * s.n, s.n, C, st))) return rc;
  if ((rc = upload_pitched(w.lx + o * s.ns, s.ns, p->l_x + o * s.n, s.n, C, st))) return rc;
  if ((rc = upload_pitched(w.ux + o * s.ns, s.ns, p->u_x + o * s.n, s.n, C, st))) return rc;
  if (s.m > 0) {
    if (s.mi > 0 && me == 0) {
      if ((rc = upload_pitched(w.M + o * w.v.sM, w.v.ldm, p->A + o * s.mi * s.n, s.n, C * s.mi, st))) return rc;
      if ((rc = upload_pitched(w.lo + o * s.ms, s.ms, p->l_A + o * s.mi, s.mi, C, st))) return rc;
      if ((rc = upload_pitched(w.up + o * s.ms, s.ms, p->u_A + o * s.mi, s.mi, C, st))) return rc;
    } else if (s.mi == 0) {
      if ((rc = upload_pitched(w.M + o * w.v.sM, w.v.ldm, p->C + o * me * s.n, s.n, C * me, st))) return rc;
      if ((rc = upload_pitched(w.lo + o * s.ms, s.ms, p->d + o * me, me, C, st))) return rc;
      if ((rc = upload_pitched(w.up + o * s.ms, s.ms, p->d + o * me, me, C, st))) return rc;
    } else {
      for (int q = q0; q < q1; ++q) {
        double* Mq = w.M + (size_t)q * w.v.sM;
        if ((rc = upload_pitched(Mq, w.v.ldm, p->A + (size_t)q * s.mi * s.n, s.n, s.mi, st))) return rc;
        if ((rc = upload_pitched(Mq + (size_t)s.mi * w.v.ldm, w.v.ldm, p->C + (size_t)q * me * s.n, s.n, me, st)))
          return rc;
        CUDA_TRY(cudaMemcpyAsync(w.lo + (size_t)q * s.ms, p->l_A + (size_t)q * s.mi, sizeof(double) * s.mi,
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(w.up + (size_t)q * s.ms, p->u_A + (size_t)q * s.mi, sizeof(double) * s.mi,
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(w.lo + (size_t)q * s.ms + s.mi, p->d + (size_t)q * me, sizeof(double) * me,
                                 cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(w.up + (size_t)q * s.ms + s.mi, p->d + (size_t)q * me, sizeof(double) * me,
                                 cudaMemcpyHostToDevice, st));
      }
    }
  }
  return IPMZ_OK;
}

static int upload_data(Workspace& w, const ipmz_problem* p) {
  const Shape& s = w.v.s;
  int rc;
  if ((rc = check_bounds(w, p, 0, w.count))) return rc;
  if ((rc = upload_range(w, p, 0, w.count, w.st))) return rc;
  if (s.m > 0) launch_transpose(w.st, w.count, w.M, w.v.ldm, w.v.sM, w.MT, w.v.ldmt, w.v.sMT, s.m, s.n);
  CUDA_TRY(cudaGetLastError());
  return IPMZ_OK;
}

static int create_workspace(Workspace** out, int count, const ipmz_problem* p, const ipmz_options* opt_in,
                            bool batch_handle = false) {
  int rc = check_problem(p);
  if (rc) return rc;
  if (count <= 0) return fail(IPMZ_ERR_ARG, "count must be positive");
  ipmz_options opt;
  if (opt_in) opt = *opt_in; else ipmz_default_options(&opt);
  if (opt.reduction != IPMZ_REDUCTION_AUGMENTED && opt.reduction != IPMZ_REDUCTION_NORMAL &&
      opt.reduction != IPMZ_REDUCTION_FULL && opt.reduction != IPMZ_REDUCTION_DUAL_NORMAL)
    return fail(IPMZ_ERR_ARG, "unknown reduction");
  if ((rc = ensure_device(opt.device))) return rc;

  const bool timing = getenv("IPMZ_CREATE_TIMING") != nullptr;  // debug: host wall time of the phases of create
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto t_start = tnow();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto t = tnow();
    fprintf(stderr, "[ipmz_create] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_start).count());
    t_start = t;
  };
  Workspace* w = new Workspace();
  std::unique_ptr<Workspace> guard(w);
  w->device = opt.device;
  w->count = count;
  w->opt = opt;
  View& v = w->v;
  fill_shape(v.s, p);
  const Shape& s = v.s;
  // EqualityHandling::None would put a symbolic zero on the diagonal of the augmented system:
  // the reference routes that to solve_indefinite_() == ASSERT(false) (Optimizer.cpp:63-75).
  if (p->m_eq > 0 && !p->equalities)
    return fail(IPMZ_ERR_INDEFINITE, "equality rows given but Settings::equalities is off");
  v.s.delta_eq = opt.delta_eq > 0.0 ? opt.delta_eq : 1e-4;
  if (s.reg_eq && opt.reduction == IPMZ_REDUCTION_FULL)
    return fail(IPMZ_ERR_ARG, "EqualityHandling::Regularization is available in the AUGMENTED and NORMAL reductions");
  // -mu I with mu -> 0: condensing it onto dx puts C^T C / mu into the matrix (the classical ill-conditioning of the
  // penalty method; the iteration stalls at res ~ 1e-4), the quasi-definite LDL^T has no such problem
  if (s.pen_eq && opt.reduction != IPMZ_REDUCTION_AUGMENTED)
    return fail(IPMZ_ERR_ARG, "EqualityHandling::PenaltyFunction* (block -mu I) is available in the AUGMENTED reduction only");
  if (opt.reduction == IPMZ_REDUCTION_DUAL_NORMAL && (s.m == 0 || s.reg_eq || s.pen_eq))
    return fail(IPMZ_ERR_ARG, "the dual-Schur normal equations need constraint rows with slacks (no rows: Hx alone is "
                              "the AUGMENTED reduction; EqualityHandling::Regularization: AUGMENTED or NORMAL)");
  if (s.hard_eq && opt.reduction != IPMZ_REDUCTION_AUGMENTED)
    return fail(IPMZ_ERR_ARG, "EqualityHandling::None (indefinite KKT, Bunch-Kaufman) is available in the AUGMENTED reduction only");
  w->Naug = s.n + s.m;
  v.normal = (opt.reduction == IPMZ_REDUCTION_NORMAL) ? 1 : 0;
  v.full = (opt.reduction == IPMZ_REDUCTION_FULL) ? 1 : 0;
  v.dual = (opt.reduction == IPMZ_REDUCTION_DUAL_NORMAL) ? 1 : 0;
  v.fl = full_layout(s);
  v.N = v.normal ? s.n : (v.full ? v.fl.N : w->Naug);
  v.ldk = pad4(v.N);
  v.sK = (size_t)v.N * v.ldk;
  if (batch_handle && v.N <= 512) v.sK = std::max(v.sK, fused_k_doubles(v.N));  // tile-major layout of the fused batch kernel
  v.ldq = s.ns; v.ldm = s.ns; v.ldmt = s.ms;
  v.sQ = (size_t)s.n * s.ns; v.sM = (size_t)s.m * s.ns; v.sMT = (size_t)s.n * s.ms;
  v.sp = (size_t)N_NSLOTS * s.ns + (size_t)N_MSLOTS * s.ms;
  v.ssol = pad4(std::max(w->Naug, v.N));
  v.tol = opt.tolerance; v.ftb = opt.fraction_to_boundary; v.sigma_pow = opt.sigma_power;
  v.max_iter = opt.max_iter;
  w->refine = v.normal ? (opt.refine_steps < 0 ? 1 : opt.refine_steps) : 0;
  w->refine_auto = v.normal && opt.refine_steps < 0;
  const int len = std::max(s.ns, s.ms);
  v.maxblk = (len + 255) / 256;

  CUDA_TRY(cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking));
  {
    const char* e = getenv("IPMZ_POOL_ALLOC");  // opt-in: measured create 22-27 ms either way at cfg3 size (the 806 MB
    w->pooled = e && atoi(e) != 0;               // upload dominates) and the pool's first allocation costs 650 ms
  }
  CUDA_TRY(cudaEventCreate(&w->ev0));
  CUDA_TRY(cudaEventCreate(&w->ev1));
  const size_t C = (size_t)count;
#define ALLOC(ptr, n) if ((rc = w->alloc(&(ptr), (n)))) return rc
#define ALLOC_ST(ptr, n) if ((rc = w->alloc_st(&(ptr), (n)))) return rc
  // The problem data first: its buffers, zero-filled on the handle's stream, then the H2D copies and the transpose on
  // that stream.  The copy engine then works (15 ms for cfg3's 806 MB) while the host builds the factorization plans
  // and allocates the rest below.
  ALLOC_ST(w->Q, C * v.sQ); ALLOC_ST(w->M, C * v.sM); ALLOC_ST(w->MT, C * v.sMT);
  ALLOC_ST(w->c, C * s.ns); ALLOC_ST(w->lx, C * s.ns); ALLOC_ST(w->ux, C * s.ns);
  ALLOC_ST(w->lo, C * s.ms); ALLOC_ST(w->up, C * s.ms);
  v.Q = w->Q; v.M = w->M; v.MT = w->MT; v.c = w->c; v.lx = w->lx; v.ux = w->ux; v.lo = w->lo; v.up = w->up;
  if ((rc = upload_data(*w, p))) return rc;
  lap("data buffers + H2D issued");
  if (count == 1) {
    const int le = lookahead_create(&w->la);
    if (le != 0) return fail(IPMZ_ERR_CUDA, std::string("lookahead_create: ") + cudaGetErrorString((cudaError_t)le));
    if (dataflow_min_n() > 0 && v.N >= dataflow_min_n() && !v.dual) {
      const int de = dataflow_plan_create(&w->df, v.N, v.ldk);
      if (de != 0) return fail(IPMZ_ERR_CUDA, std::string("dataflow_plan_create: ") + cudaGetErrorString((cudaError_t)de));
      if (v.normal && s.m > 0) {
        const int ae = dataflow_assembly_plan_create(&w->asmp, s.n, s.m);
        if (ae != 0) return fail(IPMZ_ERR_CUDA, std::string("dataflow_assembly_plan_create: ") + cudaGetErrorString((cudaError_t)ae));
      }
    }
  }
  lap("plans");
  ALLOC(v.V, C * v.sp); ALLOC(v.D, C * v.sp); ALLOC(v.DA, C * v.sp); ALLOC(v.R, C * v.sp);
  ALLOC(v.Qx, C * s.ns); ALLOC(v.MTl, C * s.ns); ALLOC(v.tn, C * s.ns);
  ALLOC(v.Mx, C * s.ms); ALLOC(v.winv, C * s.ms); ALLOC(v.W, C * s.ms); ALLOC(v.tm, C * s.ms);
  ALLOC(v.rhs, C * (s.ns + s.ms)); ALLOC(v.sol, C * v.ssol);
  ALLOC(v.out, C * (s.ns + s.ms)); ALLOC(v.resid, C * (s.ns + s.ms)); ALLOC(v.Qd, C * s.ns);
  ALLOC(v.K, C * v.sK); ALLOC(v.Dg, C * v.ldk); ALLOC(w->inv, C * factor_inv_stride(v.ldk));
  ALLOC(w->wpanel, C * factor_wpanel_stride(v.N));
  if (v.normal) ALLOC(w->MTW, C * v.sMT);
  if (v.dual) {
    w->ldS = pad4(s.m);
    ALLOC(w->S, C * (size_t)s.m * w->ldS); ALLOC(w->Dg2, C * w->ldS);
    ALLOC(w->inv2, C * factor_inv_stride(w->ldS)); ALLOC(w->wpanel2, C * factor_wpanel_stride(s.m));
  }
  ALLOC(v.sc, C); ALLOC(v.partials, C * v.maxblk * 8); ALLOC(v.counters, C);
  ALLOC(w->active_dev, C);
  ALLOC(w->refine_dev, C);
  if (s.hard_eq) ALLOC(w->ipiv, C * v.ldk);
  w->tw.cap_blocks = (v.N + 63) / 64;
  ALLOC(w->tw.flags, C * w->tw.cap_blocks); ALLOC(w->tw.ticket, 1);
  w->fused_queue_cap = batch_handle ? fused_work_words(count, opt.max_iter) : 0;
  if (w->fused_queue_cap > ((size_t)64 << 20)) w->fused_queue_cap = 0;  // absurd iteration caps: problem-granular tickets
  ALLOC(w->fused_ticket, std::max<size_t>(8, w->fused_queue_cap));
  ALLOC(w->ready_dev, 1); ALLOC(w->abort_dev, 1);
  if (count == 1) ALLOC(w->Rlast, v.sp);
  if (opt.record_steps && count == 1) ALLOC(w->steps_dev, (size_t)std::max(1, opt.max_iter) * 2 * w->Naug);
#undef ALLOC
#undef ALLOC_ST
  v.active = nullptr;
  {
    const char* e = getenv("IPMZ_BATCH_FUSED");
    w->fused = batch_handle && !v.dual && fused_batch_applicable(v) && !(e && atoi(e) == 0);
  }
  lap("device allocations + memset");
  if (!w->sc_host.resize(count)) return fail(IPMZ_ERR_ALLOC, "cudaHostAlloc of the Scal mirror failed");
  w->active_host.resize(count);
  lap("pinned Scal mirror");
  // the zero-fills of alloc() run on the legacy stream: everything the handle's (non-blocking) stream does from here
  // on -- the initial point now, every solve later -- is ordered after them
  CUDA_TRY(cudaEventRecord(w->ev0, cudaStreamLegacy));
  CUDA_TRY(cudaStreamWaitEvent(w->st, w->ev0, 0));
  launch_initial_point(w->st, v, count);
  CUDA_TRY(cudaStreamSynchronize(w->st));
  CUDA_TRY(cudaGetLastError());
  lap("H2D + transpose + initial point (wait)");
  guard.release();
  *out = w;
  return IPMZ_OK;
}

static FactorPlan plan_of(const Workspace& w, int nslots, const int* active) {
  FactorPlan fp;
  fp.N = w.v.N; fp.ld = w.v.ldk; fp.sK = w.v.sK; fp.sD = (size_t)w.v.ldk; fp.nslots = nslots; fp.active = active;
  fp.inv = w.inv; fp.sInv = factor_inv_stride(w.v.ldk);
  fp.wpanel = w.wpanel; fp.sW = factor_wpanel_stride(w.v.N);
  fp.la = (w.count == 1 && w.la.side) ? &w.la : nullptr;
  fp.df = (w.count == 1 && nslots == 1 && !active) ? w.df : nullptr;
  return fp;
}

// dual-Schur normal equations: the three factor plans over the augmented matrix K
static FactorPlan dual_plan_stage1(const Workspace& w, int nslots, const int* active) {  // eliminate the n columns of Hx
  FactorPlan fp = plan_of(w, nslots, active);
  fp.df = nullptr; fp.la = nullptr;
  fp.ncols = w.v.s.n;
  return fp;
}
static FactorPlan dual_plan_hx(const Workspace& w, int nslots, const int* active) {  // solves with the factor of Hx
  FactorPlan fp = plan_of(w, nslots, active);
  fp.df = nullptr; fp.la = nullptr;
  fp.N = w.v.s.n;
  return fp;
}
static FactorPlan dual_plan_s(const Workspace& w, int nslots, const int* active) {  // S: m x m, its own buffers
  FactorPlan fp;
  fp.N = w.v.s.m; fp.ld = w.ldS; fp.sK = (size_t)w.v.s.m * w.ldS; fp.sD = (size_t)w.ldS; fp.nslots = nslots;
  fp.active = active;
  fp.inv = w.inv2; fp.sInv = factor_inv_stride(w.ldS);
  fp.wpanel = w.wpanel2; fp.sW = factor_wpanel_stride(w.v.s.m);
  return fp;
}

// K (augmented, assembled) -> factor of Hx in its leading block, S = -(Schur complement) as its own matrix, factor of S
static void dual_factor(Workspace& w, const View& v, int nslots, bool factor_s = true) {
  const Shape& s = v.s;
  const FactorPlan f1 = dual_plan_stage1(w, nslots, v.active);
  launch_ldlt(w.st, f1, v.K, v.K, v.Dg);
  launch_negate_block(w.st, nslots, v.active, v.K + (size_t)s.n * v.ldk + s.n, v.ldk, v.sK, w.S, w.ldS,
                      (size_t)s.m * w.ldS, s.m);
  if (factor_s) launch_ldlt(w.st, dual_plan_s(w, nslots, v.active), w.S, w.S, w.Dg2);
}

// one Newton solve of the dual-Schur normal equations; the result lands in v.sol as [dx; dlam]
static void dual_solve(Workspace& w, const View& v, int nslots) {
  const Shape& s = v.s;
  const FactorPlan fh = dual_plan_hx(w, nslots, v.active), fs = dual_plan_s(w, nslots, v.active);
  launch_dual_vec(w.st, v, nslots, 0, v.rhs, v.tm);
  launch_ldlt_solve(w.st, fh, v.K, v.Dg, v.sol, v.ssol, w.tw);
  launch_matvec(w.st, nslots, v.active, v.M, v.ldm, v.sM, s.m, s.n, v.sol, v.ssol, v.Mx, s.ms);
  launch_dual_vec(w.st, v, nslots, 1, v.rhs, v.tm);
  launch_ldlt_solve(w.st, fs, w.S, w.Dg2, v.tm, s.ms, w.tw);
  launch_matvec(w.st, nslots, v.active, v.MT, v.ldmt, v.sMT, s.n, s.m, v.tm, s.ms, v.tn, s.ns);
  launch_dual_vec(w.st, v, nslots, 2, v.rhs, v.tm);
  launch_ldlt_solve(w.st, fh, v.K, v.Dg, v.sol, v.ssol, w.tw);
  launch_dual_vec(w.st, v, nslots, 3, v.rhs, v.tm);
}

// matvecs that open an iteration: Q x, M x, M^T lambda
static void iteration_matvecs(Workspace& w, const View& v, int nslots) {
  const Shape& s = v.s;
  launch_matvec(w.st, nslots, v.active, v.Q, v.ldq, v.sQ, s.n, s.n, v.V /* x slot 0 */, v.sp, v.Qx, s.ns);
  if (s.m > 0) {
    launch_matvec(w.st, nslots, v.active, v.M, v.ldm, v.sM, s.m, s.n, v.V, v.sp, v.Mx, s.ms);
    launch_matvec(w.st, nslots, v.active, v.MT, v.ldmt, v.sMT, s.n, s.m,
                  v.V + (size_t)N_NSLOTS * s.ns /* lam slot */, v.sp, v.MTl, s.ns);
  }
}

static void assemble_and_factor(Workspace& w, const View& v, int nslots) {
  const Shape& s = v.s;
  if (v.full) {
    launch_assemble_full(w.st, v, nslots);
    return;
  }
  launch_assemble(w.st, v, nslots);
  if (v.normal && s.m > 0 && w.asmp && nslots == 1 && !v.active) {
    // sign folded into the scaled operand: the UPD task subtracts
    launch_scale_cols(w.st, 1, nullptr, v.MT, w.MTW, v.ldmt, v.sMT, s.n, s.m, v.W, s.ms, -1.0);
    if (launch_assembly_dataflow(w.st, *w.asmp, v.K, v.ldk, v.MT, w.MTW, v.ldmt) == 0) return;
    // TMA path unavailable: fall through to the SYRK kernel (which adds a positively scaled operand)
  }
  if (v.normal && s.m > 0) {
    launch_scale_cols(w.st, nslots, v.active, v.MT, w.MTW, v.ldmt, v.sMT, s.n, s.m, v.W, s.ms);
    launch_syrk_ldl(w.st, nslots, v.active, v.K, v.K, v.ldk, v.sK, v.MT, v.ldmt, v.sMT, w.MTW, v.ldmt, v.sMT, s.n,
                    s.m, 1.0);
  }
}

// Normal reduction: condensed solve of the augmented right-hand side rvec = b0|b1,
//   (Hx + M^T W M) dx = b0 + M^T W b1,   dlam = W (M dx - b1),   out (+)= [dx; dlam].
static void condensed_solve(Workspace& w, const View& v, int nslots, const double* rvec, int accumulate) {
  const Shape& s = v.s;
  const FactorPlan fp = plan_of(w, nslots, v.active);
  if (s.m > 0) {
    launch_prepare_sol(w.st, v, nslots, rvec, 0);
    launch_matvec(w.st, nslots, v.active, v.MT, v.ldmt, v.sMT, s.n, s.m, v.tm, s.ms, v.tn, s.ns);
  }
  launch_prepare_sol(w.st, v, nslots, rvec, 1);
  launch_ldlt_solve(w.st, fp, v.K, v.Dg, v.sol, v.ssol, w.tw);
  if (s.m > 0) launch_matvec(w.st, nslots, v.active, v.M, v.ldm, v.sM, s.m, s.n, v.sol, v.ssol, v.Mx, s.ms);
  launch_recover_dual(w.st, v, nslots, rvec, accumulate);
}

// one Newton solve with the current factor: rhs -> direction (mode 0: affine, 1: corrector)
static void newton_direction(Workspace& w, const View& v, int nslots, int mode) {
  const Shape& s = v.s;
  if (v.full) {  // every Delta comes out of the one solve: no back-substitution
    const FactorPlan fp = plan_of(w, nslots, v.active);
    launch_full_rhs(w.st, v, nslots);
    launch_ldlt_solve(w.st, fp, v.K, v.Dg, v.sol, v.ssol, w.tw);
    launch_full_unpack(w.st, v, nslots, mode);
    return;
  }
  if (v.dual) {
    dual_solve(w, v, nslots);
  } else if (!v.normal) {
    const FactorPlan fp = plan_of(w, nslots, v.active);
    launch_prepare_sol(w.st, v, nslots, v.rhs, 0);
    if (s.hard_eq) {
      const int e = launch_bk_solve(w.st, nslots, v.active, v.K, v.ldk, v.sK, v.N, w.ipiv, (size_t)v.ldk, v.sol, v.ssol);
      if (e != 0 && w.launch_error == 0) w.launch_error = e;
    } else launch_ldlt_solve(w.st, fp, v.K, v.Dg, v.sol, v.ssol, w.tw);
  } else {
    condensed_solve(w, v, nslots, v.rhs, 0);
    // refinement runs on the problems that need it only (per-problem decision, so a QP sees the same arithmetic
    // alone, in a batch or in a sub-batch)
    View vr = v;
    int nr = nslots;
    if (!w.refine_all) { vr.active = w.refine_dev; nr = w.nref; }
    for (int r = 0; nr > 0 && r < w.refine + w.refine_extra; ++r) {
      const size_t so = (size_t)s.ns + s.ms;
      launch_matvec(w.st, nr, vr.active, v.Q, v.ldq, v.sQ, s.n, s.n, v.out, so, v.Qd, s.ns);
      if (s.m > 0) {
        launch_matvec(w.st, nr, vr.active, v.MT, v.ldmt, v.sMT, s.n, s.m, v.out + s.ns, so, v.tn, s.ns);
        // first refinement step: out's dx is still the vector the condensed solve just multiplied by M (v.Mx)
        if (r > 0) launch_matvec(w.st, nr, vr.active, v.M, v.ldm, v.sM, s.m, s.n, v.out, so, v.Mx, s.ms);
      }
      launch_aug_residual(w.st, vr, nr);
      condensed_solve(w, vr, nr, v.resid, 1);
    }
  }
  launch_backsub_step(w.st, v, nslots, mode);
}

// Which of the `nact` active problems refine their condensed solves in this iteration.  Default policy
// (options.refine_steps < 0): those with mu < 1e-3.  Without refinement the condensed step differs from the augmented
// one by about eps * max(W) ~ eps / mu: measured over every golden case and iteration <= 4e-12 relative for
// mu >= 1e-3 and 2e-11 on cfg5 at full size (n = 4096, the ill-conditioned config): 50x below the 1e-9 parity
// bound; it reaches 1e-9 .. 1e-4 for mu below 1e-5 (tests/test_refinement_policy.py),
// so the first ~4 of ~9-11 iterations skip two of their four solves and eight of their fifteen matrix passes.
// Needs w.sc_host = the Scal records of this iteration (read back by the caller).
static void select_refinement(Workspace& w, const View& v, int nact) {
  w.refine_all = true;
  w.nref = nact;
  if (!v.normal || !w.refine_auto || w.refine + w.refine_extra <= 0) return;
  if (v.s.reg_eq) return;  // EqualityHandling::Regularization: W = 1 / delta^2 is huge from the first iterate on
  const double thr = 1e-3;
  int n = 0;
  for (int i = 0; i < nact; ++i) {
    const int q = v.active ? w.active_host[i] : i;
    if (w.sc_host[q].mu < thr) ++n;
  }
  w.nref = n;
  w.refine_all = (n == nact);
  if (!w.refine_all && n > 0) launch_refine_list(w.st, v, nact, thr, w.refine_dev);
}

static void newton_iteration(Workspace& w, const View& v, int nslots, bool update, int record_iter) {
  assemble_and_factor(w, v, nslots);
  const FactorPlan fp = plan_of(w, nslots, v.active);
  // indefinite KKT (zero diagonal block): the reference's solve_indefinite_ hook (Optimizer.cpp:75), Bunch-Kaufman
  if (v.s.hard_eq) {
    const int e = launch_bk_factor(w.st, nslots, v.active, v.K, v.ldk, v.sK, v.N, w.ipiv, (size_t)v.ldk, 0);
    if (e != 0 && w.launch_error == 0) w.launch_error = e;
  } else if (v.dual) {
    dual_factor(w, v, nslots);
  } else launch_ldlt(w.st, fp, v.K, v.K, v.Dg);
  newton_direction(w, v, nslots, 0);
  launch_mu_affine(w.st, v, nslots);
  launch_residuals_rhs(w.st, v, nslots, 1);
  newton_direction(w, v, nslots, 1);
  if (record_iter >= 0 && w.steps_dev) {
    const Shape& s = v.s;
    double* dst = w.steps_dev + (size_t)record_iter * 2 * w.Naug;
    const double* packs[2] = {v.DA, v.D};
    for (int k = 0; k < 2; ++k) {
      cudaMemcpyAsync(dst + (size_t)k * w.Naug, packs[k], sizeof(double) * s.n, cudaMemcpyDeviceToDevice, w.st);
      if (s.m > 0)
        cudaMemcpyAsync(dst + (size_t)k * w.Naug + s.n, packs[k] + (size_t)N_NSLOTS * s.ns, sizeof(double) * s.m,
                        cudaMemcpyDeviceToDevice, w.st);
    }
  }
  // the next iteration's (or the final stopping test's) residual pass overwrites R: keep the corrector's r_* for
  // ipmz_get_last_iteration (the reference leaves them in the Environment, Optimizer.cpp:188-209)
  if (w.Rlast && nslots == 1 && !v.active)
    cudaMemcpyAsync(w.Rlast, v.R, sizeof(double) * v.sp, cudaMemcpyDeviceToDevice, w.st);
  if (update) launch_update(w.st, v, nslots);
}

static int run_ipm(Workspace& w, double* ms_out) {
  int rc;
  if ((rc = ensure_device(w.device))) return rc;
  const int count = w.count;
  View v = w.v;
  w.tr_f.clear(); w.tr_res.clear(); w.tr_mu.clear();
  w.tr_alpha_aff.clear(); w.tr_sigma.clear(); w.tr_alpha.clear();
  int nact = count;
  for (int i = 0; i < count; ++i) w.active_host[i] = i;
  bool identity = true;
  CUDA_TRY(cudaEventRecord(w.ev0, w.st));
  if (count > 1 || w.fused) {
    // batch handles: a solve restarts the per-problem counters and keeps the iterates (warm start, as ipmz_solve);
    // without this a second solve would see done != 0 everywhere and return the previous results at once
    CUDA_TRY(cudaMemsetAsync(v.sc, 0, sizeof(Scal) * count, w.st));
  }
  if (w.fused) {
    // one launch: every problem runs its whole predictor-corrector loop inside a persistent CTA; the host reads the
    // per-problem records once, at the end
    v.active = nullptr;
    const int e = launch_ipm_batch(w.st, v, count, w.refine_auto ? -1 : w.refine, w.fused_ticket, nullptr, nullptr, w.fused_queue_cap);
    if (e != 0) return fail(IPMZ_ERR_CUDA, std::string("launch_ipm_batch: ") + cudaGetErrorString((cudaError_t)e));
    CUDA_TRY(cudaMemcpyAsync(w.sc_host.data(), v.sc, sizeof(Scal) * count, cudaMemcpyDeviceToHost, w.st));
    CUDA_TRY(cudaEventRecord(w.ev1, w.st));
    CUDA_TRY(cudaEventSynchronize(w.ev1));
    CUDA_TRY(cudaGetLastError());
    float fms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&fms, w.ev0, w.ev1));
    if (ms_out) *ms_out = fms;
    w.tr_iters = 0;
    return IPMZ_OK;
  }
  int it = 0;
  for (;; ++it) {
    v.active = identity ? nullptr : w.active_dev;
    iteration_matvecs(w, v, nact);
    launch_residuals_rhs(w.st, v, nact, 0);
    CUDA_TRY(cudaMemcpyAsync(w.sc_host.data(), v.sc, sizeof(Scal) * count, cudaMemcpyDeviceToHost, w.st));
    CUDA_TRY(cudaStreamSynchronize(w.st));
    if (count == 1) {
      const Scal& sc = w.sc_host[0];
      if (it > 0) { w.tr_alpha_aff.push_back(sc.alpha_aff); w.tr_sigma.push_back(sc.sigma); w.tr_alpha.push_back(sc.alpha); }
      w.tr_f.push_back(sc.f); w.tr_res.push_back(sc.res); w.tr_mu.push_back(sc.mu);
    }
    int na = 0;
    for (int i = 0; i < nact; ++i) {
      const int q = w.active_host[i];
      if (w.sc_host[q].done == 0) w.active_host[na++] = q;
    }
    if (na != nact) {
      nact = na;
      identity = false;
      if (nact > 0)
        CUDA_TRY(cudaMemcpyAsync(w.active_dev, w.active_host.data(), sizeof(int) * nact, cudaMemcpyHostToDevice, w.st));
    }
    if (nact == 0) break;
    v.active = identity ? nullptr : w.active_dev;
    // NORMAL reduction, single QP, default refinement: one step of refinement leaves a relative step error of about
    // (cond * eps)^2, and cond(Hx + M^T W M) grows like 1/mu; on the last iterations of an ill-conditioned QP (cfg5:
    // res lands at 1.05e-8 against the reference's 8.19e-9 and the 1e-8 tolerance) that is not enough to follow the
    // reference's augmented step, so the late iterations take a second step (two more solves out of ~11 x 4).
    w.refine_extra = (count == 1 && w.refine_auto && w.sc_host[0].mu < 1e-6) ? 1 : 0;
    select_refinement(w, v, nact);
    newton_iteration(w, v, nact, true, (count == 1 && w.opt.record_steps) ? it : -1);
  }
  w.refine_extra = 0;
  w.refine_all = true;
  CUDA_TRY(cudaEventRecord(w.ev1, w.st));
  CUDA_TRY(cudaEventSynchronize(w.ev1));
  CUDA_TRY(cudaGetLastError());
  if (w.launch_error) {
    const int e = w.launch_error;
    w.launch_error = 0;
    return fail(IPMZ_ERR_CUDA, std::string("Bunch-Kaufman launch: ") + cudaGetErrorString((cudaError_t)e));
  }
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, w.ev0, w.ev1));
  if (ms_out) *ms_out = ms;
  w.tr_iters = it;
  if (w.asmp) {
    int aborted = 0;
    CUDA_TRY((cudaError_t)dataflow_assembly_abort_flag(w.st, *w.asmp, &aborted));
    if (aborted) return fail(IPMZ_ERR_CUDA, "condensed assembly on the dataflow kernel: a wait timed out");
  }
  if (w.df) {
    int aborted = 0;
    CUDA_TRY((cudaError_t)dataflow_abort_flag(w.st, *w.df, &aborted));
    if (aborted) return fail(IPMZ_ERR_CUDA, "dataflow factorization / streaming solve: a dependency wait timed out");
  }
  return IPMZ_OK;
}

static void fill_result(const Workspace& w, int q, ipmz_result* r, double ms) {
  const Scal& sc = w.sc_host[q];
  r->iterations = sc.iters;
  r->converged = sc.done == 1 ? 1 : 0;
  r->f = sc.f; r->res = sc.res; r->mu = sc.mu;
  r->solve_ms = ms;
  const double N = (double)w.v.N;
  r->factor_flops = (double)sc.iters * N * N * N / 3.0;
}

// ---- packed iterate <-> device packs -------------------------------------------------
// device slot order: n-slots X LAMY LAMZ Y Z ; m-slots LAM SV LAML LAMU SL SU (stacked ineq|eq)
struct HostLayout {
  int n, mi_host, me_host;  // sizes of the caller's packed iterate
  size_t off_x, off_ineq, off_eq, off_lamy, off_lamz, off_y, off_z, len;
};

static HostLayout host_layout(int n, int mi_host, int me_host) {
  HostLayout h;
  h.n = n; h.mi_host = mi_host; h.me_host = me_host;
  h.off_x = 0;
  h.off_ineq = n;
  h.off_eq = h.off_ineq + (size_t)6 * mi_host;
  h.off_lamy = h.off_eq + (size_t)6 * me_host;
  h.off_lamz = h.off_lamy + n;
  h.off_y = h.off_lamz + n;
  h.off_z = h.off_y + n;
  h.len = h.off_z + n;
  return h;
}

}  // namespace ipmz

// =========================================================================================
using namespace ipmz;

struct ipmz_solver_s {
  Workspace* w;
  int mi_host, me_host;  // row counts of the caller's ipmz_problem (packed iterate sizing)
};
struct ipmz_batch_s {
  Workspace* w;
  int mi_host, me_host;
  double last_ms;
};

extern "C" {

const char* ipmz_last_error(void) { return g_err.c_str(); }
const char* ipmz_version(void) { return "ipm-zoo_b200 0.1 (sm_100a)"; }

int ipmz_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) return 0;
  return cnt;
}

void ipmz_default_options(ipmz_options* opt) {
  opt->tolerance = 1e-8;
  opt->max_iter = 100;
  opt->fraction_to_boundary = 0.995;
  opt->sigma_power = 3.0;
  opt->reduction = IPMZ_REDUCTION_AUGMENTED;
  opt->device = 0;
  opt->record_steps = 0;
  opt->refine_steps = -1;
  opt->delta_eq = 1e-4;
}

unsigned long long ipmz_launch_count(void) { return g_launch_count.load(); }

int ipmz_iterate_len(const ipmz_problem* p) { return 5 * p->n + 6 * p->m_ineq + 6 * p->m_eq; }

// Host only: where the unknowns of the FULL reduction sit (FullLayout, ipmz_device.cuh).  offsets12 = start of
// dy dz dsl dsu dlam_y dlam_z dlam_l dlam_u ds dx dlam, then N; absent groups have zero width.
int ipmz_full_layout(const ipmz_problem* p, int* offsets12) {
  if (!p || !offsets12) return fail(IPMZ_ERR_ARG, "null argument");
  if (p->n <= 0 || p->m_ineq < 0 || p->m_eq < 0) return fail(IPMZ_ERR_ARG, "bad sizes");
  Shape s;
  fill_shape(s, p);
  const FullLayout f = full_layout(s);
  const int o[12] = {f.oy, f.oz, f.osl, f.osu, f.oly, f.olz, f.oll, f.olu, f.os, f.ox, f.olam, f.N};
  for (int i = 0; i < 12; ++i) offsets12[i] = o[i];
  return IPMZ_OK;
}

// Copy between the caller's packed iterate(s) and the device packs. dir 0: host->device.
static int move_pack(Workspace& w, int mi_host, int me_host, double* packed, int dir, double* devpack);
static int move_iterates(Workspace& w, int mi_host, int me_host, double* packed, int dir) {
  return move_pack(w, mi_host, me_host, packed, dir, w.v.V);
}
static int move_pack(Workspace& w, int mi_host, int me_host, double* packed, int dir, double* devpack) {
  const Shape& s = w.v.s;
  const int n = s.n, me_dev = s.m - s.mi;
  const HostLayout h = host_layout(n, mi_host, me_host);
  const size_t C = (size_t)w.count;
  std::vector<double> dev(C * w.v.sp, 0.0);
  if (ensure_device(w.device)) return IPMZ_ERR_CUDA;
  // start from the current device contents so untouched padding / absent groups survive
  CUDA_TRY(cudaMemcpy(dev.data(), devpack, sizeof(double) * dev.size(), cudaMemcpyDeviceToHost));
  for (size_t q = 0; q < C; ++q) {
    double* hp = packed + q * h.len;
    double* dp = dev.data() + q * w.v.sp;
    auto mv = [&](double* hptr, double* dptr, int len) {
      if (len <= 0) return;
      if (dir == 0) std::memcpy(dptr, hptr, sizeof(double) * len);
      else std::memcpy(hptr, dptr, sizeof(double) * len);
    };
    mv(hp + h.off_x, dp + (size_t)X * s.ns, n);
    mv(hp + h.off_lamy, dp + (size_t)LAMY * s.ns, n);
    mv(hp + h.off_lamz, dp + (size_t)LAMZ * s.ns, n);
    mv(hp + h.off_y, dp + (size_t)YS * s.ns, n);
    mv(hp + h.off_z, dp + (size_t)ZS * s.ns, n);
    // host m-group order: lam, s, lam_lower, lam_upper, slack_lower, slack_upper == device order
    double* dm = dp + (size_t)N_NSLOTS * s.ns;
    for (int k = 0; k < N_MSLOTS; ++k) {
      if (s.mi > 0) mv(hp + h.off_ineq + (size_t)k * mi_host, dm + (size_t)k * s.ms, s.mi);
      if (me_dev > 0) mv(hp + h.off_eq + (size_t)k * me_host, dm + (size_t)k * s.ms + s.mi, me_dev);
    }
  }
  if (dir == 0) CUDA_TRY(cudaMemcpy(devpack, dev.data(), sizeof(double) * dev.size(), cudaMemcpyHostToDevice));
  return IPMZ_OK;
}

int ipmz_create(const ipmz_problem* p, const ipmz_options* opt, ipmz_handle* out) {
  if (!out) return fail(IPMZ_ERR_ARG, "null out handle");
  Workspace* w = nullptr;
  const int rc = create_workspace(&w, 1, p, opt);
  if (rc) return rc;
  *out = new ipmz_solver_s{w, p->m_ineq, p->m_eq};
  return IPMZ_OK;
}

int ipmz_destroy(ipmz_handle h) {
  if (!h) return IPMZ_OK;
  delete h->w;
  delete h;
  return IPMZ_OK;
}

int ipmz_set_iterate(ipmz_handle h, const double* packed) {
  if (!h || !packed) return fail(IPMZ_ERR_ARG, "null argument");
  return move_iterates(*h->w, h->mi_host, h->me_host, const_cast<double*>(packed), 0);
}

int ipmz_get_iterate(ipmz_handle h, double* packed) {
  if (!h || !packed) return fail(IPMZ_ERR_ARG, "null argument");
  return move_iterates(*h->w, h->mi_host, h->me_host, packed, 1);
}

int ipmz_reset_iterate(ipmz_handle h) {
  if (!h) return fail(IPMZ_ERR_ARG, "null handle");
  Workspace& w = *h->w;
  int rc;
  if ((rc = ensure_device(w.device))) return rc;
  View v = w.v;
  v.active = nullptr;
  launch_initial_point(w.st, v, w.count);
  CUDA_TRY(cudaStreamSynchronize(w.st));
  return IPMZ_OK;
}

int ipmz_solve(ipmz_handle h, ipmz_result* res) {
  if (!h) return fail(IPMZ_ERR_ARG, "null handle");
  Workspace& w = *h->w;
  // a fresh solve restarts the per-problem counters but keeps the iterate (warm start
  // semantics of the reference: the Environment persists across solve() calls)
  int rc;
  if ((rc = ensure_device(w.device))) return rc;
  std::vector<Scal> zero(w.count);
  std::memset(zero.data(), 0, sizeof(Scal) * w.count);
  CUDA_TRY(cudaMemcpy(w.v.sc, zero.data(), sizeof(Scal) * w.count, cudaMemcpyHostToDevice));
  double ms = 0.0;
  if ((rc = run_ipm(w, &ms))) return rc;
  if (res) fill_result(w, 0, res, ms);
  return IPMZ_OK;
}

int ipmz_newton_step(ipmz_handle h, double* step_aff, double* step_cor, double* alpha_aff, double* sigma,
                     double* alpha) {
  if (!h) return fail(IPMZ_ERR_ARG, "null handle");
  Workspace& w = *h->w;
  int rc;
  if ((rc = ensure_device(w.device))) return rc;
  View v = w.v;
  v.active = nullptr;
  const Shape& s = v.s;
  iteration_matvecs(w, v, 1);
  launch_residuals_rhs(w.st, v, 1, 0);
  if (v.normal) {  // the refinement policy of run_ipm, from this iterate's mu
    CUDA_TRY(cudaMemcpyAsync(w.sc_host.data(), v.sc, sizeof(Scal), cudaMemcpyDeviceToHost, w.st));
    CUDA_TRY(cudaStreamSynchronize(w.st));
    w.refine_extra = (w.refine_auto && w.sc_host[0].mu < 1e-6) ? 1 : 0;
    select_refinement(w, v, 1);
  }
  newton_iteration(w, v, 1, false, -1);
  w.refine_extra = 0;
  w.refine_all = true;
  CUDA_TRY(cudaMemcpyAsync(w.sc_host.data(), v.sc, sizeof(Scal), cudaMemcpyDeviceToHost, w.st));
  double* outs[2] = {step_aff, step_cor};
  const double* packs[2] = {v.DA, v.D};
  for (int k = 0; k < 2; ++k) {
    if (!outs[k]) continue;
    CUDA_TRY(cudaMemcpyAsync(outs[k], packs[k], sizeof(double) * s.n, cudaMemcpyDeviceToHost, w.st));
    if (s.m > 0)
      CUDA_TRY(cudaMemcpyAsync(outs[k] + s.n, packs[k] + (size_t)N_NSLOTS * s.ns, sizeof(double) * s.m,
                               cudaMemcpyDeviceToHost, w.st));
  }
  CUDA_TRY(cudaStreamSynchronize(w.st));
  CUDA_TRY(cudaGetLastError());
  if (w.launch_error) {
    const int e = w.launch_error;
    w.launch_error = 0;
    return fail(IPMZ_ERR_CUDA, std::string("Bunch-Kaufman launch: ") + cudaGetErrorString((cudaError_t)e));
  }
  if (alpha_aff) *alpha_aff = w.sc_host[0].alpha_aff;
  if (sigma) *sigma = w.sc_host[0].sigma;
  if (alpha) *alpha = w.sc_host[0].alpha;
  return IPMZ_OK;
}

int ipmz_get_last_iteration(ipmz_handle h, double* delta, double* delta_affine, double* residuals, double* mu_centered) {
  if (!h) return fail(IPMZ_ERR_ARG, "null handle");
  Workspace& w = *h->w;
  int rc;
  if ((rc = ensure_device(w.device))) return rc;
  CUDA_TRY(cudaStreamSynchronize(w.st));
  if (delta && (rc = move_pack(w, h->mi_host, h->me_host, delta, 1, w.v.D))) return rc;
  if (delta_affine && (rc = move_pack(w, h->mi_host, h->me_host, delta_affine, 1, w.v.DA))) return rc;
  if (residuals && (rc = move_pack(w, h->mi_host, h->me_host, residuals, 1, w.Rlast))) return rc;
  if (mu_centered) {
    Scal sc;
    CUDA_TRY(cudaMemcpy(&sc, w.v.sc, sizeof(Scal), cudaMemcpyDeviceToHost));
    *mu_centered = sc.mu_c;
  }
  return IPMZ_OK;
}

int ipmz_get_trace(ipmz_handle h, int cap, double* f, double* res, double* mu, double* step_aff, double* step_cor,
                   double* alpha_aff, double* sigma, double* alpha) {
  if (!h) return fail(IPMZ_ERR_ARG, "null handle");
  Workspace& w = *h->w;
  const int nlog = (int)w.tr_f.size();
  for (int i = 0; i < nlog && i < cap; ++i) {
    if (f) f[i] = w.tr_f[i];
    if (res) res[i] = w.tr_res[i];
    if (mu) mu[i] = w.tr_mu[i];
  }
  const int nst = (int)w.tr_alpha.size();
  for (int i = 0; i < nst && i < cap; ++i) {
    if (alpha_aff) alpha_aff[i] = w.tr_alpha_aff[i];
    if (sigma) sigma[i] = w.tr_sigma[i];
    if (alpha) alpha[i] = w.tr_alpha[i];
  }
  if ((step_aff || step_cor)) {
    if (!w.steps_dev) return fail(IPMZ_ERR_ARG, "steps were not recorded (options.record_steps)");
    if (ensure_device(w.device)) return IPMZ_ERR_CUDA;
    const int k = std::min(cap, w.tr_iters);
    for (int i = 0; i < k; ++i) {
      if (step_aff)
        CUDA_TRY(cudaMemcpy(step_aff + (size_t)i * w.Naug, w.steps_dev + (size_t)i * 2 * w.Naug,
                            sizeof(double) * w.Naug, cudaMemcpyDeviceToHost));
      if (step_cor)
        CUDA_TRY(cudaMemcpy(step_cor + (size_t)i * w.Naug, w.steps_dev + ((size_t)i * 2 + 1) * w.Naug,
                            sizeof(double) * w.Naug, cudaMemcpyDeviceToHost));
    }
  }
  return IPMZ_OK;
}

int ipmz_assemble(ipmz_handle h, double* K_host, int* N_out) {
  if (!h || !K_host) return fail(IPMZ_ERR_ARG, "null argument");
  Workspace& w = *h->w;
  int rc;
  if ((rc = ensure_device(w.device))) return rc;
  View v = w.v;
  v.active = nullptr;
  iteration_matvecs(w, v, 1);
  launch_residuals_rhs(w.st, v, 1, 0);
  assemble_and_factor(w, v, 1);
  if (v.dual) {  // S = W^-1 + M Hx^-1 M^T (the matrix of the reference's normal equations, sign flipped to SPD)
    dual_factor(w, v, 1, false);
    CUDA_TRY(cudaStreamSynchronize(w.st));
    const int m = v.s.m;
    CUDA_TRY(cudaMemcpy2D(K_host, sizeof(double) * m, w.S, sizeof(double) * w.ldS, sizeof(double) * m, m,
                          cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; ++i)
      for (int j = i + 1; j < m; ++j) K_host[(size_t)i * m + j] = K_host[(size_t)j * m + i];
    if (N_out) *N_out = m;
    return IPMZ_OK;
  }
  CUDA_TRY(cudaStreamSynchronize(w.st));
  const int N = v.N;
  CUDA_TRY(cudaMemcpy2D(K_host, sizeof(double) * N, v.K, sizeof(double) * v.ldk, sizeof(double) * N, N,
                        cudaMemcpyDeviceToHost));
  if (v.normal)  // the condensed term is accumulated on the lower triangle only
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) K_host[(size_t)i * N + j] = K_host[(size_t)j * N + i];
  if (N_out) *N_out = N;
  return IPMZ_OK;
}

// HBM-side evidence for bench.py: CUDA-event time (library stream, `reps` back-to-back launches after one
// warm-up) and ALGORITHMIC bytes per launch of the streaming kernels of one iteration, at the handle's current
// iterate.  Slots: 0 k_matvec Q x | 1 k_matvec M x | 2 k_matvec M^T lambda | 3 k_assemble (+ nothing else) |
// 4 k_residuals_rhs<0> | 5 k_backsub_step<0> (k_full_unpack<0> for FULL) | 6 k_update with alpha = 0.
int ipmz_probe_kernels(ipmz_handle h, int reps, double* ms_per_launch, double* bytes_per_launch) {
  if (!h || reps <= 0 || !ms_per_launch || !bytes_per_launch) return fail(IPMZ_ERR_ARG, "bad argument");
  Workspace& w = *h->w;
  int rc;
  if ((rc = ensure_device(w.device))) return rc;
  View v = w.v;
  v.active = nullptr;
  const Shape& s = v.s;
  const double n = s.n, m = s.m, N = v.N;
  // a consistent state for the kernels that read residuals / directions
  iteration_matvecs(w, v, 1);
  launch_residuals_rhs(w.st, v, 1, 0);
  CUDA_TRY(cudaMemsetAsync(v.sol, 0, sizeof(double) * v.ssol, w.st));
  if (v.normal) CUDA_TRY(cudaMemsetAsync(v.out, 0, sizeof(double) * (s.ns + s.ms), w.st));
  auto run = [&](int slot) {
    switch (slot) {
      case 0: launch_matvec(w.st, 1, nullptr, v.Q, v.ldq, v.sQ, s.n, s.n, v.V, v.sp, v.Qx, s.ns); break;
      case 1: if (s.m > 0) launch_matvec(w.st, 1, nullptr, v.M, v.ldm, v.sM, s.m, s.n, v.V, v.sp, v.Mx, s.ms); break;
      case 2: if (s.m > 0) launch_matvec(w.st, 1, nullptr, v.MT, v.ldmt, v.sMT, s.n, s.m, v.V + (size_t)N_NSLOTS * s.ns, v.sp, v.MTl, s.ns); break;
      case 3: if (v.full) launch_assemble_full(w.st, v, 1); else launch_assemble(w.st, v, 1); break;
      case 4: launch_residuals_rhs(w.st, v, 1, 0); break;
      case 5: if (v.full) launch_full_unpack(w.st, v, 1, 0); else launch_backsub_step(w.st, v, 1, 0); break;
      case 6: launch_update(w.st, v, 1); break;
    }
  };
  const double pack = (double)v.sp * 8.0;
  const double bytes[7] = {
      (n * n + 2 * n) * 8.0, (m * n + n + m) * 8.0, (m * n + n + m) * 8.0,
      v.full ? (N * N + n * n + 2 * m * n) * 8.0 : v.normal ? (2 * n * n) * 8.0 : (N * N + n * n + 2 * m * n) * 8.0,
      3.0 * pack,   /* reads V and the matvec results, writes R and the rhs */
      3.0 * pack,   /* reads V, R and the solution, writes the direction */
      3.0 * pack};  /* reads V and D, writes V */
  // the update probe must not move the iterate: alpha = 0 for the duration
  Scal sc0;
  CUDA_TRY(cudaMemcpyAsync(&sc0, v.sc, sizeof(Scal), cudaMemcpyDeviceToHost, w.st));
  CUDA_TRY(cudaStreamSynchronize(w.st));
  Scal scz = sc0;
  scz.alpha = 0.0;
  CUDA_TRY(cudaMemcpyAsync(v.sc, &scz, sizeof(Scal), cudaMemcpyHostToDevice, w.st));
  for (int slot = 0; slot < 7; ++slot) {
    run(slot);
    CUDA_TRY(cudaEventRecord(w.ev0, w.st));
    for (int r = 0; r < reps; ++r) run(slot);
    CUDA_TRY(cudaEventRecord(w.ev1, w.st));
    CUDA_TRY(cudaEventSynchronize(w.ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, w.ev0, w.ev1));
    ms_per_launch[slot] = (double)ms / reps;
    bytes_per_launch[slot] = bytes[slot];
  }
  CUDA_TRY(cudaMemcpyAsync(v.sc, &sc0, sizeof(Scal), cudaMemcpyHostToDevice, w.st));
  CUDA_TRY(cudaStreamSynchronize(w.st));
  CUDA_TRY(cudaGetLastError());
  return IPMZ_OK;
}

// debug (library built with -DIPMZ_FUSED_CLOCKS): SM cycles per phase of the persistent batch kernel since the last call
int ipmz_debug_fused_clocks(unsigned long long* out16) {
  if (!out16) return fail(IPMZ_ERR_ARG, "null argument");
  CUDA_TRY((cudaError_t)fused_read_clocks(out16));
  return IPMZ_OK;
}

// ---- batch ------------------------------------------------------------------------------
int ipmz_batch_create(int count, const ipmz_problem* p, const ipmz_options* opt, ipmz_batch_handle* out) {
  if (!out) return fail(IPMZ_ERR_ARG, "null out handle");
  Workspace* w = nullptr;
  const int rc = create_workspace(&w, count, p, opt, true);
  if (rc) return rc;
  *out = new ipmz_batch_s{w, p->m_ineq, p->m_eq, 0.0};
  return IPMZ_OK;
}

int ipmz_batch_destroy(ipmz_batch_handle h) {
  if (!h) return IPMZ_OK;
  delete h->w;
  delete h;
  return IPMZ_OK;
}

// the caller's problem arrays must have the handle's shape and Settings
static int check_same_shape(const ipmz_batch_s* h, const ipmz_problem* data) {
  int rc = check_problem(data);
  if (rc) return rc;
  Shape s;
  fill_shape(s, data);
  const Shape& t = h->w->v.s;
  if (data->m_ineq != h->mi_host || data->m_eq != h->me_host || s.n != t.n || s.m != t.m || s.mi != t.mi ||
      s.ylo != t.ylo || s.zup != t.zup || s.ilo != t.ilo || s.iup != t.iup || s.hard_eq != t.hard_eq || s.reg_eq != t.reg_eq || s.pen_eq != t.pen_eq)
    return fail(IPMZ_ERR_ARG, "problem data does not match the shape / Settings the batch handle was created with");
  return IPMZ_OK;
}

// The copies are asynchronous (stream-ordered before the next solve): with page-locked host arrays the caller must
// leave them untouched until the next ipmz_batch_solve* / ipmz_batch_get_* on this handle has returned.
int ipmz_batch_upload(ipmz_batch_handle h, const ipmz_problem* data) {
  if (!h || !data) return fail(IPMZ_ERR_ARG, "null argument");
  Workspace& w = *h->w;
  int rc;
  if ((rc = check_same_shape(h, data))) return rc;
  if ((rc = ensure_device(w.device))) return rc;
  if ((rc = upload_data(w, data))) return rc;
  View v = w.v;
  v.active = nullptr;
  launch_initial_point(w.st, v, w.count);
  return IPMZ_OK;
}

int ipmz_batch_solve(ipmz_batch_handle h, ipmz_result* per_problem, double* ms_total) {
  if (!h) return fail(IPMZ_ERR_ARG, "null handle");
  Workspace& w = *h->w;
  double ms = 0.0;
  const int rc = run_ipm(w, &ms);
  if (rc) return rc;
  h->last_ms = ms;
  if (ms_total) *ms_total = ms;
  if (per_problem)
    for (int q = 0; q < w.count; ++q) fill_result(w, q, per_problem + q, ms);
  return IPMZ_OK;
}

// Upload + solve as ONE pipelined operation (the end-to-end path of bench.py): the persistent batch kernel is launched
// first and its CTAs pick up problems as the copy stream delivers them -- the data of `chunks` groups of problems is
// uploaded group by group, each followed by a 4-byte copy that publishes the number of resident problems; the kernel
// builds the initial point and M^T itself.  Upload and solve overlap down to the first chunk; one launch, no host
// round trip.  Host arrays as for ipmz_batch_upload (all problems back to back); ms_total = device time from the
// launch to the last result.  Only handles on the persistent-kernel path (otherwise: upload, then solve).
int ipmz_batch_solve_streamed(ipmz_batch_handle h, const ipmz_problem* data, int chunks, ipmz_result* per_problem,
                              double* ms_total) {
  if (!h || !data) return fail(IPMZ_ERR_ARG, "null argument");
  Workspace& w = *h->w;
  int rc;
  if ((rc = check_same_shape(h, data))) return rc;
  if ((rc = ensure_device(w.device))) return rc;
  if (!w.fused) {
    if ((rc = ipmz_batch_upload(h, data))) return rc;
    return ipmz_batch_solve(h, per_problem, ms_total);
  }
  const int count = w.count;
  if (chunks < 1) chunks = 1;
  if (chunks > count) chunks = count;
  if (chunks > 256) chunks = 256;
  if ((rc = check_bounds(w, data, 0, count))) return rc;
  if (!w.cps) {
    CUDA_TRY(cudaStreamCreateWithFlags(&w.cps, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&w.ev_arm, cudaEventDisableTiming));
    CUDA_TRY(cudaHostAlloc((void**)&w.ready_host, sizeof(int) * 256, cudaHostAllocDefault));
  }
  View v = w.v;
  v.active = nullptr;
  CUDA_TRY(cudaMemsetAsync(w.ready_dev, 0, sizeof(int), w.st));
  CUDA_TRY(cudaMemsetAsync(w.abort_dev, 0, sizeof(int), w.st));
  CUDA_TRY(cudaEventRecord(w.ev_arm, w.st));
  CUDA_TRY(cudaEventRecord(w.ev0, w.st));
  const int e = launch_ipm_batch(w.st, v, count, w.refine_auto ? -1 : w.refine, w.fused_ticket, w.ready_dev, w.abort_dev,
                                 w.fused_queue_cap);
  if (e != 0) return fail(IPMZ_ERR_CUDA, std::string("launch_ipm_batch: ") + cudaGetErrorString((cudaError_t)e));
  CUDA_TRY(cudaMemcpyAsync(w.sc_host.data(), v.sc, sizeof(Scal) * count, cudaMemcpyDeviceToHost, w.st));
  CUDA_TRY(cudaEventRecord(w.ev1, w.st));
  // the copy stream starts once the ready word has been cleared
  CUDA_TRY(cudaStreamWaitEvent(w.cps, w.ev_arm, 0));
  int up_rc = IPMZ_OK;
  for (int c = 0; c < chunks; ++c) {
    const int q0 = (int)((long long)c * count / chunks), q1 = (int)((long long)(c + 1) * count / chunks);
    if (up_rc == IPMZ_OK) up_rc = upload_range(w, data, q0, q1, w.cps);
    // even after a failed copy the kernel must be released (it would otherwise wait for its time-out)
    w.ready_host[c] = q1;
    if (cudaMemcpyAsync(w.ready_dev, w.ready_host + c, sizeof(int), cudaMemcpyHostToDevice, w.cps) != cudaSuccess &&
        up_rc == IPMZ_OK)
      up_rc = fail(IPMZ_ERR_CUDA, "publishing the resident-problem count failed");
  }
  CUDA_TRY(cudaEventSynchronize(w.ev1));
  CUDA_TRY(cudaStreamSynchronize(w.cps));
  CUDA_TRY(cudaGetLastError());
  if (up_rc) return up_rc;
  int aborted = 0;
  CUDA_TRY(cudaMemcpy(&aborted, w.abort_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (aborted) return fail(IPMZ_ERR_CUDA, "streamed batch solve: the kernel timed out waiting for problem data");
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, w.ev0, w.ev1));
  h->last_ms = ms;
  if (ms_total) *ms_total = ms;
  if (per_problem)
    for (int q = 0; q < count; ++q) fill_result(w, q, per_problem + q, ms);
  return IPMZ_OK;
}

// Several batches of one device concurrently: one host thread and one CUDA stream per handle, so the
// latency-bound kernels of one sub-batch overlap the throughput-bound kernels of another (SURVEY 8e:
// "one host thread + CUDA stream set per device").  ms_total = device time from the earliest start to the
// latest end over the handles' streams.
int ipmz_batch_solve_group(int g, ipmz_batch_handle* hs, double* ms_total) {
  if (g <= 0 || !hs) return fail(IPMZ_ERR_ARG, "bad argument");
  for (int i = 0; i < g; ++i)
    if (!hs[i]) return fail(IPMZ_ERR_ARG, "null handle");
  std::vector<int> rcs(g, 0);
  std::vector<std::string> errs(g);
  std::vector<std::thread> th;
  for (int i = 0; i < g; ++i)
    th.emplace_back([&, i]() {
      double ms = 0.0;
      rcs[i] = run_ipm(*hs[i]->w, &ms);
      if (rcs[i]) errs[i] = g_err;
      hs[i]->last_ms = ms;
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < g; ++i)
    if (rcs[i]) return fail(rcs[i], errs[i]);
  if (ensure_device(hs[0]->w->device)) return IPMZ_ERR_CUDA;
  float best = 0.f;
  for (int i = 0; i < g; ++i)
    for (int j = 0; j < g; ++j) {
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, hs[i]->w->ev0, hs[j]->w->ev1));
      if (ms > best) best = ms;
    }
  if (ms_total) *ms_total = best;
  return IPMZ_OK;
}

// iterations / converged / f / res / mu of every problem of the handle's last solve
int ipmz_batch_results(ipmz_batch_handle h, ipmz_result* per_problem) {
  if (!h || !per_problem) return fail(IPMZ_ERR_ARG, "null argument");
  for (int q = 0; q < h->w->count; ++q) fill_result(*h->w, q, per_problem + q, h->last_ms);
  return IPMZ_OK;
}

int ipmz_batch_get_iterates(ipmz_batch_handle h, double* packed) {
  if (!h || !packed) return fail(IPMZ_ERR_ARG, "null argument");
  return move_iterates(*h->w, h->mi_host, h->me_host, packed, 1);
}

int ipmz_batch_get_x(ipmz_batch_handle h, double* x) {
  if (!h || !x) return fail(IPMZ_ERR_ARG, "null argument");
  Workspace& w = *h->w;
  if (ensure_device(w.device)) return IPMZ_ERR_CUDA;
  CUDA_TRY(cudaMemcpy2DAsync(x, sizeof(double) * w.v.s.n, w.v.V, sizeof(double) * w.v.sp, sizeof(double) * w.v.s.n,
                             w.count, cudaMemcpyDeviceToHost, w.st));
  CUDA_TRY(cudaStreamSynchronize(w.st));
  return IPMZ_OK;
}

}  // extern "C"
