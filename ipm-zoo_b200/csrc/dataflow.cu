// ipm-zoo_b200/csrc/dataflow.cu -- persistent dataflow LDL^T: the whole blocked factorization of
// one large reduced KKT matrix (LinearSolvers::ldlt_decomposition, LinearSolvers.cpp:14-42) in
// ONE kernel launch, one CTA per SM.
//
// The matrix is cut into 128 x 128 tiles and the factorization into DIAG / TRSM / UPD tile tasks
// (ldlt_schedule.hpp).  CTAs draw tasks from a single ticket counter in the order of a list
// schedule simulated on the host; every task waits for its inputs on flags in global memory
// (release/acquire at gpu scope) -- so the critical DIAG -> TRSM -> UPD chain of panel k+1, k+2, ...
// runs ahead of the bulk trailing updates as far as the data allows (dynamic look-ahead), there
// are no launch gaps and no wave quantisation between panel steps.
//
// CTA = 16 consumer warps + 4 producer warps (warp specialisation):
//   producers  fetch the next ticket, wait for its dependencies, publish the task to the
//              consumers through a 2-deep task queue and stream the UPD operands (16-wide
//              k-slices of L and W = L D) through a 5-stage cp.async ring guarded by full/empty
//              mbarriers -- they run ahead across task boundaries, so the next task's operands
//              land while the consumers still store the previous C tile;
//   consumers  4 x 4 warps of 32 x 32 DMMA (mma.sync.m8n8k4.f64) accumulators per 128 x 128 UPD
//              tile, initialised from C so the epilogue is store-only; DIAG and TRSM tasks run
//              bulk-synchronously on the consumer warps with the ring memory re-used as tile
//              storage (one-warp 32 x 32 LDL^T, tensor-pipe panel solves with 8 x 8 inverses).
#include <vector>

#define DF_TMA 0
#include "dataflow_kernel.cuh"

namespace ipmz {

int dataflow_tma_init();
int dataflow_tma_launch(cudaStream_t st, const void* args, int ctas, const double* W);
int dataflow_tma_launch_operands(cudaStream_t st, const void* args, int ctas, const double* A, int rowsA, int ldA,
                                 const double* B, int rowsB, int ldB);

int dataflow_init() {
  const int e = df_kernel_init();
  return e ? e : dataflow_tma_init();
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}
static double env_double(const char* name, double dflt) {
  const char* s = getenv(name);
  return s ? atof(s) : dflt;
}

// Large matrices have enough bulk work to group more panels per update task (less C traffic, fewer task
// switches); below ~48 tile rows the critical chain dominates and small groups keep it fed
// (n=8192: 7.22 -> 6.97 ms, n=12288: 21.3 -> 21.0 ms; n=4096: 2.03 ms with the small groups vs 2.22).
static void df_size_policy(int N, DfModel& m) {
  if ((N + DF_TILE - 1) / DF_TILE >= 48) { m.kb = 8; m.kmax = 15; m.la = 3; }
}

bool dataflow_schedule_check(int N, int workers, int* counts3, double* makespan_us, double* work_us) {
  DfModel m;
  df_size_policy(N, m);
  if (workers > 0) m.workers = workers;
  const DfSchedule s = df_build_schedule(N, m);
  if (counts3) {
    counts3[0] = counts3[1] = counts3[2] = 0;
    for (const DfTask& t : s.tasks) {
      const int type = t.type & 0xff;
      if (type <= DF_UPD) counts3[type] += 1;
      else if (type == DF_DIAGU) counts3[DF_DIAG] += 1;  // a diagonal factorization with its last updates fused in
    }
  }
  if (makespan_us) *makespan_us = s.makespan_us;
  if (work_us) *work_us = s.work_us;
  return df_validate_schedule(N, s);
}

int dataflow_min_n() { return env_int("IPMZ_DATAFLOW_MIN_N", 512); }

int dataflow_plan_create(DataflowPlan** out, int N, int ld) {
  *out = nullptr;
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  DfModel m;
  m.workers = nsm;
  df_size_policy(N, m);
  m.kb = env_int("IPMZ_DF_KB", m.kb);
  m.la = env_int("IPMZ_DF_LA", m.la);
  m.kmax = env_int("IPMZ_DF_KMAX", m.kmax);
  if (m.kmax > 15) m.kmax = 15;  // one lane per flag in wait_deps
  m.fuse_diag = env_int("IPMZ_DF_FUSE_DIAG", m.fuse_diag ? 1 : 0) != 0;
  m.diag_us = env_double("IPMZ_DF_DIAG_US", m.diag_us);
  m.trsm_us = env_double("IPMZ_DF_TRSM_US", m.trsm_us);
  m.upd_base_us = env_double("IPMZ_DF_UPD_BASE_US", m.upd_base_us);
  m.upd_panel_us = env_double("IPMZ_DF_UPD_PANEL_US", m.upd_panel_us);
  DfSchedule s = df_build_schedule(N, m);
  if (!df_validate_schedule(N, s)) return (int)cudaErrorInvalidValue;
  DataflowPlan* p = new DataflowPlan;
  p->N = N; p->ld = ld; p->nt = s.nt; p->ntasks = (int)s.tasks.size(); p->nsm = nsm;
  p->sim_makespan_us = s.makespan_us;
  p->flag_ints = 2 + 2 * (size_t)s.nt * s.nt;
  cudaError_t e = cudaMalloc(&p->d_tasks, sizeof(int4) * s.tasks.size());
  if (e == cudaSuccess) e = cudaMalloc(&p->d_flags, sizeof(int) * p->flag_ints);
  if (e == cudaSuccess) e = cudaMalloc(&p->W, sizeof(double) * (size_t)N * ld);
  if (e == cudaSuccess) e = cudaMemset(p->W, 0, sizeof(double) * (size_t)N * ld);
  if (e == cudaSuccess) e = cudaMalloc(&p->xl, sizeof(double) * 2 * (size_t)s.nt * NB);
  if (e == cudaSuccess) e = cudaMemset(p->xl, 0xff, sizeof(double) * 2 * (size_t)s.nt * NB);  // = the "not yet" pattern
  if (e == cudaSuccess) e = cudaMalloc(&p->solve_ticket, 2 * sizeof(int));  // [0] solve ticket, [1] sticky abort flag
  if (e == cudaSuccess) e = cudaMemset(p->solve_ticket, 0, 2 * sizeof(int));
  if (e == cudaSuccess) {
    static_assert(sizeof(DfTask) == sizeof(int4), "DfTask is uploaded as int4");
    e = cudaMemcpy(p->d_tasks, s.tasks.data(), sizeof(int4) * s.tasks.size(), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    cudaFree(p->d_tasks); cudaFree(p->d_flags); cudaFree(p->W); cudaFree(p->xl);
    cudaFree(p->solve_ticket);
    delete p;
    return (int)e;
  }
  *out = p;
  return 0;
}

void dataflow_plan_destroy(DataflowPlan* p) {
  if (!p) return;
  cudaFree(p->d_tasks); cudaFree(p->d_flags); cudaFree(p->W); cudaFree(p->d_tlog);
  cudaFree(p->xl); cudaFree(p->solve_ticket);
  delete p;
}

int dataflow_plan_ntasks(const DataflowPlan* p) { return p ? p->ntasks : 0; }
double dataflow_plan_sim_us(const DataflowPlan* p) { return p ? p->sim_makespan_us : 0.0; }

static void df_launch(cudaStream_t st, const DataflowPlan& p, const double* src, double* dst, double* Dg, double* Ginv,
                      long long* tlog) {
  cudaMemsetAsync(p.d_flags, 0, sizeof(int) * p.flag_ints, st);
  DfArgs a;
  a.src = src; a.dst = dst; a.W = p.W; a.Dg = Dg; a.Ginv = Ginv; a.tasks = p.d_tasks;
  a.ticket = p.d_flags; a.abort = p.d_flags + 1; a.sticky = p.solve_ticket + 1; a.rdy = p.d_flags + 2; a.cnt = a.rdy + (size_t)p.nt * p.nt;
  a.tlog = tlog;
  a.N = p.N; a.ld = p.ld; a.nt = p.nt; a.ntasks = p.ntasks;
  const int ctas = p.nsm < p.ntasks ? p.nsm : p.ntasks;
  // TMA build by default (n = 8192: 7.95 vs 8.16 ms per factor + 2 solves); the cp.async build is the fallback when
  // the driver does not export cuTensorMapEncodeTiled, and IPMZ_DF_TMA=0 selects it for A/B runs
  static const int use_tma = env_int("IPMZ_DF_TMA", 1);
  if (!use_tma || dataflow_tma_launch(st, &a, ctas, p.W) != 0) df_kernel_launch(st, a, ctas, p.W);
  count_launch();
}

void launch_ldlt_dataflow(cudaStream_t st, const DataflowPlan& p, const double* src, double* dst, double* Dg,
                          double* Ginv) {
  df_launch(st, p, src, dst, Dg, Ginv, nullptr);
}

// 0 = ok, 1 = a dependency wait of a factorization or a solve on this plan hit its watchdog since the
// plan was created (results invalid; the flag is sticky)
int dataflow_abort_flag(cudaStream_t st, const DataflowPlan& p, int* flag) {
  cudaError_t e = cudaMemcpyAsync(flag, p.solve_ticket + 1, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return (int)e;
}

// Debug / tuning: one factorization with a per-task log (start ns, end ns, SM id, task words).
int launch_ldlt_dataflow_logged(cudaStream_t st, DataflowPlan& p, const double* src, double* dst, double* Dg,
                                double* Ginv, long long* host_log, int cap_tasks, int* ntasks) {
  cudaError_t e = cudaSuccess;
  if (!p.d_tlog) e = cudaMalloc(&p.d_tlog, sizeof(long long) * 8 * (size_t)p.ntasks);
  if (e != cudaSuccess) return (int)e;
  cudaMemsetAsync(p.d_tlog, 0, sizeof(long long) * 8 * (size_t)p.ntasks, st);
  df_launch(st, p, src, dst, Dg, Ginv, p.d_tlog);
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  const int n = p.ntasks < cap_tasks ? p.ntasks : cap_tasks;
  e = cudaMemcpy(host_log, p.d_tlog, sizeof(long long) * 8 * (size_t)n, cudaMemcpyDeviceToHost);
  *ntasks = n;
  return (int)e;
}

// ---- condensed assembly K += MT diag(W) MT^T on the dataflow kernel ----------------------------------------------
// The UPD task of the factorization computes C(i,j) -= A(i, k0..k1) B(j, k0..k1)^T on 128 x 128 tiles with TMA-fed
// operands; the condensed assembly is the same contraction with A = MT (n x m) and B = -MT diag(W), every input ready
// from the start.  So it runs as a list of UPD tasks only: `rdy` flags preset, the tile counter `cnt` orders the
// chunks of one tile (at most 15 panels each), chunk-major order keeps those chunks far apart in the ticket queue,
// and the ticket queue balances the 148 CTAs dynamically (the persistent SYRK kernel assigns tiles round-robin and
// ends with a round of 8 tiles at n = 8192).
struct AssemblyPlan {
  int n = 0, m = 0, nt = 0, ntasks = 0, nsm = 0;
  int4* d_tasks = nullptr;
  int* d_flags = nullptr;  // [0] ticket, [1] abort, [2] sticky, [3] pad, then rdy[nt*nt], cnt[nt*nt]
  size_t flag_ints = 0;
};

int dataflow_assembly_plan_create(AssemblyPlan** out, int n, int m) {
  *out = nullptr;
  static const int enabled = env_int("IPMZ_ASM_DATAFLOW", 1);
  const int nt = (n + DF_TILE - 1) / DF_TILE, P = (m + DF_TILE - 1) / DF_TILE;
  if (!enabled || nt < 8 || P < 1 || P > nt) return 0;  // rdy is indexed [panel * nt + tile row]: needs P <= nt
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  // chunks of the K range per tile: a task waits on at most 15 panels, which does not matter here (every flag is
  // preset); fewer chunks = fewer C-tile reloads, more chunks = finer balance over the CTAs
  int chunks = env_int("IPMZ_ASM_CHUNKS", (P + 14) / 15);
  if (chunks < 1) chunks = 1;
  if (chunks > P) chunks = P;
  const std::vector<DfTask> tasks = df_build_assembly_tasks(n, m, chunks);
  if (!df_validate_assembly_tasks(n, m, tasks)) return (int)cudaErrorInvalidValue;
  AssemblyPlan* p = new AssemblyPlan;
  p->n = n; p->m = m; p->nt = nt; p->ntasks = (int)tasks.size(); p->nsm = nsm;
  p->flag_ints = 4 + 2 * (size_t)nt * nt;
  cudaError_t e = cudaMalloc(&p->d_tasks, sizeof(int4) * tasks.size());
  if (e == cudaSuccess) e = cudaMalloc(&p->d_flags, sizeof(int) * p->flag_ints);
  if (e == cudaSuccess) e = cudaMemset(p->d_flags, 0, sizeof(int) * p->flag_ints);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_tasks, tasks.data(), sizeof(int4) * tasks.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(p->d_tasks); cudaFree(p->d_flags);
    delete p;
    return (int)e;
  }
  *out = p;
  return 0;
}

// host only: the task list the assembly of an n x n condensed matrix with inner dimension m would run (0 tasks when
// the path does not apply) and whether it covers every tile's K range exactly once in order
bool dataflow_assembly_schedule_check(int n, int m, int* ntasks) {
  const int nt = (n + DF_TILE - 1) / DF_TILE, P = (m + DF_TILE - 1) / DF_TILE;
  if (ntasks) *ntasks = 0;
  if (nt < 8 || P < 1 || P > nt) return true;  // not applicable: the SYRK kernel runs
  const std::vector<DfTask> tasks = df_build_assembly_tasks(n, m, (P + 14) / 15);
  if (ntasks) *ntasks = (int)tasks.size();
  return df_validate_assembly_tasks(n, m, tasks);
}

void dataflow_assembly_plan_destroy(AssemblyPlan* p) {
  if (!p) return;
  cudaFree(p->d_tasks); cudaFree(p->d_flags);
  delete p;
}

// K (n x ldk, lower triangle, already holding Hx) -= MT * NB^T with NB = -MT diag(W).  Returns non-zero when the TMA
// path is unavailable (the caller falls back to the SYRK kernel).
int launch_assembly_dataflow(cudaStream_t st, const AssemblyPlan& p, double* K, int ldk, const double* MT,
                             const double* NB, int ldmt) {
  const size_t ntnt = (size_t)p.nt * p.nt;
  cudaMemsetAsync(p.d_flags, 0, sizeof(int) * 2, st);                              // ticket, abort (sticky stays)
  cudaMemsetAsync(p.d_flags + 4, 0x02, sizeof(int) * ntnt, st);                    // rdy: every panel "finished" (>= 2)
  cudaMemsetAsync(p.d_flags + 4 + ntnt, 0, sizeof(int) * ntnt, st);                // cnt: no chunk applied yet
  DfArgs a;
  a.src = K; a.dst = K; a.W = nullptr; a.Dg = nullptr; a.Ginv = nullptr; a.tasks = p.d_tasks;
  a.ticket = p.d_flags; a.abort = p.d_flags + 1; a.sticky = p.d_flags + 2; a.rdy = p.d_flags + 4; a.cnt = a.rdy + ntnt;
  a.tlog = nullptr;
  a.N = p.n; a.ld = ldk; a.nt = p.nt; a.ntasks = p.ntasks;
  const int ctas = p.nsm < p.ntasks ? p.nsm : p.ntasks;
  const int rc = dataflow_tma_launch_operands(st, &a, ctas, MT, p.n, ldmt, NB, p.n, ldmt);
  if (rc == 0) count_launch();
  return rc;
}

int dataflow_assembly_abort_flag(cudaStream_t st, const AssemblyPlan& p, int* flag) {
  cudaError_t e = cudaMemcpyAsync(flag, p.d_flags + 2, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return (int)e;
}

}  // namespace ipmz
