// ipm-zoo_b200/csrc/ipmz_kernels.h -- host-callable launchers of the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <atomic>

#include "ipmz_device.cuh"

namespace ipmz {

// every launcher bumps this (bench.py reports it as gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;  // several handles may run on their own host threads
inline void count_launch(int n = 1) { g_launch_count.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ---- vector_kernels.cu ----
void launch_matvec(cudaStream_t st, int nslots, const int* active, const double* A, int lda, size_t sA,
                   int rows, int cols, const double* x, size_t sx, double* y, size_t sy);
void launch_transpose(cudaStream_t st, int count, const double* M, int ldm, size_t sM, double* MT, int ldmt,
                      size_t sMT, int m, int n);
void launch_initial_point(cudaStream_t st, const View& v, int nslots);
void launch_residuals_rhs(cudaStream_t st, const View& v, int nslots, int mode);
void launch_prepare_sol(cudaStream_t st, const View& v, int nslots, const double* rvec, int stage);
void launch_recover_dual(cudaStream_t st, const View& v, int nslots, const double* rvec, int accumulate);
void launch_aug_residual(cudaStream_t st, const View& v, int nslots);
void launch_backsub_step(cudaStream_t st, const View& v, int nslots, int mode);
void launch_mu_affine(cudaStream_t st, const View& v, int nslots);
void launch_update(cudaStream_t st, const View& v, int nslots);
// out <- the problems among the first nact active ones whose mu < thr (compacted, in slot order)
void launch_refine_list(cudaStream_t st, const View& v, int nact, double thr, int* out);

// ---- assemble.cu ----
// Augmented KKT [[Q + Y^-1 L_y + Z^-1 L_z, M^T],[M, -W^-1]] (full symmetric, N = n+m) or the
// diagonal-shifted copy Q + Y^-1 L_y + Z^-1 L_z that the condensed assembly accumulates
// M^T W M onto (N = n).
void launch_assemble(cudaStream_t st, const View& v, int nslots);
// out[r][c] = in[r][c] * d[c] (pre-scaled B operand MT diag(W) of the condensed assembly)
void launch_scale_cols(cudaStream_t st, int nslots, const int* active, const double* in, double* out, int ld,
                       size_t sM, int rows, int cols, const double* d, size_t sd, double sign = 1.0);

// ---- full_system.cu: the un-reduced Newton system in the FullLayout order ----
void launch_assemble_full(cudaStream_t st, const View& v, int nslots);
void launch_full_rhs(cudaStream_t st, const View& v, int nslots);             // R, V -> sol
void launch_full_unpack(cudaStream_t st, const View& v, int nslots, int mode);  // sol -> DA / D, step length

// ---- bunch_kaufman.cu: symmetric indefinite factorization (LinearSolvers.cpp:76-318), bit-exact pivots ----
// S: in = matrix (lower triangle significant), out = L / D on and below the diagonal, L^T mirrored above it.
// Both return a cudaError_t.
int launch_bk_factor(cudaStream_t st, int nslots, const int* active, double* S, int ld, size_t sS, int n, int* ipiv,
                     size_t sP, int mirror_input);
int bk_init();  // per-device opt-in shared-memory sizes of the solve kernels
int launch_bk_solve(cudaStream_t st, int nslots, const int* active, const double* S, int ld, size_t sS, int n,
                    const int* ipiv, size_t sP, double* x, size_t sx);

// ---- factor.cu ----
// Side stream + events of the look-ahead schedule of launch_ldlt (one per handle).
struct LookAhead {
  cudaStream_t side = nullptr;  // high priority: panel k+1 while the main stream applies panel k
  cudaEvent_t e_col = nullptr, e_panel = nullptr;
};
int lookahead_create(LookAhead* la);
void lookahead_destroy(LookAhead* la);

// ---- dataflow.cu: persistent single-launch LDL^T of one large matrix ----
struct DataflowPlan {  // task list (host-simulated list schedule), dependency flags, W = L D buffer
  int N = 0, ld = 0, nt = 0, ntasks = 0, nsm = 0;
  int4* d_tasks = nullptr;
  int* d_flags = nullptr;  // [0] ticket, [1] abort, [2 ..) rdy[nt*nt], cnt[nt*nt]
  size_t flag_ints = 0;
  double* W = nullptr;     // N x ld, W = L D (pre-scaled B operand of the updates)
  double* xl = nullptr;    // 2 x nt*128 self-validating exchange buffers of the solves (forward, backward)
  int* solve_ticket = nullptr;
  long long* d_tlog = nullptr;  // optional: 4 x ntasks (start ns, end ns, SM id, task words)
  double sim_makespan_us = 0.0;
};
int dataflow_init();
int dataflow_min_n();  // matrices at least this large use the dataflow kernel (IPMZ_DATAFLOW_MIN_N; 0 = never)
int dataflow_plan_create(DataflowPlan** out, int N, int ld);
// host only: build + validate the task list (true = valid topological order covering every tile)
bool dataflow_schedule_check(int N, int workers, int* counts3, double* makespan_us, double* work_us);
void dataflow_plan_destroy(DataflowPlan* p);
int dataflow_plan_ntasks(const DataflowPlan* p);
double dataflow_plan_sim_us(const DataflowPlan* p);
void launch_ldlt_dataflow(cudaStream_t st, const DataflowPlan& p, const double* src, double* dst, double* Dg,
                          double* Ginv);
int dataflow_abort_flag(cudaStream_t st, const DataflowPlan& p, int* flag);
// one factorization with a per-task log: 4 x int64 per ticket (start ns, end ns, SM id, task words)
int launch_ldlt_dataflow_logged(cudaStream_t st, DataflowPlan& p, const double* src, double* dst, double* Dg,
                                double* Ginv, long long* host_log, int cap_tasks, int* ntasks);

// condensed assembly K += MT diag(W) MT^T as UPD tasks of the dataflow kernel (one large QP, TMA build)
struct AssemblyPlan;
int dataflow_assembly_plan_create(AssemblyPlan** out, int n, int m);  // *out stays null when the path does not apply
void dataflow_assembly_plan_destroy(AssemblyPlan* p);
int launch_assembly_dataflow(cudaStream_t st, const AssemblyPlan& p, double* K, int ldk, const double* MT,
                             const double* NB /* = -MT diag(W) */, int ldmt);
int dataflow_assembly_abort_flag(cudaStream_t st, const AssemblyPlan& p, int* flag);
bool dataflow_assembly_schedule_check(int n, int m, int* ntasks);  // host only

struct FactorPlan {
  int N;        // matrix dimension
  int ld;       // leading dimension (multiple of 4)
  size_t sK;    // per-problem stride of the matrix
  size_t sD;    // per-problem stride of the pivots
  int nslots;   // problems in this launch
  const int* active;
  double* inv = nullptr;  // [nslots][(ld/8 + 4) * 96] scratch: D^-1 L^-1 of the 8 x 8 diagonal blocks
  size_t sInv = 0;
  double* wpanel = nullptr;  // [nslots][2][N x 256] W = L D of the current panels (pre-scaled B operand)
  size_t sW = 0;
  int ncols = 0;  // > 0: eliminate only the first ncols columns (the rows below them are solved and every trailing
                  // block is updated: afterwards the block [ncols, N)^2 holds the Schur complement); 0 = all N
  const LookAhead* la = nullptr;  // nullptr: single-stream schedule
  DataflowPlan* df = nullptr;     // set (single large matrix): launch_ldlt runs the persistent dataflow kernel
};
inline size_t factor_inv_stride(int ld) { return (size_t)(ld / 8 + 4) * 96; }
inline size_t factor_wpanel_stride(int N) { return (size_t)2 * N * 256; }  // double-buffered
int factor_init();  // opt-in shared memory sizes; returns cudaError_t
// L, Dg <- LDL^T(src).  src == dst factors in place; otherwise the first panel step reads src
// and writes dst so no separate copy pass is needed (the reference's ldlt_decomposition is
// out-of-place, LinearSolvers.cpp:14-42).
void launch_ldlt(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg);
// Same, with CUDA events around every launch; ms[0..2] += device time of the diagonal-block,
// panel and trailing-update (DMMA) kernels; flops_syrk += algorithmic flops of the updates.
int launch_ldlt_profiled(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg,
                         double ms[3], double* flops_syrk, int* n_syrk);
int launch_ldlt_timeline(cudaStream_t st, const FactorPlan& fp, const double* src, double* dst, double* Dg,
                         double* out, int cap, int* nrec);
// Register-resident DMMA issue-rate probe: the FP64 tensor-pipe ceiling of this device.
int fp64_peak_probe(cudaStream_t st, double* tflops);
int read_phase_clocks(long long* out16);  // debug builds (-DIPMZ_PHASE_CLOCKS)
// C (rows x rows, lower triangle) += sign * PA PB^T with PA, PB rows x kdim and PB = PA diag(d)
// pre-scaled by its producer; the DMMA kernel shared by the trailing update of the factorization
// (sign -1, PA = the panel of L, PB = W = L D written by k_trsm_panel) and the condensed assembly
// M^T W M (sign +1, PA = MT, PB = MT diag(W) written by k_scale_cols).
void launch_syrk_ldl(cudaStream_t st, int nslots, const int* active, const double* Cin, double* Cout, int ldc,
                     size_t sC, const double* PA, int lda, size_t sA, const double* PB, int ldb, size_t sB, int rows,
                     int kdim, double sign);

// ---- batch_fused.cu: the whole IPM solve of one small QP inside one persistent CTA, one launch per batch ----
bool fused_batch_applicable(const View& v);  // AUGMENTED / NORMAL, LDL^T rows only, panel fits one CTA's shared memory
size_t fused_k_doubles(int N);               // per-problem doubles of K in the fused path's tile-major layout
int fused_batch_init();                      // per-device opt-in shared-memory size; returns cudaError_t
// every problem 0..count-1 from its current iterate to convergence; refine_fixed < 0 = refinement by each problem's mu
// ready != nullptr: streamed mode -- the grid is launched before the upload and a CTA waits until *ready > its ticket
// ticket: work_words ints = control words + bucket slots of the iteration queue (fused_work_words; fewer: problem-granular
// tickets, 8 ints are enough)
int fused_ctl_words(int max_iter);
size_t fused_work_words(int count, int max_iter);
int launch_ipm_batch(cudaStream_t st, const View& v, int count, int refine_fixed, int* ticket,
                     const int* ready = nullptr, int* abort_flag = nullptr, size_t work_words = 0);
int fused_read_clocks(unsigned long long* out16);  // debug builds (-DIPMZ_FUSED_CLOCKS)

// dst[m x ldd](lower) = -src block (lower), the Schur complement left by a partial elimination, as its own matrix
void launch_negate_block(cudaStream_t st, int nslots, const int* active, const double* src, int lds, size_t sS,
                         double* dst, int ldd, size_t sD, int m);
// dual-Schur normal equations, vector steps between the solves (see solver.cu dual_solve)
void launch_dual_vec(cudaStream_t st, const View& v, int nslots, int stage, const double* rvec, double* lam);

// ---- trsv.cu ----
int trsv_init();  // opt-in shared memory size of the streaming solves
void trsv_set_debug_log(long long* dev);  // debug: [2][nblk][8] timestamps of the next streaming solves
struct TrsvWork {
  int* flags;      // [nslots][nblk]
  int* ticket;     // [1]
  int epoch;       // bumped by every launch
  int cap_blocks;  // flags capacity per slot
};
// x <- L^-1 x (unit lower), then x <- L^-T D^-1 x; L strict-lower of K, in place on x.
void launch_ldlt_solve(cudaStream_t st, const FactorPlan& fp, const double* K, const double* Dg, double* x,
                       size_t sx, TrsvWork& w);

}  // namespace ipmz
