// ipm-zoo_b200/csrc/dataflow_tma.cu -- the TMA build of the dataflow LDL^T kernel (dataflow_kernel.cuh with
// DF_TMA = 1): UPD operand slices by cp.async.bulk.tensor.2d with 128-byte swizzle, full barriers completed by
// transaction bytes, one elected producer thread.  Selected with IPMZ_DF_TMA=1 (dataflow.cu, df_launch).
#define DF_TMA 1
#include "dataflow_kernel.cuh"

namespace ipmz {

int dataflow_tma_init() { return df_kernel_init(); }

int dataflow_tma_launch(cudaStream_t st, const void* args, int ctas, const double* W) {
  return df_kernel_launch(st, *static_cast<const DfArgs*>(args), ctas, W);
}

int dataflow_tma_launch_operands(cudaStream_t st, const void* args, int ctas, const double* A, int rowsA, int ldA,
                                 const double* B, int rowsB, int ldB) {
  return df_kernel_launch_operands(st, *static_cast<const DfArgs*>(args), ctas, A, rowsA, ldA, B, rowsB, ldB);
}

}  // namespace ipmz
