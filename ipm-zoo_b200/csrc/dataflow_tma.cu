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

}  // namespace ipmz
