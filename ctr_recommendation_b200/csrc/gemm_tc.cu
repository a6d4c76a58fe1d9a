// tcgen05 / TMA / TMEM GEMM back end (FBN_PREC_TF32X3, FBN_PREC_BF16).  Placeholder until the
// tensor-core kernels land: reports an error instead of silently using another path.
#include "common.cuh"
#include "gemm.h"

namespace fbn {

bool gemm_tc_supported(const GemmArgs&, int) { return false; }

int gemm_tc(const GemmArgs&, int precision, void*, size_t, cudaStream_t) {
  set_error("precision mode %d (tcgen05) is not available in this build", precision);
  return FBN_ERR_ARG;
}

}  // namespace fbn
