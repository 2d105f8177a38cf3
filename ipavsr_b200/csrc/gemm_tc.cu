// placeholder until the tcgen05 kernel lands (next commit): report "unsupported" so ipavsr_gemm uses FP32.
#include "common.cuh"
namespace ipavsr {
bool gemm_tc_supported(int, int, int, int, int, const float*, int, const float*, int, const float*, int) { return false; }
uint64_t gemm_tc_workspace_bytes(int, int, int, int, int, int) { return 0; }
int gemm_tc(int, int, int, int, int, int, const float*, int, const float*, int, float*, int, const float*, int, int,
            void*, uint64_t, cudaStream_t) {
  set_error("gemm_tc: not built");
  return IPAVSR_ERR_UNSUPPORTED;
}
}  // namespace ipavsr
