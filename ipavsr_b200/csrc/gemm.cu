// ipavsr_gemm: mode dispatch between the FP32 CUDA-core kernel (gemm_simt.cu) and the tcgen05 kernels (gemm_tc.cu).
#include "common.cuh"

namespace ipavsr {
int gemm_simt(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
              float* C, int ldc, const float* bias, int act, int accumulate, cudaStream_t st);
int gemm_tc(int mode, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
            float* C, int ldc, const float* bias, int act, int accumulate, void* ws, uint64_t ws_bytes,
            cudaStream_t st);
uint64_t gemm_tc_workspace_bytes(int mode, int transA, int transB, int M, int N, int K);
bool gemm_tc_supported(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       const float* C, int ldc);
}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

uint64_t ipavsr_gemm_workspace_bytes(int mode, int transA, int transB, int M, int N, int K) {
  if (mode == IPAVSR_GEMM_FP32) return 0;
  return gemm_tc_workspace_bytes(mode, transA, transB, M, N, K);
}

int ipavsr_gemm(int mode, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
                int ldb, float* C, int ldc, const float* bias, int act, int accumulate, void* workspace,
                uint64_t workspace_bytes, void* stream) {
  IPAVSR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "negative size");
  IPAVSR_CHECK_ARG(A && B && C, "null pointer");
  IPAVSR_CHECK_ARG(act >= IPAVSR_ACT_LINEAR && act <= IPAVSR_ACT_ELU, "unknown nonlinearity code");
  IPAVSR_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "leading dimension too small");
  if (M == 0 || N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mode == IPAVSR_GEMM_FP32)
    return gemm_simt(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, st);
  if (mode == IPAVSR_GEMM_TF32X3 || mode == IPAVSR_GEMM_TF32) {
    // shapes the TMA/UMMA tiling cannot take (tiny, unaligned) are computed by the exact FP32 kernel instead;
    // both are CUDA paths of this library and the FP32 one is at least as accurate as either tensor-core mode.
    if (!gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb, C, ldc))
      return gemm_simt(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, st);
    return gemm_tc(mode, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, workspace,
                   workspace_bytes, st);
  }
  set_error("ipavsr_gemm: unknown mode %d", mode);
  return IPAVSR_ERR_ARG;
}

}  // extern "C"
