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
int gemm_tc_presplit(int transA, int transB, int M, int N, int K, const float* Ahi, const float* Alo, int lda,
                     const float* Bhi, const float* Blo, int ldb, float* C, int ldc, const float* bias, int act,
                     int accumulate, float* Chi, float* Clo, cudaStream_t st);
int tf32_split_launch(const float* x, float* hi, float* lo, size_t n, cudaStream_t st);
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

int ipavsr_gemm_tc_supported(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
                             int ldb, const float* C, int ldc) {
  return gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb, C, ldc) ? 1 : 0;
}

int ipavsr_tf32_split_rna(const float* x, float* hi, float* lo, uint64_t n, void* stream) {
  IPAVSR_CHECK_ARG(x && hi && lo, "null pointer");
  IPAVSR_CHECK_ARG(n % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) |
                                   reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
                   "buffers must be 16-byte aligned and a multiple of 4 floats long");
  return tf32_split_launch(x, hi, lo, n, reinterpret_cast<cudaStream_t>(stream));
}

int ipavsr_gemm_tf32x3_presplit(int transA, int transB, int M, int N, int K, const float* A_hi, const float* A_lo,
                                int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                                const float* bias, int act, int accumulate, float* C_hi, float* C_lo, void* stream) {
  IPAVSR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "negative size");
  IPAVSR_CHECK_ARG(A_hi && A_lo && B_hi && B_lo && C, "null pointer");
  IPAVSR_CHECK_ARG((C_hi == nullptr) == (C_lo == nullptr), "C_hi and C_lo go together");
  IPAVSR_CHECK_ARG(act >= IPAVSR_ACT_LINEAR && act <= IPAVSR_ACT_ELU, "unknown nonlinearity code");
  IPAVSR_CHECK_ARG(gemm_tc_supported(transA, transB, M, N, K, A_hi, lda, B_hi, ldb, C, ldc) &&
                       gemm_tc_supported(transA, transB, M, N, K, A_lo, lda, B_lo, ldb, C, ldc),
                   "shape/alignment not supported by the tensor-core path (see ipavsr_gemm_tc_supported)");
  if (M == 0 || N == 0) return IPAVSR_OK;
  return gemm_tc_presplit(transA, transB, M, N, K, A_hi, A_lo, lda, B_hi, B_lo, ldb, C, ldc, bias, act, accumulate,
                          C_hi, C_lo, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
