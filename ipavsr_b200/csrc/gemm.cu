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
int gemm_tc_f16x3(int transA, int transB, int M, int N, int K, const uint16_t* Ahi, const uint16_t* Alo, int lda,
                  const int32_t* expA, const uint16_t* Bhi, const uint16_t* Blo, int ldb, const int32_t* expB, float* C,
                  int ldc, const float* bias, int act, int accumulate, float* amax, uint16_t* C16hi, uint16_t* C16lo,
                  int c16_exp, cudaStream_t st);
bool gemm_tc_f16_supported(int M, int N, int K, const void* A, int lda, const void* B, int ldb);
int f16_split_launch(const float* x, int ldx, int rows, int cols, uint16_t* hi, uint16_t* lo, int ldo, float* amax,
                     int32_t* exp_out, int amax_ready, cudaStream_t st);

static inline size_t up8(size_t v) { return (v + 7) / 8 * 8; }
// workspace of the convenience (unsplit) fp16 mode: [amaxA, amaxB, expA, expB | pad to 64 B] Ahi Alo Bhi Blo
static size_t f16_ws_bytes(int transA, int transB, int M, int N, int K) {
  const size_t a = (size_t)(transA ? K : M) * up8(transA ? M : K), b = (size_t)(transB ? N : K) * up8(transB ? K : N);
  return 64 + 2 * (up8(a) + up8(b)) * sizeof(uint16_t);
}
}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

uint64_t ipavsr_gemm_workspace_bytes(int mode, int transA, int transB, int M, int N, int K) {
  if (mode == IPAVSR_GEMM_FP32) return 0;
  if (mode == IPAVSR_GEMM_F16X3) return f16_ws_bytes(transA, transB, M, N, K);
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
  if (mode == IPAVSR_GEMM_F16X3) {
    const int ra = transA ? K : M, ca = transA ? M : K, rb = transB ? N : K, cb = transB ? K : N;
    const int lda16 = (int)up8(ca), ldb16 = (int)up8(cb);
    if ((double)M * N * K < 4.0e6 || K < 16 || N < 8)
      return gemm_simt(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, st);
    IPAVSR_CHECK_ARG(workspace != nullptr && workspace_bytes >= f16_ws_bytes(transA, transB, M, N, K) &&
                         (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                     "workspace too small or not 16-byte aligned (ipavsr_gemm_workspace_bytes)");
    float* amax = reinterpret_cast<float*>(workspace);
    int32_t* exps = reinterpret_cast<int32_t*>(workspace) + 2;
    uint16_t* ah = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(workspace) + 64);
    uint16_t* al = ah + up8((size_t)ra * lda16);
    uint16_t* bh = al + up8((size_t)ra * lda16);
    uint16_t* bl = bh + up8((size_t)rb * ldb16);
    int rc;
    if ((rc = f16_split_launch(A, lda, ra, ca, ah, al, lda16, amax, exps, 0, st))) return rc;
    if ((rc = f16_split_launch(B, ldb, rb, cb, bh, bl, ldb16, amax + 1, exps + 1, 0, st))) return rc;
    return gemm_tc_f16x3(transA, transB, M, N, K, ah, al, lda16, exps, bh, bl, ldb16, exps + 1, C, ldc, bias, act,
                         accumulate, nullptr, nullptr, nullptr, 0, st);
  }
  set_error("ipavsr_gemm: unknown mode %d", mode);
  return IPAVSR_ERR_ARG;
}

int ipavsr_gemm_f16_supported(int M, int N, int K, const void* A, int lda, const void* B, int ldb) {
  return gemm_tc_f16_supported(M, N, K, A, lda, B, ldb) ? 1 : 0;
}

int ipavsr_gemm_f16x3(int transA, int transB, int M, int N, int K, const uint16_t* A_hi, const uint16_t* A_lo, int lda,
                      const int32_t* expA, const uint16_t* B_hi, const uint16_t* B_lo, int ldb, const int32_t* expB,
                      float* C, int ldc, const float* bias, int act, int accumulate, float* amax_out, uint16_t* C_hi,
                      uint16_t* C_lo, int c_exp, void* stream) {
  IPAVSR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "negative size");
  IPAVSR_CHECK_ARG(A_hi && A_lo && B_hi && B_lo && C && expA && expB, "null pointer");
  IPAVSR_CHECK_ARG((C_hi == nullptr) == (C_lo == nullptr) && c_exp >= -126 && c_exp <= 126, "bad C_hi / C_lo / c_exp");
  IPAVSR_CHECK_ARG(C_hi == nullptr || (ldc % 4 == 0 && ((reinterpret_cast<uintptr_t>(C_hi) | reinterpret_cast<uintptr_t>(C_lo)) & 7) == 0),
                   "C_hi / C_lo need 8-byte alignment and ldc % 4 == 0");
  IPAVSR_CHECK_ARG(act >= IPAVSR_ACT_LINEAR && act <= IPAVSR_ACT_ELU, "unknown nonlinearity code");
  IPAVSR_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "leading dimension too small");
  IPAVSR_CHECK_ARG(gemm_tc_f16_supported(M, N, K, A_hi, lda, B_hi, ldb) && gemm_tc_f16_supported(M, N, K, A_lo, lda, B_lo, ldb),
                   "shape/alignment not supported by the fp16 tensor-core path (see ipavsr_gemm_f16_supported)");
  if (M == 0 || N == 0) return IPAVSR_OK;
  return gemm_tc_f16x3(transA, transB, M, N, K, A_hi, A_lo, lda, expA, B_hi, B_lo, ldb, expB, C, ldc, bias, act,
                       accumulate, amax_out, C_hi, C_lo, c_exp, reinterpret_cast<cudaStream_t>(stream));
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
