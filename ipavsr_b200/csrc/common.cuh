// Shared helpers for the ipavsr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/ipavsr_b200.h"

namespace ipavsr {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define IPAVSR_CHECK_ARG(cond, msg)                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      ipavsr::set_error("%s: %s", __func__, msg);                     \
      return IPAVSR_ERR_ARG;                                          \
    }                                                                 \
  } while (0)

#define IPAVSR_CUDA(call)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ipavsr::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__)); \
      return IPAVSR_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

#define IPAVSR_LAUNCH_CHECK()                                                          \
  do {                                                                                 \
    ipavsr::count_launch();                                                            \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      ipavsr::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return IPAVSR_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

int sm_count();

// fp16 operand pairs of the three-product GEMM mode (f16split.cu): x 2^e = hi + lo / F16_LO_SCALE.  With the scale 1 the
// cross products lo*hi, hi*lo are on the scale of hi*hi, so one TMEM accumulator takes all three (gemm_f16p.cu) and a
// 256-wide tile can be double-buffered in the 512 TMEM columns.  The price is dynamic range: lo lives in the fp16
// subnormals for elements below 2^-18 of the tensor's maximum, i.e. the pair holds x to max(2^-22 |x|, 2^-39 max|x|).
constexpr float F16_LO_SCALE = 1.0f;
constexpr float F16_LO_INV = 1.0f / F16_LO_SCALE;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float act_fwd(float z, int act) {
  switch (act) {
    case IPAVSR_ACT_SIGMOID: return sigmoidf_(z);
    case IPAVSR_ACT_RECTIFY: return fmaxf(z, 0.0f);
    case IPAVSR_ACT_TANH: return tanhf(z);
    case IPAVSR_ACT_LEAKY: return z > 0.0f ? z : 0.01f * z;
    case IPAVSR_ACT_VERY_LEAKY: return z > 0.0f ? z : z * (1.0f / 3.0f);
    case IPAVSR_ACT_SOFTPLUS: return z > 20.0f ? z : log1pf(expf(z));
    case IPAVSR_ACT_ELU: return z > 0.0f ? z : expm1f(z);
    default: return z;
  }
}

// derivative expressed through the OUTPUT y (so the pre-activation never has to be stored).
// rectify/leaky at exactly 0 take the left slope (Theano's sub-gradient there is the mean of both slopes;
// this only differs for an input that is exactly 0 — DESIGN.md "known deviations").
__device__ __forceinline__ float act_grad_from_y(float y, int act) {
  switch (act) {
    case IPAVSR_ACT_SIGMOID: return y * (1.0f - y);
    case IPAVSR_ACT_RECTIFY: return y > 0.0f ? 1.0f : 0.0f;
    case IPAVSR_ACT_TANH: return 1.0f - y * y;
    case IPAVSR_ACT_LEAKY: return y > 0.0f ? 1.0f : 0.01f;
    case IPAVSR_ACT_VERY_LEAKY: return y > 0.0f ? 1.0f : (1.0f / 3.0f);
    case IPAVSR_ACT_SOFTPLUS: return 1.0f - expf(-y);   // sigmoid(z) with y = log(1+e^z)
    case IPAVSR_ACT_ELU: return y > 0.0f ? 1.0f : y + 1.0f;
    default: return 1.0f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ipavsr
