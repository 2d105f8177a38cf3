// the two streaming kernels auto-encoder fine-tuning adds to the encoder path
// (avletters/trimodal.py:41-89 load_dbn -> nolearn NeuralNet(objective_loss_function=squared_error, objective_l2=0.005)):
//   squared_error_kernel : loss_sum += sum (pred - target)^2,  dpred = 2 * grad_scale * (pred - target)
//   l2_penalty_kernel    : g += 2 c p,  loss_sum += loss_scale * c * p^2  with a per-tensor coefficient c over the flat arena
// Both are HBM-bound: 12 B per element (read pred, target; write dpred) and 12 B per parameter (read p, g; write g).
#include "common.cuh"

namespace ipavsr {

__device__ __forceinline__ void block_add(float v, float* out) {
  __shared__ double part[32];
  double d = (double)v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) part[wid] = d;
  __syncthreads();
  if (wid == 0) {
    d = lane < (int)(blockDim.x >> 5) ? part[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0 && d != 0.0) atomicAdd(out, (float)d);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) squared_error_kernel(const float* __restrict__ pred, int ldp,
                                                            const float* __restrict__ target, int ldt,
                                                            float* __restrict__ loss_sum, float* __restrict__ dpred,
                                                            int lddp, int64_t M, int F, float gs2) {
  float acc = 0.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    const int F4 = F >> 2;
    const int64_t total = M * F4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const int64_t r = i / F4;
      const int c = (int)(i - r * F4);
      const float4 p = __ldg(reinterpret_cast<const float4*>(pred + r * ldp) + c);
      const float4 t = __ldg(reinterpret_cast<const float4*>(target + r * ldt) + c);
      const float4 d = make_float4(p.x - t.x, p.y - t.y, p.z - t.z, p.w - t.w);
      acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
      if (dpred) reinterpret_cast<float4*>(dpred + r * lddp)[c] = make_float4(gs2 * d.x, gs2 * d.y, gs2 * d.z, gs2 * d.w);
    }
  } else {
    const int64_t total = M * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const int64_t r = i / F;
      const int c = (int)(i - r * F);
      const float d = __ldg(pred + r * ldp + c) - __ldg(target + r * ldt + c);
      acc += d * d;
      if (dpred) dpred[r * lddp + c] = gs2 * d;
    }
  }
  block_add(acc, loss_sum);
}

// seg_id maps 256-float blocks of the arena to parameter tensors (the table of ipavsr_optim_step); a block is inside one tensor
__global__ void __launch_bounds__(256) l2_penalty_kernel(const float* __restrict__ p, float* __restrict__ g, uint64_t n,
                                                         const float* __restrict__ seg_coef,
                                                         const int32_t* __restrict__ seg_id, float* __restrict__ loss_sum,
                                                         float loss_scale) {
  float acc = 0.0f;
  const uint64_t nblk = (n + 255) / 256;
  for (uint64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
    const float c = __ldg(seg_coef + __ldg(seg_id + b));
    if (c == 0.0f) continue;                               // block-uniform: biases, initial states, BN statistics
    const uint64_t i = b * 256 + threadIdx.x;
    if (i < n) {
      const float v = p[i];
      if (g) g[i] += 2.0f * c * v;
      acc += c * v * v;
    }
  }
  block_add(acc * loss_scale, loss_sum);
}

}  // namespace ipavsr

using namespace ipavsr;
#define S(s) ((cudaStream_t)(s))

extern "C" {

int ipavsr_squared_error(const float* pred, int ldp, const float* target, int ldt, float* loss_sum, float* dpred, int lddp,
                         int64_t M, int F, float grad_scale, void* stream) {
  IPAVSR_CHECK_ARG(pred && target && loss_sum && M >= 0 && F > 0, "pred, target, loss_sum and F > 0 are required");
  IPAVSR_CHECK_ARG(ldp >= F && ldt >= F && (!dpred || lddp >= F), "leading dimensions are smaller than F");
  if (M == 0) return IPAVSR_OK;
  const bool vec = F % 4 == 0 && ldp % 4 == 0 && ldt % 4 == 0 && (!dpred || lddp % 4 == 0) && ((uintptr_t)pred & 15) == 0 &&
                   ((uintptr_t)target & 15) == 0 && ((uintptr_t)dpred & 15) == 0;
  const int64_t work = vec ? M * (F / 4) : M * F, ctas = (work + 255) / 256, cap = (int64_t)sm_count() * 8;
  const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
  if (vec)
    squared_error_kernel<true><<<grid, 256, 0, S(stream)>>>(pred, ldp, target, ldt, loss_sum, dpred, lddp, M, F, 2.0f * grad_scale);
  else
    squared_error_kernel<false><<<grid, 256, 0, S(stream)>>>(pred, ldp, target, ldt, loss_sum, dpred, lddp, M, F, 2.0f * grad_scale);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_l2_penalty(const float* p, float* g, uint64_t n, const float* seg_coef, const int32_t* seg_id, float* loss_sum,
                      float loss_scale, void* stream) {
  IPAVSR_CHECK_ARG(p && seg_coef && seg_id && loss_sum, "p, seg_coef, seg_id and loss_sum are required");
  if (n == 0) return IPAVSR_OK;
  const uint64_t nblk = (n + 255) / 256, cap = (uint64_t)sm_count() * 16;
  l2_penalty_kernel<<<(unsigned)(nblk < cap ? nblk : cap), 256, 0, S(stream)>>>(p, g, n, seg_coef, seg_id, loss_sum, loss_scale);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // extern "C"
