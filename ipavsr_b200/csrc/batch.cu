// Device-side batch assembly: packed variable-length storage -> zero-padded (N, T, F) batch + mask.
//
// Replaces the per-step host loops of utils/datagen.py:92-153 (gen_lstm_batch_random: one np.concatenate per utterance
// per stream, then an H2D copy of the padded batch) and :219-229 (gen_seq_batch_from_idx) with one gather over a
// dataset that stays resident in HBM:
//     X[i, t, :] = data[integral[idx[i]] + t, :]  for t < seqlen[idx[i]],  0 beyond;   mask[i, t] = t < seqlen[idx[i]];
//     y_batch[i] = y[integral[idx[i]]]  (the label of the utterance's first frame, datagen.py:138).
// HBM-bound copy: one WARP per output row (128-bit accesses when F, the pitches and the bases allow), rows of one
// utterance are consecutive so a CTA's 8 warps read and write consecutive lines.  Algorithmic traffic per utterance:
// 4 F (len + T) bytes (+ T mask bytes).
#include "common.cuh"

namespace ipavsr {

template <bool VEC>
__global__ void __launch_bounds__(256) batch_gather_kernel(const float* __restrict__ data, int ldd,
                                                           const int64_t* __restrict__ integral,
                                                           const int32_t* __restrict__ seqlen,
                                                           const int32_t* __restrict__ idxs, const uint8_t* __restrict__ y,
                                                           float* __restrict__ X, int ldx, uint8_t* __restrict__ mask,
                                                           uint8_t* __restrict__ yb, int N, int T, int F) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)N * T;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wstride) {
    const int i = (int)(r / T), t = (int)(r - (long long)i * T);
    const int u = __ldg(idxs + i);
    const int len = __ldg(seqlen + u);
    const int64_t start = __ldg(integral + u);
    const bool live = t < len;
    float* dst = X + r * ldx;
    const float* src = data + (start + t) * ldd;
    if (VEC) {
      const int F4 = F >> 2;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      // four independent 16-byte loads in flight per lane before the first store
      for (int c = lane; c < F4; c += 128) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          v[k] = (live && c + 32 * k < F4) ? __ldcs(reinterpret_cast<const float4*>(src) + c + 32 * k) : zero;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + 32 * k < F4) __stcs(reinterpret_cast<float4*>(dst) + c + 32 * k, v[k]);
      }
    } else {
      for (int c = lane; c < F; c += 32) __stcs(dst + c, live ? __ldcs(src + c) : 0.f);
    }
    if (lane == 0) {
      if (mask) mask[r] = live ? 1 : 0;
      if (yb && y && t == 0) yb[i] = y[start];
    }
  }
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" int ipavsr_batch_gather(const float* data, int ldd, const int64_t* integral_lens, const int32_t* seqlens,
                                   const int32_t* idxs, const uint8_t* y, float* X, int ldx, uint8_t* mask,
                                   uint8_t* y_batch, int N, int T, int F, void* stream) {
  IPAVSR_CHECK_ARG(data && integral_lens && seqlens && idxs && X, "null pointer");
  IPAVSR_CHECK_ARG(N >= 0 && T >= 1 && F >= 1 && ldd >= F && ldx >= F, "bad sizes");
  IPAVSR_CHECK_ARG((y_batch == nullptr) || (y != nullptr), "y_batch needs the per-frame labels y");
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)N * T;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const bool vec = F % 4 == 0 && ldd % 4 == 0 && ldx % 4 == 0 && ((reinterpret_cast<uintptr_t>(data) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (vec)
    batch_gather_kernel<true><<<(int)blocks, 256, 0, st>>>(data, ldd, integral_lens, seqlens, idxs, y, X, ldx, mask,
                                                          y_batch, N, T, F);
  else
    batch_gather_kernel<false><<<(int)blocks, 256, 0, st>>>(data, ldd, integral_lens, seqlens, idxs, y, X, ldx, mask,
                                                           y_batch, N, T, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}
