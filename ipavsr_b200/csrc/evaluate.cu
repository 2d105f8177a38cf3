// Evaluation on the device: per-utterance majority vote over the frame-level argmax, accuracy
// count and confusion matrix.
//
// Replaces the host loops of the runners' evaluate_model / evaluate_model2 (runners/2stream_dct.py:48-81: for every
// utterance, np.argmax over the classes of its first seq_len frames, a vote count per class, np.argmax of the votes;
// then confusion_matrix[target, prediction] += 1) so that the (N, T, C) probabilities never leave HBM.  Ties break like
// np.argmax: the lowest class index wins, both for the frame argmax and for the vote.  seq_len = sum(mask[i, :])
// (the reference's np.sum(mask_val, axis=-1)); sequence-level outputs (N, C) are the T = 1 case with no mask.
// One WARP per utterance: lanes take frames, votes go to a per-warp shared-memory histogram.
#include "common.cuh"

namespace ipavsr {

constexpr int EV_WARPS = 8;

__global__ void __launch_bounds__(EV_WARPS * 32) vote_eval_kernel(const float* __restrict__ probs, int ldp,
                                                                  const uint8_t* __restrict__ mask,
                                                                  const uint8_t* __restrict__ y, int N, int T, int C,
                                                                  int32_t* __restrict__ pred,
                                                                  int32_t* __restrict__ confusion,
                                                                  int32_t* __restrict__ correct) {
  extern __shared__ int ev_votes[];            // [EV_WARPS][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* votes = ev_votes + warp * C;
  for (int i = blockIdx.x * EV_WARPS + warp; i < N; i += gridDim.x * EV_WARPS) {
    for (int c = lane; c < C; c += 32) votes[c] = 0;
    int len = T;
    if (mask) {
      int s = 0;
      for (int t = lane; t < T; t += 32) s += mask[(size_t)i * T + t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      len = s;
    }
    __syncwarp();
    for (int t = lane; t < len; t += 32) {
      const float* row = probs + ((size_t)i * T + t) * ldp;
      float best = __ldg(row);
      int arg = 0;
      for (int c = 1; c < C; ++c) {
        const float v = __ldg(row + c);
        if (v > best) {                       // strict: the first maximum wins (np.argmax)
          best = v;
          arg = c;
        }
      }
      atomicAdd(&votes[arg], 1);
    }
    __syncwarp();
    // argmax of the votes, lowest class first on ties: reduce (count, -class) lexicographically
    int bc = -1, bi = 0;
    for (int c = lane; c < C; c += 32) {
      const int v = votes[c];
      if (v > bc) {
        bc = v;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oc > bc || (oc == bc && oi < bi)) {
        bc = oc;
        bi = oi;
      }
    }
    if (lane == 0) {
      if (pred) pred[i] = bi;
      if (y) {
        const int target = y[i];
        if (confusion && target < C) atomicAdd(&confusion[(size_t)target * C + bi], 1);
        if (correct && target == bi) atomicAdd(correct, 1);
      }
    }
    __syncwarp();
  }
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" int ipavsr_vote_eval(const float* probs, int ldp, const uint8_t* mask, const uint8_t* y, int N, int T, int C,
                                int32_t* pred, int32_t* confusion, int32_t* correct, void* stream) {
  IPAVSR_CHECK_ARG(probs && N >= 0 && T >= 1 && C >= 1 && ldp >= C, "bad arguments");
  IPAVSR_CHECK_ARG(C <= 4096, "at most 4096 classes");
  IPAVSR_CHECK_ARG((confusion == nullptr && correct == nullptr) || y != nullptr, "confusion / correct need the targets y");
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)EV_WARPS * C * sizeof(int);
  if (smem > 48 * 1024)
    IPAVSR_CUDA(cudaFuncSetAttribute(vote_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = ((long long)N + EV_WARPS - 1) / EV_WARPS;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  vote_eval_kernel<<<(int)blocks, EV_WARPS * 32, smem, st>>>(probs, ldp, mask, y, N, T, C, pred, confusion, correct);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}
