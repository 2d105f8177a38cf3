// Packed / length-sorted execution of a padded batch ("skip padded work").
//
// The reference pads every utterance to the dataset-wide T with zero frames (utils/datagen.py:129-139) and runs the
// DBNF encoder over all N*T rows (modelzoo/pretrained_encoder.py:4-9 on the (N*T, D) reshape): at the usual length
// distribution a third of those rows are zeros whose encoder output is the constant c = enc(0) (SURVEY A.2 "padding
// algebra").  The engine therefore runs the encoder on the VALID frames only, packed in length-sorted utterance order,
// plus ONE zero row that yields c, and expands to the padded layout behind the bottleneck.  These are the row movers of
// that scheme; all index tables are built by the host from the utterance lengths (engine._PackPlan):
//
//   ipavsr_gather_rows      dst[r, :] = idx[r] >= 0 ? src[idx[r], :] : fill (a row, or zeros)       byte rows
//       - pack:      padded (N*T, F) original order  -> packed (M+1, F) sorted order, last row zero
//       - unpack:    packed (M+1, F)                 -> padded (N*T, F) sorted order (padding rows = row M = c)
//       - permute / un-permute whole utterances (streams that bypass the encoder, masks, targets, the network output)
//     `src` may be PINNED HOST memory (device-accessible under UVA): the kernel then reads only the valid frames over
//     PCIe — the ragged upload of the end-to-end path — with four 16-byte loads in flight per lane.
//   ipavsr_colsum_masked    out[c] (+)= sum over the rows r with (rowmask[r] != 0) != invert of X[r, c]
//       - backward of the unpack: the gradient of the constant row is the sum of the gradients of all padding rows.
//
// HBM-bound (or PCIe-bound) copies: one warp per row, consecutive rows on consecutive warps.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace ipavsr {

template <int VB>   // bytes per access: 16, 4 or 1
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* __restrict__ src, long long src_pitch,
                                                          uint8_t* __restrict__ dst, long long dst_pitch, int row_bytes,
                                                          const int32_t* __restrict__ idx,
                                                          const uint8_t* __restrict__ fill, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wstride) {
    const int i = __ldg(idx + r);
    const uint8_t* s = i >= 0 ? src + (long long)i * src_pitch : fill;
    uint8_t* d = dst + r * dst_pitch;
    if (VB == 16) {
      const int n = row_bytes >> 4;
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      for (int c = lane; c < n; c += 128) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          v[k] = (s != nullptr && c + 32 * k < n) ? __ldcs(reinterpret_cast<const uint4*>(s) + c + 32 * k) : zero;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + 32 * k < n) reinterpret_cast<uint4*>(d)[c + 32 * k] = v[k];
      }
    } else if (VB == 4) {
      const int n = row_bytes >> 2;
      for (int c = lane; c < n; c += 32)
        reinterpret_cast<uint32_t*>(d)[c] = s != nullptr ? __ldcs(reinterpret_cast<const uint32_t*>(s) + c) : 0u;
    } else {
      for (int c = lane; c < row_bytes; c += 32) d[c] = s != nullptr ? s[c] : (uint8_t)0;
    }
  }
}

// 32 columns x 8 row-lanes per block; each block walks a strided set of row groups and adds its partial sums atomically
__global__ void __launch_bounds__(256) colsum_masked_kernel(const float* __restrict__ X, int ldx,
                                                            const uint8_t* __restrict__ rowmask, int invert,
                                                            float* __restrict__ out, int M, int N) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float acc = 0.f;
  for (int r = blockIdx.y * 8 + ry; r < M; r += gridDim.y * 8) {
    const bool take = (rowmask[r] != 0) != (invert != 0);
    if (take && c < N) acc += X[(size_t)r * ldx + c];
  }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cx];
    atomicAdd(out + c, t);
  }
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" int ipavsr_gather_rows(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes,
                                  int row_bytes, const int32_t* idx, const void* fill_row, int64_t rows, void* stream) {
  IPAVSR_CHECK_ARG(dst && idx, "null pointer");
  IPAVSR_CHECK_ARG(rows >= 0 && row_bytes >= 1 && src_pitch_bytes >= 0 && dst_pitch_bytes >= row_bytes, "bad sizes");
  if (rows == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long blocks = (rows + 7) / 8;
  // A device source is an HBM-bound copy: fill the machine.  A pinned-host source is PCIe-bound (~50 GB/s needs only a
  // few hundred KB in flight) and runs on a copy stream NEXT to the compute kernels.  Measured (tools/overlap_probe.py):
  // an SM that holds one of these CTAs does not accept a GEMM / LSTM CTA launched AFTER it (a 148-CTA upload launched
  // first delays the GEMM stream by its whole 1.4 ms; launched after the GEMMs it costs them 0.35 ms), and 16-32 CTAs
  // already saturate the link — so the grid is small (IPAVSR_HOST_GATHER_CTAS, default 32) and the engine issues the
  // upload of the NEXT step after the current step's kernels (function(...).prefetch(defer=True)).
  long long cap = (long long)sm_count() * 8;
  if (src != nullptr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
      static int host_ctas = -1;
      if (host_ctas < 0) {
        const char* e = getenv("IPAVSR_HOST_GATHER_CTAS");
        host_ctas = (e && atoi(e) > 0) ? atoi(e) : 32;
      }
      cap = host_ctas;
    } else {
      (void)cudaGetLastError();
    }
  }
  if (blocks > cap) blocks = cap;
  const uintptr_t all = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
                        reinterpret_cast<uintptr_t>(fill_row) | (uintptr_t)src_pitch_bytes | (uintptr_t)dst_pitch_bytes |
                        (uintptr_t)row_bytes;
  const uint8_t* s = static_cast<const uint8_t*>(src);
  uint8_t* d = static_cast<uint8_t*>(dst);
  const uint8_t* f = static_cast<const uint8_t*>(fill_row);
  const size_t dyn = 0;
  if ((all & 15) == 0)
    gather_rows_kernel<16><<<(int)blocks, 256, dyn, st>>>(s, src_pitch_bytes, d, dst_pitch_bytes, row_bytes, idx, f, rows);
  else if ((all & 3) == 0)
    gather_rows_kernel<4><<<(int)blocks, 256, dyn, st>>>(s, src_pitch_bytes, d, dst_pitch_bytes, row_bytes, idx, f, rows);
  else
    gather_rows_kernel<1><<<(int)blocks, 256, dyn, st>>>(s, src_pitch_bytes, d, dst_pitch_bytes, row_bytes, idx, f, rows);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

extern "C" int ipavsr_colsum_masked(const float* X, int ldx, const uint8_t* rowmask, int invert, float* out, int M, int N,
                                    int accumulate, void* stream) {
  IPAVSR_CHECK_ARG(X && rowmask && out, "null pointer");
  IPAVSR_CHECK_ARG(M >= 0 && N >= 1 && ldx >= N, "bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) IPAVSR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  if (M == 0) return IPAVSR_OK;
  int gy = (M + 63) / 64;
  const int cap = sm_count() * 4;
  if (gy > cap) gy = cap;
  dim3 grid((N + 31) / 32, gy);
  colsum_masked_kernel<<<grid, 256, 0, st>>>(X, ldx, rowmask, invert, out, M, N);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

// Ragged upload by the COPY ENGINES: utterance order[i] of a padded pinned-host stream (N, T, row) -> rows
// offsets[i] .. offsets[i+1]-1 of the packed device matrix, one asynchronous copy per utterance, then the zero row.
// Unlike the gather kernel reading host memory this takes no SM away from the compute kernels it runs next to
// (tools/overlap_probe.py: +0.02 ms on a 6.7 ms GEMM stream against +0.4 .. 1.4 ms), at the price of N driver calls.
extern "C" int ipavsr_upload_ragged(const void* host_src, int64_t utt_pitch_bytes, int64_t row_bytes, void* dst,
                                    int64_t dst_pitch_bytes, const int32_t* order_host, const int64_t* offsets_host, int N,
                                    void* stream) {
  IPAVSR_CHECK_ARG(host_src && dst && order_host && offsets_host, "null pointer");
  IPAVSR_CHECK_ARG(N >= 0 && row_bytes >= 1 && dst_pitch_bytes >= row_bytes && utt_pitch_bytes >= 0, "bad sizes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint8_t* s = static_cast<const uint8_t*>(host_src);
  uint8_t* d = static_cast<uint8_t*>(dst);
  // contiguous rows: ONE batched call for all utterances (cudaMemcpyBatchAsync, CUDA 12.8+) — 512 separate
  // cudaMemcpyAsync calls cost the host ~3 ms per stream; the per-utterance loop stays as the fallback (and for padded
  // destination rows, which need 2-D copies)
  static int batch_ok = -1;
  if (batch_ok < 0) {
    const char* e = getenv("IPAVSR_UPLOAD_BATCH");
    batch_ok = (e && e[0] == '0') ? 0 : 1;
  }
  if (dst_pitch_bytes == row_bytes && batch_ok && st != nullptr && N > 0) {
    std::vector<void*> dsts, srcs;
    std::vector<size_t> sizes;
    dsts.reserve(N); srcs.reserve(N); sizes.reserve(N);
    for (int i = 0; i < N; ++i) {
      const int64_t len = offsets_host[i + 1] - offsets_host[i];
      if (len <= 0) continue;
      dsts.push_back(d + offsets_host[i] * dst_pitch_bytes);
      srcs.push_back(const_cast<uint8_t*>(s + (int64_t)order_host[i] * utt_pitch_bytes));
      sizes.push_back((size_t)(len * row_bytes));
    }
    if (!dsts.empty()) {
      cudaMemcpyAttributes attr;
      memset(&attr, 0, sizeof(attr));
      attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
      size_t attr_idx = 0, fail = 0;
      cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr, &attr_idx, 1, &fail, st);
      if (e == cudaSuccess) {
        IPAVSR_CUDA(cudaMemsetAsync(d + offsets_host[N] * dst_pitch_bytes, 0, (size_t)dst_pitch_bytes, st));
        return IPAVSR_OK;
      }
      (void)cudaGetLastError();
      batch_ok = 0;                  // not supported by this driver / stream: per-utterance copies from now on
    }
  }
  for (int i = 0; i < N; ++i) {
    const int64_t len = offsets_host[i + 1] - offsets_host[i];
    if (len <= 0) continue;
    const uint8_t* src = s + (int64_t)order_host[i] * utt_pitch_bytes;
    uint8_t* to = d + offsets_host[i] * dst_pitch_bytes;
    if (dst_pitch_bytes == row_bytes)
      IPAVSR_CUDA(cudaMemcpyAsync(to, src, (size_t)(len * row_bytes), cudaMemcpyHostToDevice, st));
    else
      IPAVSR_CUDA(cudaMemcpy2DAsync(to, (size_t)dst_pitch_bytes, src, (size_t)row_bytes, (size_t)row_bytes, (size_t)len,
                                    cudaMemcpyHostToDevice, st));
  }
  IPAVSR_CUDA(cudaMemsetAsync(d + offsets_host[N] * dst_pitch_bytes, 0, (size_t)dst_pitch_bytes, st));
  return IPAVSR_OK;
}
