// the rest of utils/preprocessing.py on the device (sm_100a).
//   compute_dct_features (:417-462)  -> dct_basis_kernel + dct_project_kernel (+ col_abs_sum / gather_cols for the
//                                        coefficient-selection methods)
//   reorder_data (:492-503)          -> reorder_kernel (per-frame d1 x d2 transpose through shared memory)
//   force_align (:607-660), multistream_force_align (:672-712) -> align_fill_kernel (row gather with a fill row)
// The reference's "2-D DCT" is scipy.fftpack.dct over the LAST axis of the (frames, D) matrix, i.e. a 1-D type-2
// orthonormal DCT of the flattened image (:427); the zigzag then walks that vector reshaped to the image shape (:431-433).
// Only the no_coeff zigzag positions are ever kept, so the kernel projects each frame onto those basis vectors only:
// 4*D bytes read and 2*D*K flops per frame (K/2 flop per byte = 15 at K=30: above the FP32-pipe balance point of this
// part, so dct_project_kernel is FFMA-bound, not HBM-bound).
#include <stdlib.h>
#include "common.cuh"

namespace ipavsr {

// dct_tc.cu: 0 = done, 1 = not a shape of the tensor-core kernel, < 0 = error
int dct_project_tc(const float* x, int ldx, const float* basis, int ldb, float* out, int ldo, int64_t frames, int D, int K,
                   cudaStream_t st);

// basis[d*ldb + k] = s(c_k) * cos(pi * (2d+1) * c_k / (2D)),  c_k = cols[k] (or k),  s(0) = sqrt(1/D), s(c>0) = sqrt(2/D)
// (scipy.fftpack.dct type 2, norm='ortho').  Evaluated in float64 with cospi (exact argument reduction), rounded once.
__global__ void dct_basis_kernel(float* __restrict__ basis, int ldb, const int32_t* __restrict__ cols, int D, int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)D * K) return;
  const int d = (int)(i / K), k = (int)(i % K);
  const int c = cols ? cols[k] : k;
  const double num = (double)(2 * (int64_t)d + 1) * (double)c;          // exact (< 2^53)
  const double v = cospi(num / (2.0 * (double)D)) * (c == 0 ? sqrt(1.0 / (double)D) : sqrt(2.0 / (double)D));
  basis[(int64_t)d * ldb + k] = (float)v;
}

// out[f, k] = sum_d x[f, d] * basis[d, k].  CTA tile: 128 frames x 32 coefficients, 256 threads, each thread 4 frames
// (fg, fg+32, fg+64, fg+96) x 4 coefficients (4cg .. 4cg+3); D is walked in chunks of 32 with the next chunk prefetched
// into registers while the current one is multiplied out of shared memory (2 LDS.128 per 16 FFMA, conflict-free:
// the four frame rows a warp touches are consecutive rows of a 36-float pitch).
constexpr int DCT_TF = 128, DCT_TK = 32, DCT_DK = 32, DCT_XP = DCT_DK + 4;

__global__ void __launch_bounds__(256) dct_project_kernel(const float* __restrict__ x, int ldx,
                                                          const float* __restrict__ basis, int ldb,
                                                          float* __restrict__ out, int ldo, int64_t frames, int D, int K) {
  __shared__ __align__(16) float xs[DCT_TF * DCT_XP];
  __shared__ __align__(16) float bs[DCT_DK * DCT_TK];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int kb = blockIdx.x;                               // coefficient block varies fastest: CTAs sharing an x tile are co-resident
  const int64_t f0 = (int64_t)blockIdx.y * DCT_TF;
  const int k0 = kb * DCT_TK;
  const int cg = tid & 7, fg = tid >> 3;

  float xr[16], br[4];
  auto fetch = [&](int d0) {
    const int d = d0 + lane;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int64_t f = f0 + wid + 8 * j;
      xr[j] = (f < frames && d < D) ? __ldg(x + f * ldx + d) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int dd = d0 + wid + 8 * j, k = k0 + lane;
      br[j] = (dd < D && k < K) ? __ldg(basis + (int64_t)dd * ldb + k) : 0.0f;
    }
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  fetch(0);
  for (int d0 = 0; d0 < D; d0 += DCT_DK) {
    __syncthreads();                                       // previous chunk fully consumed
#pragma unroll
    for (int j = 0; j < 16; ++j) xs[(wid + 8 * j) * DCT_XP + lane] = xr[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) bs[(wid + 8 * j) * DCT_TK + lane] = br[j];
    __syncthreads();
    if (d0 + DCT_DK < D) fetch(d0 + DCT_DK);
#pragma unroll
    for (int d4 = 0; d4 < DCT_DK / 4; ++d4) {
      float4 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(&xs[(fg + 32 * i) * DCT_XP + 4 * d4]);
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {
        const float4 bv = *reinterpret_cast<const float4*>(&bs[(4 * d4 + dd) * DCT_TK + 4 * cg]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xe = dd == 0 ? xv[i].x : dd == 1 ? xv[i].y : dd == 2 ? xv[i].z : xv[i].w;
          acc[i][0] = fmaf(xe, bv.x, acc[i][0]);
          acc[i][1] = fmaf(xe, bv.y, acc[i][1]);
          acc[i][2] = fmaf(xe, bv.z, acc[i][2]);
          acc[i][3] = fmaf(xe, bv.w, acc[i][3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t f = f0 + fg + 32 * i;
    if (f >= frames) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * cg + j;
      if (k < K) out[f * ldo + k] = acc[i][j];
    }
  }
}

// Aligned fast path of the same product (ldx, ldb, D multiples of 4 floats, 16-byte aligned bases): 128 threads per 128 x 32
// tile, 8 frames (fg + 16 i) x 4 coefficients per thread (32 independent FFMA chains, 12 LDS.128 per 128 FFMA), operand
// chunks brought in by 16-byte cp.async into a two-stage shared-memory ring (zero-filled past frames / D / K).
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(128, 4) dct_project_async_kernel(const float* __restrict__ x, int ldx,
                                                                   const float* __restrict__ basis, int ldb,
                                                                   float* __restrict__ out, int ldo, int64_t frames, int D,
                                                                   int K) {
  __shared__ __align__(16) float xs[2][DCT_TF * DCT_XP];
  __shared__ __align__(16) float bs[2][DCT_DK * DCT_TK];
  const int tid = threadIdx.x;
  const int kb = blockIdx.x;
  const int64_t f0 = (int64_t)blockIdx.y * DCT_TF;
  const int k0 = kb * DCT_TK;
  const int cg = tid & 7, fg = tid >> 3;                   // fg 0..15

  auto load_stage = [&](int st, int d0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {                          // 128 rows x 8 float4
      const int idx = tid + 128 * j, row = idx >> 3, c4 = idx & 7;
      const int64_t f = f0 + row;
      const int d = d0 + 4 * c4;
      const bool ok = f < frames && d < D;                 // D % 4 == 0: a float4 is inside or outside as a whole
      cp_async16_zfill(&xs[st][row * DCT_XP + 4 * c4], ok ? (const void*)(x + f * ldx + d) : (const void*)x, ok ? 16 : 0);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {                          // 32 rows x 8 float4
      const int idx = tid + 128 * j, row = idx >> 3, c4 = idx & 7;
      const int d = d0 + row, k = k0 + 4 * c4;
      int nb = (K - k) * 4;
      nb = (d < D && nb > 0) ? (nb > 16 ? 16 : nb) : 0;
      cp_async16_zfill(&bs[st][row * DCT_TK + 4 * c4], nb ? (const void*)(basis + (int64_t)d * ldb + k) : (const void*)basis, nb);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int nchunks = (D + DCT_DK - 1) / DCT_DK;
  load_stage(0, 0);
  for (int it = 0; it < nchunks; ++it) {
    const int st = it & 1;
    if (it + 1 < nchunks) {
      load_stage(st ^ 1, (it + 1) * DCT_DK);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* xq = xs[st];
    const float* bq = bs[st];
#pragma unroll
    for (int d4 = 0; d4 < DCT_DK / 4; ++d4) {
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(&xq[(fg + 16 * i) * DCT_XP + 4 * d4]);
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {
        const float4 bv = *reinterpret_cast<const float4*>(&bq[(4 * d4 + dd) * DCT_TK + 4 * cg]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xe = dd == 0 ? xv[i].x : dd == 1 ? xv[i].y : dd == 2 ? xv[i].z : xv[i].w;
          acc[i][0] = fmaf(xe, bv.x, acc[i][0]);
          acc[i][1] = fmaf(xe, bv.y, acc[i][1]);
          acc[i][2] = fmaf(xe, bv.z, acc[i][2]);
          acc[i][3] = fmaf(xe, bv.w, acc[i][3]);
        }
      }
    }
    __syncthreads();                                       // this stage is refilled by the next iteration's load
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t f = f0 + fg + 16 * i;
    if (f >= frames) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * cg + j;
      if (k < K) out[f * ldo + k] = acc[i][j];
    }
  }
}

// reorder_data: out[f, j] = x[f, src(j)] with the per-frame (d1, d2) transpose.  to_c = 1 ('f' -> 'c'):
// out[b*d2 + c] = x[b + d1*c];  to_c = 0 ('c' -> 'f'): out[b + d1*c] = x[b*d2 + c].  One frame per CTA iteration: the row is
// read coalesced into shared memory and written coalesced in the new order.
__global__ void __launch_bounds__(256) reorder_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy,
                                                      int64_t frames, int d1, int d2, int to_c, int R, int vec) {
  extern __shared__ __align__(16) float rsm[];
  const int D = d1 * d2;
  int* src_tab = reinterpret_cast<int*>(rsm);               // D ints: source position of output position j (built once)
  float* rows = rsm + ((D + 3) & ~3);                       // R frames
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    int src;
    if (to_c) {
      const int b = j / d2, c = j - b * d2;
      src = b + d1 * c;
    } else {
      const int c = j / d1, b = j - c * d1;
      src = b * d2 + c;
    }
    src_tab[j] = src;
  }
  const int64_t tiles = (frames + R - 1) / R;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t f0 = t * R;
    const int nf = (int)(frames - f0 < R ? frames - f0 : R);
    __syncthreads();                                        // table ready / previous tile written out
    if (vec) {
      const int D4 = D >> 2;
      for (int i = threadIdx.x; i < nf * D4; i += blockDim.x) {
        const int f = i / D4, c = i - f * D4;
        reinterpret_cast<float4*>(rows)[i] = __ldg(reinterpret_cast<const float4*>(x + (f0 + f) * ldx) + c);
      }
    } else {
      for (int i = threadIdx.x; i < nf * D; i += blockDim.x) {
        const int f = i / D, c = i - f * D;
        rows[i] = __ldg(x + (f0 + f) * ldx + c);
      }
    }
    __syncthreads();
    for (int f = 0; f < nf; ++f) {
      const float* rf = rows + f * D;
      float* yr = y + (f0 + f) * ldy;
      for (int j = threadIdx.x; j < D; j += blockDim.x) yr[j] = rf[src_tab[j]];
    }
  }
}

// force_align / multistream_force_align: output row r of utterance u (out_offsets[u] <= r < out_offsets[u+1]) is input row
// in_offsets[u] + j for j = r - out_offsets[u] < len_in(u), and the utterance's fill row beyond (the reference repeats the
// last frame of the shorter stream; fill_rows[u] is an absolute input row so that force_align's own indexing, which is
// relative to the OTHER stream's length (:652), can be reproduced).  One warp per output row.
template <bool VEC>
__global__ void __launch_bounds__(256) align_fill_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy,
                                                         const int64_t* __restrict__ in_off,
                                                         const int64_t* __restrict__ out_off,
                                                         const int64_t* __restrict__ fill_rows, int U, int D,
                                                         int64_t out_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < out_rows; r += nwarps) {
    int lo = 0, hi = U;                                    // largest u with out_off[u] <= r
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(out_off + mid) <= r) lo = mid; else hi = mid;
    }
    const int u = lo;
    const int64_t j = r - __ldg(out_off + u);
    const int64_t i0 = __ldg(in_off + u), len_in = __ldg(in_off + u + 1) - i0;
    const int64_t src = j < len_in ? i0 + j : (fill_rows ? __ldg(fill_rows + u) : i0 + len_in - 1);
    if (VEC) {
      const float4* s = reinterpret_cast<const float4*>(x + src * ldx);
      float4* d = reinterpret_cast<float4*>(y + r * ldy);
      for (int c = lane; c < D / 4; c += 32) d[c] = __ldg(s + c);
    } else {
      const float* s = x + src * ldx;
      float* d = y + r * ldy;
      for (int c = lane; c < D; c += 32) d[c] = __ldg(s + c);
    }
  }
}

// out[f, k] = x[f, idx[k]]  (the coefficient-selection methods of compute_dct_features keep the no_coeff best columns)
__global__ void __launch_bounds__(256) gather_cols_kernel(const float* __restrict__ x, int ldx,
                                                          const int32_t* __restrict__ idx, float* __restrict__ out,
                                                          int ldo, int64_t frames, int K) {
  const int64_t total = frames * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / K;
    const int k = (int)(i - f * K);
    out[f * ldo + k] = __ldg(x + f * ldx + __ldg(idx + k));
  }
}

// sums[c] += sum_f |x[f, c]|   (method 'energy', :455-458).  Row stripes per CTA, float partials, double atomics.
__global__ void __launch_bounds__(256) col_abs_sum_kernel(const float* __restrict__ x, int ldx, double* __restrict__ sums,
                                                          int64_t frames, int F, int rows_per_cta) {
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < frames ? r0 + rows_per_cta : frames;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < F; c += gridDim.x * blockDim.x) {
    double s = 0.0;
    float part = 0.0f;
    int n = 0;
    for (int64_t r = r0; r < r1; ++r) {
      part += fabsf(__ldg(x + r * ldx + c));
      if (++n == 64) { s += (double)part; part = 0.0f; n = 0; }
    }
    s += (double)part;
    atomicAdd(sums + c, s);
  }
}

}  // namespace ipavsr

using namespace ipavsr;
#define S(s) ((cudaStream_t)(s))

extern "C" {

int ipavsr_zigzag_indices(int rows, int cols, int32_t* order_host) {
  IPAVSR_CHECK_ARG(rows > 0 && cols > 0 && order_host, "rows, cols > 0 and a host output buffer are required");
  // the traversal of utils/preprocessing.py:280-338 (first move is to the right, then diagonally down-left, ...)
  const int size = rows * cols;
  int r = 0, c = 0;
  bool down = false;
  for (int i = 0; i < size; ++i) {
    order_host[i] = r * cols + c;
    if (r == 0) {
      if (c % 2) { down = true; ++r; --c; }
      else if (c == cols - 1) { down = true; ++r; }
      else ++c;
    } else if (c == 0) {
      if (r % 2) {
        if (r == rows - 1) { down = false; ++c; }
        else ++r;
      } else { down = false; --r; ++c; }
    } else if (!down) {
      if (c == cols - 1) { down = true; ++r; }
      else { --r; ++c; }
    } else {
      if (r == rows - 1) { down = false; ++c; }
      else { ++r; --c; }
    }
    if (i + 1 < size && (r < 0 || r >= rows || c < 0 || c >= cols)) {
      // the reference walks off the array here too (IndexError for single-row / single-column shapes)
      set_error("%s: the zigzag walk leaves a %d x %d array at step %d", __func__, rows, cols, i + 1);
      return IPAVSR_ERR_ARG;
    }
  }
  return IPAVSR_OK;
}

int ipavsr_dct_basis(float* basis, int ldb, const int32_t* cols, int D, int K, void* stream) {
  IPAVSR_CHECK_ARG(basis && D > 0 && K > 0 && ldb >= K, "basis, D, K > 0 and ldb >= K are required");
  const int64_t n = (int64_t)D * K;
  dct_basis_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(basis, ldb, cols, D, K);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_dct_project(const float* x, int ldx, const float* basis, int ldb, float* out, int ldo, int64_t frames, int D,
                       int K, void* stream) {
  IPAVSR_CHECK_ARG(x && basis && out && frames >= 0 && D > 0 && K > 0, "x, basis, out, D, K > 0 are required");
  IPAVSR_CHECK_ARG(ldx >= D && ldb >= K && ldo >= K, "leading dimensions are smaller than the rows");
  if (frames == 0) return IPAVSR_OK;
  const int64_t tiles = (frames + DCT_TF - 1) / DCT_TF;
  IPAVSR_CHECK_ARG(tiles <= 65535, "at most 65535 * 128 frames per call");
  // the usual case — a few zig-zag coefficients of long frames — goes to the tensor cores (dct_tc.cu); IPAVSR_DCT_TC=0
  // keeps the FFMA kernels
  static int use_tc = -1;
  if (use_tc < 0) {
    const char* e = getenv("IPAVSR_DCT_TC");
    use_tc = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_tc) {
    const int rc = dct_project_tc(x, ldx, basis, ldb, out, ldo, frames, D, K, S(stream));
    if (rc <= 0) return rc;           // done (0) or a real error (< 0); 1 = not its shape
  }
  dim3 grid((K + DCT_TK - 1) / DCT_TK, (unsigned)tiles);
  const bool aligned = D % 4 == 0 && ldx % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)basis & 15) == 0;
  if (aligned)
    dct_project_async_kernel<<<grid, 128, 0, S(stream)>>>(x, ldx, basis, ldb, out, ldo, frames, D, K);
  else
    dct_project_kernel<<<grid, 256, 0, S(stream)>>>(x, ldx, basis, ldb, out, ldo, frames, D, K);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_reorder(const float* x, int ldx, float* y, int ldy, int64_t frames, int d1, int d2, int to_c, void* stream) {
  IPAVSR_CHECK_ARG(x && y && x != y && frames >= 0 && d1 > 0 && d2 > 0, "x, y (distinct), d1, d2 > 0 are required");
  const int64_t D = (int64_t)d1 * d2;
  IPAVSR_CHECK_ARG(ldx >= D && ldy >= D, "leading dimensions are smaller than d1*d2");
  IPAVSR_CHECK_ARG(D * 8 <= 200 * 1024, "d1*d2 floats (+ the index table) must fit one CTA's shared memory (200 KB)");
  if (frames == 0) return IPAVSR_OK;
  const int64_t Dp = (D + 3) & ~(int64_t)3;
  int R = (int)((40 * 1024 - Dp * 4) / (D * 4));            // ~40 KB per CTA: 5 CTAs per SM
  if (R < 1) R = 1;
  if (R > 16) R = 16;
  const size_t smem = (size_t)(Dp + (int64_t)R * D) * 4;
  if (smem > 48 * 1024)
    IPAVSR_CUDA(cudaFuncSetAttribute(reorder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vec = D % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)x & 15) == 0;
  const int64_t tiles = (frames + R - 1) / R, want = (int64_t)sm_count() * 5;
  const unsigned grid = (unsigned)(tiles < want ? tiles : want);
  reorder_kernel<<<grid, 256, smem, S(stream)>>>(x, ldx, y, ldy, frames, d1, d2, to_c ? 1 : 0, R, vec);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_align_fill(const float* x, int ldx, float* y, int ldy, const int64_t* in_offsets, const int64_t* out_offsets,
                      const int64_t* fill_rows, int U, int D, int64_t out_rows, void* stream) {
  IPAVSR_CHECK_ARG(x && y && in_offsets && out_offsets && U > 0 && D > 0 && out_rows >= 0,
                   "x, y, in_offsets, out_offsets, U, D > 0 are required");
  IPAVSR_CHECK_ARG(ldx >= D && ldy >= D, "leading dimensions are smaller than D");
  if (out_rows == 0) return IPAVSR_OK;
  const bool vec = D % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0;
  const int64_t ctas = (out_rows + 7) / 8, cap = (int64_t)sm_count() * 16;
  const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
  if (vec)
    align_fill_kernel<true><<<grid, 256, 0, S(stream)>>>(x, ldx, y, ldy, in_offsets, out_offsets, fill_rows, U, D, out_rows);
  else
    align_fill_kernel<false><<<grid, 256, 0, S(stream)>>>(x, ldx, y, ldy, in_offsets, out_offsets, fill_rows, U, D, out_rows);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_gather_cols(const float* x, int ldx, const int32_t* idx, float* out, int ldo, int64_t frames, int K,
                       void* stream) {
  IPAVSR_CHECK_ARG(x && idx && out && frames >= 0 && K > 0 && ldo >= K, "x, idx, out, K > 0 and ldo >= K are required");
  if (frames == 0) return IPAVSR_OK;
  const int64_t total = frames * K, ctas = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  gather_cols_kernel<<<(unsigned)(ctas < cap ? ctas : cap), 256, 0, S(stream)>>>(x, ldx, idx, out, ldo, frames, K);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_col_abs_sum(const float* x, int ldx, double* sums, int64_t frames, int F, void* stream) {
  IPAVSR_CHECK_ARG(x && sums && frames >= 0 && F > 0 && ldx >= F, "x, sums, F > 0 and ldx >= F are required");
  IPAVSR_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)F, S(stream)));
  if (frames == 0) return IPAVSR_OK;
  const int gx = (F + 255) / 256;
  int64_t gy = ((int64_t)sm_count() * 8 + gx - 1) / gx;
  if (gy > frames) gy = frames;
  if (gy > 65535) gy = 65535;
  const int rows_per_cta = (int)((frames + gy - 1) / gy);
  gy = (frames + rows_per_cta - 1) / rows_per_cta;
  col_abs_sum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, S(stream)>>>(x, ldx, sums, frames, F, rows_per_cta);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // extern "C"
