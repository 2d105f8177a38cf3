// FP32 CUDA-core GEMM with fused bias + nonlinearity epilogue (mode IPAVSR_GEMM_FP32).
//
// Replaces Theano's `T.dot(x, W) + b` -> nonlinearity of Lasagne's DenseLayer
// (reference modelzoo/pretrained_encoder.py:4-9) and every dgrad/wgrad that T.grad derives from it, in exact
// float32 arithmetic (FFMA, fp32 accumulate).  This is the bit-faithful mode the tensor-core modes
// (gemm_tc.cu) are checked against; it is also what small / oddly shaped products use.
//
// Tiling: BMxBNx16 block tile (128x128 or 64x64), 256 threads, (BM/16)x(BN/16) register tile per thread,
// operands staged k-major in shared memory so the inner loop is two LDS.128 per 16/64 FFMA, register-staged
// double buffering (one __syncthreads per k-tile).  Optional split-K over gridDim.z with float atomics for
// the skinny wgrad shapes (K = N*T frames, tiny MxN).
#include "common.cuh"

namespace ipavsr {

constexpr int GK = 16;

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                        int ldc, const float* __restrict__ bias, int act,
                                                        int accumulate, int k_per_split, int a_vec, int b_vec,
                                                        int c_vec) {
  constexpr int TM = BM / 16, TN = BN / 16;       // 8 or 4
  constexpr int CM = TM / 4, CN = TN / 4;         // chunks of 4 (stride 64)
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[2][GK][BM + PAD];
  __shared__ __align__(16) float Bs[2][GK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);

  constexpr int A_F4 = BM * GK / 4 / 256;   // float4 loads per thread for A (2 or 1)
  constexpr int B_F4 = BN * GK / 4 / 256;
  float4 ra[A_F4], rb[B_F4];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int idx = tid + i * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!TA) {               // A[M,K]: float4 along k
        int row = idx >> 2, kq = (idx & 3) * 4;
        int gm = m0 + row, gk = k0 + kq;
        if (gm < M) {
          const float* p = A + (size_t)gm * lda + gk;
          if (a_vec && gk + 3 < k_end) v = *reinterpret_cast<const float4*>(p);
          else {
            if (gk + 0 < k_end) v.x = p[0];
            if (gk + 1 < k_end) v.y = p[1];
            if (gk + 2 < k_end) v.z = p[2];
            if (gk + 3 < k_end) v.w = p[3];
          }
        }
      } else {                 // A stored [K,M]: float4 along m
        int kk = idx / (BM / 4), mq = (idx % (BM / 4)) * 4;
        int gk = k0 + kk, gm = m0 + mq;
        if (gk < k_end) {
          const float* p = A + (size_t)gk * lda + gm;
          if (a_vec && gm + 3 < M) v = *reinterpret_cast<const float4*>(p);
          else {
            if (gm + 0 < M) v.x = p[0];
            if (gm + 1 < M) v.y = p[1];
            if (gm + 2 < M) v.z = p[2];
            if (gm + 3 < M) v.w = p[3];
          }
        }
      }
      ra[i] = v;
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int idx = tid + i * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!TB) {               // B[K,N]: float4 along n
        int kk = idx / (BN / 4), nq = (idx % (BN / 4)) * 4;
        int gk = k0 + kk, gn = n0 + nq;
        if (gk < k_end) {
          const float* p = B + (size_t)gk * ldb + gn;
          if (b_vec && gn + 3 < N) v = *reinterpret_cast<const float4*>(p);
          else {
            if (gn + 0 < N) v.x = p[0];
            if (gn + 1 < N) v.y = p[1];
            if (gn + 2 < N) v.z = p[2];
            if (gn + 3 < N) v.w = p[3];
          }
        }
      } else {                 // B stored [N,K]: float4 along k
        int row = idx >> 2, kq = (idx & 3) * 4;
        int gn = n0 + row, gk = k0 + kq;
        if (gn < N) {
          const float* p = B + (size_t)gn * ldb + gk;
          if (b_vec && gk + 3 < k_end) v = *reinterpret_cast<const float4*>(p);
          else {
            if (gk + 0 < k_end) v.x = p[0];
            if (gk + 1 < k_end) v.y = p[1];
            if (gk + 2 < k_end) v.z = p[2];
            if (gk + 3 < k_end) v.w = p[3];
          }
        }
      }
      rb[i] = v;
    }
  };
  auto store_ab = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int idx = tid + i * 256;
      if (!TA) {
        int row = idx >> 2, kq = (idx & 3) * 4;
        As[buf][kq + 0][row] = ra[i].x;
        As[buf][kq + 1][row] = ra[i].y;
        As[buf][kq + 2][row] = ra[i].z;
        As[buf][kq + 3][row] = ra[i].w;
      } else {
        int kk = idx / (BM / 4), mq = (idx % (BM / 4)) * 4;
        *reinterpret_cast<float4*>(&As[buf][kk][mq]) = ra[i];
      }
    }
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int idx = tid + i * 256;
      if (!TB) {
        int kk = idx / (BN / 4), nq = (idx % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = rb[i];
      } else {
        int row = idx >> 2, kq = (idx & 3) * 4;
        Bs[buf][kq + 0][row] = rb[i].x;
        Bs[buf][kq + 1][row] = rb[i].y;
        Bs[buf][kq + 2][row] = rb[i].z;
        Bs[buf][kq + 3][row] = rb[i].w;
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (k_begin < k_end) {
    load_a(k_begin);
    load_b(k_begin);
    store_ab(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
    const bool has_next = (k0 + GK) < k_end;
    if (has_next) {
      load_a(k0 + GK);
      load_b(k0 + GK);
    }
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int c = 0; c < CM; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][c * 64 + ty * 4]);
        a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][c * 64 + tx * 4]);
        b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) store_ab(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int ci = 0; ci < CM; ++ci)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int gm = m0 + ci * 64 + ty * 4 + i;
      if (gm >= M) continue;
#pragma unroll
      for (int cj = 0; cj < CN; ++cj) {
        int gn = n0 + cj * 64 + tx * 4;
        if (gn >= N) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[ci * 4 + i][cj * 4 + j];
        float* cp = C + (size_t)gm * ldc + gn;
        if (split) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gn + j < N) {
              float add = v[j];
              if (bias != nullptr && blockIdx.z == 0) add += bias[gn + j];
              atomicAdd(cp + j, add);
            }
        } else {
          if (c_vec && gn + 3 < N) {
            float4 o;
            if (accumulate) {
              float4 old = *reinterpret_cast<const float4*>(cp);
              v[0] += old.x; v[1] += old.y; v[2] += old.z; v[3] += old.w;
            }
            if (bias != nullptr) {
              float4 bb = *reinterpret_cast<const float4*>(bias + gn);
              v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
            }
            o.x = act_fwd(v[0], act); o.y = act_fwd(v[1], act); o.z = act_fwd(v[2], act); o.w = act_fwd(v[3], act);
            *reinterpret_cast<float4*>(cp) = o;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (gn + j < N) {
                float o = v[j];
                if (accumulate) o += cp[j];
                if (bias != nullptr) o += bias[gn + j];
                cp[j] = act_fwd(o, act);
              }
          }
        }
      }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BM, int BN>
static int launch_simt(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       float* C, int ldc, const float* bias, int act, int accumulate, cudaStream_t st) {
  int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  int tiles = tiles_m * tiles_n;
  int sms = sm_count();
  int splitk = 1;
  if (act == IPAVSR_ACT_LINEAR && tiles * 2 <= sms && K >= 1024) {
    splitk = (sms + tiles - 1) / tiles;
    int maxs = K / 256;
    if (splitk > maxs) splitk = maxs;
    if (splitk > 64) splitk = 64;
    if (splitk < 1) splitk = 1;
  }
  int kps = ((K + splitk - 1) / splitk + GK - 1) / GK * GK;
  splitk = (K + kps - 1) / kps;
  if (splitk > 1 && !accumulate) {
    IPAVSR_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st));
  }
  int a_vec = (lda % 4 == 0) && aligned16(A);
  int b_vec = (ldb % 4 == 0) && aligned16(B);
  int c_vec = (ldc % 4 == 0) && aligned16(C) && (bias == nullptr || aligned16(bias));
  dim3 grid(tiles_n, tiles_m, splitk);
#define IPAVSR_SIMT_LAUNCH(TA, TB)                                                                              \
  gemm_simt_kernel<BM, BN, TA, TB><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, \
                                                         kps, a_vec, b_vec, c_vec)
  if (!transA && !transB) IPAVSR_SIMT_LAUNCH(false, false);
  else if (!transA && transB) IPAVSR_SIMT_LAUNCH(false, true);
  else if (transA && !transB) IPAVSR_SIMT_LAUNCH(true, false);
  else IPAVSR_SIMT_LAUNCH(true, true);
#undef IPAVSR_SIMT_LAUNCH
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int gemm_simt(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
              float* C, int ldc, const float* bias, int act, int accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0) return IPAVSR_OK;
  // small-output products waste most of a 128x128 tile: use 64x64 there
  long tiles128 = (long)((M + 127) / 128) * ((N + 127) / 128);
  if (N <= 64 || M <= 64 || tiles128 < sm_count())
    return launch_simt<64, 64>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, st);
  return launch_simt<128, 128>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, st);
}

// out[k] (+)= sum_j W[k * ldw + j] * s[j]  (k < rows, j < cols): one warp per row of W.  The d(hid_init) product of the
// LSTM backward after the sum over the utterances (lstm*.cu): 1 x 4H by 4H x H.
__global__ void __launch_bounds__(256) gemv_rows_kernel(const float* __restrict__ W, int ldw, const float* __restrict__ s,
                                                        float* __restrict__ out, int rows, int cols, int accumulate) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* w = W + (size_t)warp * ldw;
  float acc = 0.f;
  for (int j = lane; j < cols; j += 32) acc = fmaf(__ldg(w + j), __ldg(s + j), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[warp] = accumulate ? out[warp] + acc : acc;
}

int gemv_rows(const float* W, int ldw, const float* s, float* out, int rows, int cols, int accumulate, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return IPAVSR_OK;
  gemv_rows_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(W, ldw, s, out, rows, cols, accumulate);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // namespace ipavsr
