// Memory-bound kernels of the hot path: DenseLayer backward prep, column sums, fusion (sum / adasum / concat),
// slice, dropout, BatchNorm, softmax + losses, fused optimiser step, TF32 split.  All are coalesced streaming
// kernels judged by HBM GB/s; grids are sized from the SM count.
#include <cuda_fp16.h>
#include "common.cuh"

namespace ipavsr {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------------
// dZ = dY * act'(Y), db += colsum(dZ).  Block = 32 columns x 8 row-lanes, ROWS rows per block.
// ------------------------------------------------------------------------------------------------------
constexpr int PREP_ROWS = 128;

template <bool WITH_ACT>
__global__ void __launch_bounds__(256) colsum_prep_kernel(const float* __restrict__ dY, int lddy,
                                                          const float* __restrict__ Y, int ldy,
                                                          float* __restrict__ dZ, int lddz, float* __restrict__ db,
                                                          int M, int N, int act, float* __restrict__ amax) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * PREP_ROWS;
  const int r1 = min(M, r0 + PREP_ROWS);
  float s = 0.f, mx = 0.f;
  if (c < N) {
    for (int r = r0 + rl; r < r1; r += 8) {
      float g = dY[(size_t)r * lddy + c];
      if (WITH_ACT) {
        g *= act_grad_from_y(Y[(size_t)r * ldy + c], act);
        dZ[(size_t)r * lddz + c] = g;
      }
      s += g;
      mx = fmaxf(mx, fabsf(g));
    }
  }
  if (amax != nullptr) {
    mx = warp_max(mx);
    if (lane == 0 && mx > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(mx));
  }
  if (db == nullptr) return;
  red[rl][lane] = s;
  __syncthreads();
  if (rl == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    atomicAdd(db + c, t);
  }
}

// Vectorised form (N % 4 == 0, 16-byte aligned rows): a block owns a 128-column x PREP_ROWS-row tile, a thread 4 adjacent
// columns of every 8th row, so a warp streams 512 contiguous bytes per row; column sums are combined in shared memory
// and leave the block as one atomicAdd per column.
template <bool WITH_ACT>
__global__ void __launch_bounds__(256) colsum_prep_vec_kernel(const float* __restrict__ dY, int lddy,
                                                              const float* __restrict__ Y, int ldy,
                                                              float* __restrict__ dZ, int lddz, float* __restrict__ db,
                                                              int M, int N, int act, float* __restrict__ amax) {
  __shared__ float4 red[8][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * PREP_ROWS;
  const int r1 = min(M, r0 + PREP_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  float mx = 0.f;
  if (c < N) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 8) {
      float4 g = __ldcs(reinterpret_cast<const float4*>(dY + (size_t)r * lddy + c));
      if (WITH_ACT) {
        const float4 y = __ldcs(reinterpret_cast<const float4*>(Y + (size_t)r * ldy + c));
        g.x *= act_grad_from_y(y.x, act); g.y *= act_grad_from_y(y.y, act);
        g.z *= act_grad_from_y(y.z, act); g.w *= act_grad_from_y(y.w, act);
        *reinterpret_cast<float4*>(dZ + (size_t)r * lddz + c) = g;
      }
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w))));
    }
  }
  if (amax != nullptr) {
    mx = warp_max(mx);
    if (lane == 0 && mx > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(mx));
  }
  if (db == nullptr) return;
  red[rl][lane] = s;
  __syncthreads();
  if (rl == 0 && c < N) {
    float4 t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 v = red[i][lane];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    atomicAdd(db + c, t.x);
    atomicAdd(db + c + 1, t.y);
    atomicAdd(db + c + 2, t.z);
    atomicAdd(db + c + 3, t.w);
  }
}

// Fused form for the fp16 three-product GEMM mode: dZ = dY * act'(Y) leaves the kernel ONLY as the fp16 hi/lo operand pair
// the weight-gradient and data-gradient GEMMs read (gemm_tc.cu, f16split.cu), together with the bias gradient.  The scale
// comes from an upper bound of |dY| the producer left behind (the dgrad GEMM's epilogue |C|max): |act'| <= 1 for every
// nonlinearity of custom/nonlinearities.py, so |dZ| <= bound and one pass suffices — no float32 dZ, no max pass, no split
// pass (12 bytes per element instead of 28).
__global__ void __launch_bounds__(256) prep_f16_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ Y,
                                                       int ldy, float* __restrict__ db, int M, int N, int act,
                                                       const float* __restrict__ bound, __half* __restrict__ hi,
                                                       __half* __restrict__ lo, int ldo, int32_t* __restrict__ exp_out) {
  __shared__ float4 red[8][33];
  const float b = __ldg(bound);
  int e = 0;
  if (b > 0.f && !isinf(b)) {
    e = 14 - ilogbf(b);
    e = e < -126 ? -126 : (e > 126 ? 126 : e);
  }
  const int e1 = e / 2, e2 = e - e1;
  const float s1 = __int_as_float((127 + e1) << 23), s2 = __int_as_float((127 + e2) << 23);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *exp_out = e;
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * PREP_ROWS;
  const int r1 = min(M, r0 + PREP_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < N) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 8) {
      float4 g = __ldcs(reinterpret_cast<const float4*>(dY + (size_t)r * lddy + c));
      const float4 y = __ldcs(reinterpret_cast<const float4*>(Y + (size_t)r * ldy + c));
      g.x *= act_grad_from_y(y.x, act); g.y *= act_grad_from_y(y.y, act);
      g.z *= act_grad_from_y(y.z, act); g.w *= act_grad_from_y(y.w, act);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      const float v[4] = {g.x * s1 * s2, g.y * s1 * s2, g.z * s1 * s2, g.w * s1 * s2};
      __half h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        h[k] = __float2half_rn(v[k]);
        l[k] = __float2half_rn((v[k] - __half2float(h[k])) * F16_LO_SCALE);
      }
      *reinterpret_cast<uint2*>(hi + (size_t)r * ldo + c) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(lo + (size_t)r * ldo + c) = *reinterpret_cast<const uint2*>(l);
    }
  }
  if (db == nullptr) return;
  red[rl][lane] = s;
  __syncthreads();
  if (rl == 0 && c < N) {
    float4 t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 v = red[i][lane];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    atomicAdd(db + c, t.x);
    atomicAdd(db + c + 1, t.y);
    atomicAdd(db + c + 2, t.z);
    atomicAdd(db + c + 3, t.w);
  }
}

// ------------------------------------------------------------------------------------------------------
// out = sum_s coeff_s * in_s
// ------------------------------------------------------------------------------------------------------
struct PtrPack {
  const float* p[8];
  int ld[8];
  int n;
};

__global__ void fuse_sum_kernel(PtrPack ins, const float* __restrict__ coeffs, float* __restrict__ out, int ldo,
                                int M, int F4) {
  // F4 = F/4 float4 columns; requires F%4==0 and 16B alignment (checked by the host wrapper)
  const size_t total = (size_t)M * F4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / F4), c = (int)(i % F4) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ins.n; ++s) {
      float4 v = *reinterpret_cast<const float4*>(ins.p[s] + (size_t)r * ins.ld[s] + c);
      float w = coeffs ? coeffs[s] : 1.0f;
      a.x += w * v.x; a.y += w * v.y; a.z += w * v.z; a.w += w * v.w;
    }
    *reinterpret_cast<float4*>(out + (size_t)r * ldo + c) = a;
  }
}

__global__ void fuse_sum_scalar_kernel(PtrPack ins, const float* __restrict__ coeffs, float* __restrict__ out,
                                       int ldo, int M, int F) {
  const size_t total = (size_t)M * F;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / F), c = (int)(i % F);
    float a = 0.f;
    for (int s = 0; s < ins.n; ++s) a += (coeffs ? coeffs[s] : 1.0f) * ins.p[s][(size_t)r * ins.ld[s] + c];
    out[(size_t)r * ldo + c] = a;
  }
}

__global__ void __launch_bounds__(256) adasum_coeff_kernel(const float* __restrict__ dout, int lddo, PtrPack ins,
                                                           float* __restrict__ dcoeff, int M, int F) {
  __shared__ float red[8][8];
  float acc[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) acc[s] = 0.f;
  const size_t total = (size_t)M * F;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / F), c = (int)(i % F);
    float g = dout[(size_t)r * lddo + c];
#pragma unroll
    for (int s = 0; s < 8; ++s)
      if (s < ins.n) acc[s] += g * ins.p[s][(size_t)r * ins.ld[s] + c];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    float v = warp_sum(acc[s]);
    if (lane == 0) red[w][s] = v;
  }
  __syncthreads();
  if (threadIdx.x < ins.n) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(dcoeff + threadIdx.x, t);
  }
}

template <bool VEC>
__global__ void copy2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int M, int F,
                              const float* __restrict__ alpha_dev, int accumulate) {
  const float alpha = alpha_dev ? *alpha_dev : 1.0f;
  const int W = VEC ? F / 4 : F;
  const size_t total = (size_t)M * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / W), c = (int)(i % W);
    if (VEC) {
      float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * lds + c * 4);
      float4* d = reinterpret_cast<float4*>(dst + (size_t)r * ldd + c * 4);
      v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
      if (accumulate) { float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *d = v;
    } else {
      float v = alpha * src[(size_t)r * lds + c];
      float* d = dst + (size_t)r * ldd + c;
      *d = accumulate ? *d + v : v;
    }
  }
}

__global__ void slice_last_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int N,
                                  int T, int F, int backward, int accumulate) {
  const size_t total = (size_t)N * F;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n = (int)(i / F), c = (int)(i % F);
    size_t big = ((size_t)n * T + (T - 1));
    if (!backward) {
      float v = src[big * lds + c];
      float* d = dst + (size_t)n * ldd + c;
      *d = accumulate ? *d + v : v;
    } else {
      float v = src[(size_t)n * lds + c];
      float* d = dst + big * ldd + c;
      *d = accumulate ? *d + v : v;
    }
  }
}

__global__ void dropout_kernel(const float* __restrict__ x, int ldx, const uint8_t* __restrict__ keep,
                               float* __restrict__ y, int ldy, int M, int F, float scale) {
  const size_t total = (size_t)M * F;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / F), c = (int)(i % F);
    y[(size_t)r * ldy + c] = keep[i] ? x[(size_t)r * ldx + c] * scale : 0.0f;
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, uint64_t n, float p, uint64_t seed, uint64_t offset) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t h = splitmix64(seed ^ splitmix64(offset + i));
    float u = (float)(h >> 40) * (1.0f / 16777216.0f);
    keep[i] = u >= p ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------------
// BatchNorm over rows (F small, M = N*T rows).  Block = 32 columns x 8 row-lanes.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int ldx, double* __restrict__ stats,
                                                       int M, int F) {
  __shared__ double r1[8][33], r2[8][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int rb = blockIdx.y * PREP_ROWS, re = min(M, rb + PREP_ROWS);
  double s = 0.0, q = 0.0;
  if (c < F)
    for (int r = rb + rl; r < re; r += 8) {
      double v = (double)x[(size_t)r * ldx + c];
      s += v;
      q += v * v;
    }
  r1[rl][lane] = s;
  r2[rl][lane] = q;
  __syncthreads();
  if (rl == 0 && c < F) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) { a += r1[i][lane]; b += r2[i][lane]; }
    atomicAdd(stats + c, a);
    atomicAdd(stats + F + c, b);
  }
}

__global__ void bn_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy,
                              const float* __restrict__ beta, const float* __restrict__ gamma,
                              float* __restrict__ run_mean, float* __restrict__ run_istd,
                              const double* __restrict__ stats, float* __restrict__ save_mean,
                              float* __restrict__ save_istd, int M, int F, double inv_total, float eps, float alpha,
                              int deterministic, int update_running) {
  const size_t total = (size_t)M * F;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / F), c = (int)(i % F);
    float mean, istd;
    if (deterministic) {
      mean = run_mean[c];
      istd = run_istd[c];
    } else {
      double m = stats[c] * inv_total;
      double var = stats[F + c] * inv_total - m * m;
      if (var < 0) var = 0;
      mean = (float)m;
      istd = (float)(1.0 / sqrt(var + (double)eps));
      if (r == 0) {
        save_mean[c] = mean;
        save_istd[c] = istd;
      }
    }
    y[(size_t)r * ldy + c] = (x[(size_t)r * ldx + c] - mean) * (gamma[c] * istd) + beta[c];
  }
}

// running-stat update is a separate tiny kernel so that it happens exactly once and after every reader
__global__ void bn_running_kernel(float* run_mean, float* run_istd, const float* save_mean, const float* save_istd,
                                  int F, float alpha) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < F) {
    run_mean[c] = (1.0f - alpha) * run_mean[c] + alpha * save_mean[c];
    run_istd[c] = (1.0f - alpha) * run_istd[c] + alpha * save_istd[c];
  }
}

__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float* __restrict__ dy, int lddy,
                                                           const float* __restrict__ x, int ldx,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ istd, double* __restrict__ bstats,
                                                           int M, int F) {
  __shared__ double r1[8][33], r2[8][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int rb = blockIdx.y * PREP_ROWS, re = min(M, rb + PREP_ROWS);
  double s = 0.0, q = 0.0;
  if (c < F) {
    float mu = mean[c], is = istd[c];
    for (int r = rb + rl; r < re; r += 8) {
      float g = dy[(size_t)r * lddy + c];
      float xh = (x[(size_t)r * ldx + c] - mu) * is;
      s += (double)g;
      q += (double)(g * xh);
    }
  }
  r1[rl][lane] = s;
  r2[rl][lane] = q;
  __syncthreads();
  if (rl == 0 && c < F) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) { a += r1[i][lane]; b += r2[i][lane]; }
    atomicAdd(bstats + c, a);
    atomicAdd(bstats + F + c, b);
  }
}

__global__ void bn_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                              const float* __restrict__ gamma, const float* __restrict__ mean,
                              const float* __restrict__ istd, const double* __restrict__ bstats,
                              float* __restrict__ dx, int lddx, float* __restrict__ dbeta, float* __restrict__ dgamma,
                              int M, int F, float inv_total, int accumulate_params) {
  const size_t total = (size_t)M * F;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / F), c = (int)(i % F);
    float db = (float)bstats[c], dg = (float)bstats[F + c];
    float is = istd[c];
    float xh = (x[(size_t)r * ldx + c] - mean[c]) * is;
    float g = dy[(size_t)r * lddy + c];
    dx[(size_t)r * lddx + c] = gamma[c] * is * (g - (db + xh * dg) * inv_total);
    if (r == 0) {
      dbeta[c] = accumulate_params ? dbeta[c] + db : db;
      dgamma[c] = accumulate_params ? dgamma[c] + dg : dg;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// softmax + losses: one warp per row
// ------------------------------------------------------------------------------------------------------
__global__ void softmax_kernel(const float* __restrict__ z, int ldz, float* __restrict__ p, int ldp, int M, int C) {
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < M; r += gridDim.x * warps) {
    const float* zr = z + (size_t)r * ldz;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, zr[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(zr[c] - mx);
    s = warp_sum(s);
    float inv = 1.0f / s;
    for (int c = lane; c < C; c += 32) p[(size_t)r * ldp + c] = expf(zr[c] - mx) * inv;
  }
}

// reference custom/objectives.py:27-37 — softmax applied to the already-softmaxed network output
__global__ void temporal_loss_kernel(const float* __restrict__ probs, int ldp, const int32_t* __restrict__ y,
                                     const uint8_t* __restrict__ mask, float* __restrict__ loss_sum,
                                     float* __restrict__ dlogits, int lddl, int M, int C, float inv_norm,
                                     const float* __restrict__ count_dev) {
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  float local = 0.f;
  if (count_dev != nullptr) inv_norm = inv_norm / *count_dev;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < M; r += gridDim.x * warps) {
    const float* pr = probs + (size_t)r * ldp;
    const float m = mask[r] ? 1.0f : 0.0f;
    const int yr = y[r];
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, pr[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(pr[c] - mx);
    s = warp_sum(s);
    const float inv = 1.0f / s;
    // dL/dp_j = m*(q_j - [j==y])*inv_norm ; dL/dz_j = p_j*(dL/dp_j - sum_k p_k dL/dp_k)
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      float q = expf(pr[c] - mx) * inv;
      float dp = m * (q - (c == yr ? 1.0f : 0.0f)) * inv_norm;
      dot += pr[c] * dp;
    }
    dot = warp_sum(dot);
    if (dlogits != nullptr)
      for (int c = lane; c < C; c += 32) {
        float q = expf(pr[c] - mx) * inv;
        float dp = m * (q - (c == yr ? 1.0f : 0.0f)) * inv_norm;
        dlogits[(size_t)r * lddl + c] = pr[c] * (dp - dot);
      }
    if (lane == 0 && m != 0.0f && yr >= 0 && yr < C) local += -(pr[yr] - mx - logf(s));
  }
  __shared__ float red[32];
  if (lane == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < warps; ++i) t += red[i];
    if (t != 0.f) atomicAdd(loss_sum, t);
  }
}

__global__ void xent_kernel(const float* __restrict__ probs, int ldp, const int32_t* __restrict__ y,
                            float* __restrict__ loss_sum, float* __restrict__ dlogits, int lddl, int M, int C,
                            float inv_norm, const float* __restrict__ count_dev) {
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  float local = 0.f;
  if (count_dev != nullptr) inv_norm = inv_norm / *count_dev;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < M; r += gridDim.x * warps) {
    const float* pr = probs + (size_t)r * ldp;
    const int yr = y[r];
    if (dlogits != nullptr)
      for (int c = lane; c < C; c += 32)
        dlogits[(size_t)r * lddl + c] = (pr[c] - (c == yr ? 1.0f : 0.0f)) * inv_norm;
    if (lane == 0 && yr >= 0 && yr < C) local += -logf(pr[yr]);
  }
  __shared__ float red[32];
  if (lane == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < warps; ++i) t += red[i];
    atomicAdd(loss_sum, t);
  }
}

// ------------------------------------------------------------------------------------------------------
// fused multi-tensor optimiser step over the flat parameter arena: 16 B read + 12 B written per parameter
// ------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) optim_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ s1, float* __restrict__ s2, uint64_t n4,
                                                    float lr, const float* __restrict__ seg_lr,
                                                    const int32_t* __restrict__ seg_id, float step_scalar, float hp1,
                                                    float hp2, float eps, float grad_scale) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
    float my_lr = lr;
    if (seg_lr != nullptr) my_lr = seg_lr[seg_id[i >> 6]];      // 64 float4 = 256 floats per segment block
    float4 P = reinterpret_cast<float4*>(p)[i];
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float pv[4] = {P.x, P.y, P.z, P.w}, gv[4] = {G.x, G.y, G.z, G.w};
    float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if (KIND != IPAVSR_OPT_SGD) {
      float4 A = reinterpret_cast<float4*>(s1)[i];
      a[0] = A.x; a[1] = A.y; a[2] = A.z; a[3] = A.w;
    }
    if (KIND == IPAVSR_OPT_ADAM || KIND == IPAVSR_OPT_ADADELTA) {
      float4 B = reinterpret_cast<float4*>(s2)[i];
      b[0] = B.x; b[1] = B.y; b[2] = B.z; b[3] = B.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = gv[j] * grad_scale;
      if (KIND == IPAVSR_OPT_ADAM) {               // custom/updates.py:83-96
        float a_t = my_lr * step_scalar;
        a[j] = hp1 * a[j] + (1.0f - hp1) * gg;
        b[j] = hp2 * b[j] + (1.0f - hp2) * gg * gg;
        pv[j] -= a_t * a[j] / (sqrtf(b[j]) + eps);
      } else if (KIND == IPAVSR_OPT_ADADELTA) {    // lasagne.updates.adadelta
        a[j] = hp1 * a[j] + (1.0f - hp1) * gg * gg;
        float upd = gg * sqrtf(b[j] + eps) / sqrtf(a[j] + eps);
        pv[j] -= my_lr * upd;
        b[j] = hp1 * b[j] + (1.0f - hp1) * upd * upd;
      } else if (KIND == IPAVSR_OPT_SGD) {
        pv[j] -= my_lr * gg;
      } else if (KIND == IPAVSR_OPT_MOMENTUM) {    // apply_momentum
        a[j] = hp1 * a[j] - my_lr * gg;
        pv[j] += a[j];
      } else {                                     // apply_nesterov_momentum
        a[j] = hp1 * a[j] - my_lr * gg;
        pv[j] += hp1 * a[j] - my_lr * gg;
      }
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    if (KIND != IPAVSR_OPT_SGD) reinterpret_cast<float4*>(s1)[i] = make_float4(a[0], a[1], a[2], a[3]);
    if (KIND == IPAVSR_OPT_ADAM || KIND == IPAVSR_OPT_ADADELTA)
      reinterpret_cast<float4*>(s2)[i] = make_float4(b[0], b[1], b[2], b[3]);
  }
}

__global__ void fill_kernel(float* p, uint64_t n, float v) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void tf32_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                  uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    if (hi) hi[i] = h;
    lo[i] = v - h;
  }
}

static inline int grid_for(size_t total, int threads = 256, int per_sm = 8) {
  size_t want = (total + threads - 1) / threads;
  size_t cap = (size_t)sm_count() * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

int ipavsr_dense_bwd_prep(const float* dY, int lddy, const float* Y, int ldy, float* dZ, int lddz, float* db, int M,
                          int N, int act, int accumulate_db, float* amax, void* stream) {
  IPAVSR_CHECK_ARG(M >= 0 && N >= 0 && dY && Y && dZ, "bad arguments");
  if (M == 0 || N == 0) return IPAVSR_OK;
  if (db && !accumulate_db) IPAVSR_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, S(stream)));
  if (N % 4 == 0 && lddy % 4 == 0 && ldy % 4 == 0 && lddz % 4 == 0 && aligned16(dY) && aligned16(Y) && aligned16(dZ)) {
    dim3 vgrid((N + 127) / 128, (M + PREP_ROWS - 1) / PREP_ROWS);
    colsum_prep_vec_kernel<true><<<vgrid, 256, 0, S(stream)>>>(dY, lddy, Y, ldy, dZ, lddz, db, M, N, act, amax);
    IPAVSR_LAUNCH_CHECK();
    return IPAVSR_OK;
  }
  dim3 grid((N + 31) / 32, (M + PREP_ROWS - 1) / PREP_ROWS);
  colsum_prep_kernel<true><<<grid, 256, 0, S(stream)>>>(dY, lddy, Y, ldy, dZ, lddz, db, M, N, act, amax);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_dense_bwd_prep_f16_supported(const float* dY, int lddy, const float* Y, int ldy, int N, const void* hi,
                                        const void* lo, int ldo) {
  return (N % 4 == 0 && lddy % 4 == 0 && ldy % 4 == 0 && ldo % 4 == 0 && ldo >= N && aligned16(dY) && aligned16(Y) &&
          (reinterpret_cast<uintptr_t>(hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7) == 0) ? 1 : 0;
}

int ipavsr_dense_bwd_prep_f16(const float* dY, int lddy, const float* Y, int ldy, float* db, int M, int N, int act,
                              int accumulate_db, const float* bound, uint16_t* dZ_hi, uint16_t* dZ_lo, int ldo,
                              int32_t* exp_out, void* stream) {
  IPAVSR_CHECK_ARG(M >= 0 && N >= 0 && dY && Y && bound && dZ_hi && dZ_lo && exp_out, "bad arguments");
  IPAVSR_CHECK_ARG(ipavsr_dense_bwd_prep_f16_supported(dY, lddy, Y, ldy, N, dZ_hi, dZ_lo, ldo),
                   "needs N % 4 == 0, leading dimensions % 4 == 0 and 16-byte aligned rows");
  if (M == 0 || N == 0) return IPAVSR_OK;
  if (db && !accumulate_db) IPAVSR_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, S(stream)));
  dim3 grid((N + 127) / 128, (M + PREP_ROWS - 1) / PREP_ROWS);
  prep_f16_kernel<<<grid, 256, 0, S(stream)>>>(dY, lddy, Y, ldy, db, M, N, act, bound, reinterpret_cast<__half*>(dZ_hi),
                                               reinterpret_cast<__half*>(dZ_lo), ldo, exp_out);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_colsum(const float* X, int ldx, float* out, int M, int N, int accumulate, void* stream) {
  IPAVSR_CHECK_ARG(M >= 0 && N >= 0 && X && out, "bad arguments");
  if (N == 0) return IPAVSR_OK;
  if (!accumulate) IPAVSR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, S(stream)));
  if (M == 0) return IPAVSR_OK;
  if (N % 4 == 0 && ldx % 4 == 0 && aligned16(X)) {
    dim3 vgrid((N + 127) / 128, (M + PREP_ROWS - 1) / PREP_ROWS);
    colsum_prep_vec_kernel<false><<<vgrid, 256, 0, S(stream)>>>(X, ldx, nullptr, 0, nullptr, 0, out, M, N, 0, nullptr);
    IPAVSR_LAUNCH_CHECK();
    return IPAVSR_OK;
  }
  dim3 grid((N + 31) / 32, (M + PREP_ROWS - 1) / PREP_ROWS);
  colsum_prep_kernel<false><<<grid, 256, 0, S(stream)>>>(X, ldx, nullptr, 0, nullptr, 0, out, M, N, 0, nullptr);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_fuse_sum(const float* const* ins, const int* lds, int Sn, const float* coeffs, float* out, int ldo, int M,
                    int F, void* stream) {
  IPAVSR_CHECK_ARG(Sn >= 1 && Sn <= 8 && ins && lds && out, "1..8 inputs required");
  if (M == 0 || F == 0) return IPAVSR_OK;
  PtrPack pk;
  pk.n = Sn;
  bool vec = (F % 4 == 0) && (ldo % 4 == 0) && aligned16(out);
  for (int s = 0; s < Sn; ++s) {
    pk.p[s] = ins[s];
    pk.ld[s] = lds[s];
    vec = vec && (lds[s] % 4 == 0) && aligned16(ins[s]);
  }
  if (vec)
    fuse_sum_kernel<<<grid_for((size_t)M * F / 4), 256, 0, S(stream)>>>(pk, coeffs, out, ldo, M, F / 4);
  else
    fuse_sum_scalar_kernel<<<grid_for((size_t)M * F), 256, 0, S(stream)>>>(pk, coeffs, out, ldo, M, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_adasum_bwd_coeff(const float* dout, int lddo, const float* const* ins, const int* lds, int Sn,
                            float* dcoeff, int M, int F, int accumulate, void* stream) {
  IPAVSR_CHECK_ARG(Sn >= 1 && Sn <= 8 && ins && lds && dcoeff && dout, "1..8 inputs required");
  if (!accumulate) IPAVSR_CUDA(cudaMemsetAsync(dcoeff, 0, sizeof(float) * Sn, S(stream)));
  if (M == 0 || F == 0) return IPAVSR_OK;
  PtrPack pk;
  pk.n = Sn;
  for (int s = 0; s < Sn; ++s) { pk.p[s] = ins[s]; pk.ld[s] = lds[s]; }
  adasum_coeff_kernel<<<grid_for((size_t)M * F, 256, 4), 256, 0, S(stream)>>>(dout, lddo, pk, dcoeff, M, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_copy2d(const float* src, int lds, float* dst, int ldd, int M, int F, const float* alpha_dev,
                  int accumulate, void* stream) {
  IPAVSR_CHECK_ARG(src && dst && M >= 0 && F >= 0, "bad arguments");
  if (M == 0 || F == 0) return IPAVSR_OK;
  bool vec = (F % 4 == 0) && (lds % 4 == 0) && (ldd % 4 == 0) && aligned16(src) && aligned16(dst);
  if (vec)
    copy2d_kernel<true><<<grid_for((size_t)M * F / 4), 256, 0, S(stream)>>>(src, lds, dst, ldd, M, F, alpha_dev, accumulate);
  else
    copy2d_kernel<false><<<grid_for((size_t)M * F), 256, 0, S(stream)>>>(src, lds, dst, ldd, M, F, alpha_dev, accumulate);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_slice_last(const float* src, int lds, float* dst, int ldd, int N, int T, int F, int backward,
                      int accumulate, void* stream) {
  IPAVSR_CHECK_ARG(src && dst && N >= 0 && T >= 1 && F >= 0, "bad arguments");
  if (N == 0 || F == 0) return IPAVSR_OK;
  slice_last_kernel<<<grid_for((size_t)N * F), 256, 0, S(stream)>>>(src, lds, dst, ldd, N, T, F, backward, accumulate);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_dropout(const float* x, int ldx, const uint8_t* keep, float* y, int ldy, int M, int F, float scale,
                   void* stream) {
  IPAVSR_CHECK_ARG(x && keep && y, "bad arguments");
  if (M == 0 || F == 0) return IPAVSR_OK;
  dropout_kernel<<<grid_for((size_t)M * F), 256, 0, S(stream)>>>(x, ldx, keep, y, ldy, M, F, scale);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_dropout_mask(uint8_t* keep, uint64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  IPAVSR_CHECK_ARG(keep, "bad arguments");
  if (n == 0) return IPAVSR_OK;
  dropout_mask_kernel<<<grid_for(n), 256, 0, S(stream)>>>(keep, n, p, seed, offset);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_bn_stats(const float* x, int ldx, double* stats, int M, int F, void* stream) {
  IPAVSR_CHECK_ARG(x && stats && F > 0, "bad arguments");
  IPAVSR_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * F, S(stream)));
  if (M == 0) return IPAVSR_OK;
  dim3 grid((F + 31) / 32, (M + PREP_ROWS - 1) / PREP_ROWS);
  bn_stats_kernel<<<grid, 256, 0, S(stream)>>>(x, ldx, stats, M, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_bn_fwd(const float* x, int ldx, float* y, int ldy, const float* beta, const float* gamma, float* run_mean,
                  float* run_istd, const double* stats, float* save_mean, float* save_istd, int M, int F,
                  int64_t M_total, float eps, float alpha, int deterministic, int update_running, void* stream) {
  IPAVSR_CHECK_ARG(x && y && beta && gamma && run_mean && run_istd, "bad arguments");
  IPAVSR_CHECK_ARG(deterministic || (stats && save_mean && save_istd && M_total > 0), "training mode needs stats");
  if (M == 0 || F == 0) return IPAVSR_OK;
  bn_fwd_kernel<<<grid_for((size_t)M * F), 256, 0, S(stream)>>>(x, ldx, y, ldy, beta, gamma, run_mean, run_istd, stats,
                                                              save_mean, save_istd, M, F,
                                                              deterministic ? 0.0 : 1.0 / (double)M_total, eps, alpha,
                                                              deterministic, update_running);
  IPAVSR_LAUNCH_CHECK();
  if (!deterministic && update_running) {
    bn_running_kernel<<<(F + 127) / 128, 128, 0, S(stream)>>>(run_mean, run_istd, save_mean, save_istd, F, alpha);
    IPAVSR_LAUNCH_CHECK();
  }
  return IPAVSR_OK;
}

int ipavsr_bn_bwd_stats(const float* dy, int lddy, const float* x, int ldx, const float* save_mean,
                        const float* save_istd, double* bstats, int M, int F, void* stream) {
  IPAVSR_CHECK_ARG(dy && x && save_mean && save_istd && bstats, "bad arguments");
  IPAVSR_CUDA(cudaMemsetAsync(bstats, 0, sizeof(double) * 2 * F, S(stream)));
  if (M == 0) return IPAVSR_OK;
  dim3 grid((F + 31) / 32, (M + PREP_ROWS - 1) / PREP_ROWS);
  bn_bwd_stats_kernel<<<grid, 256, 0, S(stream)>>>(dy, lddy, x, ldx, save_mean, save_istd, bstats, M, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_bn_bwd(const float* dy, int lddy, const float* x, int ldx, const float* gamma, const float* save_mean,
                  const float* save_istd, const double* bstats, float* dx, int lddx, float* dbeta, float* dgamma,
                  int M, int F, int64_t M_total, int accumulate_params, void* stream) {
  IPAVSR_CHECK_ARG(dy && x && gamma && save_mean && save_istd && bstats && dx && dbeta && dgamma && M_total > 0,
                   "bad arguments");
  if (M == 0 || F == 0) return IPAVSR_OK;
  bn_bwd_kernel<<<grid_for((size_t)M * F), 256, 0, S(stream)>>>(dy, lddy, x, ldx, gamma, save_mean, save_istd, bstats,
                                                              dx, lddx, dbeta, dgamma, M, F, 1.0f / (float)M_total,
                                                              accumulate_params);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_softmax(const float* logits, int ldl, float* probs, int ldp, int M, int C, void* stream) {
  IPAVSR_CHECK_ARG(logits && probs && C > 0, "bad arguments");
  if (M == 0) return IPAVSR_OK;
  softmax_kernel<<<grid_for((size_t)M * 32), 256, 0, S(stream)>>>(logits, ldl, probs, ldp, M, C);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_temporal_softmax_loss(const float* probs, int ldp, const int32_t* y, const uint8_t* mask, float* loss_sum,
                                 float* dlogits, int lddl, int M, int C, float inv_norm, const float* count_dev,
                                 void* stream) {
  IPAVSR_CHECK_ARG(probs && y && mask && loss_sum && C > 0, "bad arguments");
  if (M == 0) return IPAVSR_OK;
  temporal_loss_kernel<<<grid_for((size_t)M * 32), 256, 0, S(stream)>>>(probs, ldp, y, mask, loss_sum, dlogits, lddl,
                                                                        M, C, inv_norm, count_dev);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_categorical_crossentropy(const float* probs, int ldp, const int32_t* y, float* loss_sum, float* dlogits,
                                    int lddl, int M, int C, float inv_norm, const float* count_dev, void* stream) {
  IPAVSR_CHECK_ARG(probs && y && loss_sum && C > 0, "bad arguments");
  if (M == 0) return IPAVSR_OK;
  xent_kernel<<<grid_for((size_t)M * 32), 256, 0, S(stream)>>>(probs, ldp, y, loss_sum, dlogits, lddl, M, C, inv_norm,
                                                               count_dev);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_optim_step(int kind, float* p, const float* g, float* s1, float* s2, uint64_t n, float lr,
                      const float* seg_lr, const int32_t* seg_id, float step_scalar, float hp1, float hp2, float eps,
                      float grad_scale, void* stream) {
  IPAVSR_CHECK_ARG(p && g, "bad arguments");
  IPAVSR_CHECK_ARG(n % 4 == 0 && aligned16(p) && aligned16(g), "arena must be 16-byte aligned and a multiple of 4 floats");
  IPAVSR_CHECK_ARG((seg_lr == nullptr) == (seg_id == nullptr), "seg_lr and seg_id go together");
  if (n == 0) return IPAVSR_OK;
  uint64_t n4 = n / 4;
  int grid = grid_for(n4);
  cudaStream_t st = S(stream);
  switch (kind) {
    case IPAVSR_OPT_ADAM:
      IPAVSR_CHECK_ARG(s1 && s2, "adam needs two state arenas");
      optim_kernel<IPAVSR_OPT_ADAM><<<grid, 256, 0, st>>>(p, g, s1, s2, n4, lr, seg_lr, seg_id, step_scalar, hp1, hp2, eps, grad_scale);
      break;
    case IPAVSR_OPT_ADADELTA:
      IPAVSR_CHECK_ARG(s1 && s2, "adadelta needs two state arenas");
      optim_kernel<IPAVSR_OPT_ADADELTA><<<grid, 256, 0, st>>>(p, g, s1, s2, n4, lr, seg_lr, seg_id, step_scalar, hp1, hp2, eps, grad_scale);
      break;
    case IPAVSR_OPT_SGD:
      optim_kernel<IPAVSR_OPT_SGD><<<grid, 256, 0, st>>>(p, g, s1, s2, n4, lr, seg_lr, seg_id, step_scalar, hp1, hp2, eps, grad_scale);
      break;
    case IPAVSR_OPT_MOMENTUM:
      IPAVSR_CHECK_ARG(s1, "momentum needs a velocity arena");
      optim_kernel<IPAVSR_OPT_MOMENTUM><<<grid, 256, 0, st>>>(p, g, s1, s2, n4, lr, seg_lr, seg_id, step_scalar, hp1, hp2, eps, grad_scale);
      break;
    case IPAVSR_OPT_NESTEROV:
      IPAVSR_CHECK_ARG(s1, "nesterov needs a velocity arena");
      optim_kernel<IPAVSR_OPT_NESTEROV><<<grid, 256, 0, st>>>(p, g, s1, s2, n4, lr, seg_lr, seg_id, step_scalar, hp1, hp2, eps, grad_scale);
      break;
    default:
      set_error("ipavsr_optim_step: unknown kind %d", kind);
      return IPAVSR_ERR_ARG;
  }
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_fill(float* p, uint64_t n, float v, void* stream) {
  IPAVSR_CHECK_ARG(p || n == 0, "bad arguments");
  if (n == 0) return IPAVSR_OK;
  fill_kernel<<<grid_for(n), 256, 0, S(stream)>>>(p, n, v);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_tf32_split(const float* x, float* hi, float* lo, uint64_t n, void* stream) {
  IPAVSR_CHECK_ARG(x && lo, "bad arguments");
  if (n == 0) return IPAVSR_OK;
  tf32_split_kernel<<<grid_for(n), 256, 0, S(stream)>>>(x, hi, lo, n);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // extern "C"
