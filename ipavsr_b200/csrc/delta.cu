// DeltaLayer forward/backward as one fused streaming kernel each.
//
// Replaces the reference's three nested theano.scan loops (custom/layers.py:105-121 -> utils/signal.py:59-80
// -> delta_t :26-39 -> delta_theta :7-23).  Per utterance, with Theta edge-replicated frames on both sides:
//     d[t] = sum_{theta=1..Theta} float32( acc + float64(theta) * float64(x[t+theta]-x[t-theta]) / float64(2 theta^2) )
// (float32 subtraction, float64 product/quotient/add, float32 round after every theta), then the same operator
// on d gives a; output row = [x | d | a].  The layer is mask-agnostic and runs over the whole padded T.
//
// General kernel (any T, any Theta): a CTA stages UPC utterances' (T x F) tiles in shared memory (coalesced row loads), every thread owns a
// (utterance, feature, 4-frame run) and slides a register window of 2*Theta+4 frames over it (window loads are
// bank-conflict-free: lanes walk the feature axis), writes d back to shared memory, repeats for a, then the CTA
// streams the [x|d|a] rows out coalesced.  Algorithmic traffic 16*F bytes/frame (read 4F, write 12F): HBM-bound.
// EXACT=true reproduces the reference's float64 operation sequence bit for bit (correctly rounded quotient, float64
// add, float32 round per theta; FP64-pipe bound); EXACT=false is a plain float32 FMA chain (few-ulp deviation, stated
// in DESIGN.md; HBM bound).
#include "common.cuh"

namespace ipavsr {

constexpr int DELTA_THREADS = 256;
constexpr int DELTA_RUN = 4;

template <bool EXACT>
__device__ __forceinline__ float delta_step(float acc, float diff, int th) {
  if (EXACT) {
    // The reference evaluates, in float64:  term = (theta*diff) / (2*theta*theta)  [= RN53(diff / (2 theta)), the
    // numerator is exact],  s = RN53(acc + term),  acc = RN24(s)   (utils/signal.py:19-21).  Exact float32 ties of s
    // are common (whenever 2*theta divides diff's mantissa), so the quotient has to be the correctly rounded one:
    // for a power-of-two theta the product diff * (1/(2 theta)) is exact; otherwise one Markstein correction step
    // (q = d*r; rem = fma(-q, c, d) exact; q' = fma(rem, r, q)) gives the correctly rounded quotient for these small
    // integer divisors c.
    const double c = 2.0 * (double)th;
    const double r = 1.0 / c;
    const double d = (double)diff;
    double q = d * r;
    if ((th & (th - 1)) != 0) {
      const double rem = fma(-q, c, d);
      q = fma(rem, r, q);
    }
    return (float)((double)acc + q);
  } else {
    return fmaf(diff, 1.0f / (2.0f * (float)th), acc);
  }
}

// src/dst: shared-memory tiles of one utterance, element (t,f) at [t*F + f]
template <int TH, bool EXACT>
__device__ __forceinline__ void delta_tile(const float* __restrict__ src, float* __restrict__ dst, int T, int F,
                                           int theta, int tid, int nthreads) {
  const int runs = (T + DELTA_RUN - 1) / DELTA_RUN;
  const int items = runs * F;
  for (int it = tid; it < items; it += nthreads) {
    const int f = it % F;
    const int t0 = (it / F) * DELTA_RUN;
    if (TH > 0) {
      float w[2 * TH + DELTA_RUN];
#pragma unroll
      for (int i = 0; i < 2 * TH + DELTA_RUN; ++i) {
        int t = min(max(t0 - TH + i, 0), T - 1);
        w[i] = src[t * F + f];
      }
#pragma unroll
      for (int r = 0; r < DELTA_RUN; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int th = 1; th <= TH; ++th) acc = delta_step<EXACT>(acc, w[r + TH + th] - w[r + TH - th], th);
        if (t0 + r < T) dst[(t0 + r) * F + f] = acc;
      }
    } else {
      for (int r = 0; r < DELTA_RUN && t0 + r < T; ++r) {
        const int t = t0 + r;
        float acc = 0.f;
        for (int th = 1; th <= theta; ++th) {
          float hi = src[min(t + th, T - 1) * F + f];
          float lo = src[max(t - th, 0) * F + f];
          acc = delta_step<EXACT>(acc, hi - lo, th);
        }
        dst[t * F + f] = acc;
      }
    }
  }
}

template <int TH, bool EXACT>
__global__ void __launch_bounds__(DELTA_THREADS) delta_fwd_kernel(const float* __restrict__ x, int ldx,
                                                                  float* __restrict__ y, int ldy, int N, int T, int F,
                                                                  int theta, int upc) {
  extern __shared__ __align__(16) float sm[];
  const int TF = T * F;
  float* sx = sm;                      // [upc][T][F]
  float* sd = sm + (size_t)upc * TF;   // [upc][T][F]
  float* sa = sd + (size_t)upc * TF;   // [upc][T][F]
  const int tid = threadIdx.x;
  const bool vec_in = (F % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int u0 = blockIdx.x * upc; u0 < N; u0 += gridDim.x * upc) {
    const int nu = min(upc, N - u0);
    const int rows = nu * T;
    // ---- load rows (coalesced) ----
    if (vec_in) {
      const int F4 = F / 4;
      for (int i = tid; i < rows * F4; i += DELTA_THREADS) {
        int r = i / F4, c = (i % F4) * 4;
        float4 v = __ldcs(reinterpret_cast<const float4*>(x + ((size_t)u0 * T + r) * ldx + c));
        *reinterpret_cast<float4*>(sx + r * F + c) = v;
      }
    } else {
      for (int i = tid; i < rows * F; i += DELTA_THREADS) {
        int r = i / F, c = i % F;
        sx[r * F + c] = __ldcs(x + ((size_t)u0 * T + r) * ldx + c);
      }
    }
    __syncthreads();
    for (int u = 0; u < nu; ++u) delta_tile<TH, EXACT>(sx + u * TF, sd + u * TF, T, F, theta, tid, DELTA_THREADS);
    __syncthreads();
    for (int u = 0; u < nu; ++u) delta_tile<TH, EXACT>(sd + u * TF, sa + u * TF, T, F, theta, tid, DELTA_THREADS);
    __syncthreads();
    // ---- store [x | d | a] rows (coalesced along the 3F row) ----
    const int F3 = 3 * F;
    for (int i = tid; i < rows * F3; i += DELTA_THREADS) {
      int r = i / F3, c = i % F3;
      float v = c < F ? sx[r * F + c] : (c < 2 * F ? sd[r * F + (c - F)] : sa[r * F + (c - 2 * F)]);
      __stcs(y + ((size_t)u0 * T + r) * ldy + c, v);
    }
    __syncthreads();
  }
}

// Backward: gx = g_x + D^T (g_d + D^T g_a), D^T applied as a scatter with the clamp folded in.
__device__ __forceinline__ void delta_T_column(const float* __restrict__ v, float* __restrict__ out, int T, int F,
                                               int theta, int f) {
  // out[s] += sum_t sum_th ( [clamp(t+th)==s] - [clamp(t-th)==s] ) * v[t] / (2 th);  thread owns column f
  for (int t = 0; t < T; ++t) {
    float vt = v[t * F + f];
    for (int th = 1; th <= theta; ++th) {
      float w = vt * (1.0f / (2.0f * (float)th));
      out[min(t + th, T - 1) * F + f] += w;
      out[max(t - th, 0) * F + f] -= w;
    }
  }
}

__global__ void __launch_bounds__(DELTA_THREADS) delta_bwd_kernel(const float* __restrict__ gy, int ldgy,
                                                                  float* __restrict__ gx, int ldgx, int N, int T,
                                                                  int F, int theta, int upc, int accumulate) {
  extern __shared__ __align__(16) float sm[];
  const int TF = T * F;
  float* s0 = sm;                      // g_a, later scratch
  float* s1 = sm + (size_t)upc * TF;   // g_d + D^T g_a
  float* s2 = s1 + (size_t)upc * TF;   // g_x + D^T (...)
  const int tid = threadIdx.x;
  const int F3 = 3 * F;
  for (int u0 = blockIdx.x * upc; u0 < N; u0 += gridDim.x * upc) {
    const int nu = min(upc, N - u0);
    const int rows = nu * T;
    for (int i = tid; i < rows * F3; i += DELTA_THREADS) {
      int r = i / F3, c = i % F3;
      float v = gy[((size_t)u0 * T + r) * ldgy + c];
      if (c < F) s2[r * F + c] = v;
      else if (c < 2 * F) s1[r * F + (c - F)] = v;
      else s0[r * F + (c - 2 * F)] = v;
    }
    __syncthreads();
    for (int i = tid; i < nu * F; i += DELTA_THREADS) {
      int u = i / F, f = i % F;
      delta_T_column(s0 + u * TF, s1 + u * TF, T, F, theta, f);
      delta_T_column(s1 + u * TF, s2 + u * TF, T, F, theta, f);
    }
    __syncthreads();
    for (int i = tid; i < rows * F; i += DELTA_THREADS) {
      int r = i / F, c = i % F;
      float* o = gx + ((size_t)u0 * T + r) * ldgx + c;
      float v = s2[r * F + c];
      *o = accumulate ? *o + v : v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Column kernel (the fast path: T <= 48 and Theta in {1,4,9}).  One THREAD owns one (utterance, feature) column and
// keeps the whole column in registers: TMAX + 2*Theta edge-replicated frames of x (the clamp is folded into the load
// address, so every register index is static), then d, then a.  No shared memory, no barriers; a warp's loads and
// stores at a fixed t cover 32 consecutive features = one 128-byte line; every thread has ~60 independent loads in
// flight, which is what keeps HBM busy.  Frames t >= T of the TMAX template are computed on clamped data and
// discarded.
// ---------------------------------------------------------------------------------------------------------
template <int TH, int TMAX, bool EXACT>
__global__ void __launch_bounds__(128) delta_fwd_col_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                            int ldy, int N, int T, int F) {
  const long long total = (long long)N * F;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total;
       c += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(c / F), f = (int)(c % F);
    const float* xc = x + (size_t)n * T * ldx + f;
    float* yc = y + (size_t)n * T * ldy + f;
    float xs[TMAX + 2 * TH];
#pragma unroll
    for (int i = 0; i < TMAX + 2 * TH; ++i) {
      const int t = min(max(i - TH, 0), T - 1);
      xs[i] = __ldg(xc + (size_t)t * ldx);
    }
    float d[TMAX];
    float dlast = 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      float acc = 0.f;
#pragma unroll
      for (int th = 1; th <= TH; ++th) acc = delta_step<EXACT>(acc, xs[t + TH + th] - xs[t + TH - th], th);
      d[t] = acc;
      if (t == T - 1) dlast = acc;
      if (t < T) {
        __stcs(yc + (size_t)t * ldy, xs[t + TH]);
        __stcs(yc + (size_t)t * ldy + F, acc);
      }
    }
    // a[t] from the edge-replicated d: index clamp(t +- th, 0, T-1); static register indices, runtime select for >= T
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      float acc = 0.f;
#pragma unroll
      for (int th = 1; th <= TH; ++th) {
        const int ip = t + th, im = t - th;
        const float hi = ip < TMAX ? (ip < T ? d[ip < TMAX ? ip : 0] : dlast) : dlast;
        const float lo = d[im > 0 ? im : 0];
        acc = delta_step<EXACT>(acc, hi - lo, th);
      }
      if (t < T) __stcs(yc + (size_t)t * ldy + 2 * F, acc);
    }
  }
}

template <int TH, int TMAX, bool EXACT>
static int launch_delta_col(const float* x, int ldx, float* y, int ldy, int N, int T, int F, cudaStream_t st) {
  const long long total = (long long)N * F;
  long long blocks = (total + 127) / 128;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  delta_fwd_col_kernel<TH, TMAX, EXACT><<<(int)blocks, 128, 0, st>>>(x, ldx, y, ldy, N, T, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

template <int TH, bool EXACT>
static int dispatch_delta_col(const float* x, int ldx, float* y, int ldy, int N, int T, int F, cudaStream_t st) {
  if (T <= 24) return launch_delta_col<TH, 24, EXACT>(x, ldx, y, ldy, N, T, F, st);
  if (T <= 40) return launch_delta_col<TH, 40, EXACT>(x, ldx, y, ldy, N, T, F, st);
  return launch_delta_col<TH, 48, EXACT>(x, ldx, y, ldy, N, T, F, st);
}

template <int TH, bool EXACT>
static int launch_delta_fwd(const float* x, int ldx, float* y, int ldy, int N, int T, int F, int theta, int upc,
                            int grid, size_t smem, cudaStream_t st) {
  auto k = delta_fwd_kernel<TH, EXACT>;
  if (smem > 48 * 1024) IPAVSR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, DELTA_THREADS, smem, st>>>(x, ldx, y, ldy, N, T, F, theta, upc);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

int ipavsr_delta_fwd(const float* x, int ldx, float* y, int ldy, int N, int T, int F, int theta, int exact,
                     void* stream) {
  IPAVSR_CHECK_ARG(x && y && N >= 0 && T >= 1 && F >= 1 && theta >= 0, "bad arguments");
  IPAVSR_CHECK_ARG(ldx >= F && ldy >= 3 * F, "leading dimensions too small");
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (T <= 48 && (theta == 1 || theta == 4 || theta == 9)) {
    // register-resident column kernel (every shipped configuration: T <= 40, theta = 9)
    if (theta == 1) return exact ? dispatch_delta_col<1, true>(x, ldx, y, ldy, N, T, F, st)
                                 : dispatch_delta_col<1, false>(x, ldx, y, ldy, N, T, F, st);
    if (theta == 4) return exact ? dispatch_delta_col<4, true>(x, ldx, y, ldy, N, T, F, st)
                                 : dispatch_delta_col<4, false>(x, ldx, y, ldy, N, T, F, st);
    return exact ? dispatch_delta_col<9, true>(x, ldx, y, ldy, N, T, F, st)
                 : dispatch_delta_col<9, false>(x, ldx, y, ldy, N, T, F, st);
  }
  const size_t per_utt = (size_t)3 * T * F * sizeof(float);
  IPAVSR_CHECK_ARG(per_utt <= 200 * 1024, "one utterance tile (3*T*F floats) must fit in 200 KB of shared memory");
  // utterances per CTA: aim for ~36 KB of shared memory per CTA so ~6 CTAs stay resident per SM
  int upc = (int)((36 * 1024) / per_utt);
  if (upc < 1) upc = 1;
  if (upc > 8) upc = 8;
  size_t smem = per_utt * upc;
  int blocks_needed = (N + upc - 1) / upc;
  int cap = sm_count() * 6;
  int grid = blocks_needed < cap ? blocks_needed : cap;
#define IPAVSR_DELTA_CASE(TH)                                                                              \
  return exact ? launch_delta_fwd<TH, true>(x, ldx, y, ldy, N, T, F, theta, upc, grid, smem, st)           \
               : launch_delta_fwd<TH, false>(x, ldx, y, ldy, N, T, F, theta, upc, grid, smem, st)
  switch (theta) {
    case 1: IPAVSR_DELTA_CASE(1);
    case 2: IPAVSR_DELTA_CASE(2);
    case 3: IPAVSR_DELTA_CASE(3);
    case 4: IPAVSR_DELTA_CASE(4);
    case 9: IPAVSR_DELTA_CASE(9);
    default: IPAVSR_DELTA_CASE(0);
  }
#undef IPAVSR_DELTA_CASE
}

int ipavsr_delta_bwd(const float* gy, int ldgy, float* gx, int ldgx, int N, int T, int F, int theta, int accumulate,
                     void* stream) {
  IPAVSR_CHECK_ARG(gy && gx && N >= 0 && T >= 1 && F >= 1 && theta >= 0, "bad arguments");
  IPAVSR_CHECK_ARG(ldgy >= 3 * F && ldgx >= F, "leading dimensions too small");
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t per_utt = (size_t)3 * T * F * sizeof(float);
  IPAVSR_CHECK_ARG(per_utt <= 200 * 1024, "one utterance tile (3*T*F floats) must fit in 200 KB of shared memory");
  int upc = (int)((36 * 1024) / per_utt);
  if (upc < 1) upc = 1;
  if (upc > 8) upc = 8;
  size_t smem = per_utt * upc;
  int blocks_needed = (N + upc - 1) / upc;
  int cap = sm_count() * 6;
  int grid = blocks_needed < cap ? blocks_needed : cap;
  if (smem > 48 * 1024)
    IPAVSR_CUDA(cudaFuncSetAttribute(delta_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  delta_bwd_kernel<<<grid, DELTA_THREADS, smem, st>>>(gy, ldgy, gx, ldgx, N, T, F, theta, upc, accumulate);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // extern "C"
