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
#include <stdlib.h>
#include "common.cuh"
#include "delta_common.cuh"

namespace ipavsr {

constexpr int DELTA_THREADS = 256;
constexpr int DELTA_RUN = 4;

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

// Backward as a GATHER (the production kernel; the scatter form above is kept as the cross-check).  With c_th = 1/(2 th)
// and v zero outside [0, T):
//     (D^T v)[s] = sum_th c_th (v[s-th] - v[s+th])                          0 < s < T-1
//     (D^T v)[T-1] = sum_th c_th sum_{t = max(0, T-1-th)}^{T-1} v[t]         (everything the forward clamped onto T-1)
//     (D^T v)[0]   = - sum_th c_th sum_{t = 0}^{min(th, T-1)} v[t]          (everything clamped onto 0)
// A CTA stages the [g_x | g_d | g_a] rows of `upc` utterances in shared memory (coalesced), every thread owns a run of 4
// frames of one feature and slides a register window over it (lanes walk the feature axis: conflict-free), twice.
constexpr int DBW_RUN = 4;

template <int TH>
__device__ __forceinline__ void delta_T_gather(const float* __restrict__ v, const float* __restrict__ base,
                                               float* __restrict__ out, int T, int F, int theta, int tid, int nthreads) {
  // out[s][f] = base[s][f] + (D^T v)[s][f]   (out may alias base)
  const int runs = (T + DBW_RUN - 1) / DBW_RUN;
  const int items = runs * F;
  const int th_n = TH > 0 ? TH : theta;
  for (int it = tid; it < items; it += nthreads) {
    const int f = it % F;
    const int s0 = (it / F) * DBW_RUN;
    float w[2 * (TH > 0 ? TH : 1) + DBW_RUN];
    if (TH > 0) {
#pragma unroll
      for (int i = 0; i < 2 * TH + DBW_RUN; ++i) {
        const int t = s0 - TH + i;
        w[i] = (t >= 0 && t < T) ? v[t * F + f] : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < DBW_RUN; ++r) {
      const int s = s0 + r;
      if (s >= T) break;
      float acc = 0.f;
      if (T > 1 && s == T - 1) {
        float run = v[(T - 1) * F + f];
        for (int th = 1; th <= th_n; ++th) {
          if (T - 1 - th >= 0) run += v[(T - 1 - th) * F + f];
          acc = fmaf(run, 1.0f / (2.0f * (float)th), acc);
        }
      } else if (T > 1 && s == 0) {
        float run = v[f];
        for (int th = 1; th <= th_n; ++th) {
          if (th <= T - 1) run += v[th * F + f];
          acc = fmaf(-run, 1.0f / (2.0f * (float)th), acc);
        }
      } else if (T > 1) {
        if (TH > 0) {
#pragma unroll
          for (int th = 1; th <= TH; ++th) acc = fmaf(w[r + TH - th] - w[r + TH + th], 1.0f / (2.0f * (float)th), acc);
        } else {
          for (int th = 1; th <= theta; ++th) {
            const float lo = s - th >= 0 ? v[(s - th) * F + f] : 0.f;
            const float hi = s + th < T ? v[(s + th) * F + f] : 0.f;
            acc = fmaf(lo - hi, 1.0f / (2.0f * (float)th), acc);
          }
        }
      }
      out[s * F + f] = base[s * F + f] + acc;
    }
  }
}

template <int TH>
__global__ void __launch_bounds__(DELTA_THREADS) delta_bwd_gather_kernel(const float* __restrict__ gy, int ldgy,
                                                                         float* __restrict__ gx, int ldgx, int N, int T,
                                                                         int F, int theta, int upc, int accumulate) {
  extern __shared__ __align__(16) float sm[];
  const int TF = T * F;
  float* s_a = sm;                      // g_a
  float* s_d = sm + (size_t)upc * TF;   // g_d, then u = g_d + D^T g_a
  float* s_x = s_d + (size_t)upc * TF;  // g_x, then g_x + D^T u
  const int tid = threadIdx.x;
  const int F3 = 3 * F;
  const bool vec = (F % 4 == 0) && (ldgy % 4 == 0) && ((reinterpret_cast<uintptr_t>(gy) & 15) == 0);
  for (int u0 = blockIdx.x * upc; u0 < N; u0 += gridDim.x * upc) {
    const int nu = min(upc, N - u0);
    const int rows = nu * T;
    if (vec) {
      const int F34 = F3 / 4;
      for (int i = tid; i < rows * F34; i += DELTA_THREADS) {
        const int r = i / F34, c = (i - r * F34) * 4;
        const float4 v = __ldcs(reinterpret_cast<const float4*>(gy + ((size_t)u0 * T + r) * ldgy + c));
        float* dst = c < F ? s_x + r * F + c : (c < 2 * F ? s_d + r * F + (c - F) : s_a + r * F + (c - 2 * F));
        *reinterpret_cast<float4*>(dst) = v;
      }
    } else {
      for (int i = tid; i < rows * F3; i += DELTA_THREADS) {
        const int r = i / F3, c = i - r * F3;
        const float v = __ldcs(gy + ((size_t)u0 * T + r) * ldgy + c);
        if (c < F) s_x[r * F + c] = v;
        else if (c < 2 * F) s_d[r * F + (c - F)] = v;
        else s_a[r * F + (c - 2 * F)] = v;
      }
    }
    __syncthreads();
    for (int u = 0; u < nu; ++u) delta_T_gather<TH>(s_a + u * TF, s_d + u * TF, s_d + u * TF, T, F, theta, tid, DELTA_THREADS);
    __syncthreads();
    for (int u = 0; u < nu; ++u) delta_T_gather<TH>(s_d + u * TF, s_x + u * TF, s_x + u * TF, T, F, theta, tid, DELTA_THREADS);
    __syncthreads();
    for (int i = tid; i < rows * F; i += DELTA_THREADS) {
      const int r = i / F, c = i - r * F;
      float* o = gx + ((size_t)u0 * T + r) * ldgx + c;
      const float v = s_x[r * F + c];
      *o = accumulate ? *o + v : v;
    }
    __syncthreads();
  }
}

// Register-resident column form of the gather (T <= 48, Theta in {1,4,9}): one THREAD owns one (utterance, feature)
// column, zero-padded by Theta on both sides so that every register index is static.  Rows 0 and T-1 collect what the
// forward clamped onto them: row 0 has a static formula; row T-1 sits at a runtime index, so its value is formed as
// sum_t v[t] * wtab[T-1-t] with a small constant table (the index is uniform over the launch: constant-cache broadcast)
// and selected when the row is written.  No shared memory, no barriers.
__constant__ float c_dbw_tab[64];      // wtab[d] = sum_{th = max(d,1)}^{Theta} 1/(2 th) for 0 <= d <= Theta, else 0

template <int TH, int TMAX>
__device__ __forceinline__ void delta_T_col(const float (&v)[TMAX + 2 * TH], float (&y)[TMAX], int T) {
  // v[i] holds frame i - TH (zero outside [0, T)); y[s] = (D^T v)[s] for s < T (garbage beyond)
  float last = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    const int d = T - 1 - t;
    last = fmaf(v[t + TH], (d >= 0 && d <= TH) ? c_dbw_tab[d] : 0.f, last);
  }
#pragma unroll
  for (int s = 0; s < TMAX; ++s) {
    float acc = 0.f;
    if (s == 0) {
      float run = v[TH];
#pragma unroll
      for (int th = 1; th <= TH; ++th) {
        run += v[TH + th];
        acc = fmaf(-run, 1.0f / (2.0f * (float)th), acc);
      }
    } else {
#pragma unroll
      for (int th = 1; th <= TH; ++th) acc = fmaf(v[s + TH - th] - v[s + TH + th], 1.0f / (2.0f * (float)th), acc);
    }
    y[s] = (T > 1) ? ((s == T - 1) ? last : acc) : 0.f;
  }
}

template <int TH, int TMAX>
__global__ void __launch_bounds__(128) delta_bwd_col_kernel(const float* __restrict__ gy, int ldgy, float* __restrict__ gx,
                                                            int ldgx, int N, int T, int F, int accumulate) {
  const long long total = (long long)N * F;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total;
       c += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(c / F), f = (int)(c % F);
    const float* gc = gy + (size_t)n * T * ldgy + f;
    float* oc = gx + (size_t)n * T * ldgx + f;
    float v[TMAX + 2 * TH], y[TMAX], w[TMAX];
    // both gradient columns are requested up front (80 independent loads in flight per thread), g_x right after
#pragma unroll
    for (int i = 0; i < TMAX + 2 * TH; ++i) {
      const int t = i - TH;
      v[i] = (t >= 0 && t < T) ? __ldg(gc + (size_t)t * ldgy + 2 * F) : 0.f;      // g_a
    }
#pragma unroll
    for (int t = 0; t < TMAX; ++t) w[t] = t < T ? __ldg(gc + (size_t)t * ldgy + F) : 0.f;      // g_d
    delta_T_col<TH, TMAX>(v, y, T);
#pragma unroll
    for (int i = 0; i < TMAX + 2 * TH; ++i) {
      const int t = i - TH;
      v[i] = (t >= 0 && t < TMAX && t < T) ? w[t >= 0 && t < TMAX ? t : 0] + y[t >= 0 && t < TMAX ? t : 0] : 0.f;   // u
    }
#pragma unroll
    for (int t = 0; t < TMAX; ++t) w[t] = t < T ? __ldg(gc + (size_t)t * ldgy) : 0.f;          // g_x
    delta_T_col<TH, TMAX>(v, y, T);
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < T) {
        const float o = w[t] + y[t];
        float* dst = oc + (size_t)t * ldgx;
        if (accumulate) *dst += o;
        else __stcs(dst, o);
      }
  }
}

// which Theta the constant table currently holds: ONE flag for every instantiation (a per-template static would let
// <9,24> skip the upload after <4,24> had replaced the table)
static int g_dbw_tab_theta = -1;

template <int TH, int TMAX>
static int launch_delta_bwd_col(const float* gy, int ldgy, float* gx, int ldgx, int N, int T, int F, int accumulate,
                                cudaStream_t st) {
  int& tab_theta = g_dbw_tab_theta;
  if (tab_theta != TH) {
    float tab[64];
    for (int d = 0; d < 64; ++d) {
      float s = 0.f;
      if (d <= TH)
        for (int th = (d > 1 ? d : 1); th <= TH; ++th) s += 1.0f / (2.0f * (float)th);
      tab[d] = s;
    }
    IPAVSR_CUDA(cudaMemcpyToSymbolAsync(c_dbw_tab, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, st));
    tab_theta = TH;
  }
  const long long total = (long long)N * F;
  long long blocks = (total + 127) / 128;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  delta_bwd_col_kernel<TH, TMAX><<<(int)blocks, 128, 0, st>>>(gy, ldgy, gx, ldgx, N, T, F, accumulate);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

template <int TH>
static int dispatch_delta_bwd_col(const float* gy, int ldgy, float* gx, int ldgx, int N, int T, int F, int accumulate,
                                  cudaStream_t st) {
  if (T <= 24) return launch_delta_bwd_col<TH, 24>(gy, ldgy, gx, ldgx, N, T, F, accumulate, st);
  if (T <= 40) return launch_delta_bwd_col<TH, 40>(gy, ldgy, gx, ldgx, N, T, F, accumulate, st);
  return launch_delta_bwd_col<TH, 48>(gy, ldgy, gx, ldgx, N, T, F, accumulate, st);
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

// ---------------------------------------------------------------------------------------------------------
// Bulk-copy kernel (the streaming path for large batches).  One utterance = one contiguous (T x ldx) input tile and
// one contiguous (T x ldy) output tile, so both directions are single TMA bulk copies (cp.async.bulk): a persistent
// CTA keeps a 3-deep ring of input tiles in flight (mbarrier complete_tx) and double-buffers the output tile, whose
// store is issued by one thread and drains asynchronously (bulk_group) while the next utterance is computed.  The
// arithmetic is the same register sliding window over the shared tile as in the general kernel.  Nothing but TMA
// touches global memory: every byte moves in full, aligned, coalesced lines.
// ---------------------------------------------------------------------------------------------------------
constexpr int DB_THREADS = 256;
constexpr int DB_STAGES = 3;
constexpr int DB_RUN = 8;

__device__ __forceinline__ uint32_t db_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void db_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DB_DONE;\n"
      "bra DB_WAIT;\n"
      "DB_DONE:\n"
      "}\n" ::"r"(db_smem(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void db_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(db_smem(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   db_smem(dst)),
               "l"(src), "r"(bytes), "r"(db_smem(bar))
               : "memory");
}
__device__ __forceinline__ void db_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(db_smem(src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int TH, bool EXACT>
__device__ __forceinline__ void db_pass(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int T,
                                        int F, int tid) {
  const int runs = (T + DB_RUN - 1) / DB_RUN;
  const int items = runs * F;
  for (int it = tid; it < items; it += DB_THREADS) {
    const int f = it % F;
    const int t0 = (it / F) * DB_RUN;
    float w[2 * TH + DB_RUN];
#pragma unroll
    for (int i = 0; i < 2 * TH + DB_RUN; ++i) {
      int t = min(max(t0 - TH + i, 0), T - 1);
      w[i] = src[t * lds + f];
    }
#pragma unroll
    for (int r = 0; r < DB_RUN; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int th = 1; th <= TH; ++th) acc = delta_step<EXACT>(acc, w[r + TH + th] - w[r + TH - th], th);
      if (t0 + r < T) dst[(t0 + r) * ldd + f] = acc;
    }
  }
}

// fast (float32) mode, two adjacent features per thread with Blackwell's packed FFMA2:
//   acc += c*x[t+th];  acc -= c*x[t-th]   (2 packed FMAs per theta for 2 features; 64-bit shared loads)
template <int TH, int DB_RUN_X2>
__device__ __forceinline__ void db_pass_x2(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                           int T, int F, int tid) {
  const int runs = (T + DB_RUN_X2 - 1) / DB_RUN_X2;
  const int F2 = F >> 1;
  const int items = runs * F2;
  for (int it = tid; it < items; it += DB_THREADS) {
    const int f = 2 * (it % F2);
    const int t0 = (it / F2) * DB_RUN_X2;
    float2 w[2 * TH + DB_RUN_X2];
#pragma unroll
    for (int i = 0; i < 2 * TH + DB_RUN_X2; ++i) {
      int t = min(max(t0 - TH + i, 0), T - 1);
      w[i] = *reinterpret_cast<const float2*>(src + t * lds + f);
    }
#pragma unroll
    for (int r = 0; r < DB_RUN_X2; ++r) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int th = 1; th <= TH; ++th) {
        const float c = 1.0f / (2.0f * (float)th);
        acc = __ffma2_rn(w[r + TH + th], make_float2(c, c), acc);
        acc = __ffma2_rn(w[r + TH - th], make_float2(-c, -c), acc);
      }
      if (t0 + r < T) *reinterpret_cast<float2*>(dst + (t0 + r) * ldd + f) = acc;
    }
  }
}

template <int TH, bool EXACT, bool X2>
__global__ void __launch_bounds__(DB_THREADS) delta_fwd_bulk_kernel(const float* __restrict__ x, int ldx,
                                                                    float* __restrict__ y, int ldy, int N, int T, int F) {
  extern __shared__ __align__(16) float sm[];
  __shared__ __align__(8) uint64_t full[DB_STAGES];
  const int in_f = T * ldx, out_f = T * ldy;          // floats per tile (multiples of 4)
  float* sin = sm;                                     // [DB_STAGES][in_f]
  float* sout = sm + (size_t)DB_STAGES * in_f;         // [2][out_f]
  const int tid = threadIdx.x;
  const uint32_t in_bytes = (uint32_t)in_f * 4, out_bytes = (uint32_t)out_f * 4;
  if (tid == 0) {
    for (int s = 0; s < DB_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(db_smem(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // padding columns of the output tiles are written once (zeros) and never touched again
  for (int i = tid; i < 2 * out_f; i += DB_THREADS) sout[i] = 0.f;
  __syncthreads();
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_count = first < N ? (N - first + stride - 1) / stride : 0;
  if (tid == 0)
    for (int i = 0; i < DB_STAGES && i < my_count; ++i)
      db_load(sin + (size_t)i * in_f, x + (size_t)(first + (size_t)i * stride) * in_f, in_bytes, &full[i]);
  for (int i = 0; i < my_count; ++i) {
    const int stage = i % DB_STAGES;
    const uint32_t phase = (uint32_t)(i / DB_STAGES) & 1u;
    const int n = first + i * stride;
    float* so = sout + (size_t)(i & 1) * out_f;
    const float* si = sin + (size_t)stage * in_f;
    // the bulk store that last read this output buffer (utterance i-2) must have finished reading shared memory
    if (tid == 0 && i >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    db_mbar_wait(&full[stage], phase);
    __syncthreads();
    // x block and d block
    for (int j = tid; j < T * F; j += DB_THREADS) {
      const int t = j / F, f = j - t * F;
      so[t * ldy + f] = si[t * ldx + f];
    }
    // short runs when there would otherwise be too few (feature pair, run) items for the 256 threads
    const bool short_runs = ((T + DB_RUN - 1) / DB_RUN) * (F >> 1) < 100;
    if (X2) {
      if (short_runs) db_pass_x2<TH, 4>(si, ldx, so + F, ldy, T, F, tid);
      else db_pass_x2<TH, 8>(si, ldx, so + F, ldy, T, F, tid);
    }
    else db_pass<TH, EXACT>(si, ldx, so + F, ldy, T, F, tid);
    __syncthreads();
    // the input stage is free again: prefetch utterance i + DB_STAGES into it
    if (tid == 0 && i + DB_STAGES < my_count)
      db_load(sin + (size_t)stage * in_f, x + (size_t)(first + (size_t)(i + DB_STAGES) * stride) * in_f, in_bytes,
              &full[stage]);
    if (X2) {
      if (short_runs) db_pass_x2<TH, 4>(so + F, ldy, so + 2 * F, ldy, T, F, tid);
      else db_pass_x2<TH, 8>(so + F, ldy, so + 2 * F, ldy, T, F, tid);
    }
    else db_pass<TH, EXACT>(so + F, ldy, so + 2 * F, ldy, T, F, tid);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) db_store(y + (size_t)n * out_f, so, out_bytes);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int TH, bool EXACT, bool X2>
static int launch_delta_bulk_k(const float* x, int ldx, float* y, int ldy, int N, int T, int F, cudaStream_t st);

template <int TH, bool EXACT>
static int launch_delta_bulk(const float* x, int ldx, float* y, int ldy, int N, int T, int F, cudaStream_t st) {
  // packed two-feature path: float32 mode, even F (so that F, 2F and every row start are 8-byte aligned)
  if (!EXACT && F % 2 == 0) return launch_delta_bulk_k<TH, false, true>(x, ldx, y, ldy, N, T, F, st);
  return launch_delta_bulk_k<TH, EXACT, false>(x, ldx, y, ldy, N, T, F, st);
}

template <int TH, bool EXACT, bool X2>
static int launch_delta_bulk_k(const float* x, int ldx, float* y, int ldy, int N, int T, int F, cudaStream_t st) {
  const size_t smem = ((size_t)DB_STAGES * T * ldx + (size_t)2 * T * ldy) * sizeof(float);
  auto k = delta_fwd_bulk_kernel<TH, EXACT, X2>;
  IPAVSR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((220 * 1024) / (smem + 2048));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 6) per_sm = 6;
  int grid = sm_count() * per_sm;
  if (grid > N) grid = N;
  k<<<grid, DB_THREADS, smem, st>>>(x, ldx, y, ldy, N, T, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
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

int delta_fwd_stream(const float* x, int ldx, float* y, int ldy, int N, int T, int F, int theta, int exact,
                     cudaStream_t st);
int delta_bwd_stream(const float* gy, int ldgy, float* gx, int ldgx, int N, int T, int F, int theta, int accumulate,
                     cudaStream_t st);

}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

int ipavsr_delta_fwd(const float* x, int ldx, float* y, int ldy, int N, int T, int F, int theta, int exact,
                     void* stream) {
  IPAVSR_CHECK_ARG(x && y && N >= 0 && T >= 1 && F >= 1 && theta >= 0, "bad arguments");
  IPAVSR_CHECK_ARG(ldx >= F && ldy >= 3 * F, "leading dimensions too small");
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool tiles_ok = (ldx % 4 == 0) && (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ldy - 3 * F < 8 &&
                        ((size_t)DB_STAGES * T * ldx + (size_t)2 * T * ldy) * sizeof(float) <= 200 * 1024;
  static int force_path = -1;   // IPAVSR_DELTA_PATH=stream|bulk|col|general (benchmarking aid); default: automatic
  if (force_path < 0) {
    const char* e = getenv("IPAVSR_DELTA_PATH");
    force_path = !e ? 0 : (e[0] == 's' ? 4 : (e[0] == 'b' ? 1 : (e[0] == 'c' ? 2 : 3)));
  }
  if (force_path == 0 || force_path == 4) {
    // streaming kernel (delta_stream.cu): even F, T <= 48, Theta in {1,4,9} -- every shipped configuration
    const int rc = delta_fwd_stream(x, ldx, y, ldy, N, T, F, theta, exact, st);
    if (rc != -1) return rc;
  }
  // exact mode with a wide window is FP64-pipe bound: the register-resident column kernel is the faster one there
  const bool prefer_col = exact && theta >= 4 && T <= 48;
  if (tiles_ok && (theta == 1 || theta == 4 || theta == 9) && N >= 64 &&
      ((force_path == 0 && !prefer_col) || force_path == 1)) {
    // TMA bulk-copy streaming kernel
    if (theta == 1) return exact ? launch_delta_bulk<1, true>(x, ldx, y, ldy, N, T, F, st)
                                 : launch_delta_bulk<1, false>(x, ldx, y, ldy, N, T, F, st);
    if (theta == 4) return exact ? launch_delta_bulk<4, true>(x, ldx, y, ldy, N, T, F, st)
                                 : launch_delta_bulk<4, false>(x, ldx, y, ldy, N, T, F, st);
    return exact ? launch_delta_bulk<9, true>(x, ldx, y, ldy, N, T, F, st)
                 : launch_delta_bulk<9, false>(x, ldx, y, ldy, N, T, F, st);
  }
  if (T <= 48 && (theta == 1 || theta == 4 || theta == 9) && force_path != 3) {
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
  static int use_scatter = -1;          // IPAVSR_DELTA_BWD=scatter|tile|col selects the earlier kernels
  if (use_scatter < 0) {
    const char* e = getenv("IPAVSR_DELTA_BWD");
    use_scatter = (e && e[0] == 's') ? 1 : ((e && e[0] == 't') ? 2 : ((e && e[0] == 'c') ? 3 : 0));
  }
  if (use_scatter == 0) {
    const int rc = delta_bwd_stream(gy, ldgy, gx, ldgx, N, T, F, theta, accumulate, st);
    if (rc != -1) return rc;
  }
  if (use_scatter == 3) use_scatter = 0;
  if (use_scatter == 0 && T <= 48 && (theta == 1 || theta == 4 || theta == 9)) {
    if (theta == 1) return dispatch_delta_bwd_col<1>(gy, ldgy, gx, ldgx, N, T, F, accumulate, st);
    if (theta == 4) return dispatch_delta_bwd_col<4>(gy, ldgy, gx, ldgx, N, T, F, accumulate, st);
    return dispatch_delta_bwd_col<9>(gy, ldgy, gx, ldgx, N, T, F, accumulate, st);
  }
  if (use_scatter == 1) {
    if (smem > 48 * 1024)
      IPAVSR_CUDA(cudaFuncSetAttribute(delta_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    delta_bwd_kernel<<<grid, DELTA_THREADS, smem, st>>>(gy, ldgy, gx, ldgx, N, T, F, theta, upc, accumulate);
    IPAVSR_LAUNCH_CHECK();
    return IPAVSR_OK;
  }
#define IPAVSR_DBW_CASE(TH)                                                                                        \
  do {                                                                                                            \
    auto k = delta_bwd_gather_kernel<TH>;                                                                         \
    if (smem > 48 * 1024) IPAVSR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k<<<grid, DELTA_THREADS, smem, st>>>(gy, ldgy, gx, ldgx, N, T, F, theta, upc, accumulate);                     \
  } while (0)
  switch (theta) {
    case 1: IPAVSR_DBW_CASE(1); break;
    case 4: IPAVSR_DBW_CASE(4); break;
    case 9: IPAVSR_DBW_CASE(9); break;
    default: IPAVSR_DBW_CASE(0); break;
  }
#undef IPAVSR_DBW_CASE
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // extern "C"
