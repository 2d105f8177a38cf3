// DeltaLayer streaming kernels (the production path for even F, T <= 48, Theta in {1,4,9}).
//
// Same operator as delta.cu (custom/layers.py:105-121 -> utils/signal.py:7-80), organised for HBM throughput:
//   * loads: a persistent CTA stages tiles of G whole utterances (one contiguous G*T*ld block of the input) in shared
//     memory with ONE cp.async.bulk per tile on a 2-stage mbarrier ring, so every SM keeps >= one tile (40-120 KB) of
//     reads in flight while it computes the previous one -- the register-column kernels could not (their loads sit
//     behind their own arithmetic: 18-24 % warps active, 46 % of HBM);
//   * compute: one THREAD owns a PAIR of adjacent feature columns of one utterance (float2, Blackwell packed FFMA2) and
//     streams down the T frames once: a register window of 2*Theta+1 frames of x yields d[j], a second window of d
//     yields a[j-Theta] in the same step (backward: g_a -> w = g_d + D^T g_a -> g_x + D^T w).  The loop is fully
//     unrolled, so the windows are register renames; T is a run-time bound inside a TMAX-step template;
//   * stores: straight from registers, 8 bytes per lane, lanes of a warp on consecutive feature pairs of a row, the
//     three row segments [x | d | a] written in the same step (full lines by the time L2 writes them back).
// No block-wide barrier inside a tile: one __syncthreads per tile hands the stage back to the TMA.
// Algorithmic traffic: 16*F bytes/frame in both directions (forward reads 4F writes 12F; backward the reverse).
#include <stdio.h>
#include <stdlib.h>
#include <utility>
#include "common.cuh"
#include "delta_common.cuh"

namespace ipavsr {

namespace {

__device__ __forceinline__ uint32_t ds_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ds_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DS_DONE;\n"
      "bra DS_WAIT;\n"
      "DS_DONE:\n"
      "}\n" ::"r"(ds_smem(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void ds_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ds_smem(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ds_smem(dst)),
               "l"(src), "r"(bytes), "r"(ds_smem(bar))
               : "memory");
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ void stg2(float* p, float2 v) { __stcs(reinterpret_cast<float2*>(p), v); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

// ---- bit-exact mode: the reference's float64 chain (delta_common.cuh) with fewer conversions --------------------
// The conversion instructions (F2F, quarter-rate XU pipe) bound the exact kernels, so:
//  * between two consecutive non-power-of-two thetas the float32 round stays in float64: RN24(s) = (s + B) - B with
//    B = sign(s) * 2^(exponent(s) + 29) (the sum's ulp is then the float32 ulp of s, ties to even on the same bit),
//    which replaces a float64->float32->float64 round trip by two DADDs and two integer operations;
//    which replaces a float64->float32->float64 round trip by two DADDs and two integer operations (15 -> 11
//    conversions per output).  Widening the differences with integer instructions instead of F2F was tried and is not
//    worth it: the special-case fallback keeps the F2F in the instruction stream and the extra ALU work makes the kernel
//    issue-bound.  Doing the sandwiched power-of-two thetas (4, 8) in float64 as well (9 conversions) is slower too
//    (0.50 vs 0.45 ms at F = 50): the extra FP64 operations cost more than the two conversions they save.
__device__ __forceinline__ double round24(double s) {
  const int hi = __double2hiint(s);
  const double big = __hiloint2double((hi & (int)0xfff00000) + (29 << 20), 0);
  return (s + big) - big;
}
__device__ __forceinline__ constexpr bool ds_pow2(int th) { return (th & (th - 1)) == 0; }

template <int TH, bool LOW>
__device__ __forceinline__ float exact_chain(const float2 (&w)[2 * TH + 1]) {
  float af = 0.f;
  double ad = 0.0;
  bool in_double = false;       // compile-time after unrolling
#pragma unroll
  for (int th = 1; th <= TH; ++th) {
    const float diff = LOW ? (w[TH + th].x - w[TH - th].x) : (w[TH + th].y - w[TH - th].y);
    if (ds_pow2(th)) {
      if (in_double) {
        af = (float)ad;
        in_double = false;
      }
      af = fmaf(diff, 1.0f / (2.0f * (float)th), af);
    } else {
      if (!in_double) ad = (double)af;
      const double c = 2.0 * (double)th;
      const double r = 1.0 / c;
      const double d = (double)diff;
      double q = d * r;
      const double rem = fma(-q, c, d);
      q = fma(rem, r, q);
      const double sum = ad + q;
      if (th + 1 <= TH && !ds_pow2(th + 1)) {
        ad = round24(sum);
        in_double = true;
      } else {
        af = (float)sum;
        in_double = false;
      }
    }
  }
  return af;
}

// d = sum_th (w[TH+th] - w[TH-th]) / (2 th) for a pair of columns
template <int TH, bool EXACT>
__device__ __forceinline__ float2 taps_fwd(const float2 (&w)[2 * TH + 1]) {
  if (EXACT) {
    return f2(exact_chain<TH, true>(w), exact_chain<TH, false>(w));
  } else {
    // three interleaved partial sums: 6 independent FFMA2 chains per thread (d and a) cover the FMA latency
    float2 acc[3] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
    for (int th = 1; th <= TH; ++th) {
      const float c = 1.0f / (2.0f * (float)th);
      acc[th % 3] = __ffma2_rn(w[TH + th], f2(c, c), acc[th % 3]);
      acc[th % 3] = __ffma2_rn(w[TH - th], f2(-c, -c), acc[th % 3]);
    }
    if (TH == 1) return acc[1];
    if (TH == 2) return add2(acc[1], acc[2]);
    return add2(add2(acc[1], acc[2]), acc[0]);
  }
}

// interior part of the adjoint: sum_th (w[TH-th] - w[TH+th]) / (2 th), window zero-extended outside [0, T)
template <int TH>
__device__ __forceinline__ float2 taps_bwd(const float2 (&w)[2 * TH + 1], float2 base) {
  float2 acc[3] = {base, f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
  for (int th = 1; th <= TH; ++th) {
    const float c = 1.0f / (2.0f * (float)th);
    acc[th % 3] = __ffma2_rn(w[TH - th], f2(c, c), acc[th % 3]);
    acc[th % 3] = __ffma2_rn(w[TH + th], f2(-c, -c), acc[th % 3]);
  }
  if (TH == 1) return add2(acc[0], acc[1]);
  return add2(add2(acc[1], acc[2]), acc[0]);
}
// what the forward clamped onto row 0 beyond the interior formula:  - sum_th c_th sum_{t=0}^{th-1} v[t],  v[t] = w[TH+t]
template <int TH>
__device__ __forceinline__ float2 edge_low(const float2 (&w)[2 * TH + 1]) {
  float2 run = f2(0.f, 0.f), acc = f2(0.f, 0.f);
#pragma unroll
  for (int th = 1; th <= TH; ++th) {
    run = add2(run, w[TH + th - 1]);
    const float c = -1.0f / (2.0f * (float)th);
    acc = __ffma2_rn(run, f2(c, c), acc);
  }
  return acc;
}
constexpr int DS_R = 4;            // rows per output chunk (one bulk store per utterance per chunk)
constexpr int DS_TMAX = 48;        // longest utterance the run-time-T instantiation takes

struct EdgeTab {                   // v[k] = sum_{th=k+1}^{Theta} 1/(2 th) for 0 <= k < Theta, else 0
  float v[64];
};

__device__ __forceinline__ float2 lds2s(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts2s(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}

// Output side of a tile.  STAGED: rows are assembled in a 2-deep ring of shared-memory chunks ([G][DS_R][ldo] floats
// each) and leave as whole contiguous row blocks through cp.async.bulk (full lines; the alignment-padding columns of the
// rows are written as zeros).  Otherwise: 8-byte stores straight from registers (row pitches wider than the padding,
// accumulate mode).
template <bool STAGED>
struct OutSink {
  float* gbase;       // global: row 0 of this lane's utterance, at this lane's column pair
  float* gtile;       // global: row 0 of the tile's first utterance, column 0
  uint32_t ring;      // shared address of the chunk ring
  uint32_t lane_off;  // byte offset of this lane's (utterance, column pair) inside a chunk
  uint32_t cur;       // shared address of this lane's position in the current chunk
  uint32_t chunk_b, ldo4, seg4;     // bytes per chunk / per row / per row segment (F floats)
  int ldo, G, nu;
  uint32_t cc;        // running chunk counter of this CTA (selects the ring slot)
  bool active;

  __device__ __forceinline__ void begin_tile() { cur = ring + (cc & 1u) * chunk_b + lane_off; }
  // forward row: [x | d | a]
  __device__ __forceinline__ void put3(int i, float2 x, float2 d, float2 a) const {
    if (!active) return;
    if (STAGED) {
      const uint32_t rp = cur + (uint32_t)(i % DS_R) * ldo4;
      sts2s(rp, x);
      sts2s(rp + seg4, d);
      sts2s(rp + 2 * seg4, a);
    } else {
      float* rp = gbase + (size_t)i * ldo;
      stg2(rp, x);
      stg2(rp + (seg4 >> 2), d);
      stg2(rp + (seg4 >> 1), a);
    }
  }
  // backward row
  __device__ __forceinline__ void put1(int i, float2 o, int accumulate) const {
    if (!active) return;
    if (STAGED) {
      sts2s(cur + (uint32_t)(i % DS_R) * ldo4, o);
    } else {
      float2* rp = reinterpret_cast<float2*>(gbase + (size_t)i * ldo);
      if (accumulate) *rp = add2(*rp, o);
      else __stcs(rp, o);
    }
  }
  // after row i of T: if it closes a chunk, hand the chunk to the TMA
  __device__ __forceinline__ void row_done(int i, int T) {
    if (!STAGED) return;
    if ((i % DS_R) != DS_R - 1 && i != T - 1) return;
    const int r0 = (i / DS_R) * DS_R;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // the bulk stores this thread issued earlier (previous chunk: the other ring slot's last reader) are done reading
    if ((int)threadIdx.x < G) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    if ((int)threadIdx.x < nu) {
      const uint32_t u = threadIdx.x;
      const uint32_t bytes = (uint32_t)min(DS_R, T - r0) * ldo4;
      const uint32_t s = ring + (cc & 1u) * chunk_b + u * (uint32_t)DS_R * ldo4;
      float* d = gtile + ((size_t)u * T + r0) * ldo;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(s), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    ++cc;
    cur = ring + (cc & 1u) * chunk_b + lane_off;
  }
};

// ---- forward: one pair of columns, frames 0..T-1; src = shared address of the staged column (row pitch lds4 bytes).
//      Row i = j - Theta is complete at step j: x[i] and d[i] are still in the windows, a[i] is formed from
//      d[i-Theta .. i+Theta = j].  TS > 0: T is the compile-time constant TS (every bound below folds). ----
template <int TH, int TS, bool EXACT, bool STAGED>
__device__ __forceinline__ void stream_fwd_column(uint32_t src, uint32_t lds4, OutSink<STAGED>& out, int Trt) {
  const int T = TS > 0 ? TS : Trt;
  constexpr int TMAX = TS > 0 ? TS : DS_TMAX;
  float2 xw[2 * TH + 1], dw[2 * TH + 1];
  // window for step j holds x[clamp(j - TH + k)], k = 0..2TH; the newest element is loaded inside the step
  const float2 x0 = lds2s(src);
#pragma unroll
  for (int k = 0; k <= TH; ++k) xw[k] = x0;
#pragma unroll
  for (int k = 1; k < TH; ++k) xw[TH + k] = lds2s(src + (uint32_t)min(k, T - 1) * lds4);
  xw[2 * TH] = x0;
#pragma unroll
  for (int k = 0; k <= 2 * TH; ++k) dw[k] = f2(0.f, 0.f);
  float2 dlast = f2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < TMAX + TH; ++j) {
    if (j >= T + TH) break;
    if (j + TH <= T - 1) xw[2 * TH] = lds2s(src + (uint32_t)(j + TH) * lds4);
    else if (j + TH == 1 || TH == 1) xw[2 * TH] = lds2s(src + (uint32_t)(T - 1) * lds4);
    else xw[2 * TH] = xw[2 * TH - 1];                 // edge replication of x beyond T-1 (the previous newest frame)
    float2 dj;
    if (j < T) {
      dj = taps_fwd<TH, EXACT>(xw);
      dlast = dj;
    } else {
      dj = dlast;                       // edge replication of d beyond T-1
    }
    if (j == 0) {
#pragma unroll
      for (int k = 0; k <= 2 * TH; ++k) dw[k] = dj;     // ... and below 0
    } else {
      dw[2 * TH] = dj;
    }
    const int i = j - TH;
    if (i >= 0) {
      const float2 ai = taps_fwd<TH, EXACT>(dw);
      out.put3(i, xw[0], dw[TH], ai);
      out.row_done(i, T);
    }
#pragma unroll
    for (int k = 0; k < 2 * TH; ++k) {
      xw[k] = xw[k + 1];
      dw[k] = dw[k + 1];
    }
  }
}

// ---- backward: src = staged [g_x | g_d | g_a] column; row i of gx is complete at step j = i + Theta.  The steps are
//      instantiated one by one (fold over an index sequence): left to "#pragma unroll" the compiler keeps this loop
//      rolled and shifts the two windows with register moves. ----
template <int TH>
struct BwdCol {
  float2 aw[2 * TH + 1], ww[2 * TH + 1];
  float2 corr_w, corr_x;     // what the forward clamped onto row T-1, collected as the operands go by
  uint32_t src, gd, ga, lds4;
  int T, accumulate;
};

template <int TH, int TS, bool STAGED, int J>
__device__ __forceinline__ void bwd_step(BwdCol<TH>& c, OutSink<STAGED>& out, const EdgeTab& tab) {
  const int T = TS > 0 ? TS : c.T;
  if (J >= T + TH) return;
  const float2 zero = f2(0.f, 0.f);
  const bool edges = T > 1;             // T == 1: the forward operator is identically zero
  c.aw[2 * TH] = zero;
  if (J + TH < T) {
    c.aw[2 * TH] = lds2s(c.ga + (uint32_t)(J + TH) * c.lds4);
    if (T - 1 - (J + TH) < TH) {
      const float k = tab.v[T - 1 - (J + TH)];
      c.corr_w = __ffma2_rn(c.aw[2 * TH], f2(k, k), c.corr_w);
    }
  }
  float2 wj = zero;
  if (J < T) {
    wj = lds2s(c.gd + (uint32_t)J * c.lds4);
    if (edges) {
      wj = taps_bwd<TH>(c.aw, wj);
      if (J == 0) wj = add2(wj, edge_low<TH>(c.aw));
      if (J == T - 1) wj = add2(wj, c.corr_w);
      if (T - 1 - J < TH) {
        const float k = tab.v[T - 1 - J];
        c.corr_x = __ffma2_rn(wj, f2(k, k), c.corr_x);
      }
    }
  }
  c.ww[2 * TH] = wj;
  constexpr int I = J - TH;
  if (I >= 0) {
    float2 o = lds2s(c.src + (uint32_t)(I >= 0 ? I : 0) * c.lds4);
    if (edges) {
      o = taps_bwd<TH>(c.ww, o);
      if (I == 0) o = add2(o, edge_low<TH>(c.ww));
      if (I == T - 1) o = add2(o, c.corr_x);
    }
    out.put1(I >= 0 ? I : 0, o, c.accumulate);
    out.row_done(I >= 0 ? I : 0, T);
  }
#pragma unroll
  for (int k = 0; k < 2 * TH; ++k) {
    c.aw[k] = c.aw[k + 1];
    c.ww[k] = c.ww[k + 1];
  }
}

template <int TH, int TS, bool STAGED, int... Js>
__device__ __forceinline__ void bwd_steps(BwdCol<TH>& c, OutSink<STAGED>& out, const EdgeTab& tab,
                                          std::integer_sequence<int, Js...>) {
  (bwd_step<TH, TS, STAGED, Js>(c, out, tab), ...);
}

// run-time T: the plain loop (bounds are uniform branches)
template <int TH, bool STAGED>
__device__ __forceinline__ void stream_bwd_column_rt(uint32_t src, uint32_t lds4, uint32_t seg4, OutSink<STAGED>& out,
                                                     int T, int accumulate, const EdgeTab& tab) {
  float2 aw[2 * TH + 1], ww[2 * TH + 1];
  const float2 zero = f2(0.f, 0.f);
  const uint32_t ga = src + 2 * seg4;
  const uint32_t gd = src + seg4;
  float2 corr_w = zero, corr_x = zero;
#pragma unroll
  for (int k = 0; k <= 2 * TH; ++k) {
    aw[k] = zero;
    ww[k] = zero;
  }
#pragma unroll
  for (int k = 0; k < TH; ++k) {
    if (k < T) {
      aw[TH + k] = lds2s(ga + (uint32_t)k * lds4);
      const float c = tab.v[T - 1 - k];
      corr_w = __ffma2_rn(aw[TH + k], f2(c, c), corr_w);
    }
  }
  const bool edges = T > 1;
#pragma unroll
  for (int j = 0; j < DS_TMAX + TH; ++j) {
    if (j >= T + TH) break;
    aw[2 * TH] = zero;
    if (j + TH < T) {
      aw[2 * TH] = lds2s(ga + (uint32_t)(j + TH) * lds4);
      const float c = tab.v[T - 1 - (j + TH)];
      corr_w = __ffma2_rn(aw[2 * TH], f2(c, c), corr_w);
    }
    float2 wj = zero;
    if (j < T) {
      wj = lds2s(gd + (uint32_t)j * lds4);
      if (edges) {
        wj = taps_bwd<TH>(aw, wj);
        if (j == 0) wj = add2(wj, edge_low<TH>(aw));
        if (j == T - 1) wj = add2(wj, corr_w);
        const float c = tab.v[T - 1 - j];
        corr_x = __ffma2_rn(wj, f2(c, c), corr_x);
      }
    }
    ww[2 * TH] = wj;
    const int i = j - TH;
    if (i >= 0) {
      float2 o = lds2s(src + (uint32_t)i * lds4);
      if (edges) {
        o = taps_bwd<TH>(ww, o);
        if (i == 0) o = add2(o, edge_low<TH>(ww));
        if (i == T - 1) o = add2(o, corr_x);
      }
      out.put1(i, o, accumulate);
      out.row_done(i, T);
    }
#pragma unroll
    for (int k = 0; k < 2 * TH; ++k) {
      aw[k] = aw[k + 1];
      ww[k] = ww[k + 1];
    }
  }
}

template <int TH, int TS, bool STAGED>
__device__ __forceinline__ void stream_bwd_column(uint32_t src, uint32_t lds4, uint32_t seg4, OutSink<STAGED>& out,
                                                  int Trt, int accumulate, const EdgeTab& tab) {
  if constexpr (TS == 0) {
    stream_bwd_column_rt<TH, STAGED>(src, lds4, seg4, out, Trt, accumulate, tab);
  } else {
    constexpr int T = TS;
    BwdCol<TH> c;
    const float2 zero = f2(0.f, 0.f);
    c.src = src;
    c.gd = src + seg4;
    c.ga = src + 2 * seg4;
    c.lds4 = lds4;
    c.T = T;
    c.accumulate = accumulate;
    c.corr_w = zero;
    c.corr_x = zero;
#pragma unroll
    for (int k = 0; k <= 2 * TH; ++k) {
      c.aw[k] = zero;
      c.ww[k] = zero;
    }
#pragma unroll
    for (int k = 0; k < TH; ++k) {
      if (k < T) {
        c.aw[TH + k] = lds2s(c.ga + (uint32_t)k * lds4);
        if (T - 1 - k < TH) {
          const float w = tab.v[T - 1 - k];
          c.corr_w = __ffma2_rn(c.aw[TH + k], f2(w, w), c.corr_w);
        }
      }
    }
    bwd_steps<TH, TS, STAGED>(c, out, tab, std::make_integer_sequence<int, TS + TH>{});
  }
}

// MODE 0: forward fast, 1: forward exact, 2: backward.  TS > 0: instantiation for exactly T == TS frames.
template <int TH, int TS, int MODE, bool STAGED>
__global__ void __launch_bounds__(256, MODE == 1 ? 1 : 2)
    delta_stream_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, int N, int Trt, int F,
                        int G, int nstages, int accumulate, const __grid_constant__ EdgeTab tab) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t full[2];
  const int T = TS > 0 ? TS : Trt;
  const int tid = threadIdx.x;
  const int F2 = F >> 1;
  const int utt_f = T * ldi;                  // floats of one staged utterance
  const int tile_f = G * utt_f;
  const int g = tid / F2, fp = tid - g * F2;
  const int ntiles = (N + G - 1) / G;
  float* ring = sm + (size_t)nstages * tile_f;
  OutSink<STAGED> sink;
  sink.ring = ds_smem(ring);
  sink.ldo4 = (uint32_t)ldo * 4u;
  sink.chunk_b = (uint32_t)(G * DS_R) * sink.ldo4;
  sink.seg4 = (uint32_t)F * 4u;
  sink.ldo = ldo;
  sink.G = G;
  sink.cc = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ds_smem(&full[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ds_smem(&full[1])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (STAGED)       // the padding columns of the chunk rows are zero and stay zero
    for (int i = tid; i < 2 * G * DS_R * ldo; i += blockDim.x) ring[i] = 0.f;
  __syncthreads();
  const int first = blockIdx.x, stride = gridDim.x;
  const int my_count = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
  auto issue = [&](int i) {
    const int n0 = (first + i * stride) * G;
    const int nu = min(G, N - n0);
    const int s = nstages == 2 ? (i & 1) : 0;
    ds_load(sm + (size_t)s * tile_f, in + (size_t)n0 * utt_f, (uint32_t)nu * (uint32_t)utt_f * 4u, &full[s]);
  };
  if (tid == 0) {
    if (my_count > 0) issue(0);
    if (my_count > 1 && nstages == 2) issue(1);
  }
  for (int i = 0; i < my_count; ++i) {
    const int stage = nstages == 2 ? (i & 1) : 0;
    const uint32_t phase = (uint32_t)(nstages == 2 ? (i >> 1) : i) & 1u;
    const int n0 = (first + i * stride) * G;
    const int nu = min(G, N - n0);
    ds_mbar_wait(&full[stage], phase);
    // lanes without a column (g >= nu) walk utterance 0 of the tile with their stores disabled: the chunk hand-over
    // inside the column loop is a block-wide barrier
    sink.active = g < nu;
    sink.nu = nu;
    const int gg = sink.active ? g : 0;
    const int col = sink.active ? 2 * fp : 0;
    sink.lane_off = (uint32_t)(gg * DS_R) * sink.ldo4 + (uint32_t)col * 4u;
    sink.gtile = out + (size_t)n0 * T * ldo;
    sink.gbase = sink.gtile + (size_t)gg * T * ldo + col;
    sink.begin_tile();
    const uint32_t src = ds_smem(sm + (size_t)stage * tile_f + (size_t)gg * utt_f + col);
    if ((tid & ~31) >= nu * F2) {
      // a warp with no column at all only keeps the chunk hand-over barriers company
      if (STAGED)
        for (int r = DS_R - 1; r < T + DS_R - 1; r += DS_R) sink.row_done(r < T ? r : T - 1, T);
    } else if (MODE == 2) {
      stream_bwd_column<TH, TS, STAGED>(src, (uint32_t)ldi * 4u, sink.seg4, sink, T, accumulate, tab);
    } else {
      stream_fwd_column<TH, TS, MODE == 1, STAGED>(src, (uint32_t)ldi * 4u, sink, T);
    }
    __syncthreads();                          // every thread is done reading this stage
    if (tid == 0 && i + nstages < my_count) issue(i + nstages);
  }
  if (STAGED && tid < G) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct StreamCfg {
  int threads, G, grid, nstages;
  size_t smem;
};

// threads / utterances per tile / input stages: as many whole utterances as the CTA has lanes for (one lane per feature
// pair), one or two input stages of G*T*ldi floats (+ two output chunks of G*DS_R*ldo floats when staged), and as many
// CTAs per SM as the kernel's registers and this shared memory allow (occupancy API).  Score = lane efficiency x
// resident warps (saturating at 16) with a small bonus for the second input stage.
template <typename K>
static bool stream_config(K kernel, int N, int T, int F, int ldi, int ldo, bool staged, StreamCfg* cfg) {
  const int F2 = F / 2;
  static int force_threads = -1, force_stages = -1;
  if (force_threads < 0) {
    const char* e = getenv("IPAVSR_DELTA_STREAM_THREADS");
    force_threads = e ? atoi(e) : 0;
    e = getenv("IPAVSR_DELTA_STREAM_STAGES");
    force_stages = e ? atoi(e) : 0;
  }
  double best = -1.0;
  const int cand[3] = {64, 128, 256};
  for (int ci = 0; ci < 3; ++ci) {
    const int threads = cand[ci];
    if (force_threads > 0 && threads != force_threads) continue;
    for (int nst = 2; nst >= 1; --nst) {
      if (force_stages > 0 && nst != force_stages) continue;
      const int Gmax = threads / F2;
      if (Gmax < 1) continue;
      const size_t utt = ((size_t)nst * T * ldi + (staged ? (size_t)2 * DS_R * ldo : 0)) * sizeof(float);
      const size_t budget = 224 * 1024;
      // fewer utterances per tile can buy another resident CTA: scan down to half the lanes
      for (int G = Gmax; G >= 1 && 2 * G >= Gmax; --G) {
        const size_t smem = G * utt;
        if (smem > budget) continue;
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
          cudaGetLastError();
          continue;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess ||
            per_sm < 1) {
          cudaGetLastError();
          continue;
        }
        const int act_warps = (G * F2 + 31) / 32;       // warps of a CTA that own columns (the others idle at barriers)
        const double eff = (double)G * F2 / (act_warps * 32.0);
        const double warps = (double)per_sm * act_warps;
        const int pipes = per_sm * nst;               // tiles that can be in flight or in compute per SM
        const double overlap = pipes >= 3 ? 1.0 : (pipes == 2 ? 0.85 : 0.5);
        const double score = eff * (warps >= 16.0 ? 1.0 : warps / 16.0) * overlap;
        if (score > best + 1e-9) {
          best = score;
          cfg->threads = threads;
          cfg->G = G;
          cfg->nstages = nst;
          cfg->smem = smem;
          const long long tiles = ((long long)N + G - 1) / G;
          const long long cap = (long long)sm_count() * per_sm;
          cfg->grid = (int)(tiles < cap ? tiles : cap);
        }
      }
    }
  }
  if (best > 0.0 && getenv("IPAVSR_DELTA_STREAM_VERBOSE"))
    fprintf(stderr, "delta_stream: T=%d F=%d threads=%d G=%d stages=%d smem=%zu grid=%d\n", T, F, cfg->threads, cfg->G,
            cfg->nstages, cfg->smem, cfg->grid);
  return best > 0.0;
}

static EdgeTab make_tab(int theta) {
  EdgeTab t;
  for (int k = 0; k < 64; ++k) {
    float s = 0.f;
    for (int th = theta; th >= k + 1; --th) s += 1.0f / (2.0f * (float)th);
    t.v[k] = s;
  }
  return t;
}

// one-entry cache of the chosen configuration per instantiation (the occupancy queries are host-side but not free)
struct CfgKey {
  int N, T, F, ldi, ldo;
  bool operator==(const CfgKey& o) const { return N == o.N && T == o.T && F == o.F && ldi == o.ldi && ldo == o.ldo; }
};

template <int TH, int TS, int MODE, bool STAGED>
static int launch_stream(const float* in, int ldi, float* out, int ldo, int N, int T, int F, int accumulate,
                         cudaStream_t st) {
  auto k = delta_stream_kernel<TH, TS, MODE, STAGED>;
  static CfgKey key = {-1, -1, -1, -1, -1};
  static StreamCfg c;
  static bool ok = false;
  const CfgKey now = {N, T, F, ldi, ldo};
  if (!(key == now)) {
    ok = stream_config(k, N, T, F, ldi, ldo, STAGED, &c);
    key = now;
    if (ok) IPAVSR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  }
  if (!ok) return -1;
  k<<<c.grid, c.threads, c.smem, st>>>(in, ldi, out, ldo, N, T, F, c.G, c.nstages, accumulate, make_tab(TH));
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

template <int TH, int MODE>
static int dispatch_stream(const float* in, int ldi, float* out, int ldo, int N, int T, int F, int accumulate,
                           bool staged, cudaStream_t st) {
#define IPAVSR_DS_T(TS)                                                                                \
  return staged ? launch_stream<TH, TS, MODE, true>(in, ldi, out, ldo, N, T, F, accumulate, st)        \
                : launch_stream<TH, TS, MODE, false>(in, ldi, out, ldo, N, T, F, accumulate, st)
  if (T == 40) IPAVSR_DS_T(40);       // the padded length of every shipped configuration: compile-time bounds
  IPAVSR_DS_T(0);
#undef IPAVSR_DS_T
}

static bool stream_shape_ok(const void* in, int ldi, const void* out, int ldo, int T, int F, int theta) {
  return (theta == 1 || theta == 4 || theta == 9) && T >= 1 && T <= 48 && F >= 2 && F % 2 == 0 && F / 2 <= 256 &&
         ldi % 4 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(out) & 7) == 0;
}
// whole output rows may leave as bulk copies when the pitch only adds alignment padding (which is then zero-filled)
static bool stream_staged_ok(const void* out, int ldo, int cols) {
  static int no_stage = -1;
  if (no_stage < 0) {
    const char* e = getenv("IPAVSR_DELTA_STREAM_DIRECT");
    no_stage = (e && e[0] == '1') ? 1 : 0;
  }
  return !no_stage && ldo % 4 == 0 && ldo - cols < 8 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

}  // namespace

// Returns IPAVSR_OK / an error code, or -1 when the shape is not one this path takes (caller falls through).
int delta_fwd_stream(const float* x, int ldx, float* y, int ldy, int N, int T, int F, int theta, int exact,
                     cudaStream_t st) {
  if (!stream_shape_ok(x, ldx, y, ldy, T, F, theta)) return -1;
  const bool staged = stream_staged_ok(y, ldy, 3 * F);
#define IPAVSR_DS_CASE(TH)                                                                     \
  return exact ? dispatch_stream<TH, 1>(x, ldx, y, ldy, N, T, F, 0, staged, st)                \
               : dispatch_stream<TH, 0>(x, ldx, y, ldy, N, T, F, 0, staged, st)
  if (theta == 1) IPAVSR_DS_CASE(1);
  if (theta == 4) IPAVSR_DS_CASE(4);
  IPAVSR_DS_CASE(9);
#undef IPAVSR_DS_CASE
}

int delta_bwd_stream(const float* gy, int ldgy, float* gx, int ldgx, int N, int T, int F, int theta, int accumulate,
                     cudaStream_t st) {
  if (!stream_shape_ok(gy, ldgy, gx, ldgx, T, F, theta)) return -1;
  const bool staged = !accumulate && stream_staged_ok(gx, ldgx, F);
  if (theta == 1) return dispatch_stream<1, 2>(gy, ldgy, gx, ldgx, N, T, F, accumulate, staged, st);
  if (theta == 4) return dispatch_stream<4, 2>(gy, ldgy, gx, ldgx, N, T, F, accumulate, staged, st);
  return dispatch_stream<9, 2>(gy, ldgy, gx, ldgx, N, T, F, accumulate, staged, st);
}

}  // namespace ipavsr
