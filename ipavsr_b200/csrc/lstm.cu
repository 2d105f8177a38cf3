// Masked Lasagne-LSTM recurrence, forward and BPTT (reference custom/layers.py:10-80 -> lasagne LSTMLayer;
// semantics restated in SURVEY.md Appendix A.3): peepholes added after the gradient-clip node, sigmoid gates,
// tanh cell/output, `switch`-masked state carry, learned initial state, optional time reversal.
//
// The input projection x W_in + b is hoisted out (one ipavsr_gemm over all N*T frames).  What is left is the
// sequential part, T dependent steps of  g = xw[t] + h_{t-1} W_hid  followed by ~30 H elementwise flops.
//
// impl 0 — persistent thread-block-cluster kernel.  A cluster of CS = ceil(H/32) CTAs owns a tile of 32
// utterances for all T steps; CTA r owns hidden units [32r, 32r+32) for all four gates, i.e. the contiguous
// column block [128r, 128r+128) of the gate-interleaved W_hid, which it keeps resident in shared memory
// (H x 128 floats; 125 KB at H=250) for the whole sequence.  Each step every thread accumulates a 4-utterance x
// 4-gate register tile over k (two LDS.128 per 16 FFMA, both conflict-free), applies the cell update in
// registers, and broadcasts its 4 new h values to the h buffers of all CS CTAs through distributed shared
// memory; one cluster barrier per step (double-buffered h) is the only synchronisation.  No global-memory
// round trip for the state, no per-step launch.  When H x 128 floats do not fit (H=500) W_hid is streamed from
// L2 instead (same kernel, W_SMEM=false).  The backward kernel mirrors it: CTA r keeps the transposed slice
// W_hid^T[128r:128r+128, :] resident, computes its partial dh_{t-1} = dg[:, own cols] W^T for all H units and
// reduce-scatters the partials into the owners' shared-memory inboxes (DSMEM), again one barrier per step.
// Judged by per-step latency, not by a roofline.
//
// impl 1 — one GEMM + one elementwise launch per step (the plain form; kept as an in-library cross-check and
// as the path for H > 512).
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace ipavsr {

int gemv_rows(const float* W, int ldw, const float* s, float* out, int rows, int cols, int accumulate, cudaStream_t st);
int gemm_simt(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
              float* C, int ldc, const float* bias, int act, int accumulate, cudaStream_t st);

// ------------------------------------------------------------------------------------------------------
// cell arithmetic shared by both implementations
// ------------------------------------------------------------------------------------------------------
struct CellOut {
  float i, f, cin, o, c, h;
};

__device__ __forceinline__ CellOut lstm_cell_fwd(float gi, float gf, float gc, float go, float c_prev, float h_prev,
                                                 bool m, bool has_peep, float w_ci, float w_cf, float w_co) {
  CellOut r;
  if (has_peep) {
    gi = fmaf(c_prev, w_ci, gi);
    gf = fmaf(c_prev, w_cf, gf);
  }
  r.i = sigmoidf_(gi);
  r.f = sigmoidf_(gf);
  r.cin = tanhf(gc);
  float c_u = r.f * c_prev + r.i * r.cin;
  if (has_peep) go = fmaf(c_u, w_co, go);
  r.o = sigmoidf_(go);
  float h_u = r.o * tanhf(c_u);
  r.c = m ? c_u : c_prev;
  r.h = m ? h_u : h_prev;
  return r;
}

struct CellGrad {
  float dgi, dgf, dgc, dgo;   // clipped, w.r.t. pre-peephole gate pre-activations
  float dc_prev, dh_pass;
  float pci, pcf, pco;        // peephole gradient contributions
};

__device__ __forceinline__ CellGrad lstm_cell_bwd(float dh, float dc, float i, float f, float cin, float o, float c,
                                                  float c_prev, bool m, bool has_peep, float w_ci, float w_cf,
                                                  float w_co, float clip) {
  CellGrad g;
  if (!m) {
    g.dgi = g.dgf = g.dgc = g.dgo = 0.f;
    g.dc_prev = dc;
    g.dh_pass = dh;
    g.pci = g.pcf = g.pco = 0.f;
    return g;
  }
  float tc = tanhf(c);
  float dgo = dh * tc * o * (1.f - o);
  float dcu = dc + dh * o * (1.f - tc * tc);
  if (has_peep) dcu = fmaf(dgo, w_co, dcu);
  float dgi = dcu * cin * i * (1.f - i);
  float dgf = dcu * c_prev * f * (1.f - f);
  float dgc = dcu * i * (1.f - cin * cin);
  g.dc_prev = dcu * f;
  if (has_peep) g.dc_prev += dgi * w_ci + dgf * w_cf;
  g.pci = dgi * c_prev;
  g.pcf = dgf * c_prev;
  g.pco = dgo * c;
  if (clip > 0.f) {
    dgi = fminf(fmaxf(dgi, -clip), clip);
    dgf = fminf(fmaxf(dgf, -clip), clip);
    dgc = fminf(fmaxf(dgc, -clip), clip);
    dgo = fminf(fmaxf(dgo, -clip), clip);
  }
  g.dgi = dgi; g.dgf = dgf; g.dgc = dgc; g.dgo = dgo;
  g.dh_pass = 0.f;
  return g;
}

// ======================================================================================================
// impl 1: step-wise kernels
// ======================================================================================================
__global__ void lstm_step_fwd_kernel(const float* __restrict__ xw, const float* __restrict__ rec /*[N,4H]*/,
                                     const float* __restrict__ peep, const float* __restrict__ c_prev_buf,
                                     const float* __restrict__ h_prev_buf, const float* __restrict__ cell_init,
                                     const float* __restrict__ hid_init, int first,
                                     const uint8_t* __restrict__ mask, float* __restrict__ c_next_buf,
                                     float* __restrict__ h_next_buf, float* __restrict__ out,
                                     float* __restrict__ gates, float* __restrict__ cell, float* __restrict__ hprev,
                                     int N, int T, int H, int ldh, int t) {
  const int total = N * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int n = idx / H, u = idx % H;
    size_t row = (size_t)n * T + t;
    float4 x4 = *reinterpret_cast<const float4*>(xw + row * 4 * H + 4 * u);
    float4 r4 = first ? make_float4(0, 0, 0, 0) : *reinterpret_cast<const float4*>(rec + (size_t)n * 4 * H + 4 * u);
    float c_prev = first ? cell_init[u] : c_prev_buf[idx];
    float h_prev = first ? hid_init[u] : h_prev_buf[idx];
    bool m = mask[row] != 0;
    bool hp = peep != nullptr;
    CellOut r = lstm_cell_fwd(x4.x + r4.x, x4.y + r4.y, x4.z + r4.z, x4.w + r4.w, c_prev, h_prev, m, hp,
                              hp ? peep[u] : 0.f, hp ? peep[H + u] : 0.f, hp ? peep[2 * H + u] : 0.f);
    c_next_buf[idx] = r.c;
    h_next_buf[idx] = r.h;
    out[row * ldh + u] = r.h;
    if (gates) *reinterpret_cast<float4*>(gates + row * 4 * H + 4 * u) = make_float4(r.i, r.f, r.cin, r.o);
    if (cell) cell[row * H + u] = r.c;
    if (hprev) hprev[row * ldh + u] = h_prev;
  }
}

// recurrent term needs h_prev W_hid at the very first step too when hid_init != 0: handled by seeding h buffer
__global__ void lstm_seed_kernel(const float* __restrict__ cell_init, const float* __restrict__ hid_init,
                                 float* __restrict__ cbuf, float* __restrict__ hbuf, int N, int H) {
  const int total = N * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int u = idx % H;
    if (cbuf) cbuf[idx] = cell_init ? cell_init[u] : 0.f;
    if (hbuf) hbuf[idx] = hid_init ? hid_init[u] : 0.f;
  }
}

__global__ void lstm_step_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ dh_rec /*[N,H]*/,
                                     float* __restrict__ dh_pass_buf, float* __restrict__ dc_buf,
                                     const float* __restrict__ peep, const float* __restrict__ cell_init,
                                     const uint8_t* __restrict__ mask, const float* __restrict__ gates,
                                     const float* __restrict__ cell, float* __restrict__ dgates,
                                     float* __restrict__ dpeep_part /*[3,N,H] accum*/, int N, int T, int H, int ldh, int t,
                                     int t_prev /* -1 if this is the first processed step */, float clip, int last) {
  const int total = N * H;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int n = idx / H, u = idx % H;
    size_t row = (size_t)n * T + t;
    float dh = dout[row * ldh + u];
    if (!last) dh += dh_rec[idx] + dh_pass_buf[idx];
    float dc = last ? 0.f : dc_buf[idx];
    float4 g4 = *reinterpret_cast<const float4*>(gates + row * 4 * H + 4 * u);
    float c = cell[row * H + u];
    float c_prev = t_prev < 0 ? cell_init[u] : cell[((size_t)n * T + t_prev) * H + u];
    bool m = mask[row] != 0;
    bool hp = peep != nullptr;
    CellGrad g = lstm_cell_bwd(dh, dc, g4.x, g4.y, g4.z, g4.w, c, c_prev, m, hp, hp ? peep[u] : 0.f,
                               hp ? peep[H + u] : 0.f, hp ? peep[2 * H + u] : 0.f, clip);
    *reinterpret_cast<float4*>(dgates + row * 4 * H + 4 * u) = make_float4(g.dgi, g.dgf, g.dgc, g.dgo);
    dc_buf[idx] = g.dc_prev;
    dh_pass_buf[idx] = g.dh_pass;
    if (hp) {
      dpeep_part[idx] += g.pci;
      dpeep_part[(size_t)total + idx] += g.pcf;
      dpeep_part[2 * (size_t)total + idx] += g.pco;
    }
  }
}

__global__ void add_vec_kernel(float* __restrict__ a, const float* __restrict__ b, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] += b[i];
}

// ======================================================================================================
// impl 0: persistent cluster kernels
// ======================================================================================================
constexpr int LU = 32;         // hidden units per CTA
constexpr int LTHREADS = 256;

// A cluster tile holds LN = 8 * UPT utterances (UPT = 4 or 5).  Only ~15 clusters of 8 CTAs are co-resident on a
// B200, so the host picks the UPT that needs the fewest waves (512 utterances: 16 tiles of 32 = two waves, 13 tiles of
// 40 = one).  Utterance "positions" inside a tile: thread-group tg (0..7) owns positions 4tg..4tg+3 and, for UPT = 5,
// position 32 + tg, so that the first four are one aligned float4 of a shared-memory row.
__device__ __forceinline__ int lstm_pos(int tg, int i) { return i < 4 ? 4 * tg + i : 32 + tg; }

// Forward: 512 threads = two K-groups of 256.  Group g accumulates the k = g (mod 2) half of the reduction for the
// whole LN x 128 tile (UPT utterances x 4 gates per thread), the two halves are exchanged through shared memory so that
// every thread finalises 2 (or 3) utterances x 1 unit, and 16 warps per SM hide the shared-memory latency of the inner
// loop (which is bound by the shared-memory -> register return path: 2-3 LDS per 16-20 FFMA).
constexpr int LFWD_THREADS = 512;

template <bool W_SMEM, int UPT>
__global__ void __launch_bounds__(LFWD_THREADS, 1)
lstm_fwd_persistent(const float* __restrict__ xw, const float* __restrict__ w_hid, const float* __restrict__ peep,
                    const float* __restrict__ cell_init, const float* __restrict__ hid_init,
                    const uint8_t* __restrict__ mask, float* __restrict__ out, float* __restrict__ gates,
                    float* __restrict__ cell, float* __restrict__ hprev, int N, int T, int H, int ldh, int backwards) {
  constexpr int LN = 8 * UPT;
  constexpr int NFM = UPT - 2;                         // most utterances one thread finalises (K-group 0)
  constexpr int X0 = 8, X1 = 4 * (UPT - 2);            // floats a K-group-0 / K-group-1 thread hands to its partner
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / CS;
  const int Hpad = CS * LU;
  const int H4 = 4 * H;

  extern __shared__ __align__(16) float smem[];
  float* hT = smem;                                    // [2][Hpad][LN]
  float* xch0 = smem + 2 * (size_t)Hpad * LN;          // [256][X0]  partials of positions 2,3 (written by K-group 0)
  float* xch1 = xch0 + 256 * X0;                       // [256][X1]  partials of positions 0,1(,4) (written by K-group 1)
  float* Ws = xch1 + 256 * X1;                         // [H][128]  (only if W_SMEM)

  const int tid = threadIdx.x;
  const int grp = tid >> 8;                            // K-group 0/1
  const int t8 = tid & 255;
  const int w = t8 >> 5, l = t8 & 31;
  const int ul = (w & 3) * 8 + (l & 7);                // local unit 0..31
  const int tg = (w >> 2) * 4 + (l >> 3);              // utterance group 0..7
  const int ug = rank * LU + ul;                       // global unit
  const bool u_ok = ug < H;
  const int col0 = rank * 4 * LU;
  const int nf = grp ? 2 : NFM;                        // utterances this thread finalises
  int fpos[NFM];                                       // their positions in the tile
#pragma unroll
  for (int i = 0; i < NFM; ++i) fpos[i] = grp ? lstm_pos(tg, 2 + (i & 1)) : lstm_pos(tg, i < 2 ? i : 4);

  if (W_SMEM) {
    for (int i = tid; i < H * 32; i += LFWD_THREADS) {
      int k = i >> 5, c4 = (i & 31) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0 + c4 < H4) v = *reinterpret_cast<const float4*>(w_hid + (size_t)k * H4 + col0 + c4);
      *reinterpret_cast<float4*>(Ws + (size_t)k * 128 + c4) = v;
    }
  }
  for (int i = tid; i < Hpad * LN; i += LFWD_THREADS) {
    int k = i / LN;
    hT[i] = k < H ? hid_init[k] : 0.f;
  }
  const bool has_peep = peep != nullptr;
  const float w_ci = (has_peep && u_ok) ? peep[ug] : 0.f;
  const float w_cf = (has_peep && u_ok) ? peep[H + ug] : 0.f;
  const float w_co = (has_peep && u_ok) ? peep[2 * H + ug] : 0.f;
  float c_prev[NFM];
  int ng[NFM];
  bool n_ok[NFM];
#pragma unroll
  for (int i = 0; i < NFM; ++i) {
    c_prev[i] = u_ok ? cell_init[ug] : 0.f;
    ng[i] = tile * LN + fpos[i];
    n_ok[i] = i < nf && ng[i] < N;
  }
  cluster.sync();   // everyone's hT[0] / Ws initialised before any remote write can land

  // software pipeline over time: the gate pre-activations (and mask) of step s+1 are fetched while step s computes
  float4 xn[NFM];
  bool mnext[NFM];
  auto fetch = [&](int t) {
#pragma unroll
    for (int i = 0; i < NFM; ++i) {
      xn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      mnext[i] = false;
      if (n_ok[i] && u_ok) {
        size_t row = (size_t)ng[i] * T + t;
        xn[i] = __ldg(reinterpret_cast<const float4*>(xw + row * H4 + 4 * ug));
        mnext[i] = mask[row] != 0;
      }
    }
  };
  fetch(backwards ? (T - 1) : 0);

  int cur = 0;
  for (int s = 0; s < T; ++s) {
    const int t = backwards ? (T - 1 - s) : s;
    float4 xg[NFM];
    bool m[NFM];
#pragma unroll
    for (int i = 0; i < NFM; ++i) { xg[i] = xn[i]; m[i] = mnext[i]; }
    if (s + 1 < T) fetch(backwards ? (T - 2 - s) : (s + 1));
    float acc[UPT][4];
#pragma unroll
    for (int i = 0; i < UPT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* hcur = hT + (size_t)cur * Hpad * LN;
#pragma unroll 4
    for (int k = grp; k < H; k += 2) {
      float4 wv;
      if (W_SMEM) wv = *reinterpret_cast<const float4*>(Ws + (size_t)k * 128 + 4 * ul);
      else wv = u_ok ? __ldg(reinterpret_cast<const float4*>(w_hid + (size_t)k * H4 + 4 * ug)) : make_float4(0, 0, 0, 0);
      const float4 hv = *reinterpret_cast<const float4*>(hcur + (size_t)k * LN + 4 * tg);
      float hh[UPT];
      hh[0] = hv.x; hh[1] = hv.y; hh[2] = hv.z; hh[3] = hv.w;
      if (UPT == 5) hh[UPT - 1] = hcur[(size_t)k * LN + 32 + tg];
#pragma unroll
      for (int i = 0; i < UPT; ++i) {
        acc[i][0] = fmaf(hh[i], wv.x, acc[i][0]);
        acc[i][1] = fmaf(hh[i], wv.y, acc[i][1]);
        acc[i][2] = fmaf(hh[i], wv.z, acc[i][2]);
        acc[i][3] = fmaf(hh[i], wv.w, acc[i][3]);
      }
    }
    // exchange (warp-uniform branches, static register indices): K-group 0 keeps positions 0,1(,4) and hands its
    // partial sums of positions 2,3 to the partner thread (same t8) of K-group 1, and vice versa.
    float keep[NFM][4];
    float hpv[NFM];
    if (grp == 0) {
      float4* dst = reinterpret_cast<float4*>(xch0 + (size_t)t8 * X0);
      dst[0] = make_float4(acc[2][0], acc[2][1], acc[2][2], acc[2][3]);
      dst[1] = make_float4(acc[3][0], acc[3][1], acc[3][2], acc[3][3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        keep[0][j] = acc[0][j];
        keep[1][j] = acc[1][j];
        if (UPT == 5) keep[NFM - 1][j] = acc[UPT - 1][j];
      }
    } else {
      float4* dst = reinterpret_cast<float4*>(xch1 + (size_t)t8 * X1);
      dst[0] = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
      dst[1] = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
      if (UPT == 5) dst[2] = make_float4(acc[UPT - 1][0], acc[UPT - 1][1], acc[UPT - 1][2], acc[UPT - 1][3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        keep[0][j] = acc[2][j];
        keep[1][j] = acc[3][j];
        if (UPT == 5) keep[NFM - 1][j] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NFM; ++i) hpv[i] = hcur[(size_t)(u_ok ? ug : 0) * LN + fpos[i]];
    __syncthreads();
    CellOut r[NFM];
    {
      const float4* src = grp == 0 ? reinterpret_cast<const float4*>(xch1 + (size_t)t8 * X1)
                                   : reinterpret_cast<const float4*>(xch0 + (size_t)t8 * X0);
#pragma unroll
      for (int i = 0; i < NFM; ++i) {
        if (i < nf) {
          const float4 pv = src[i];
          r[i] = lstm_cell_fwd(keep[i][0] + pv.x + xg[i].x, keep[i][1] + pv.y + xg[i].y, keep[i][2] + pv.z + xg[i].z,
                               keep[i][3] + pv.w + xg[i].w, c_prev[i], hpv[i], m[i], has_peep, w_ci, w_cf, w_co);
          c_prev[i] = r[i].c;
        }
      }
    }
    // broadcast the new h of my utterances to every CTA of the cluster, then arrive on the cluster barrier BEFORE the
    // global stores of this step: the barrier's release then only has to cover the DSMEM writes, and the global
    // stores drain underneath the next step's matmul.
    if (u_ok) {
      float* mine = hT + (size_t)(cur ^ 1) * Hpad * LN + (size_t)ug * LN;
      const float2 hv2 = make_float2(r[0].h, r[1].h);
      for (int rr = 0; rr < CS; ++rr) {
        float* base = cluster.map_shared_rank(mine, rr);
        *reinterpret_cast<float2*>(base + fpos[0]) = hv2;
        if (UPT == 5 && grp == 0) base[fpos[NFM - 1]] = r[NFM - 1].h;
      }
    }
    cluster.barrier_arrive();
#pragma unroll
    for (int i = 0; i < NFM; ++i)
      if (n_ok[i] && u_ok) {
        size_t row = (size_t)ng[i] * T + t;
        out[row * ldh + ug] = r[i].h;
        if (gates) *reinterpret_cast<float4*>(gates + row * H4 + 4 * ug) = make_float4(r[i].i, r[i].f, r[i].cin, r[i].o);
        if (cell) cell[row * H + ug] = r[i].c;
        if (hprev) hprev[row * ldh + ug] = hpv[i];
      }
    cluster.barrier_wait();
    cur ^= 1;
  }
}

// Backward.  wT is W_hid^T with zero-padded rows: [4H][ldt], element (j,k) = W_hid[k][j].
template <bool W_SMEM, int NGRP, int UPT>
__global__ void __launch_bounds__(LTHREADS, 1)
lstm_bwd_persistent(const float* __restrict__ dout, const float* __restrict__ wT, int ldt,
                    const float* __restrict__ peep, const float* __restrict__ cell_init,
                    const uint8_t* __restrict__ mask, const float* __restrict__ gates, const float* __restrict__ cell,
                    float* __restrict__ dgates, float* __restrict__ dpeep, float* __restrict__ dc_fin,
                    float* __restrict__ dh_fin, int N, int T, int H, int ldh, int backwards, float clip, int WROW) {
  constexpr int LN = 8 * UPT;
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / CS;
  const int Hpad = CS * LU;
  const int H4 = 4 * H;

  extern __shared__ __align__(16) float smem[];
  float* dgT = smem;                                   // [128][LN]
  float* inbox = dgT + 128 * LN;                       // [2][CS][LN][LU]
  float* Wt = inbox + 2 * (size_t)CS * LN * LU;        // [128][WROW] (+4 floats of slack; only if W_SMEM)

  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  // elementwise mapping (same as forward): UPT utterances x 1 unit
  const int ul = (w & 3) * 8 + (l & 7);
  const int tg = (w >> 2) * 4 + (l >> 3);
  const int ug = rank * LU + ul;
  const bool u_ok = ug < H;
  const int col0 = rank * 4 * LU;
  // matmul mapping: UPT utterances x NGRP groups of 4 consecutive output units k (group q -> k = 4*(kg + 32*q))
  const int kg = ul;                                   // 0..31
  const int ngroups = Hpad / 4;                        // = 8*CS

  if (W_SMEM) {
    const int w4 = WROW / 4;
    for (int i = tid; i < 128 * w4; i += LTHREADS) {
      int jj = i / w4, k4 = (i % w4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0 + jj < H4 && k4 < ldt) v = *reinterpret_cast<const float4*>(wT + (size_t)(col0 + jj) * ldt + k4);
      *reinterpret_cast<float4*>(Wt + (size_t)jj * WROW + k4) = v;
    }
    if (tid < 4) Wt[(size_t)128 * WROW + tid] = 0.f;   // slack read by the last row when WROW < Hpad
  }
  const bool has_peep = peep != nullptr;
  const float w_ci = (has_peep && u_ok) ? peep[ug] : 0.f;
  const float w_cf = (has_peep && u_ok) ? peep[H + ug] : 0.f;
  const float w_co = (has_peep && u_ok) ? peep[2 * H + ug] : 0.f;
  int ng[UPT];
  bool n_ok[UPT];
#pragma unroll
  for (int i = 0; i < UPT; ++i) {
    ng[i] = tile * LN + lstm_pos(tg, i);
    n_ok[i] = ng[i] < N;
  }
  float dh_next[UPT], dc_next[UPT], dh_pass[UPT];
#pragma unroll
  for (int i = 0; i < UPT; ++i) dh_next[i] = dc_next[i] = dh_pass[i] = 0.f;
  float pci = 0.f, pcf = 0.f, pco = 0.f;
  cluster.sync();

  // software pipeline over time: the saved tensors of the next processed step are fetched during the matmul
  float pf_dout[UPT], pf_c[UPT], pf_cp[UPT];
  float4 pf_g[UPT];
  bool pf_m[UPT];
  auto fetch = [&](int s) {
    const int t = backwards ? (T - 1 - s) : s;
    const int t_prev = (s == 0) ? -1 : (backwards ? t + 1 : t - 1);
#pragma unroll
    for (int i = 0; i < UPT; ++i) {
      pf_dout[i] = pf_c[i] = pf_cp[i] = 0.f;
      pf_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      pf_m[i] = false;
      if (n_ok[i] && u_ok) {
        size_t row = (size_t)ng[i] * T + t;
        pf_dout[i] = __ldg(dout + row * ldh + ug);
        pf_g[i] = __ldg(reinterpret_cast<const float4*>(gates + row * H4 + 4 * ug));
        pf_c[i] = __ldg(cell + row * H + ug);
        pf_cp[i] = t_prev < 0 ? cell_init[ug] : __ldg(cell + ((size_t)ng[i] * T + t_prev) * H + ug);
        pf_m[i] = mask[row] != 0;
      }
    }
  };
  fetch(T - 1);

  int par = 0;
  for (int s = T - 1; s >= 0; --s) {                   // reverse of the processing order
    const int t = backwards ? (T - 1 - s) : s;
    // ---- 1. elementwise gate gradients for my (UPT utterances, 1 unit) ----
    float4 dgv[UPT];
#pragma unroll
    for (int i = 0; i < UPT; ++i) {
      dgv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      dh_pass[i] = 0.f;
      if (n_ok[i] && u_ok) {
        size_t row = (size_t)ng[i] * T + t;
        float dh = pf_dout[i] + dh_next[i];
        CellGrad g = lstm_cell_bwd(dh, dc_next[i], pf_g[i].x, pf_g[i].y, pf_g[i].z, pf_g[i].w, pf_c[i], pf_cp[i],
                                   pf_m[i], has_peep, w_ci, w_cf, w_co, clip);
        dgv[i] = make_float4(g.dgi, g.dgf, g.dgc, g.dgo);
        dc_next[i] = g.dc_prev;
        dh_pass[i] = g.dh_pass;
        pci += g.pci; pcf += g.pcf; pco += g.pco;
        *reinterpret_cast<float4*>(dgates + row * H4 + 4 * ug) = dgv[i];
      }
    }
    if (s > 0) fetch(s - 1);
    // dgT[jj][position], jj = 4*ul + gate
    *reinterpret_cast<float4*>(dgT + (size_t)(4 * ul + 0) * LN + 4 * tg) = make_float4(dgv[0].x, dgv[1].x, dgv[2].x, dgv[3].x);
    *reinterpret_cast<float4*>(dgT + (size_t)(4 * ul + 1) * LN + 4 * tg) = make_float4(dgv[0].y, dgv[1].y, dgv[2].y, dgv[3].y);
    *reinterpret_cast<float4*>(dgT + (size_t)(4 * ul + 2) * LN + 4 * tg) = make_float4(dgv[0].z, dgv[1].z, dgv[2].z, dgv[3].z);
    *reinterpret_cast<float4*>(dgT + (size_t)(4 * ul + 3) * LN + 4 * tg) = make_float4(dgv[0].w, dgv[1].w, dgv[2].w, dgv[3].w);
    if (UPT == 5) {
      dgT[(size_t)(4 * ul + 0) * LN + 32 + tg] = dgv[UPT - 1].x;
      dgT[(size_t)(4 * ul + 1) * LN + 32 + tg] = dgv[UPT - 1].y;
      dgT[(size_t)(4 * ul + 2) * LN + 32 + tg] = dgv[UPT - 1].z;
      dgT[(size_t)(4 * ul + 3) * LN + 32 + tg] = dgv[UPT - 1].w;
    }
    __syncthreads();
    if (s == 0) break;   // the recurrent gradient of the first processed step goes to hid_init: handled below
    // ---- 2. partial dh_prev[n][k] = sum_jj dg[n][jj] * W_hid[k][col0+jj] ----
    float pacc[NGRP][UPT][4];
#pragma unroll
    for (int q = 0; q < NGRP; ++q)
#pragma unroll
      for (int i = 0; i < UPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) pacc[q][i][j] = 0.f;
#pragma unroll 2
    for (int jj = 0; jj < 128; ++jj) {
      const float4 d4 = *reinterpret_cast<const float4*>(dgT + (size_t)jj * LN + 4 * tg);
      float dd[UPT];
      dd[0] = d4.x; dd[1] = d4.y; dd[2] = d4.z; dd[3] = d4.w;
      if (UPT == 5) dd[UPT - 1] = dgT[(size_t)jj * LN + 32 + tg];
#pragma unroll
      for (int q = 0; q < NGRP; ++q) {
        const int k0 = 4 * (kg + 32 * q);
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kg + 32 * q < ngroups) {
          if (W_SMEM) wv = *reinterpret_cast<const float4*>(Wt + (size_t)jj * WROW + k0);
          else if (col0 + jj < H4 && k0 < ldt)
            wv = __ldg(reinterpret_cast<const float4*>(wT + (size_t)(col0 + jj) * ldt + k0));
        }
#pragma unroll
        for (int i = 0; i < UPT; ++i) {
          pacc[q][i][0] = fmaf(dd[i], wv.x, pacc[q][i][0]);
          pacc[q][i][1] = fmaf(dd[i], wv.y, pacc[q][i][1]);
          pacc[q][i][2] = fmaf(dd[i], wv.z, pacc[q][i][2]);
          pacc[q][i][3] = fmaf(dd[i], wv.w, pacc[q][i][3]);
        }
      }
    }
    // ---- 3. reduce-scatter: my partials for units owned by CTA `owner` go to its inbox[par][rank] ----
    {
      float* box = inbox + ((size_t)par * CS + rank) * LN * LU;   // same offset in every CTA
#pragma unroll
      for (int q = 0; q < NGRP; ++q) {
        if (kg + 32 * q < ngroups) {
          const int k0 = 4 * (kg + 32 * q);
          const int owner = k0 / LU, uo = k0 % LU;
          float* rbox = cluster.map_shared_rank(box, owner);
#pragma unroll
          for (int i = 0; i < UPT; ++i)
            *reinterpret_cast<float4*>(rbox + (size_t)lstm_pos(tg, i) * LU + uo) =
                make_float4(pacc[q][i][0], pacc[q][i][1], pacc[q][i][2], pacc[q][i][3]);
        }
      }
    }
    cluster.sync();
    // ---- 4. sum the CS partials for my (utterances, unit) ----
#pragma unroll
    for (int i = 0; i < UPT; ++i) {
      float sacc = dh_pass[i];
      for (int src = 0; src < CS; ++src)
        sacc += inbox[(((size_t)par * CS + src) * LN + lstm_pos(tg, i)) * LU + ul];
      dh_next[i] = sacc;
    }
    par ^= 1;
  }
  // after the first processed step: dc_next is d(cell_init) per utterance; d(hid_init) = dg W^T + dh_pass, which
  // the host finishes with one GEMM over the first-step dgates (it needs every column, not just this CTA's).
#pragma unroll
  for (int i = 0; i < UPT; ++i)
    if (n_ok[i] && u_ok) {
      dc_fin[(size_t)ng[i] * H + ug] = dc_next[i];
      dh_fin[(size_t)ng[i] * H + ug] = dh_pass[i];
    }
  if (has_peep && u_ok) {
    atomicAdd(dpeep + ug, pci);
    atomicAdd(dpeep + H + ug, pcf);
    atomicAdd(dpeep + 2 * H + ug, pco);
  }
}

__global__ void transpose_pad_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst,
                                     int ldt) {
  // dst[c][r] = src[r][c]; dst row stride ldt >= rows, padding zeroed
  __shared__ float tile[32][33];
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < ldt) dst[(size_t)c * ldt + r] = tile[threadIdx.x][i];
  }
}

template <typename K>
static int launch_cluster(K kernel, int grid, int cs, size_t smem, cudaStream_t st, void** args, int threads = LTHREADS) {
  IPAVSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) IPAVSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  IPAVSR_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), args));
  count_launch();
  return IPAVSR_OK;
}

// number of clusters of `cs` CTAs of this kernel that can be resident at once (0 when the query is unavailable)
template <typename K>
static int max_active_clusters(K kernel, int cs, size_t smem, int threads) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cs > 8) cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * 64);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, reinterpret_cast<const void*>(kernel), &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// UPT (utterances per thread group, tile = 8*UPT utterances): the one that needs the fewest waves of co-resident
// clusters; a 40-utterance tile costs ~20 % more per step than a 32-utterance one.  IPAVSR_LSTM_UPT=4|5 overrides.
static int choose_upt(int N, int maxc4, int maxc5, bool fits5) {
  static int force = -1;
  if (force < 0) {
    const char* e = getenv("IPAVSR_LSTM_UPT");
    force = e ? atoi(e) : 0;
  }
  if (!fits5 || force == 4) return 4;
  if (force == 5) return 5;
  if (maxc4 <= 0 || maxc5 <= 0) return 4;
  const int t4 = (N + 31) / 32, t5 = (N + 39) / 40;
  const double c4 = (double)((t4 + maxc4 - 1) / maxc4), c5 = 1.2 * (double)((t5 + maxc5 - 1) / maxc5);
  return c5 < c4 ? 5 : 4;
}

static int max_smem_optin() {
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return v > 0 ? v : 227 * 1024;
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

uint64_t ipavsr_lstm_workspace_bytes(int N, int T, int H) {
  (void)T;
  size_t n = (size_t)N, h = (size_t)H;
  size_t floats = 4 * h * n        // recurrent pre-activations (impl 1)
                  + 8 * n * h      // state / gradient ping-pong buffers, dc_fin, dh_fin
                  + 3 * n * h      // per-utterance peephole partials (impl 1)
                  + 4 * h * (h + 8)  // W_hid^T (impl 0 backward)
                  + 4096;
  return floats * sizeof(float);
}

int ipavsr_lstm_fwd(const float* xw, const float* w_hid, const float* peep, const float* cell_init,
                    const float* hid_init, const uint8_t* mask, float* out, float* gates, float* cell, float* hprev,
                    int N, int T, int H, int ldh, int backwards, int impl, void* workspace,
                    uint64_t workspace_bytes, void* stream) {
  IPAVSR_CHECK_ARG(xw && w_hid && cell_init && hid_init && mask && out, "null pointer");
  IPAVSR_CHECK_ARG(N >= 0 && T >= 1 && H >= 1 && ldh >= H, "bad sizes");
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int cs = (H + LU - 1) / LU;
  if (impl == 0 && cs > 16) impl = 1;
  if (impl == 0) {
    const size_t optin = (size_t)max_smem_optin();
    const size_t w_bytes = (size_t)H * 128 * sizeof(float);
    // shared memory of the UPT = 4 / 5 variants: hT [2][Hpad][8*UPT] + exchange 256 * (8 + 4*(UPT-2)) floats (+ W slice)
    const size_t h4 = (2 * (size_t)cs * LU * 32 + 256 * 16) * sizeof(float);
    const size_t h5 = (2 * (size_t)cs * LU * 40 + 256 * 20) * sizeof(float);
    const bool w_smem = h4 + w_bytes <= optin;
    const bool fits5 = w_smem ? (h5 + w_bytes <= optin) : (h5 <= optin);
    const size_t smem4 = h4 + (w_smem ? w_bytes : 0), smem5 = h5 + (w_smem ? w_bytes : 0);
    // co-resident clusters per variant (cached per cluster size / shared-memory footprint)
    static int c_cs = -1, c_m4 = 0, c_m5 = 0;
    static size_t c_s4 = 0, c_s5 = 0;
    if (c_cs != cs || c_s4 != smem4 || c_s5 != smem5) {
      c_m4 = w_smem ? max_active_clusters(lstm_fwd_persistent<true, 4>, cs, smem4, LFWD_THREADS)
                    : max_active_clusters(lstm_fwd_persistent<false, 4>, cs, smem4, LFWD_THREADS);
      c_m5 = !fits5 ? 0 : (w_smem ? max_active_clusters(lstm_fwd_persistent<true, 5>, cs, smem5, LFWD_THREADS)
                                  : max_active_clusters(lstm_fwd_persistent<false, 5>, cs, smem5, LFWD_THREADS));
      c_cs = cs; c_s4 = smem4; c_s5 = smem5;
    }
    const int upt = choose_upt(N, c_m4, c_m5, fits5);
    const int tiles = (N + 8 * upt - 1) / (8 * upt);
    void* args[] = {&xw, &w_hid, &peep, &cell_init, &hid_init, &mask, &out, &gates, &cell, &hprev, &N, &T, &H, &ldh, &backwards};
    if (upt == 5) {
      if (w_smem) return launch_cluster(lstm_fwd_persistent<true, 5>, tiles * cs, cs, smem5, st, args, LFWD_THREADS);
      return launch_cluster(lstm_fwd_persistent<false, 5>, tiles * cs, cs, smem5, st, args, LFWD_THREADS);
    }
    if (w_smem) return launch_cluster(lstm_fwd_persistent<true, 4>, tiles * cs, cs, smem4, st, args, LFWD_THREADS);
    return launch_cluster(lstm_fwd_persistent<false, 4>, tiles * cs, cs, smem4, st, args, LFWD_THREADS);
  }
  // ---- impl 1 ----
  IPAVSR_CHECK_ARG(workspace && workspace_bytes >= ipavsr_lstm_workspace_bytes(N, T, H), "workspace too small");
  float* ws = reinterpret_cast<float*>(workspace);
  float* rec = ws;                                  // [N,4H]
  float* hb[2] = {rec + (size_t)4 * H * N, rec + (size_t)4 * H * N + (size_t)N * H};
  float* cb[2] = {hb[1] + (size_t)N * H, hb[1] + 2 * (size_t)N * H};
  const int total = N * H;
  const int grid = (total + 255) / 256 < sm_count() * 8 ? (total + 255) / 256 : sm_count() * 8;
  lstm_seed_kernel<<<grid, 256, 0, st>>>(cell_init, hid_init, cb[0], hb[0], N, H);
  IPAVSR_LAUNCH_CHECK();
  int cur = 0;
  for (int s = 0; s < T; ++s) {
    int t = backwards ? T - 1 - s : s;
    int rc = gemm_simt(0, 0, N, 4 * H, H, hb[cur], H, w_hid, 4 * H, rec, 4 * H, nullptr, IPAVSR_ACT_LINEAR, 0, st);
    if (rc) return rc;
    lstm_step_fwd_kernel<<<grid, 256, 0, st>>>(xw, rec, peep, cb[cur], hb[cur], cell_init, hid_init, 0, mask,
                                               cb[cur ^ 1], hb[cur ^ 1], out, gates, cell, hprev, N, T, H, ldh, t);
    IPAVSR_LAUNCH_CHECK();
    cur ^= 1;
  }
  return IPAVSR_OK;
}

int ipavsr_lstm_bwd(const float* dout, const float* w_hid, const float* peep, const float* cell_init,
                    const uint8_t* mask, const float* gates, const float* cell, float* dgates, float* dpeep,
                    float* dcell_init, float* dhid_init, int N, int T, int H, int ldh, int backwards, float clip,
                    int accumulate, int impl, void* workspace, uint64_t workspace_bytes, void* stream) {
  IPAVSR_CHECK_ARG(dout && w_hid && cell_init && mask && gates && cell && dgates && dcell_init && dhid_init,
                   "null pointer");
  IPAVSR_CHECK_ARG((peep == nullptr) == (dpeep == nullptr), "peep and dpeep go together");
  IPAVSR_CHECK_ARG(N >= 0 && T >= 1 && H >= 1, "bad sizes");
  IPAVSR_CHECK_ARG(workspace && workspace_bytes >= ipavsr_lstm_workspace_bytes(N, T, H), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) {
    IPAVSR_CUDA(cudaMemsetAsync(dcell_init, 0, sizeof(float) * H, st));
    IPAVSR_CUDA(cudaMemsetAsync(dhid_init, 0, sizeof(float) * H, st));
    if (dpeep) IPAVSR_CUDA(cudaMemsetAsync(dpeep, 0, sizeof(float) * 3 * H, st));
  }
  if (N == 0) return IPAVSR_OK;
  float* ws = reinterpret_cast<float*>(workspace);
  const size_t NH = (size_t)N * H;
  float* rec = ws;                    // [N,4H] region reused: dh_rec [N,H]
  float* base = ws + 4 * NH;
  float* dh_pass = base;              // [N,H]
  float* dc_buf = base + NH;          // [N,H]
  float* dc_fin = base + 2 * NH;
  float* dh_fin = base + 3 * NH;
  float* ppart = base + 8 * NH;       // [3,N,H]
  float* wT = ws + ((4 * NH + 8 * NH + 3 * NH + 3) / 4) * 4;   // [4H][ldt], 16-byte aligned for float4 loads
  const int cs = (H + LU - 1) / LU;
  if (impl == 0 && cs > 16) impl = 1;
  const int t_first = backwards ? T - 1 : 0;   // the first *processed* step
  if (impl == 0) {
    const int ldt = (H + 3) / 4 * 4;
    dim3 tb(32, 8), tg((4 * H + 31) / 32, (H + 31) / 32);
    transpose_pad_kernel<<<tg, tb, 0, st>>>(w_hid, H, 4 * H, wT, ldt);
    IPAVSR_LAUNCH_CHECK();
    const size_t optin = (size_t)max_smem_optin();
    // shared memory: dgT [128][LN] + inbox [2][CS][LN][LU] + W^T slice [128][WROW] (+4 floats slack), LN = 8*UPT.
    // WROW = Hpad + 4 when it fits, else the tight ldt (H rounded up to 4).
    const size_t f4 = (128 * 32 + 2 * (size_t)cs * 32 * LU) * sizeof(float);
    const size_t f5 = (128 * 40 + 2 * (size_t)cs * 40 * LU) * sizeof(float);
    auto wbytes = [&](int wrow) { return ((size_t)128 * wrow + 4) * sizeof(float); };
    const int wide = cs * LU + 4;
    const bool w_smem = f4 + wbytes(ldt) <= optin;
    const int wrow4 = (f4 + wbytes(wide) <= optin) ? wide : ldt;
    const int wrow5 = (f5 + wbytes(wide) <= optin) ? wide : ldt;
    const bool fits5 = w_smem ? (f5 + wbytes(wrow5) <= optin) : (f5 <= optin);
    const size_t smem4 = f4 + (w_smem ? wbytes(wrow4) : 0), smem5 = f5 + (w_smem ? wbytes(wrow5) : 0);
    const int ngrp = (cs * 8 + 31) / 32;   // groups of 4 output units per matmul thread
    int rc = IPAVSR_OK;
    int wrow = 0;
    void* args[] = {&dout, &wT, (void*)&ldt, &peep, &cell_init, &mask, &gates, &cell, &dgates, &dpeep,
                    &dc_fin, &dh_fin, &N, &T, &H, &ldh, &backwards, &clip, &wrow};
#define IPAVSR_BWD_RUN(G, U, SM, WR)                                                                        \
  do {                                                                                                     \
    wrow = (WR);                                                                                           \
    const int tiles = (N + 8 * U - 1) / (8 * U);                                                           \
    rc = w_smem ? launch_cluster(lstm_bwd_persistent<true, G, U>, tiles * cs, cs, SM, st, args)            \
                : launch_cluster(lstm_bwd_persistent<false, G, U>, tiles * cs, cs, SM, st, args);          \
  } while (0)
#define IPAVSR_BWD_CASE(G)                                                                                  \
  do {                                                                                                     \
    static int c_cs = -1, c_m4 = 0, c_m5 = 0;                                                              \
    static size_t c_s4 = 0, c_s5 = 0;                                                                      \
    if (c_cs != cs || c_s4 != smem4 || c_s5 != smem5) {                                                    \
      c_m4 = w_smem ? max_active_clusters(lstm_bwd_persistent<true, G, 4>, cs, smem4, LTHREADS)            \
                    : max_active_clusters(lstm_bwd_persistent<false, G, 4>, cs, smem4, LTHREADS);          \
      c_m5 = !fits5 ? 0 : (w_smem ? max_active_clusters(lstm_bwd_persistent<true, G, 5>, cs, smem5, LTHREADS)   \
                                  : max_active_clusters(lstm_bwd_persistent<false, G, 5>, cs, smem5, LTHREADS)); \
      c_cs = cs; c_s4 = smem4; c_s5 = smem5;                                                               \
    }                                                                                                      \
    if (choose_upt(N, c_m4, c_m5, fits5) == 5) IPAVSR_BWD_RUN(G, 5, smem5, wrow5);                         \
    else IPAVSR_BWD_RUN(G, 4, smem4, wrow4);                                                               \
  } while (0)
    if (ngrp <= 1) { IPAVSR_BWD_CASE(1); }
    else if (ngrp == 2) { IPAVSR_BWD_CASE(2); }
    else if (ngrp == 3) { IPAVSR_BWD_CASE(3); }
    else { IPAVSR_BWD_CASE(4); }
#undef IPAVSR_BWD_CASE
#undef IPAVSR_BWD_RUN
    if (rc) return rc;
  } else {
    const int total = N * H;
    const int grid = (total + 255) / 256 < sm_count() * 8 ? (total + 255) / 256 : sm_count() * 8;
    if (peep) IPAVSR_CUDA(cudaMemsetAsync(ppart, 0, sizeof(float) * 3 * NH, st));
    for (int s = T - 1; s >= 0; --s) {
      int t = backwards ? T - 1 - s : s;
      int t_prev = s == 0 ? -1 : (backwards ? t + 1 : t - 1);
      lstm_step_bwd_kernel<<<grid, 256, 0, st>>>(dout, rec, dh_pass, dc_buf, peep, cell_init, mask, gates, cell,
                                                 dgates, ppart, N, T, H, ldh, t, t_prev, clip, s == T - 1);
      IPAVSR_LAUNCH_CHECK();
      if (s > 0) {
        // dh_rec[N,H] = dg_t[N,4H] (rows strided by T*4H) * W_hid^T
        int rc = gemm_simt(0, 1, N, H, 4 * H, dgates + (size_t)t * 4 * H, T * 4 * H, w_hid, 4 * H, rec, H, nullptr,
                           IPAVSR_ACT_LINEAR, 0, st);
        if (rc) return rc;
      }
    }
    if (peep) {
      for (int k = 0; k < 3; ++k) {
        int rc = ipavsr_colsum(ppart + k * NH, H, dpeep + k * H, N, H, 1, stream);
        if (rc) return rc;
      }
    }
    dc_fin = dc_buf;
    dh_fin = dh_pass;
  }
  // d(hid_init) = sum_n [ dg_first W_hid^T + dh_pass ],  d(cell_init) = sum_n dc_prev at the first processed step; the
  // product is linear, so it is taken after the sum over the utterances (one column sum + a vector-matrix product)
  float* dg_sum = rec;                // the recurrent pre-activation region is free once the steps are done
  int rc = ipavsr_colsum(dgates + (size_t)t_first * 4 * H, T * 4 * H, dg_sum, N, 4 * H, 0, stream);
  if (rc) return rc;
  rc = ipavsr_colsum(dh_fin, H, dhid_init, N, H, 1, stream);
  if (rc) return rc;
  rc = gemv_rows(w_hid, 4 * H, dg_sum, dhid_init, H, 4 * H, 1, st);
  if (rc) return rc;
  rc = ipavsr_colsum(dc_fin, H, dcell_init, N, H, 1, stream);
  return rc;
}

}  // extern "C"
