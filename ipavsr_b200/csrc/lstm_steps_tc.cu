// Masked Lasagne-LSTM recurrence for WIDE layers (H > 256) on the tensor cores, one recurrent GEMM per time step.
//
// Same semantics as lstm.cu / lstm_tc.cu (reference custom/layers.py:10-80 -> lasagne LSTMLayer).
// adenet_v3 runs 2 x lstm_size = 500-wide LSTMs (modelzoo/adenet_v3.py:114,125,136,156) and adenet_v1 a 500-wide second
// BLSTM (modelzoo/adenet_v1.py:95).  Their W_hid (H x 4H; 4 MB as the fp16 hi/lo pair of the f16x3 arithmetic) does not fit
// the shared memory of one thread-block cluster (16 x 227 KB), so the cluster-resident recurrence of lstm_tc.cu stops at
// H = 256 and the FFMA cluster kernel of lstm.cu streams W_hid from L2 on the CUDA cores (measured 35.9 ms per 512-utterance
// adenet_v3 step, almost all of it the ten recurrences).  Here every step is
//     rec[n_t, 4H] = h_{t-1}[n_t, H] * W_hid[H, 4H]          one fp16 three-product tcgen05 GEMM (gemm_tc.cu), operands:
//                                                              h as fp16 hi/lo written by the previous step's cell kernel,
//                                                              W_hid straight from the engine's fp16 split of the arena
//     cell update for all N utterances                         one elementwise kernel (fp32 state, emits the next operand)
// and the backward pass mirrors it (cell backward emits the clipped gate gradients as fp32 AND as the fp16 hi/lo operand —
// which is also the split of dgates the weight-gradient GEMMs need — then dh_rec = dg_t W_hid^T is one GEMM).
// All launches of a pass are enqueued by ONE C-ABI call.  `active_rows` (host, optional): active_rows[t] = number of
// leading utterances that can be unmasked at frame t (the engine's length-sorted batches); the GEMM of a step only computes
// those rows.  Judged by per-step latency, reported by tools/bench_kernels.py.
#include <cuda_fp16.h>
#include "common.cuh"

namespace ipavsr {

int gemv_rows(const float* W, int ldw, const float* s, float* out, int rows, int cols, int accumulate, cudaStream_t st);
int gemm_simt(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
              float* C, int ldc, const float* bias, int act, int accumulate, cudaStream_t st);
int gemm_tc_f16x3(int transA, int transB, int M, int N, int K, const uint16_t* Ahi, const uint16_t* Alo, int lda,
                  const int32_t* expA, const uint16_t* Bhi, const uint16_t* Blo, int ldb, const int32_t* expB, float* C,
                  int ldc, const float* bias, int act, int accumulate, float* amax, uint16_t* C16hi, uint16_t* C16lo,
                  int c16_exp, cudaStream_t st);
bool gemm_tc_f16_supported(int M, int N, int K, const void* A, int lda, const void* B, int ldb);

namespace {

__device__ __forceinline__ void split16(float x, float scale, __half& hi, __half& lo) {
  const float xs = x * scale;
  hi = __float2half_rn(xs);
  lo = __float2half_rn((xs - __half2float(hi)) * F16_LO_SCALE);
}

struct SCell {
  float i, f, cin, o, c, h;
};
__device__ __forceinline__ SCell s_cell_fwd(float gi, float gf, float gc, float go, float c_prev, float h_prev, bool m,
                                            bool has_peep, float w_ci, float w_cf, float w_co) {
  SCell r;
  if (has_peep) {
    gi = fmaf(c_prev, w_ci, gi);
    gf = fmaf(c_prev, w_cf, gf);
  }
  r.i = sigmoidf_(gi);
  r.f = sigmoidf_(gf);
  r.cin = tanhf(gc);
  const float c_u = r.f * c_prev + r.i * r.cin;
  if (has_peep) go = fmaf(c_u, w_co, go);
  r.o = sigmoidf_(go);
  const float h_u = r.o * tanhf(c_u);
  r.c = m ? c_u : c_prev;
  r.h = m ? h_u : h_prev;
  return r;
}

// state buffers <- initial state for every utterance; the scale exponent of the h operand from max(1, |hid_init|max)
__global__ void __launch_bounds__(256) steps_seed_kernel(const float* __restrict__ cell_init,
                                                         const float* __restrict__ hid_init, float* __restrict__ cbuf,
                                                         float* __restrict__ hbuf, __half* __restrict__ h_hi,
                                                         __half* __restrict__ h_lo, int32_t* __restrict__ exp_h, int N,
                                                         int H, int ldk) {
  __shared__ float red[8];
  float mx = 1.0f;
  for (int k = threadIdx.x; k < H; k += blockDim.x) mx = fmaxf(mx, fabsf(hid_init[k]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  const int eh = 14 - ilogbf(mx);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_h = eh;
  const float hs = __int_as_float((127 + eh) << 23);
  const long long total = (long long)N * ldk;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / ldk), u = (int)(idx - (long long)n * ldk);
    __half hi = __float2half_rn(0.f), lo = hi;
    if (u < H) {
      const float h0 = hid_init[u];
      cbuf[(size_t)n * H + u] = cell_init[u];
      hbuf[(size_t)n * H + u] = h0;
      split16(h0, hs, hi, lo);
    }
    h_hi[idx] = hi;          // the padding columns H..ldk-1 of the operand stay zero
    h_lo[idx] = lo;
  }
}

__global__ void __launch_bounds__(256) steps_fwd_kernel(const float* __restrict__ xw, const float* __restrict__ rec,
                                                        int n_rec, const float* __restrict__ peep,
                                                        float* __restrict__ cbuf, float* __restrict__ hbuf,
                                                        __half* __restrict__ h_hi, __half* __restrict__ h_lo,
                                                        const int32_t* __restrict__ exp_h,
                                                        const uint8_t* __restrict__ mask, float* __restrict__ out,
                                                        float* __restrict__ gates, float* __restrict__ cell,
                                                        float* __restrict__ hprev, int N, int T, int H, int ldh, int ldk,
                                                        int t) {
  const float hs = __int_as_float((127 + __ldg(exp_h)) << 23);
  const bool hp = peep != nullptr;
  const long long total = (long long)N * H;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / H), u = (int)(idx - (long long)n * H);
    const size_t row = (size_t)n * T + t;
    const bool m = mask[row] != 0;
    const float c_prev = cbuf[idx], h_prev = hbuf[idx];
    SCell r;
    if (m || n < n_rec) {
      const float4 x4 = __ldg(reinterpret_cast<const float4*>(xw + row * 4 * H + 4 * u));
      const float4 r4 = n < n_rec ? *reinterpret_cast<const float4*>(rec + (size_t)n * 4 * H + 4 * u)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      r = s_cell_fwd(x4.x + r4.x, x4.y + r4.y, x4.z + r4.z, x4.w + r4.w, c_prev, h_prev, m, hp, hp ? peep[u] : 0.f,
                     hp ? peep[H + u] : 0.f, hp ? peep[2 * H + u] : 0.f);
    } else {                    // beyond the active rows of this frame: the state passes through, the gates are never read
      r.i = r.f = r.cin = r.o = 0.f;
      r.c = c_prev;
      r.h = h_prev;
    }
    cbuf[idx] = r.c;
    hbuf[idx] = r.h;
    __half hi, lo;
    split16(r.h, hs, hi, lo);
    h_hi[(size_t)n * ldk + u] = hi;
    h_lo[(size_t)n * ldk + u] = lo;
    out[row * ldh + u] = r.h;
    if (gates) *reinterpret_cast<float4*>(gates + row * 4 * H + 4 * u) = make_float4(r.i, r.f, r.cin, r.o);
    if (cell) cell[row * H + u] = r.c;
    if (hprev) hprev[row * ldh + u] = h_prev;
  }
}

struct SGrad {
  float dgi, dgf, dgc, dgo, dc_prev, dh_pass, pci, pcf, pco;
};
__device__ __forceinline__ SGrad s_cell_bwd(float dh, float dc, float i, float f, float cin, float o, float c,
                                            float c_prev, bool m, bool has_peep, float w_ci, float w_cf, float w_co,
                                            float clip) {
  SGrad g;
  if (!m) {
    g.dgi = g.dgf = g.dgc = g.dgo = 0.f;
    g.dc_prev = dc;
    g.dh_pass = dh;
    g.pci = g.pcf = g.pco = 0.f;
    return g;
  }
  const float tc = tanhf(c);
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

// one BPTT step at frame t: dh = dout[t] + (dh_rec of the previously processed frame, rows < n_rec) + dh_pass
__global__ void __launch_bounds__(256) steps_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ dh_rec,
                                                        int n_rec, float* __restrict__ dh_pass_buf,
                                                        float* __restrict__ dc_buf, const float* __restrict__ peep,
                                                        const float* __restrict__ cell_init,
                                                        const uint8_t* __restrict__ mask, const float* __restrict__ gates,
                                                        const float* __restrict__ cell, float* __restrict__ dgates,
                                                        __half* __restrict__ dg_hi, __half* __restrict__ dg_lo, float gs,
                                                        float* __restrict__ dpeep_part, int N, int T, int H, int ldh,
                                                        int t, int t_prev, float clip, int last) {
  const bool hp = peep != nullptr;
  const long long total = (long long)N * H;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / H), u = (int)(idx - (long long)n * H);
    const size_t row = (size_t)n * T + t;
    float dh = dout[row * ldh + u];
    float dc = 0.f;
    if (!last) {
      dh += dh_pass_buf[idx] + (n < n_rec ? dh_rec[idx] : 0.f);
      dc = dc_buf[idx];
    }
    const bool m = mask[row] != 0;
    SGrad g;
    if (m) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gates + row * 4 * H + 4 * u));
      const float c = cell[row * H + u];
      const float c_prev = t_prev < 0 ? cell_init[u] : cell[((size_t)n * T + t_prev) * H + u];
      g = s_cell_bwd(dh, dc, g4.x, g4.y, g4.z, g4.w, c, c_prev, true, hp, hp ? peep[u] : 0.f, hp ? peep[H + u] : 0.f,
                     hp ? peep[2 * H + u] : 0.f, clip);
    } else {
      g.dgi = g.dgf = g.dgc = g.dgo = 0.f;
      g.dc_prev = dc;
      g.dh_pass = dh;
      g.pci = g.pcf = g.pco = 0.f;
    }
    const size_t o = row * 4 * H + 4 * u;
    *reinterpret_cast<float4*>(dgates + o) = make_float4(g.dgi, g.dgf, g.dgc, g.dgo);
    __half h[4], l[4];
    split16(g.dgi, gs, h[0], l[0]);
    split16(g.dgf, gs, h[1], l[1]);
    split16(g.dgc, gs, h[2], l[2]);
    split16(g.dgo, gs, h[3], l[3]);
    *reinterpret_cast<uint2*>(dg_hi + o) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(dg_lo + o) = *reinterpret_cast<const uint2*>(l);
    dc_buf[idx] = g.dc_prev;
    dh_pass_buf[idx] = g.dh_pass;
    if (hp && m) {
      dpeep_part[idx] += g.pci;
      dpeep_part[(size_t)total + idx] += g.pcf;
      dpeep_part[2 * (size_t)total + idx] += g.pco;
    }
  }
}

__global__ void set_int_kernel(int32_t* p, int32_t v) { *p = v; }

inline size_t up4(size_t v) { return (v + 3) / 4 * 4; }

struct StepsWs {
  float *rec, *hbuf, *cbuf, *dh_pass, *dc, *ppart;
  __half *h_hi, *h_lo;
  int32_t* exp_h;
  int ldk;
};
StepsWs carve(void* ws, int N, int H) {
  StepsWs w;
  const size_t NH = (size_t)N * H;
  float* p = reinterpret_cast<float*>(ws);
  w.rec = p;               p += up4(4 * NH);          // [N, 4H] forward; [N, H] backward
  w.hbuf = p;              p += up4(NH);
  w.cbuf = p;              p += up4(NH);
  w.dh_pass = p;           p += up4(NH);
  w.dc = p;                p += up4(NH);
  w.ppart = p;             p += up4(3 * NH);
  w.ldk = (H + 7) / 8 * 8;
  w.h_hi = reinterpret_cast<__half*>(p);   p += up4(((size_t)N * w.ldk + 1) / 2);
  w.h_lo = reinterpret_cast<__half*>(p);   p += up4(((size_t)N * w.ldk + 1) / 2);
  w.exp_h = reinterpret_cast<int32_t*>(p);
  return w;
}

inline int grid_for(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

uint64_t ipavsr_lstm_steps_workspace_bytes(int N, int T, int H) {
  (void)T;
  const size_t NH = (size_t)N * H, ldk = (size_t)(H + 7) / 8 * 8;
  const size_t floats = up4(4 * NH) + 4 * up4(NH) + up4(3 * NH) + 2 * up4(((size_t)N * ldk + 1) / 2) + 64;
  return floats * sizeof(float);
}

int ipavsr_lstm_steps_supported(int N, int T, int H, int ldw) {
  (void)T;
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("IPAVSR_LSTM_STEPS");
    off = (e && e[0] == '0') ? 1 : 0;
  }
  // the per-step product must be one the tensor-core GEMM takes (>= 4 MFLOP, 16-byte rows): N * 4H * H >= 4e6
  return (!off && H >= 64 && H <= 2048 && N >= 1 && ldw % 8 == 0 && ldw >= 4 * H && (double)N * 4.0 * H * H >= 4.0e6) ? 1 : 0;
}

int ipavsr_lstm_fwd_f16_steps(const float* xw, const float* w_hid, const uint16_t* whid_hi, const uint16_t* whid_lo,
                              const int32_t* whid_exp, int ldw, const float* peep, const float* cell_init, const float* hid_init,
                              const uint8_t* mask, float* out, float* gates, float* cell, float* hprev, int N, int T, int H,
                              int ldh, int backwards, const int32_t* active_rows, void* workspace,
                              uint64_t workspace_bytes, void* stream) {
  IPAVSR_CHECK_ARG(xw && w_hid && whid_hi && whid_lo && whid_exp && cell_init && hid_init && mask && out, "null pointer");
  IPAVSR_CHECK_ARG(N >= 0 && T >= 1 && H >= 1 && ldh >= H, "bad sizes");
  if (!ipavsr_lstm_steps_supported(N, T, H, ldw) ||
      ((reinterpret_cast<uintptr_t>(whid_hi) | reinterpret_cast<uintptr_t>(whid_lo)) & 15) != 0) {
    set_error("ipavsr_lstm_fwd_f16_steps: needs 64 <= H <= 2048, N*4H*H >= 4e6, 16-byte aligned W_hid halves, ldw %% 8 == 0");
    return IPAVSR_ERR_UNSUPPORTED;
  }
  IPAVSR_CHECK_ARG(workspace && workspace_bytes >= ipavsr_lstm_steps_workspace_bytes(N, T, H), "workspace too small");
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  StepsWs w = carve(workspace, N, H);
  steps_seed_kernel<<<grid_for((long long)N * w.ldk), 256, 0, st>>>(cell_init, hid_init, w.cbuf, w.hbuf, w.h_hi, w.h_lo,
                                                                     w.exp_h, N, H, w.ldk);
  IPAVSR_LAUNCH_CHECK();
  const int grid = grid_for((long long)N * H);
  for (int s = 0; s < T; ++s) {
    const int t = backwards ? T - 1 - s : s;
    int n_act = active_rows ? active_rows[t] : N;
    if (n_act < 0) n_act = 0;
    if (n_act > N) n_act = N;
    int n_rec = 0;
    if (n_act > 0) {
      // rec[n_act, 4H] = h[n_act, H] W_hid[H, 4H]; tiny tails fall back to the exact CUDA-core product on the fp32 state
      int rc;
      if (gemm_tc_f16_supported(n_act, 4 * H, H, w.h_hi, w.ldk, whid_hi, ldw))
        rc = gemm_tc_f16x3(0, 0, n_act, 4 * H, H, reinterpret_cast<const uint16_t*>(w.h_hi),
                           reinterpret_cast<const uint16_t*>(w.h_lo), w.ldk, w.exp_h, whid_hi, whid_lo, ldw, whid_exp,
                           w.rec, 4 * H, nullptr, IPAVSR_ACT_LINEAR, 0, nullptr, nullptr, nullptr, 0, st);
      else      // a handful of rows (the last frames of the longest utterances): exact product on the fp32 state
        rc = gemm_simt(0, 0, n_act, 4 * H, H, w.hbuf, H, w_hid, 4 * H, w.rec, 4 * H, nullptr, IPAVSR_ACT_LINEAR, 0, st);
      if (rc) return rc;
      n_rec = n_act;
    }
    steps_fwd_kernel<<<grid, 256, 0, st>>>(xw, w.rec, n_rec, peep, w.cbuf, w.hbuf, w.h_hi, w.h_lo, w.exp_h, mask, out, gates,
                                           cell, hprev, N, T, H, ldh, w.ldk, t);
    IPAVSR_LAUNCH_CHECK();
  }
  return IPAVSR_OK;
}

int ipavsr_lstm_bwd_f16_steps(const float* dout, const float* w_hid, const uint16_t* whid_hi, const uint16_t* whid_lo,
                              const int32_t* whid_exp, int ldw, const float* peep, const float* cell_init,
                              const uint8_t* mask, const float* gates, const float* cell, float* dgates, float* dpeep,
                              float* dcell_init, float* dhid_init, int N, int T, int H, int ldh, int backwards, float clip,
                              int accumulate, uint16_t* dg_hi, uint16_t* dg_lo, int32_t* dg_exp,
                              const int32_t* active_rows, void* workspace, uint64_t workspace_bytes, void* stream) {
  IPAVSR_CHECK_ARG(dout && w_hid && whid_hi && whid_lo && whid_exp && cell_init && mask && gates && cell && dgates &&
                       dcell_init && dhid_init && dg_hi && dg_lo && dg_exp,
                   "null pointer");
  IPAVSR_CHECK_ARG((peep == nullptr) == (dpeep == nullptr), "peep and dpeep go together");
  IPAVSR_CHECK_ARG(N >= 0 && T >= 1 && H >= 1 && ldh >= H, "bad sizes");
  if (!ipavsr_lstm_steps_supported(N, T, H, ldw) || !(clip > 0.f && clip < 16384.f) ||
      ((reinterpret_cast<uintptr_t>(whid_hi) | reinterpret_cast<uintptr_t>(whid_lo) | reinterpret_cast<uintptr_t>(dg_hi) |
        reinterpret_cast<uintptr_t>(dg_lo)) & 15) != 0) {
    set_error("ipavsr_lstm_bwd_f16_steps: needs 64 <= H <= 2048, N*4H*H >= 4e6, clip > 0, 16-byte aligned halves");
    return IPAVSR_ERR_UNSUPPORTED;
  }
  IPAVSR_CHECK_ARG(workspace && workspace_bytes >= ipavsr_lstm_steps_workspace_bytes(N, T, H), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) {
    IPAVSR_CUDA(cudaMemsetAsync(dcell_init, 0, sizeof(float) * H, st));
    IPAVSR_CUDA(cudaMemsetAsync(dhid_init, 0, sizeof(float) * H, st));
    if (dpeep) IPAVSR_CUDA(cudaMemsetAsync(dpeep, 0, sizeof(float) * 3 * H, st));
  }
  if (N == 0) return IPAVSR_OK;
  StepsWs w = carve(workspace, N, H);
  const size_t NH = (size_t)N * H;
  // |dg| <= clip: static scale 2^eg with clip * 2^eg < 2^15 (as lstm_tc.cu)
  const int eg = 14 - (ilogbf(clip) + 1);
  const float gs = ldexpf(1.0f, eg);
  set_int_kernel<<<1, 1, 0, st>>>(dg_exp, eg);
  IPAVSR_LAUNCH_CHECK();
  if (peep) IPAVSR_CUDA(cudaMemsetAsync(w.ppart, 0, sizeof(float) * 3 * NH, st));
  const int grid = grid_for((long long)N * H);
  const int t_first = backwards ? T - 1 : 0;
  int n_rec = 0;
  for (int s = T - 1; s >= 0; --s) {
    const int t = backwards ? T - 1 - s : s;
    const int t_prev = s == 0 ? -1 : (backwards ? t + 1 : t - 1);
    steps_bwd_kernel<<<grid, 256, 0, st>>>(dout, w.rec, n_rec, w.dh_pass, w.dc, peep, cell_init, mask, gates, cell, dgates,
                                           reinterpret_cast<__half*>(dg_hi), reinterpret_cast<__half*>(dg_lo), gs, w.ppart, N,
                                           T, H, ldh, t, t_prev, clip, s == T - 1);
    IPAVSR_LAUNCH_CHECK();
    n_rec = 0;
    if (s > 0) {
      int n_act = active_rows ? active_rows[t] : N;
      if (n_act < 0) n_act = 0;
      if (n_act > N) n_act = N;
      if (n_act > 0) {
        // dh_rec[n_act, H] = dg_t[n_act, 4H] (rows of frame t: stride T*4H) W_hid^T
        const size_t off = (size_t)t * 4 * H;
        int rc;
        if (gemm_tc_f16_supported(n_act, H, 4 * H, dg_hi + off, T * 4 * H, whid_hi, ldw))
          rc = gemm_tc_f16x3(0, 1, n_act, H, 4 * H, dg_hi + off, dg_lo + off, T * 4 * H, dg_exp, whid_hi, whid_lo, ldw,
                             whid_exp, w.rec, H, nullptr, IPAVSR_ACT_LINEAR, 0, nullptr, nullptr, nullptr, 0, st);
        else
          rc = gemm_simt(0, 1, n_act, H, 4 * H, dgates + off, T * 4 * H, w_hid, 4 * H, w.rec, H, nullptr, IPAVSR_ACT_LINEAR,
                         0, st);
        if (rc) return rc;
        n_rec = n_act;
      }
    }
  }
  if (peep)
    for (int k = 0; k < 3; ++k) {
      int rc = ipavsr_colsum(w.ppart + k * NH, H, dpeep + k * H, N, H, 1, stream);
      if (rc) return rc;
    }
  // d(hid_init) = sum_n [ dg_first W_hid^T + dh_pass ],  d(cell_init) = sum_n dc_prev at the first processed step; the
  // product is linear, so it is taken after the sum over the utterances (one column sum + a vector-matrix product)
  int rc;
  const size_t off_first = (size_t)t_first * 4 * H;
  float* dg_sum = w.rec;              // free once the steps are done
  rc = ipavsr_colsum(dgates + off_first, T * 4 * H, dg_sum, N, 4 * H, 0, stream);
  if (rc) return rc;
  rc = ipavsr_colsum(w.dh_pass, H, dhid_init, N, H, 1, stream);
  if (rc) return rc;
  rc = gemv_rows(w_hid, 4 * H, dg_sum, dhid_init, H, 4 * H, 1, st);
  if (rc) return rc;
  return ipavsr_colsum(w.dc, H, dcell_init, N, H, 1, stream);
}

}  // extern "C"
