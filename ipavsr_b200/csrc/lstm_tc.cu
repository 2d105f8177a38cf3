// Masked Lasagne-LSTM recurrence on the tensor cores (forward), fp32-parity three-product fp16 arithmetic.
//
// Same semantics as lstm.cu (reference custom/layers.py:10-80 -> lasagne LSTMLayer); this is
// the fast path for H <= 256 when the fp16 hi/lo split of W_hid is available (engine mode f16x3).  The FFMA kernel in
// lstm.cu is bound by the shared-memory -> register return path (13-16 us per time step); here the per-step product
//     g^T[128 gate columns, 32 utterances] = W_hid^T[128, K] * h^T[K, 32]
// is 48 tcgen05.mma (kind::f16, M=128, N=32, K=16; lo*hi + hi*lo -> cross accumulator, hi*hi -> main accumulator, as in
// gemm_tc.cu) issued by one thread, with the accumulators in TMEM.
//
// A cluster of CS = ceil(H/32) CTAs owns a tile of 32 utterances for all T steps.  CTA r owns hidden units
// [32r, 32r+32), i.e. gate columns [128r, 128r+128) of the gate-interleaved W_hid:
//   * its W_hid^T slice (fp16 hi and lo, 2 x 64 KB at K = 256) is loaded ONCE by TMA, straight from the engine's fp16
//     split of the parameter arena (MN-major SWIZZLE_128B operand: no transposed copy), and stays in shared memory;
//   * h_t (fp16 hi/lo, scale 2^eH with eH from max(1, |hid_init|)) lives in a double-buffered K-major SWIZZLE_64B
//     operand tile; after the cell update every CTA converts its 32 units x 32 utterances of h_{t+1} into that layout,
//     writes it straight into its own slot of its next-step tile and into a staging buffer that it pushes into the
//     next-step tile of the CS-1 OTHER CTAs with cp.async.bulk (shared::cta -> shared::cluster), completing on the
//     destination's mbarrier (complete_tx) — the MMA thread of each CTA waits for (CS-1) x 4 KB to land plus one arrival of
//     its own epilogue: there is no cluster barrier in the time loop.  (The bulk copy is specified for REMOTE destinations
//     only: compute-sanitizer flagged the earlier self-push, profiles/r02_sanitizer_summary.md.)
//   * frames at which every utterance of the tile is masked are skipped (no MMAs, no exchange): with the engine's
//     length-sorted batches a tile runs max(len) steps instead of T;
//   * 128 epilogue threads own one TMEM lane each (gate column 4u+g): tcgen05.ld, add the hoisted input projection
//     xw[t] (coalesced: a warp reads 32 consecutive gate columns of one frame), a 4x4 shuffle transpose brings the four
//     gates of a (unit, utterance) cell into one thread, which then owns 8 cells.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
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

constexpr int QN = 32;            // utterances per cluster tile (MMA N)
constexpr int QU = 32;            // hidden units per CTA (128 gate columns = MMA M)
constexpr int QTHREADS = 288;     // warps 0-7: epilogue (TMEM lane quadrant w%4, utterance half w/4); warp 8: TMA / MMA
constexpr int QEPI = 256;         // epilogue threads
constexpr int QC = 16;            // utterances (TMEM columns) per epilogue thread
// The tensor core adds into its fp32 TMEM accumulator with truncation (gemm_tc.cu), a bias that compounds over the T
// recurrent steps.  The hi*hi products are therefore spread over QACC accumulators (two k-steps = 32 hidden units
// each at H <= 256) that the epilogue sums with round-to-nearest; the tiny cross terms share one accumulator.
constexpr int QACC = 8;
constexpr int QCROSS = 2;         // cross-term accumulators (alternating k-steps): halves the dependent MMA chain
constexpr int QTMEM_COLS = 512;   // (QACC + QCROSS) * 32 = 320 columns, allocated as the next power of two

__device__ __forceinline__ uint32_t q_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void q_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(q_smem(bar)), "r"(count));
}
__device__ __forceinline__ void q_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(q_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void q_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "Q_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra Q_DONE;\n"
      "bra Q_WAIT;\n"
      "Q_DONE:\n"
      "}\n" ::"r"(q_smem(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t q_mapa(uint32_t addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
  return out;
}
__device__ __forceinline__ uint32_t q_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t q_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void q_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void q_tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          q_smem(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(q_smem(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared::cta -> (remote) shared::cluster bulk copy, completing `bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void q_bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void q_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void q_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(q_smem(bar)) : "memory");
}
__device__ __forceinline__ void q_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void q_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void q_tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void q_tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (see gemm_tc.cu): layout 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t q_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// byte offset of element (row = utterance j, k-in-chunk c in 0..31) inside one 32 x 64-byte SWIZZLE_64B chunk
__device__ __forceinline__ uint32_t q_sw64(int j, int c) {
  return (uint32_t)(j * 64 + ((((c >> 3) ^ ((j >> 1) & 3)) & 3) << 4) + ((c & 7) << 1));
}

// sigmoid / tanh from ex2.approx + rcp.approx (<= 2 ulp each): ~3e-7 relative for sigmoid; tanh as 1 - 2 / (e^{2x} + 1)
// with an odd polynomial below |x| = 0.15 (where that form cancels): absolute error < 2e-7 everywhere.  Saturates
// correctly at +-inf.  (The FFMA kernel in lstm.cu keeps libm's expf / tanhf.)
__device__ __forceinline__ float q_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float q_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float q_sigmoid(float x) { return q_rcp(1.0f + q_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float q_tanh(float x) {
  const float x2 = x * x;
  const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.053968254f, 0.133333333f), -0.333333333f), 1.0f);
  const float big = 1.0f - 2.0f * q_rcp(1.0f + q_ex2(2.8853900817779268f * x));
  return fabsf(x) < 0.15f ? poly : big;
}

// bit t of the result: some utterance of the 32-utterance tile is unmasked at frame t (T <= 64); called by one full warp
__device__ __forceinline__ unsigned long long q_tile_active(const uint8_t* __restrict__ mask, int tile, int N, int T,
                                                            int lane) {
  const int n = tile * QN + lane;
  unsigned long long bits = 0ull;
  if (n < N)
    for (int t = 0; t < T; ++t)
      if (mask[(size_t)n * T + t]) bits |= 1ull << t;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
  return bits;
}
// the next processed step after s whose frame is active, or T
__device__ __forceinline__ int q_next_active(unsigned long long active, int s, int T, int backwards) {
  for (int s2 = s + 1; s2 < T; ++s2)
    if ((active >> (backwards ? (T - 1 - s2) : s2)) & 1ull) return s2;
  return T;
}

struct CellF {
  float i, f, cin, o, c, h;
};
__device__ __forceinline__ CellF q_cell(float gi, float gf, float gc, float go, float c_prev, float h_prev, bool m,
                                        bool has_peep, float w_ci, float w_cf, float w_co) {
  CellF r;
  if (has_peep) {
    gi = fmaf(c_prev, w_ci, gi);
    gf = fmaf(c_prev, w_cf, gf);
  }
  r.i = q_sigmoid(gi);
  r.f = q_sigmoid(gf);
  r.cin = q_tanh(gc);
  const float c_u = r.f * c_prev + r.i * r.cin;
  if (has_peep) go = fmaf(c_u, w_co, go);
  r.o = q_sigmoid(go);
  const float h_u = r.o * q_tanh(c_u);
  r.c = m ? c_u : c_prev;
  r.h = m ? h_u : h_prev;
  return r;
}

__device__ __forceinline__ long long q_clock() { return clock64(); }
}  // namespace

unsigned long long* g_lstm_dbg = nullptr;      // profiling aid (ipavsr_debug_lstm_timestamps)

__global__ void __launch_bounds__(QTHREADS, 1)
lstm_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo,
                   const float* __restrict__ xw, const int32_t* __restrict__ w_exp, const float* __restrict__ peep,
                   const float* __restrict__ cell_init, const float* __restrict__ hid_init,
                   const uint8_t* __restrict__ mask, float* __restrict__ out, float* __restrict__ gates,
                   float* __restrict__ cell, float* __restrict__ hprev, int N, int T, int H, int ldh, int backwards,
                   unsigned long long* dbg) {
  const int CS = (int)q_cluster_size();
  const int rank = (int)q_cluster_rank();
  const int tile = blockIdx.x / CS;
  const int H4 = 4 * H;
  const int KB = (CS * QU + 63) / 64;                  // 64-row k-boxes of the W operand
  const int KSTEPS = CS * 2;                           // MMA k-steps of 16
  const uint32_t WBYTES = (uint32_t)KB * 16384u;       // one W array (hi or lo): KB boxes x {2 m-boxes x 8 KB}
  const uint32_t HBYTES = (uint32_t)CS * 2048u;        // one h array (hi or lo) of one buffer: CS chunks of 32 x 64 B

  extern __shared__ uint8_t q_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(q_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sWhi = smem;
  uint8_t* sWlo = sWhi + WBYTES;
  uint8_t* sH = sWlo + WBYTES;                         // [2 buffers][hi | lo][CS chunks][2048]
  uint8_t* sStage = sH + 4 * HBYTES;                   // [2][hi 2048 | lo 2048]
  __shared__ __align__(8) uint64_t hbar[2], accbar, wbar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_red[4];
  __shared__ int s_eh;
  __shared__ unsigned long long s_active;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    // two arrivals per phase: the control thread's expect_tx (bytes pushed by the OTHER CTAs) and the epilogue's arrival
    // after it has written this CTA's own chunk of h straight into the tile
    q_mbar_init(&hbar[0], 2);
    q_mbar_init(&hbar[1], 2);
    q_mbar_init(&accbar, 1);
    q_mbar_init(&wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(q_smem(&tmem_base_smem)),
                 "r"((uint32_t)QTMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // scale exponent of the h operand: |h_t| <= max(1, max|hid_init|)
  if (warp < 4) {
    float mx = 1.0f;
    for (int k = tid; k < H; k += 128) mx = fmaxf(mx, fabsf(hid_init[k]));
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
  }
  // frames at which at least one utterance of the tile is unmasked; at the other frames every state is carried through
  // (Lasagne's switch) and the step needs neither the recurrent product nor the exchange of h.  Every CTA of the
  // cluster derives the same bit set from the same 32 mask rows.
  if (warp == 4) {
    const unsigned long long bits = q_tile_active(mask, tile, N, T, lane);
    if (lane == 0) s_active = bits;
  }
  q_fence_before();
  __syncthreads();
  q_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const unsigned long long active = s_active;
  if (tid == 0) {
    const float mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    s_eh = 14 - ilogbf(mx);
    // W_hid^T slice: gate columns [128 rank, +128) x all k, straight from the fp16 split of the parameter arena
    q_mbar_expect_tx(&wbar, 2 * WBYTES);
    for (int kb = 0; kb < KB; ++kb)
      for (int mb = 0; mb < 2; ++mb) {
        q_tma_load_2d(sWhi + (size_t)(kb * 2 + mb) * 8192, &mapWhi, &wbar, rank * 128 + mb * 64, kb * 64);
        q_tma_load_2d(sWlo + (size_t)(kb * 2 + mb) * 8192, &mapWlo, &wbar, rank * 128 + mb * 64, kb * 64);
      }
  }
  __syncthreads();
  const int eh = s_eh;
  const float hs = __int_as_float((127 + eh) << 23);   // 2^eH
  // h_0 = hid_init for every utterance of the tile (buffer 0, all CS chunks), written by every CTA for itself
  {
    const int Kpad = CS * QU;
    for (int i = tid; i < QN * Kpad; i += QTHREADS) {
      const int j = i / Kpad, k = i - j * Kpad;
      const float xs = (k < H ? hid_init[k] : 0.f) * hs;
      const __half hi = __float2half_rn(xs);
      const __half lo = __float2half_rn((xs - __half2float(hi)) * F16_LO_SCALE);
      const uint32_t off = (uint32_t)(k >> 5) * 2048u + q_sw64(j, k & 31);
      *reinterpret_cast<__half*>(sH + off) = hi;
      *reinterpret_cast<__half*>(sH + HBYTES + off) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  q_cluster_sync();      // every CTA's barriers are initialised before any remote copy can complete on them

  if (warp == 8) {
    // ===================== control warp: one thread issues the MMAs =====================
    if (lane == 0) {
      // M = 128, N = 32, A MN-major, B K-major, fp16 x fp16 -> fp32
      constexpr uint32_t idesc = (1u << 4) | (1u << 15) | ((uint32_t)(QN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      q_mbar_wait(&wbar, 0);
      uint32_t hphase[2] = {0, 0};
      long long c_wait = 0, c_issue = 0;
      for (int s = q_next_active(active, -1, T, backwards), k = 0; s < T; s = q_next_active(active, s, T, backwards), ++k) {
        const int b = k & 1;               // k counts the active steps: buffers and barrier phases alternate on it
        const long long c0 = q_clock();
        if (k > 0) {
          q_mbar_expect_tx(&hbar[b], (uint32_t)(CS - 1) * 4096u);
          q_mbar_wait(&hbar[b], hphase[b]);
          hphase[b] ^= 1;
        }
        q_fence_after();
        const long long c1 = q_clock();
        c_wait += c1 - c0;
        // descriptors differ from the k-step-0 ones only in the 14-bit start-address field (units of 16 bytes)
        const uint64_t a_hi0 = q_desc(q_smem(sWhi), 8192, 1024, 2), a_lo0 = q_desc(q_smem(sWlo), 8192, 1024, 2);
        const uint32_t bH = q_smem(sH + (size_t)b * 2 * HBYTES);
        const uint64_t b_hi0 = q_desc(bH, 16, 512, 4), b_lo0 = q_desc(bH + HBYTES, 16, 512, 4);
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          if (ks < KSTEPS) {
            const uint64_t a_off = (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 128);     // 16 KB per k-box, 2 KB per k-step
            const uint64_t b_off = (uint64_t)((ks >> 1) * 128 + (ks & 1) * 2);        // 2 KB per chunk, 32 B per k-step
            const uint32_t tc = tmem_base + (QACC + (ks % QCROSS)) * 32;
            q_mma_f16(tc, a_lo0 + a_off, b_hi0 + b_off, idesc, ks < QCROSS ? 0u : 1u);     // cross terms
            q_mma_f16(tc, a_hi0 + a_off, b_lo0 + b_off, idesc, 1u);
            q_mma_f16(tmem_base + ((ks >> 1) % QACC) * 32, a_hi0 + a_off, b_hi0 + b_off, idesc,
                      (ks < 2 * QACC && (ks & 1) == 0) ? 0u : 1u);                         // main
          }
        }
        q_commit(&accbar);
        c_issue += q_clock() - c1;
      }
      if (dbg != nullptr && blockIdx.x == 0) { dbg[0] = (unsigned long long)c_wait; dbg[1] = (unsigned long long)c_issue; }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: one TMEM lane (gate column) x 16 utterances per thread =====================
    const int q4 = warp & 3, half = warp >> 2;         // TMEM lane quadrant, utterance half of this warp
    const int m = q4 * 32 + lane;                      // local gate column = TMEM lane
    const int ul = m >> 2, g = m & 3;                  // local unit, gate
    const int ug = rank * QU + ul;
    const bool u_ok = ug < H;
    const int gcol = rank * 128 + m;                   // global gate column (4 ug + g)
    const bool col_ok = gcol < H4;
    const int j0 = half * QC;                          // first utterance (TMEM column) of this thread
    const bool has_peep = peep != nullptr;
    const float w_ci = (has_peep && u_ok) ? peep[ug] : 0.f;
    const float w_cf = (has_peep && u_ok) ? peep[H + ug] : 0.f;
    const float w_co = (has_peep && u_ok) ? peep[2 * H + ug] : 0.f;
    // result scale: (main + cross 2^-11) 2^-(eW + eH)
    const int e = -(__ldg(w_exp) + eh);
    const int e1 = (e > 126 || e < -115) ? e / 2 : e, e2 = e - e1;
    const float ms1 = __int_as_float((127 + e1) << 23), ms2 = __int_as_float((127 + e2) << 23);
    // the 4 cells this thread owns: utterances j0 + 4 i + g of the tile, unit ug
    float c_prev[4], h_prev[4];
    unsigned long long mbits[4];                       // mask of the utterance, bit t
    int n_own[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      c_prev[i] = u_ok ? cell_init[ug] : 0.f;
      h_prev[i] = u_ok ? hid_init[ug] : 0.f;
      n_own[i] = tile * QN + j0 + 4 * i + g;
      mbits[i] = 0ull;
      if (n_own[i] < N)
        for (int t = 0; t < T; ++t)
          if (mask[(size_t)n_own[i] * T + t]) mbits[i] |= 1ull << t;
    }
    // hoisted input projection of the current step, for the 16 utterances of this thread's gate column
    float xv[QC];
    auto fetch = [&](int t) {
#pragma unroll
      for (int j = 0; j < QC; ++j) {
        const int n = tile * QN + j0 + j;
        xv[j] = (col_ok && n < N) ? __ldg(xw + ((size_t)n * T + t) * H4 + gcol) : 0.f;
      }
    };
    {
      const int s0 = q_next_active(active, -1, T, backwards);
      if (s0 < T) fetch(backwards ? (T - 1 - s0) : s0);
    }
    long long e_wait = 0, e_ld = 0, e_math = 0, e_push = 0;
    const int nacc = (KSTEPS / 2) < QACC ? (KSTEPS / 2) : QACC;
    int k = 0;                                         // active steps done so far
    for (int s = 0; s < T; ++s) {
      const int t = backwards ? (T - 1 - s) : s;
      if (!((active >> t) & 1ull)) {
        // every utterance of the tile is masked at this frame: the states pass through unchanged
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (u_ok && n_own[i] < N) {
            const size_t row = (size_t)n_own[i] * T + t;
            out[row * ldh + ug] = h_prev[i];
            if (gates) *reinterpret_cast<float4*>(gates + row * H4 + 4 * ug) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cell) cell[row * H + ug] = c_prev[i];
            if (hprev) hprev[row * ldh + ug] = h_prev[i];
          }
        continue;
      }
      const int s_next = q_next_active(active, s, T, backwards);
      const long long k0 = q_clock();
      q_mbar_wait(&accbar, (uint32_t)(k & 1));
      q_fence_after();
      const long long k1 = q_clock();
      e_wait += k1 - k0;
      float a[QC];
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)j0;
        float c0[QC], c1[QC], c2[QC], c3[QC];
        q_tmem_ld16(taddr + QACC * 32, c0);              // cross terms (smallest) ...
        q_tmem_ld16(taddr + (QACC + 1) * 32, c1);
        q_tmem_ld16(taddr, c2);                          // ... and the first two main accumulators
        q_tmem_ld16(taddr + 32, c3);
        q_tmem_wait();
#pragma unroll
        for (int j = 0; j < QC; ++j) {
          const float cr = KSTEPS >= QCROSS ? c0[j] + c1[j] : c0[j];
          a[j] = fmaf(cr, F16_LO_INV, c2[j]);
          if (nacc > 1) a[j] += c3[j];
        }
#pragma unroll 1
        for (int q = 2; q + 1 < nacc; q += 2) {          // two more accumulators per TMEM round trip
          q_tmem_ld16(taddr + q * 32, c0);
          q_tmem_ld16(taddr + (q + 1) * 32, c1);
          q_tmem_wait();
#pragma unroll
          for (int j = 0; j < QC; ++j) a[j] += c0[j] + c1[j];
        }
        if ((nacc & 1) && nacc > 2) {
          q_tmem_ld16(taddr + (nacc - 1) * 32, c0);
          q_tmem_wait();
#pragma unroll
          for (int j = 0; j < QC; ++j) a[j] += c0[j];
        }
      }
      q_fence_before();
      const long long k2 = q_clock();
      e_ld += k2 - k1;
#pragma unroll
      for (int j = 0; j < QC; ++j) a[j] = a[j] * ms1 * ms2 + xv[j];
      if (s_next < T) fetch(backwards ? (T - 1 - s_next) : s_next);
      // 4x4 transpose inside each group of 4 lanes (the 4 gates of a unit): afterwards a[4i + q] = gate q of utterance j0 + 4i + g
      const bool b0 = (g & 1) != 0, b1 = (g & 2) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x0 = a[4 * i], x1 = a[4 * i + 1], x2 = a[4 * i + 2], x3 = a[4 * i + 3];
        float r0 = __shfl_xor_sync(0xffffffffu, b0 ? x0 : x1, 1);
        float r1 = __shfl_xor_sync(0xffffffffu, b0 ? x2 : x3, 1);
        if (b0) { x0 = r0; x2 = r1; } else { x1 = r0; x3 = r1; }
        r0 = __shfl_xor_sync(0xffffffffu, b1 ? x0 : x2, 2);
        r1 = __shfl_xor_sync(0xffffffffu, b1 ? x1 : x3, 2);
        if (b1) { x0 = r0; x1 = r1; } else { x2 = r0; x3 = r1; }
        a[4 * i] = x0; a[4 * i + 1] = x1; a[4 * i + 2] = x2; a[4 * i + 3] = x3;
      }
      uint8_t* stg = sStage + (size_t)(k & 1) * 4096;
      // this CTA's own slot of the next-step operand tile (buffer (k+1)&1 was last read by the MMAs of step k-1, which
      // completed before the accumulator of step k did): written directly, the other CTAs get the staged copy pushed
      uint8_t* own = sH + (size_t)((k + 1) & 1) * 2 * HBYTES + (size_t)rank * 2048;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool mk = (mbits[i] >> t) & 1ull;
        const CellF r = q_cell(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3], c_prev[i], h_prev[i], mk, has_peep, w_ci,
                               w_cf, w_co);
        if (u_ok && n_own[i] < N) {
          const size_t row = (size_t)n_own[i] * T + t;
          out[row * ldh + ug] = r.h;
          if (gates) *reinterpret_cast<float4*>(gates + row * H4 + 4 * ug) = make_float4(r.i, r.f, r.cin, r.o);
          if (cell) cell[row * H + ug] = r.c;
          if (hprev) hprev[row * ldh + ug] = h_prev[i];
        }
        c_prev[i] = r.c;
        h_prev[i] = r.h;
        // h_{t+1} of (utterance j0 + 4i + g, unit ul) into the staging chunk, fp16 hi/lo, SWIZZLE_64B K-major layout
        const float xs = (u_ok ? r.h : 0.f) * hs;
        const __half hi = __float2half_rn(xs);
        const __half lo = __float2half_rn((xs - __half2float(hi)) * F16_LO_SCALE);
        const uint32_t off = q_sw64(j0 + 4 * i + g, ul);
        *reinterpret_cast<__half*>(stg + off) = hi;
        *reinterpret_cast<__half*>(stg + 2048 + off) = lo;
        *reinterpret_cast<__half*>(own + off) = hi;
        *reinterpret_cast<__half*>(own + HBYTES + off) = lo;
      }
      const long long k3 = q_clock();
      e_math += k3 - k2;
      if (s_next < T) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // push my chunk of h_{t+1} into the next-step operand tile of every OTHER CTA of the cluster (bulk shared::cta ->
        // shared::cluster copies completing on the destination's barrier); my own tile already holds it: one arrival
        const int nb = (k + 1) & 1;
        if (tid < CS && tid != rank) {
          const uint32_t dst_hi = q_smem(sH + (size_t)nb * 2 * HBYTES + (size_t)rank * 2048);
          const uint32_t bar = q_mapa(q_smem(&hbar[nb]), (uint32_t)tid);
          q_bulk_s2s(q_mapa(dst_hi, (uint32_t)tid), q_smem(stg), 2048, bar);
          q_bulk_s2s(q_mapa(dst_hi + HBYTES, (uint32_t)tid), q_smem(stg + 2048), 2048, bar);
        } else if (tid == rank) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(q_smem(&hbar[nb])) : "memory");
        }
      }
      e_push += q_clock() - k3;
      ++k;
    }
    if (dbg != nullptr && blockIdx.x == 0 && tid == 0) {
      dbg[2] = (unsigned long long)e_wait; dbg[3] = (unsigned long long)e_ld;
      dbg[4] = (unsigned long long)e_math; dbg[5] = (unsigned long long)e_push;
    }
  }
  q_fence_before();
  q_cluster_sync();      // no CTA leaves while copies into / out of its shared memory may still be in flight
  if (warp == 8) {
    q_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)QTMEM_COLS)
                 : "memory");
  }
}

// =========================================================================================================
// Backward (BPTT) on the tensor cores.  CTA r owns units U_r = [32r, 32r+32) and their gate columns J_r.  Per processed
// step (in reverse order):
//   1. 256 threads run the cell backward for the CTA's 32 units x 32 utterances (4 cells each): the recurrent gradient
//      dh_next is the sum of the CS partial blocks that landed in the inbox, dg (clipped) goes to global memory and,
//      as fp16 hi/lo (static scale: |dg| <= clip), into a K-major SWIZZLE_128B operand tile [32 utterances x 128 j];
//   2. one thread issues  D[k, n] = sum_{j in J_r} W_hid[k, j] dg[n, j]  for ALL k: 2 row blocks x 8 k-steps x 3 products,
//      A = the CTA's column slice of W_hid (K-major, resident, TMA-loaded from the arena split);
//   3. every thread reads its k row (32 utterances) from TMEM and stores it into the inbox of the CTA that owns unit k
//      (distributed shared memory, 8 x 16-byte stores), i.e. a reduce-scatter of the partial products;
//   4. one cluster barrier makes the inboxes visible.
// =========================================================================================================
struct CellB {
  float dgi, dgf, dgc, dgo, dc_prev, dh_pass, pci, pcf, pco;
};
__device__ __forceinline__ CellB q_cell_bwd(float dh, float dc, float i, float f, float cin, float o, float c, float c_prev,
                                            bool m, bool has_peep, float w_ci, float w_cf, float w_co, float clip) {
  CellB g;
  if (!m) {
    g.dgi = g.dgf = g.dgc = g.dgo = 0.f;
    g.dc_prev = dc;
    g.dh_pass = dh;
    g.pci = g.pcf = g.pco = 0.f;
    return g;
  }
  const float tc = q_tanh(c);
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

constexpr int BACC = 4;           // main accumulators per row block (2 k-steps each at K = 128)
constexpr int BCOLS = (BACC + QCROSS) * 32;      // TMEM columns per row block

__global__ void __launch_bounds__(QTHREADS, 1)
lstm_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo,
                   const float* __restrict__ dout, const int32_t* __restrict__ w_exp, const float* __restrict__ peep,
                   const float* __restrict__ cell_init, const uint8_t* __restrict__ mask, const float* __restrict__ gates,
                   const float* __restrict__ cell, float* __restrict__ dgates, float* __restrict__ dpeep,
                   float* __restrict__ dc_fin, float* __restrict__ dh_fin, int N, int T, int H, int ldh, int backwards,
                   float clip, int eg, float* __restrict__ db, uint16_t* __restrict__ dg_hi, uint16_t* __restrict__ dg_lo,
                   int32_t* __restrict__ dg_exp) {
  const int CS = (int)q_cluster_size();
  const int rank = (int)q_cluster_rank();
  const int tile = blockIdx.x / CS;
  const int H4 = 4 * H;
  const int MB = (CS * QU + 127) / 128;                // 128-row blocks of the k dimension (1 or 2)

  extern __shared__ uint8_t q_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(q_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sWhi = smem;                                // [MB][2 j-chunks][128 rows x 128 B]
  uint8_t* sWlo = sWhi + (size_t)MB * 32768;
  uint8_t* sB = sWlo + (size_t)MB * 32768;             // dg operand: [hi | lo][2 j-chunks][32 rows x 128 B]
  float* inbox = reinterpret_cast<float*>(sB + 16384); // [2 parities][CS sources][32 units][32 utterances]
  __shared__ __align__(8) uint64_t accbar, wbar, bready;
  __shared__ uint32_t tmem_base_smem;
  __shared__ unsigned long long s_active;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    q_mbar_init(&accbar, 1);
    q_mbar_init(&wbar, 1);
    q_mbar_init(&bready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(q_smem(&tmem_base_smem)),
                 "r"((uint32_t)QTMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // frames at which some utterance of the tile is unmasked (as in the forward kernel): at the others the gate gradients
  // are zero and the state gradients pass through, so the step needs no recurrent product and no cluster exchange
  if (warp == 4) {
    const unsigned long long bits = q_tile_active(mask, tile, N, T, lane);
    if (lane == 0) s_active = bits;
  }
  q_fence_before();
  __syncthreads();
  q_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const unsigned long long active = s_active;
  if (tid == 0) {
    // W_hid[:, J_r]: all k rows x the CTA's 128 gate columns, K-major (j contiguous), from the fp16 split of the arena
    q_mbar_expect_tx(&wbar, (uint32_t)MB * 65536u);
    for (int mb = 0; mb < MB; ++mb)
      for (int jc = 0; jc < 2; ++jc) {
        q_tma_load_2d(sWhi + (size_t)(mb * 2 + jc) * 16384, &mapWhi, &wbar, rank * 128 + jc * 64, mb * 128);
        q_tma_load_2d(sWlo + (size_t)(mb * 2 + jc) * 16384, &mapWlo, &wbar, rank * 128 + jc * 64, mb * 128);
      }
  }
  q_cluster_sync();

  if (warp == 8) {
    // ===================== control warp =====================
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(QN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // K-major A and B
    if (lane == 0) q_mbar_wait(&wbar, 0);
    for (int s = T - 1, k = 0; s >= 1; --s) {
      if (!((active >> (backwards ? (T - 1 - s) : s)) & 1ull)) continue;
      if (lane == 0) {
        q_mbar_wait(&bready, (uint32_t)(k & 1));
        q_fence_after();
        const uint64_t b_hi0 = q_desc(q_smem(sB), 16, 1024, 2), b_lo0 = q_desc(q_smem(sB + 8192), 16, 1024, 2);
        for (int mb = 0; mb < MB; ++mb) {
          const uint64_t a_hi0 = q_desc(q_smem(sWhi + (size_t)mb * 32768), 16, 1024, 2);
          const uint64_t a_lo0 = q_desc(q_smem(sWlo + (size_t)mb * 32768), 16, 1024, 2);
          const uint32_t tb = tmem_base + (uint32_t)(mb * BCOLS);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a_off = (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2);      // 16 KB per j-chunk, 32 B per k-step
            const uint64_t b_off = (uint64_t)((ks >> 2) * 256 + (ks & 3) * 2);       // 4 KB per j-chunk
            const uint32_t tc = tb + (BACC + (ks % QCROSS)) * 32;
            q_mma_f16(tc, a_lo0 + a_off, b_hi0 + b_off, idesc, ks < QCROSS ? 0u : 1u);
            q_mma_f16(tc, a_hi0 + a_off, b_lo0 + b_off, idesc, 1u);
            q_mma_f16(tb + (ks >> 1) * 32, a_hi0 + a_off, b_hi0 + b_off, idesc, (ks & 1) ? 1u : 0u);
          }
        }
        q_commit(&accbar);
      }
      __syncwarp();
      q_cluster_sync();
      ++k;
    }
  } else {
    // ===================== 256 worker threads =====================
    // cell mapping: unit ul = tid >> 3, utterances n = 4 (tid & 7) + i (one float4 of an inbox row)
    const int ul = tid >> 3, nq = tid & 7;
    const int ug = rank * QU + ul;
    const bool u_ok = ug < H;
    // TMEM mapping: row block mbk = warp >> 2, lane quadrant warp & 3 -> k row; its owner CTA is the same for the warp
    const int mbk = warp >> 2;
    const int krow = mbk * 128 + (warp & 3) * 32 + lane;
    const int owner = krow >> 5;
    const bool has_peep = peep != nullptr;
    const float w_ci = (has_peep && u_ok) ? peep[ug] : 0.f;
    const float w_cf = (has_peep && u_ok) ? peep[H + ug] : 0.f;
    const float w_co = (has_peep && u_ok) ? peep[2 * H + ug] : 0.f;
    const int e = -(__ldg(w_exp) + eg);
    const int e1 = (e > 126 || e < -115) ? e / 2 : e, e2 = e - e1;
    const float ms1 = __int_as_float((127 + e1) << 23), ms2 = __int_as_float((127 + e2) << 23);
    const float gs = __int_as_float((127 + eg) << 23);          // 2^eG: scale of the dg operand
    int ng[4];
    bool n_ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ng[i] = tile * QN + 4 * nq + i;
      n_ok[i] = ng[i] < N;
    }
    float dh_next[4] = {0.f, 0.f, 0.f, 0.f}, dc_next[4] = {0.f, 0.f, 0.f, 0.f}, dh_pass[4] = {0.f, 0.f, 0.f, 0.f};
    float pci = 0.f, pcf = 0.f, pco = 0.f;
    float4 dbs = make_float4(0.f, 0.f, 0.f, 0.f);      // bias gradient of my unit's 4 gates over my utterances and all steps
    if (dg_exp != nullptr && blockIdx.x == 0 && tid == 0) *dg_exp = eg;
    float pf_dout[4], pf_c[4], pf_cp[4];
    float4 pf_g[4];
    bool pf_m[4];
    auto fetch = [&](int s) {
      const int t = backwards ? (T - 1 - s) : s;
      const int t_prev = (s == 0) ? -1 : (backwards ? t + 1 : t - 1);
      const bool act = (active >> t) & 1ull;           // an inactive frame only needs its dout
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pf_dout[i] = pf_c[i] = pf_cp[i] = 0.f;
        pf_g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pf_m[i] = false;
        if (n_ok[i] && u_ok) {
          const size_t row = (size_t)ng[i] * T + t;
          pf_dout[i] = __ldg(dout + row * ldh + ug);
          if (!act) continue;
          pf_g[i] = __ldg(reinterpret_cast<const float4*>(gates + row * H4 + 4 * ug));
          pf_c[i] = __ldg(cell + row * H + ug);
          pf_cp[i] = t_prev < 0 ? cell_init[ug] : __ldg(cell + ((size_t)ng[i] * T + t_prev) * H + ug);
          pf_m[i] = mask[row] != 0;
        }
      }
    };
    fetch(T - 1);
    int par = 0, k = 0;                                // k counts the active steps with a recurrent product (s > 0)
    for (int s = T - 1; s >= 0; --s) {
      const int t = backwards ? (T - 1 - s) : s;
      if (!((active >> t) & 1ull)) {
        // tile-wide masked frame: dg = 0, the hidden-state gradient collects this frame's dout and passes through
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dh_pass[i] = 0.f;
          if (n_ok[i] && u_ok) {
            dh_pass[i] = pf_dout[i] + dh_next[i];
            dh_next[i] = dh_pass[i];
            const size_t o = ((size_t)ng[i] * T + t) * H4 + 4 * ug;
            *reinterpret_cast<float4*>(dgates + o) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dg_hi != nullptr) {
              *reinterpret_cast<uint2*>(dg_hi + o) = make_uint2(0u, 0u);
              *reinterpret_cast<uint2*>(dg_lo + o) = make_uint2(0u, 0u);
            }
          }
        }
        if (s == 0) break;
        fetch(s - 1);
        continue;
      }
      // ---- 1. cell backward for my 4 cells; dg -> global (fp32) and -> the fp16 operand tile ----
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
        dh_pass[i] = 0.f;
        if (n_ok[i] && u_ok) {
          const size_t row = (size_t)ng[i] * T + t;
          const CellB g = q_cell_bwd(pf_dout[i] + dh_next[i], dc_next[i], pf_g[i].x, pf_g[i].y, pf_g[i].z, pf_g[i].w,
                                     pf_c[i], pf_cp[i], pf_m[i], has_peep, w_ci, w_cf, w_co, clip);
          dg = make_float4(g.dgi, g.dgf, g.dgc, g.dgo);
          dc_next[i] = g.dc_prev;
          dh_pass[i] = g.dh_pass;
          pci += g.pci; pcf += g.pcf; pco += g.pco;
          *reinterpret_cast<float4*>(dgates + row * H4 + 4 * ug) = dg;
          dbs.x += dg.x; dbs.y += dg.y; dbs.z += dg.z; dbs.w += dg.w;
        }
        // operand tile: row n = 4 nq + i, columns j = 4 ul .. 4 ul + 3 (8 bytes), 128-byte rows, 128B swizzle
        const float v[4] = {dg.x * gs, dg.y * gs, dg.z * gs, dg.w * gs};
        __half h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          h[q] = __float2half_rn(v[q]);
          l[q] = __float2half_rn((v[q] - __half2float(h[q])) * F16_LO_SCALE);
        }
        if (dg_hi != nullptr && n_ok[i] && u_ok) {
          // the same fp16 hi/lo pair is the split of dgates for the weight-gradient GEMMs that follow: no split pass
          const size_t o = ((size_t)ng[i] * T + t) * H4 + 4 * ug;
          *reinterpret_cast<uint2*>(dg_hi + o) = *reinterpret_cast<const uint2*>(h);
          *reinterpret_cast<uint2*>(dg_lo + o) = *reinterpret_cast<const uint2*>(l);
        }
        const int n = 4 * nq + i;
        const uint32_t off = (uint32_t)(ul >> 4) * 4096u + (uint32_t)n * 128u +
                             (uint32_t)((((ul & 15) >> 1) ^ (n & 7)) << 4) + (uint32_t)(ul & 1) * 8u;
        *reinterpret_cast<uint2*>(sB + off) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(sB + 8192 + off) = *reinterpret_cast<const uint2*>(l);
      }
      if (s == 0) break;   // the recurrent gradient of the first processed step goes to hid_init (host side)
      fetch(s - 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(q_smem(&bready)) : "memory");
      // ---- 2./3. my k row of the partial product -> the owner's inbox ----
      q_mbar_wait(&accbar, (uint32_t)(k & 1));
      ++k;
      q_fence_after();
      if (mbk < MB) {
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mbk * BCOLS);
        if (owner < CS) {
          float* box = inbox + (((size_t)par * CS + rank) * QU + lane) * QN;      // same offset in every CTA
          const uint32_t remote = q_mapa(q_smem(box), (uint32_t)owner);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float a[QC], c0[QC], c1[QC], c2[QC], c3[QC], c4[QC];
            q_tmem_ld16(taddr + BACC * 32 + hf * QC, c0);
            q_tmem_ld16(taddr + (BACC + 1) * 32 + hf * QC, c1);
            q_tmem_ld16(taddr + hf * QC, a);
            q_tmem_ld16(taddr + 32 + hf * QC, c2);
            q_tmem_ld16(taddr + 64 + hf * QC, c3);
            q_tmem_ld16(taddr + 96 + hf * QC, c4);
            q_tmem_wait();
#pragma unroll
            for (int j = 0; j < QC; ++j)
              a[j] = (fmaf(c0[j] + c1[j], F16_LO_INV, a[j]) + c2[j] + (c3[j] + c4[j])) * ms1 * ms2;
#pragma unroll
            for (int j = 0; j < QC; j += 4)
              asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote + (uint32_t)(hf * QC + j) * 4u),
                           "f"(a[j]), "f"(a[j + 1]), "f"(a[j + 2]), "f"(a[j + 3])
                           : "memory");
          }
        }
      }
      q_fence_before();
      // ---- 4. the inboxes of this step are complete ----
      q_cluster_sync();
      // ---- sum the CS partial blocks for my (unit, 4 utterances) ----
      {
        float4 acc = make_float4(dh_pass[0], dh_pass[1], dh_pass[2], dh_pass[3]);
        for (int src = 0; src < CS; ++src) {
          const float4 v = *reinterpret_cast<const float4*>(inbox + (((size_t)par * CS + src) * QU + ul) * QN + 4 * nq);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        dh_next[0] = acc.x; dh_next[1] = acc.y; dh_next[2] = acc.z; dh_next[3] = acc.w;
      }
      par ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n_ok[i] && u_ok) {
        dc_fin[(size_t)ng[i] * H + ug] = dc_next[i];
        dh_fin[(size_t)ng[i] * H + ug] = dh_pass[i];
      }
    if (has_peep && u_ok) {
      atomicAdd(dpeep + ug, pci);
      atomicAdd(dpeep + H + ug, pcf);
      atomicAdd(dpeep + 2 * H + ug, pco);
    }
    if (db != nullptr && u_ok) {
      atomicAdd(db + 4 * ug, dbs.x);
      atomicAdd(db + 4 * ug + 1, dbs.y);
      atomicAdd(db + 4 * ug + 2, dbs.z);
      atomicAdd(db + 4 * ug + 3, dbs.w);
    }
  }
  q_fence_before();
  q_cluster_sync();
  if (warp == 8) {
    q_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)QTMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*QEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static QEncodeFn q_get_encode() {
  static QEncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<QEncodeFn>(ptr);
  }
  return fn;
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

int ipavsr_lstm_fwd_f16_supported(int N, int T, int H, int ldw) {
  (void)N;
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("IPAVSR_LSTM_TC");
    off = (e && e[0] == '0') ? 1 : 0;
  }
  return (!off && H >= 8 && H <= 256 && T >= 1 && T <= 64 && ldw % 8 == 0 && ldw >= 4 * H) ? 1 : 0;
}

int ipavsr_lstm_fwd_f16(const float* xw, const uint16_t* whid_hi, const uint16_t* whid_lo, const int32_t* whid_exp,
                        int ldw, const float* peep, const float* cell_init, const float* hid_init, const uint8_t* mask,
                        float* out, float* gates, float* cell, float* hprev, int N, int T, int H, int ldh, int backwards,
                        void* stream) {
  IPAVSR_CHECK_ARG(xw && whid_hi && whid_lo && whid_exp && cell_init && hid_init && mask && out, "null pointer");
  IPAVSR_CHECK_ARG(N >= 0 && T >= 1 && H >= 1 && ldh >= H, "bad sizes");
  if (!ipavsr_lstm_fwd_f16_supported(N, T, H, ldw) ||
      ((reinterpret_cast<uintptr_t>(whid_hi) | reinterpret_cast<uintptr_t>(whid_lo)) & 15) != 0) {
    set_error("ipavsr_lstm_fwd_f16: needs 8 <= H <= 256, T <= 64, 16-byte aligned W_hid halves with ldw %% 8 == 0");
    return IPAVSR_ERR_UNSUPPORTED;
  }
  if (N == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  QEncodeFn enc = q_get_encode();
  if (enc == nullptr) {
    set_error("ipavsr_lstm_fwd_f16: cuTensorMapEncodeTiled is not available from the driver");
    return IPAVSR_ERR_CUDA;
  }
  // W_hid split: fp16 [H rows (k)][4H gate columns], row stride ldw halves; box {64 columns, 64 k-rows}, 128B swizzle
  CUtensorMap maps[2];
  const uint16_t* base[2] = {whid_hi, whid_lo};
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)(4 * H), (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)ldw * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(base[i]), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("ipavsr_lstm_fwd_f16: cuTensorMapEncodeTiled failed with %d", (int)r);
      return IPAVSR_ERR_CUDA;
    }
  }
  const int cs = (H + QU - 1) / QU;
  const int kb = (cs * QU + 63) / 64;
  const size_t smem = (size_t)2 * kb * 16384 + (size_t)4 * cs * 2048 + 2 * 4096 + 1024;
  const int tiles = (N + QN - 1) / QN;
  IPAVSR_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * cs);
  cfg.blockDim = dim3(QTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  IPAVSR_CUDA(cudaLaunchKernelEx(&cfg, lstm_fwd_tc_kernel, maps[0], maps[1], xw, whid_exp, peep, cell_init, hid_init, mask,
                                 out, gates, cell, hprev, N, T, H, ldh, backwards, g_lstm_dbg));
  count_launch();
  return IPAVSR_OK;
}

}  // extern "C"

extern "C" {

int ipavsr_lstm_bwd_f16_supported(int N, int T, int H, int ldw, float clip) {
  return (ipavsr_lstm_fwd_f16_supported(N, T, H, ldw) && clip > 0.f && clip < 16384.f) ? 1 : 0;
}

int ipavsr_lstm_bwd_f16(const float* dout, const float* w_hid, const uint16_t* whid_hi, const uint16_t* whid_lo,
                        const int32_t* whid_exp, int ldw, const float* peep, const float* cell_init, const uint8_t* mask,
                        const float* gates, const float* cell, float* dgates, float* dpeep, float* dcell_init,
                        float* dhid_init, int N, int T, int H, int ldh, int backwards, float clip, int accumulate,
                        float* db, uint16_t* dg_hi, uint16_t* dg_lo, int32_t* dg_exp, void* workspace,
                        uint64_t workspace_bytes, void* stream) {
  IPAVSR_CHECK_ARG((dg_hi == nullptr) == (dg_lo == nullptr) && (dg_hi == nullptr) == (dg_exp == nullptr),
                   "dg_hi, dg_lo and dg_exp go together");
  IPAVSR_CHECK_ARG(dout && w_hid && whid_hi && whid_lo && whid_exp && cell_init && mask && gates && cell && dgates &&
                       dcell_init && dhid_init,
                   "null pointer");
  IPAVSR_CHECK_ARG((peep == nullptr) == (dpeep == nullptr), "peep and dpeep go together");
  IPAVSR_CHECK_ARG(N >= 0 && T >= 1 && H >= 1 && ldh >= H, "bad sizes");
  IPAVSR_CHECK_ARG(workspace && workspace_bytes >= (uint64_t)2 * N * H * sizeof(float), "workspace too small (2*N*H floats)");
  if (!ipavsr_lstm_bwd_f16_supported(N, T, H, ldw, clip) ||
      ((reinterpret_cast<uintptr_t>(whid_hi) | reinterpret_cast<uintptr_t>(whid_lo)) & 15) != 0) {
    set_error("ipavsr_lstm_bwd_f16: needs 8 <= H <= 256, T <= 64, clip > 0, 16-byte aligned W_hid halves, ldw %% 8 == 0");
    return IPAVSR_ERR_UNSUPPORTED;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) {
    IPAVSR_CUDA(cudaMemsetAsync(dcell_init, 0, sizeof(float) * H, st));
    IPAVSR_CUDA(cudaMemsetAsync(dhid_init, 0, sizeof(float) * H, st));
    if (dpeep) IPAVSR_CUDA(cudaMemsetAsync(dpeep, 0, sizeof(float) * 3 * H, st));
    if (db) IPAVSR_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 4 * H, st));
  }
  if (N == 0) return IPAVSR_OK;
  QEncodeFn enc = q_get_encode();
  if (enc == nullptr) {
    set_error("ipavsr_lstm_bwd_f16: cuTensorMapEncodeTiled is not available from the driver");
    return IPAVSR_ERR_CUDA;
  }
  // W_hid split, K-major A operand: box {64 gate columns (128 bytes), 128 k-rows}
  CUtensorMap maps[2];
  const uint16_t* base[2] = {whid_hi, whid_lo};
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)(4 * H), (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)ldw * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(base[i]), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("ipavsr_lstm_bwd_f16: cuTensorMapEncodeTiled failed with %d", (int)r);
      return IPAVSR_ERR_CUDA;
    }
  }
  float* dc_fin = reinterpret_cast<float*>(workspace);
  float* dh_fin = dc_fin + (size_t)N * H;
  const int cs = (H + QU - 1) / QU;
  const int mb = (cs * QU + 127) / 128;
  const size_t smem = (size_t)2 * mb * 32768 + 16384 + (size_t)2 * cs * QU * QN * sizeof(float) + 1024;
  const int tiles = (N + QN - 1) / QN;
  // |dg| <= clip: static scale 2^eg with clip * 2^eg < 2^15
  const int eg = 14 - (ilogbf(clip) + 1);
  IPAVSR_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * cs);
  cfg.blockDim = dim3(QTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  IPAVSR_CUDA(cudaLaunchKernelEx(&cfg, lstm_bwd_tc_kernel, maps[0], maps[1], dout, whid_exp, peep, cell_init, mask, gates,
                                 cell, dgates, dpeep, dc_fin, dh_fin, N, T, H, ldh, backwards, clip, eg, db, dg_hi, dg_lo,
                                 dg_exp));
  count_launch();
  // d(hid_init) = sum_n [ dg_first W_hid^T + dh_pass ],  d(cell_init) = sum_n dc_prev at the first processed step.
  // Only the sum over the utterances is wanted, so the product is taken AFTER the sum (it is linear): one column sum of
  // the first processed frame's gate gradients (rows n, stride T*4H) and a 1 x 4H by 4H x H vector-matrix product, in
  // place of the N x H x 4H GEMM this used to run on the critical path of every backward step (95 us at H = 250: its
  // output pitch H = 250 took the GEMM's scalar epilogue).
  const int t_first = backwards ? T - 1 : 0;
  int rc;
  const size_t off_first = (size_t)t_first * 4 * H;
  float* dg_sum = dh_fin + (size_t)N * H;          // 4H floats behind dc_fin / dh_fin (the workspace holds 8 N H floats)
  rc = ipavsr_colsum(dgates + off_first, T * 4 * H, dg_sum, N, 4 * H, 0, stream);
  if (rc) return rc;
  rc = ipavsr_colsum(dh_fin, H, dhid_init, N, H, 1, stream);
  if (rc) return rc;
  rc = gemv_rows(w_hid, 4 * H, dg_sum, dhid_init, H, 4 * H, 1, st);
  if (rc) return rc;
  return ipavsr_colsum(dc_fin, H, dcell_init, N, H, 1, stream);
}

}  // extern "C"

extern "C" int ipavsr_debug_lstm_timestamps(unsigned long long* buf) {
  ipavsr::g_lstm_dbg = buf;
  return IPAVSR_OK;
}
