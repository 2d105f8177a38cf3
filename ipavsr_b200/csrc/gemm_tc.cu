// Tensor-core GEMM for sm_100a: TMA-fed tcgen05.mma (kind::tf32) with the accumulator in TMEM and a fused
// bias + nonlinearity epilogue.  Modes: IPAVSR_GEMM_TF32 (one MMA per product) and IPAVSR_GEMM_TF32X3 — the fp32
// parity mode: every operand is split a = hi + lo with hi, lo exactly representable in tf32 and the product is
// accumulated as lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator (the dropped lo*lo term is ~2^-22 relative).
//
// Replaces the cuBLAS/BLAS sgemm behind Theano's T.dot in the DBNF encoder (reference
// modelzoo/pretrained_encoder.py:4-9), the hoisted LSTM input projections and all their dgrad/wgrad products.
//
// Structure (one CTA per 128 x BN output tile, optional split-K over gridDim.z):
//   warp 0  TMA producer   cp.async.bulk.tensor.2d global -> 128B-swizzled shared tiles, mbarrier complete_tx
//   warp 1  MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8),
//                          tcgen05.commit releases the shared stage / signals the epilogue
//   warp 2  TMEM allocator tcgen05.alloc / dealloc of BN columns
//   warps 4-7 epilogue     tcgen05.ld 32x32b (one TMEM lane = one output row per thread), bias + activation,
//                          vectorised global stores (or float atomics for split-K)
// IPAVSR_GEMM_F16X3 is the same three-product scheme on 16-bit operands (kind::f16, twice the MMA rate and half the
// operand bytes of tf32): x * 2^e = hi + lo * 2^-11 with hi, lo in fp16 and e a per-tensor power-of-two scale chosen
// from the tensor's amax (f16split.cu); hi*hi goes to the main accumulators, lo*hi + hi*lo to the cross accumulator,
// and the epilogue forms (main + cross * 2^-11) * 2^-(eA+eB).
//
// Operand layouts: a K-major operand (reduction dim contiguous) is one TMA box of [rows x 32 floats]; an MN-major
// operand (stored transposed: A as [K,M] for wgrad, B as [K,N] for the forward x*W) is BLOCK/32 boxes of
// [32 k-rows x 32 floats] (TMA swizzle 128B_ATOM_32B) and is consumed through MN-major UMMA descriptors
// (SWIZZLE_128B_BASE32B) — no transposed copies are made.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ipavsr {

// CG = 2: a CTA pair (cluster (2,1,1), same TPC) computes a 256 x BN tile with one cta_group::2 MMA per k-step: each CTA
// keeps its own 128 rows of A and HALF of the B tile in shared memory (the tensor core of each SM reads both halves),
// which cuts the L2->SM operand traffic per flop by a third and the shared-memory reads per MMA by a third.
// (Measured and removed in round 2: two CTAs of 128-wide pair tiles per SM, each with half of the shared memory and of the
//  TMEM columns, so that one tile's epilogue runs under the other's mainloop — no gain, 0.176 vs 0.175 ms for the fc1
//  forward: the 128-wide tile's mainloop runs at 78 % of the 256-wide tile's rate.  The persistent kernel of gemm_f16p.cu
//  overlaps the epilogue instead.)
template <int BN, bool A_MN, bool B_MN, int NPROD, bool F16, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo, TcParams p) {
  constexpr int BKE = F16 ? 2 * TC_BK : TC_BK;        // elements per stage along K (128 bytes either way)
  constexpr int MNBOX = F16 ? 64 : 32;                // MN-major TMA box: 128 bytes of M/N x BKE k-rows
  constexpr int MNBOX_BYTES = BKE * 128;
  constexpr int A_BYTES = TC_BM * TC_BK * 4;          // 16 KB
  constexpr int BNL = BN / CG;                        // B-tile columns this CTA loads
  constexpr int B_BYTES = BNL * TC_BK * 4;
  static_assert(BNL % MNBOX == 0, "the per-CTA share of the B tile must be whole TMA boxes");
  constexpr int NOPER = (NPROD == 3) ? 2 : 1;         // hi (+ lo) copies of each operand
  constexpr int STAGE_BYTES = NOPER * (A_BYTES + B_BYTES);
  constexpr int SMEM_BUDGET = 200 * 1024;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES < 8 ? SMEM_BUDGET / STAGE_BYTES : 8;
  static_assert(STAGES >= 2, "need at least a double buffer");
  static_assert(!F16 || NPROD == 3, "the fp16 path is the three-product mode");
  // TMEM accumulators.  The tensor core adds into its fp32 accumulator with truncation, so the error grows with the
  // number of accumulations into one accumulator.  In the 3xTF32 mode the tiny cross terms (lo*hi + hi*lo) get their
  // own accumulator and the hi*hi terms are spread round-robin (by k-block) over NMAIN accumulators; the epilogue sums
  // them in registers with round-to-nearest.
  constexpr int TMEM_COLS = (NPROD == 3) ? 512 : BN;
  constexpr int NMAIN = (NPROD == 3) ? (TMEM_COLS / BN - 1) : 1;
  static_assert(NMAIN >= 1, "main + cross accumulators must fit the TMEM share of this CTA");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], accum_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[256];        // bias of this tile's columns (zero where absent)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* dbg = nullptr;
  if (p.dbg != nullptr)
    dbg = p.dbg + 8ull * (blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z));
  if (dbg != nullptr && threadIdx.x == 0) dbg[0] = gtimer();       // CTA start
  // pair variant: grid.x = 2 * column tiles (the two CTAs of a cluster (2,1,1) are x = 2j, 2j+1: the two row halves of
  // column tile j), grid.y = row-tile pairs.  Column tiles are the fastest-varying index in both variants, so the CTAs
  // that share an A row panel run at the same time and the panel is read from DRAM once (ncu: 751 MB -> DRAM reads
  // for the 98 MB A operand of the fc1 product when row tiles varied fastest).
  const int m0 = (CG == 2 ? (blockIdx.y * 2 + (blockIdx.x & 1)) : blockIdx.y) * TC_BM;
  const int n0 = (CG == 2 ? (blockIdx.x >> 1) : blockIdx.x) * BN;
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;     // 0 = leader of the pair
  const int nl0 = n0 + (int)crank * BNL;                         // first B column this CTA loads
  const int num_kb_total = (p.K + BKE - 1) / BKE;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
  const int num_kb = max(kb_end - kb_begin, 0);

  if (threadIdx.x < BN) {
    const int c = (CG == 2 ? (blockIdx.x >> 1) : blockIdx.x) * BN + threadIdx.x;
    s_bias[threadIdx.x] = (p.bias != nullptr && c < p.N) ? __ldg(p.bias + c) : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();     // the peer's barriers must be initialised before anything can arrive on them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (dbg != nullptr && threadIdx.x == 0) dbg[1] = gtimer();       // setup done (barriers, TMEM, cluster sync)

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sA = smem + (size_t)stage * STAGE_BYTES;
      uint8_t* sB = sA + NOPER * A_BYTES;
      const int k0 = (kb_begin + kb) * BKE;
      if (CG == 2) {
        // both CTAs' loads complete on the leader's barrier, which therefore expects the bytes of both
        if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
        const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
#pragma unroll
        for (int o = 0; o < NOPER; ++o) {
          const CUtensorMap* ma = o == 0 ? &mapA : &mapAlo;
          const CUtensorMap* mb = o == 0 ? &mapB : &mapBlo;
          if (!A_MN) {
            tma_load_2d_cg2(sA + o * A_BYTES, ma, fb, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / MNBOX; ++j)
              tma_load_2d_cg2(sA + o * A_BYTES + j * MNBOX_BYTES, ma, fb, m0 + MNBOX * j, k0);
          }
          if (!B_MN) {
            tma_load_2d_cg2(sB + o * B_BYTES, mb, fb, k0, nl0);
          } else {
#pragma unroll
            for (int j = 0; j < BNL / MNBOX; ++j)
              tma_load_2d_cg2(sB + o * B_BYTES + j * MNBOX_BYTES, mb, fb, nl0 + MNBOX * j, k0);
          }
        }
      } else {
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
        for (int o = 0; o < NOPER; ++o) {
          const CUtensorMap* ma = o == 0 ? &mapA : &mapAlo;
          const CUtensorMap* mb = o == 0 ? &mapB : &mapBlo;
          if (!A_MN) {
            tma_load_2d(sA + o * A_BYTES, ma, &full_bar[stage], k0, m0);                 // box {32 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < TC_BM / MNBOX; ++j)                                      // boxes {128 B of m, BKE k-rows}
              tma_load_2d(sA + o * A_BYTES + j * MNBOX_BYTES, ma, &full_bar[stage], m0 + MNBOX * j, k0);
          }
          if (!B_MN) {
            tma_load_2d(sB + o * B_BYTES, mb, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / MNBOX; ++j)
              tma_load_2d(sB + o * B_BYTES + j * MNBOX_BYTES, mb, &full_bar[stage], n0 + MNBOX * j, k0);
          }
        }
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && crank == 0) {
    // ===================== MMA issuer (leader CTA only in the pair variant) =====================
    constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN, F16, TC_BM * CG);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (dbg != nullptr && kb == 0) dbg[2] = gtimer();            // first stage landed
      const uint32_t sA = smem_u32(smem + (size_t)stage * STAGE_BYTES);
      const uint32_t sB = sA + NOPER * A_BYTES;
#pragma unroll
      for (int k = 0; k < TC_BK / 8; ++k) {
        // K-major (SWIZZLE_128B): advance 32 bytes inside the swizzled 128-byte row; SBO = 1024 (8 rows).
        // MN-major (SWIZZLE_128B_BASE32B, atoms of 4 k-rows x 128 bytes): advance 8 k-rows (1024 bytes);
        // SBO = 512 (next 4-row atom along K), LBO = 4096 (next 32-wide chunk along M/N = next TMA box).
        // fp16 MN-major is the plain SWIZZLE_128B canonical layout: atoms of 8 k-rows x 128 bytes (64 halves of M/N),
        // one MMA (K = 16) spans two atoms: advance 2048 bytes, SBO = 1024 (next atom along K), LBO = next TMA box.
        constexpr uint32_t MN_STEP = F16 ? 2048 : 1024, MN_SBO = F16 ? 1024 : 512, MN_LT = F16 ? 2 : 1;
        const uint32_t a_off = A_MN ? k * MN_STEP : k * 32;
        const uint32_t b_off = B_MN ? k * MN_STEP : k * 32;
        const uint32_t a_lbo = A_MN ? MNBOX_BYTES : 16, b_lbo = B_MN ? MNBOX_BYTES : 16;
        const uint32_t a_sbo = A_MN ? MN_SBO : 1024, b_sbo = B_MN ? MN_SBO : 1024;
        const uint32_t a_lt = A_MN ? MN_LT : 2, b_lt = B_MN ? MN_LT : 2;
        const uint64_t a_hi = make_smem_desc(sA + a_off, a_lbo, a_sbo, a_lt);
        const uint64_t b_hi = make_smem_desc(sB + b_off, b_lbo, b_sbo, b_lt);
        const uint32_t first = (kb | k) == 0 ? 0u : 1u;
        if (NPROD == 3) {
          const uint64_t a_lo = make_smem_desc(sA + A_BYTES + a_off, a_lbo, a_sbo, a_lt);
          const uint64_t b_lo = make_smem_desc(sB + B_BYTES + b_off, b_lbo, b_sbo, b_lt);
          const uint32_t t_cross = tmem_base + (uint32_t)(NMAIN * BN);
          const uint32_t t_main = tmem_base + (uint32_t)((kb % NMAIN) * BN);
          const uint32_t main_acc = (kb < NMAIN && k == 0) ? 0u : 1u;
          if (F16 && p.oneacc) {
            if (CG == 2) {
              tcgen05_mma_cg2<F16>(t_main, a_lo, b_hi, idesc, main_acc);
              tcgen05_mma_cg2<F16>(t_main, a_hi, b_lo, idesc, 1u);
              tcgen05_mma_cg2<F16>(t_main, a_hi, b_hi, idesc, 1u);
            } else {
              tcgen05_mma_f16(t_main, a_lo, b_hi, idesc, main_acc);
              tcgen05_mma_f16(t_main, a_hi, b_lo, idesc, 1u);
              tcgen05_mma_f16(t_main, a_hi, b_hi, idesc, 1u);
            }
          } else if (CG == 2) {
            tcgen05_mma_cg2<F16>(t_cross, a_lo, b_hi, idesc, first);
            tcgen05_mma_cg2<F16>(t_cross, a_hi, b_lo, idesc, 1u);
            tcgen05_mma_cg2<F16>(t_main, a_hi, b_hi, idesc, main_acc);
          } else if (F16) {
            tcgen05_mma_f16(t_cross, a_lo, b_hi, idesc, first);
            tcgen05_mma_f16(t_cross, a_hi, b_lo, idesc, 1u);
            tcgen05_mma_f16(t_main, a_hi, b_hi, idesc, main_acc);
          } else {
            tcgen05_mma_tf32(t_cross, a_lo, b_hi, idesc, first);
            tcgen05_mma_tf32(t_cross, a_hi, b_lo, idesc, 1u);
            tcgen05_mma_tf32(t_main, a_hi, b_hi, idesc, main_acc);
          }
        } else {
          if (CG == 2) tcgen05_mma_cg2<false>(tmem_base, a_hi, b_hi, idesc, first);
          else tcgen05_mma_tf32(tmem_base, a_hi, b_hi, idesc, first);
        }
      }
      // frees this shared stage (in both CTAs of a pair) once the MMAs above have read it
      if (CG == 2) tcgen05_commit_cg2(&empty_bar[stage]);
      else tcgen05_commit(&empty_bar[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // accumulator complete
    if (CG == 2) tcgen05_commit_cg2(&accum_bar);
    else tcgen05_commit(&accum_bar);
    if (dbg != nullptr) dbg[3] = gtimer();                         // last MMA issued
  }
  __syncwarp();
  {
    // ===================== epilogue: all 8 warps =====================
    // Warp w reads TMEM lanes 32*(w%4).. (one output row per thread); warps 4-7 take the even 32-column chunks and
    // warps 0-3 (whose producer / MMA-issuer lanes have finished by now) the odd ones, so every SM sub-partition has
    // two warps to hide the epilogue's latencies behind each other.
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int row = m0 + q * 32 + lane;
    if (num_kb > 0) {
      mbar_wait(&accum_bar, 0);
      tcgen05_fence_after();
    }
    if (dbg != nullptr && threadIdx.x == 128) dbg[4] = gtimer();   // accumulator complete: epilogue starts
    const bool split = p.splits > 1;
    const bool vec = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    // f16 mode: result = (main + cross * 2^-11) * 2^-(eA+eB); one exact factor when it is a normal float, else two
    float ms1 = 1.f, ms2 = 1.f;
    if (F16) {
      const int e = -(__ldg(p.expA) + __ldg(p.expB));
      const int e1 = (e > 126 || e < -115) ? e / 2 : e, e2 = e - e1;
      ms1 = __int_as_float((127 + e1) << 23);
      ms2 = __int_as_float((127 + e2) << 23);
    }
    const float cs1 = F16 ? ms1 * (F16_LO_INV) : 1.f;
    float tile_max = 0.f;
    // (a k-split tile takes the fast path too: its partial sums go out as float4 atomics from the staging tile, 4 rows x
    //  128 contiguous bytes per instruction instead of 32 rows x 16 bytes)
    const bool fast_ok = vec && num_kb > 0;                // warp-uniform
#pragma unroll 1
    for (int c0 = (warp >= 4 ? 0 : 32); c0 < BN; c0 += 64) {
      if (n0 + c0 >= p.N) break;                   // warp-uniform
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if (fast_ok && n0 + c0 + 32 <= p.N) {
        // ---------------- fast path: a full 32-column chunk, vector stores, compile-time activation ----------------
        float v[32];
        if (NPROD == 3) {
          float u[32];
          if (F16 && p.oneacc) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          } else {
            tmem_ld32_issue(lane_base + (uint32_t)(NMAIN * BN), v);         // cross terms
          }
          tmem_ld32_issue(lane_base, u);                                    // first main accumulator
          tmem_ld_wait();
          if (F16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], cs1, u[j] * ms1);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
          const int used = num_kb < NMAIN ? num_kb : NMAIN;
#pragma unroll 1
          for (int a = 1; a < used; ++a) {
            tmem_ld32_issue(lane_base + (uint32_t)(a * BN), u);
            tmem_ld_wait();
            if (F16) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaf(u[j], ms1, v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
          }
          if (F16 && ms2 != 1.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= ms2;
          }
        } else {
          tmem_ld32_issue(lane_base, v);
          tmem_ld_wait();
        }
        {
          // Coalesced store of the 32 x 32 chunk through this warp's staging tile in the (now idle) pipeline memory:
          // thread = row writes its 8 float4 at XOR-swizzled positions (conflict-free), then every store instruction
          // covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes.
          float* stg = reinterpret_cast<float*>(smem) + warp * 1024;
          if (row < p.M) {
            float* cpo = p.C + (size_t)row * p.ldc + n0 + c0;
            if (p.accumulate && !split) {
#pragma unroll
              for (int j4 = 0; j4 < 32; j4 += 4) {
                const float4 old = *reinterpret_cast<const float4*>(cpo + j4);
                v[j4] += old.x; v[j4 + 1] += old.y; v[j4 + 2] += old.z; v[j4 + 3] += old.w;
              }
            }
            if (p.bias != nullptr && (!split || blockIdx.z == 0)) {
#pragma unroll
              for (int j4 = 0; j4 < 32; j4 += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + j4]);
                v[j4] += b4.x; v[j4 + 1] += b4.y; v[j4 + 2] += b4.z; v[j4 + 3] += b4.w;
              }
            }
            switch (p.act) {
              case IPAVSR_ACT_LINEAR: break;
              case IPAVSR_ACT_SIGMOID:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = sigmoid_fast(v[j]);
                break;
              case IPAVSR_ACT_RECTIFY:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                break;
              default:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], p.act);
                break;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          {
            const int rbase = m0 + q * 32, jj = lane & 7;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const int rr = 4 * qq + (lane >> 3);
              const float4 o = *reinterpret_cast<const float4*>(stg + rr * 32 + ((jj ^ (rr & 7)) << 2));
              if (rbase + rr < p.M) {
                float4* dst = reinterpret_cast<float4*>(p.C + (size_t)(rbase + rr) * p.ldc + n0 + c0 + 4 * jj);
                if (split) atomicAdd(dst, o);
                else *dst = o;
              }
            }
            if (p.C16hi != nullptr) {
              // fp16 hi/lo split of the chunk under the static scale, from the same staging tile: 8 lanes cover 64
              // contiguous bytes of a row in each array
              const float sc = __int_as_float((127 + p.c16_exp) << 23);
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                const int rr = 4 * qq + (lane >> 3);
                const float4 o = *reinterpret_cast<const float4*>(stg + rr * 32 + ((jj ^ (rr & 7)) << 2));
                if (rbase + rr < p.M) {
                  const float ov[4] = {o.x * sc, o.y * sc, o.z * sc, o.w * sc};
                  __half h[4], l[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    h[j] = __float2half_rn(ov[j]);
                    l[j] = __float2half_rn((ov[j] - __half2float(h[j])) * F16_LO_SCALE);
                  }
                  const size_t off = (size_t)(rbase + rr) * p.ldc + n0 + c0 + 4 * jj;
                  *reinterpret_cast<uint2*>(p.C16hi + off) = *reinterpret_cast<const uint2*>(h);
                  *reinterpret_cast<uint2*>(p.C16lo + off) = *reinterpret_cast<const uint2*>(l);
                }
              }
            }
          }
          __syncwarp();
        }
        if (row < p.M) {
          if (p.amax != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tile_max = fmaxf(tile_max, fabsf(v[j]));
          }
          if (p.Chi != nullptr) {
            const size_t off = (size_t)row * p.ldc + n0 + c0;
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              float4 h, l;
              tf32_hi_lo(v[j4], h.x, l.x); tf32_hi_lo(v[j4 + 1], h.y, l.y);
              tf32_hi_lo(v[j4 + 2], h.z, l.z); tf32_hi_lo(v[j4 + 3], h.w, l.w);
              *reinterpret_cast<float4*>(p.Chi + off + j4) = h;
              *reinterpret_cast<float4*>(p.Clo + off + j4) = l;
            }
          }
        }
        continue;
      }
      // ---------------- general path: ragged edges, split-K atomics, unaligned C ----------------
      float v[32];
      if (num_kb > 0) {
        if (NPROD == 3) {
          if (F16 && p.oneacc) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          } else {
            tmem_ld32(lane_base + (uint32_t)(NMAIN * BN), v);               // cross terms first (smallest)
          }
          if (F16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= (F16_LO_INV);          // lo carries a 2^11 scale
          }
          const int used = num_kb < NMAIN ? num_kb : NMAIN;
#pragma unroll 1
          for (int a = 0; a < used; ++a) {
            float u[32];
            tmem_ld32(lane_base + (uint32_t)(a * BN), u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += u[j];
          }
          if (F16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] * ms1 * ms2;
          }
        } else {
          tmem_ld32(lane_base, v);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (row < p.M) {
        float* cp = p.C + (size_t)row * p.ldc + n0 + c0;
        if (split) {
          if (vec && n0 + c0 + 32 <= p.N) {
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              float4 add = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
              if (p.bias != nullptr && blockIdx.z == 0) {
                const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + j4]);
                add.x += b4.x; add.y += b4.y; add.z += b4.z; add.w += b4.w;
              }
              atomicAdd(reinterpret_cast<float4*>(cp + j4), add);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < p.N) {
                float add = v[j];
                if (p.bias != nullptr && blockIdx.z == 0) add += s_bias[c0 + j];
                atomicAdd(cp + j, add);
              }
          }
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const int col = n0 + c0 + j4;
            if (col >= p.N) break;
            if (vec && col + 3 < p.N) {
              float4 o = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
              if (p.accumulate) {
                float4 old = *reinterpret_cast<const float4*>(cp + j4);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              if (p.bias != nullptr) {
                o.x += s_bias[c0 + j4]; o.y += s_bias[c0 + j4 + 1];
                o.z += s_bias[c0 + j4 + 2]; o.w += s_bias[c0 + j4 + 3];
              }
              o.x = act_epi(o.x, p.act); o.y = act_epi(o.y, p.act); o.z = act_epi(o.z, p.act); o.w = act_epi(o.w, p.act);
              *reinterpret_cast<float4*>(cp + j4) = o;
              tile_max = fmaxf(tile_max, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
              if (p.Chi != nullptr) {
                float4 h, l;
                tf32_hi_lo(o.x, h.x, l.x); tf32_hi_lo(o.y, h.y, l.y); tf32_hi_lo(o.z, h.z, l.z); tf32_hi_lo(o.w, h.w, l.w);
                const size_t off = (size_t)row * p.ldc + col;
                *reinterpret_cast<float4*>(p.Chi + off) = h;
                *reinterpret_cast<float4*>(p.Clo + off) = l;
              }
              if (p.C16hi != nullptr) {
                const float sc = __int_as_float((127 + p.c16_exp) << 23);
                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float xs = ov[j] * sc;
                  const __half h = __float2half_rn(xs);
                  reinterpret_cast<__half*>(p.C16hi)[(size_t)row * p.ldc + col + j] = h;
                  reinterpret_cast<__half*>(p.C16lo)[(size_t)row * p.ldc + col + j] =
                      __float2half_rn((xs - __half2float(h)) * F16_LO_SCALE);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col + j < p.N) {
                  float o = v[j4 + j];
                  if (p.accumulate) o += cp[j4 + j];
                  if (p.bias != nullptr) o += s_bias[c0 + j4 + j];
                  o = act_epi(o, p.act);
                  cp[j4 + j] = o;
                  tile_max = fmaxf(tile_max, fabsf(o));
                  if (p.Chi != nullptr) {
                    float h, l;
                    tf32_hi_lo(o, h, l);
                    p.Chi[(size_t)row * p.ldc + col + j] = h;
                    p.Clo[(size_t)row * p.ldc + col + j] = l;
                  }
                  if (p.C16hi != nullptr) {
                    const float xs = o * __int_as_float((127 + p.c16_exp) << 23);
                    const __half h = __float2half_rn(xs);
                    reinterpret_cast<__half*>(p.C16hi)[(size_t)row * p.ldc + col + j] = h;
                    reinterpret_cast<__half*>(p.C16lo)[(size_t)row * p.ldc + col + j] =
                        __float2half_rn((xs - __half2float(h)) * F16_LO_SCALE);
                  }
                }
            }
          }
        }
      }
    }
    if (p.amax != nullptr && !split) {
      tile_max = warp_max(tile_max);
      if (lane == 0 && tile_max > 0.f) atomicMax(reinterpret_cast<unsigned int*>(p.amax), __float_as_uint(tile_max));
    }
    tcgen05_fence_before();
    if (dbg != nullptr && threadIdx.x == 128) dbg[5] = gtimer();   // epilogue done (even chunks)
    if (dbg != nullptr && threadIdx.x == 0) dbg[6] = gtimer();     // epilogue done (odd chunks)
  }
  if (CG == 2) cluster_sync_all();     // the leader's MMAs read the peer's shared memory and write its TMEM
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
  }
}

// hi = rna_tf32(x), lo = rna_tf32(x - hi): both exactly representable in tf32
__global__ void tf32_split_rna_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                      size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    float in[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tf32_hi_lo(in[j], h[j], l[j]);
    reinterpret_cast<float4*>(hi)[i] = make_float4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<float4*>(lo)[i] = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D row-major tensor [outer][inner] (fp32, or fp16 when f16) with row stride ld elements; box {box_inner, box_outer};
// 128B swizzle (the 32B-atom variant for MN-major 32-bit operands)
int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
             uint32_t box_outer, bool mn_major, bool f16 = false, bool sw64 = false) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error("gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    return IPAVSR_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw64 ? CU_TENSOR_MAP_SWIZZLE_64B
                        : ((mn_major && !f16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc: cuTensorMapEncodeTiled failed with %d (inner %llu outer %llu ld %llu)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
    return IPAVSR_ERR_CUDA;
  }
  return IPAVSR_OK;
}

unsigned long long* g_gemm_dbg = nullptr;      // set by ipavsr_debug_gemm_timestamps (profiling aid)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool gemm_tc_supported(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       const float* C, int ldc) {
  (void)transA; (void)transB; (void)C; (void)ldc;
  if (!aligned16(A) || !aligned16(B) || lda % 4 != 0 || ldb % 4 != 0) return false;
  if (K < 16 || M < 1 || N < 8) return false;
  // tiny products are launch-latency bound either way; the FP32 kernel handles them exactly
  if ((double)M * N * K < 4.0e6) return false;
  return true;
}

static size_t operand_floats(int trans_rows, int ld) { return up((size_t)trans_rows * ld, 4); }

uint64_t gemm_tc_workspace_bytes(int mode, int transA, int transB, int M, int N, int K) {
  if (mode != IPAVSR_GEMM_TF32X3) return 0;
  // hi/lo copies of both operands, sized with padded leading dimensions (ld <= extent rounded up to 4 ... the caller's
  // ld can be larger; ipavsr_gemm re-checks against the actual ld)
  size_t a = (size_t)(transA ? K : M) * up(transA ? M : K, 4) + 64;
  size_t b = (size_t)(transB ? N : K) * up(transB ? K : N, 4) + 64;
  return 2 * (a + b) * sizeof(float) * 2;   // x2 head-room for leading dimensions up to twice the logical width
}

template <int BN, bool A_MN, bool B_MN, int NPROD, bool F16, int CG>
static int launch_tc(const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB, const CUtensorMap& mBlo,
                     TcParams p, cudaStream_t st) {
  constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = (BN / CG) * TC_BK * 4;
  constexpr int NOPER = (NPROD == 3) ? 2 : 1;
  constexpr int STAGE_BYTES = NOPER * (A_BYTES + B_BYTES);
  constexpr int SMEM_BUDGET = 200 * 1024;
  constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES < 8 ? SMEM_BUDGET / STAGE_BYTES : 8;
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, NPROD, F16, CG>;
  IPAVSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int mtiles = (p.M + TC_BM - 1) / TC_BM;
  dim3 grid((p.N + BN - 1) / BN, mtiles, p.splits);
  if (CG == 2) grid = dim3(2 * ((p.N + BN - 1) / BN), (mtiles + 1) / 2, p.splits);
  if (CG == 1) {
    kern<<<grid, TC_THREADS, smem, st>>>(mA, mAlo, mB, mBlo, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    IPAVSR_CUDA(cudaLaunchKernelEx(&cfg, kern, mA, mAlo, mB, mBlo, p));
  }
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

template <int BN, int NPROD, bool F16, int CG>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB,
                          const CUtensorMap& mBlo, TcParams p, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_tc<BN, false, false, NPROD, F16, CG>(mA, mAlo, mB, mBlo, p, st);
  if (!a_mn && b_mn) return launch_tc<BN, false, true, NPROD, F16, CG>(mA, mAlo, mB, mBlo, p, st);
  if (a_mn && !b_mn) return launch_tc<BN, true, false, NPROD, F16, CG>(mA, mAlo, mB, mBlo, p, st);
  return launch_tc<BN, true, true, NPROD, F16, CG>(mA, mAlo, mB, mBlo, p, st);
}

int amax_launch(const float* x, int ldx, int rows, int cols, float* amax, cudaStream_t st);   // f16split.cu
int gemm_f16p_launch(int bn, int kb, bool a_mn, bool b_mn, const CUtensorMap& mA, const CUtensorMap& mAlo,
                     const CUtensorMap& mB, const CUtensorMap& mBlo, const TcParams& p, cudaStream_t st);   // gemm_f16p.cu

// core: operands already split (3xTF32: fp32 hi/lo arrays; f16: fp16 hi/lo arrays + scale exponents) or raw (TF32:
// *lo ignored).  kind: 0 = single TF32, 1 = 3xTF32, 2 = fp16 three-product.
static int gemm_tc_core(int kind, int transA, int transB, int M, int N, int K, const void* Ahi, const void* Alo, int lda,
                        const void* Bhi, const void* Blo, int ldb, float* C, int ldc, const float* bias, int act,
                        int accumulate, float* Chi, float* Clo, const int32_t* expA, const int32_t* expB, float* amax,
                        cudaStream_t st, uint16_t* C16hi = nullptr, uint16_t* C16lo = nullptr, int c16_exp = 0) {
  const bool x3 = kind != 0, f16 = kind == 2;
  const bool a_mn = transA != 0;     // A stored [K,M]: M contiguous
  const bool b_mn = transB == 0;     // B stored [K,N]: N contiguous
  int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int bke = f16 ? 2 * TC_BK : TC_BK;      // elements per k-block = 128 bytes
  const int mnbox = f16 ? 64 : 32;              // MN-major box: 128 bytes of M/N, bke k-rows
  CUtensorMap mA, mAlo, mB, mBlo;
  int rc;
  // A: K-major -> tensor [M][K], box {kb, 128};  MN-major -> tensor [K][M], box {mnbox, kb};  B likewise with its share
  // bnl of the tile's columns.  kb = elements along K per stage: one 128-byte swizzle row, or half of one (64-byte
  // swizzle for the K-major operand) in the persistent kernel's deep ring.
  auto make_maps = [&](int kb, int bnl) -> int {
    const bool sw64 = f16 && kb == 32;
    int r;
    if (!a_mn) {
      if ((r = make_map(&mA, Ahi, K, M, lda, kb, TC_BM, false, f16, sw64))) return r;
      if ((r = make_map(&mAlo, Alo, K, M, lda, kb, TC_BM, false, f16, sw64))) return r;
    } else {
      if ((r = make_map(&mA, Ahi, M, K, lda, mnbox, kb, true, f16))) return r;
      if ((r = make_map(&mAlo, Alo, M, K, lda, mnbox, kb, true, f16))) return r;
    }
    if (!b_mn) {
      if ((r = make_map(&mB, Bhi, K, N, ldb, kb, bnl, false, f16, sw64))) return r;
      if ((r = make_map(&mBlo, Blo, K, N, ldb, kb, bnl, false, f16, sw64))) return r;
    } else {
      if ((r = make_map(&mB, Bhi, N, K, ldb, mnbox, kb, true, f16))) return r;
      if ((r = make_map(&mBlo, Blo, N, K, ldb, mnbox, kb, true, f16))) return r;
    }
    return IPAVSR_OK;
  };
  // CTA pairs (cta_group::2) for wide outputs with at least two row tiles; IPAVSR_GEMM_CG=1 forces single-CTA tiles
  static int cg_env = -1;
  if (cg_env < 0) {
    const char* e = getenv("IPAVSR_GEMM_CG");
    cg_env = (e && e[0] == '1') ? 1 : 2;
  }
  int cg = (BN == 256 && M > TC_BM && cg_env == 2) ? 2 : 1;
  TcParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias; p.act = act; p.accumulate = accumulate;
  p.Chi = Chi; p.Clo = Clo; p.expA = expA; p.expB = expB; p.amax = amax;
  p.dbg = g_gemm_dbg;
  static int oneacc_env = -1;
  if (oneacc_env < 0) {
    const char* e = getenv("IPAVSR_GEMM_ONEACC");
    oneacc_env = (e && e[0] == '1') ? 1 : 0;
  }
  p.oneacc = (f16 && oneacc_env) ? 1 : 0;
  p.C16hi = C16hi; p.C16lo = C16lo; p.c16_exp = c16_exp;
  const int num_kb = (K + bke - 1) / bke;
  const int tiles = ((M + TC_BM * cg - 1) / (TC_BM * cg)) * cg * ((N + BN - 1) / BN);
  int splits = 1;
  if (act == IPAVSR_ACT_LINEAR && Chi == nullptr && C16hi == nullptr) {
    if (tiles * 2 <= sm_count() && num_kb >= 32 && amax == nullptr) {         // fill the machine for skinny outputs
      splits = sm_count() / tiles;
      if (splits > num_kb / 8) splits = num_kb / 8;
    }
    // bound the accumulations into one TMEM accumulator (truncating adds): at most 64 k-blocks (K = 2048) per split
    const int acc_splits = (num_kb + 63) / 64;
    if (x3 && acc_splits > splits) splits = acc_splits;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    // Wave quantisation: when the product is split anyway, take the split count (from the minimum up to 3x it, at least
    // 8 k-blocks each) whose CTA count fills whole waves of the machine best — 32 tiles x 5 splits = 160 CTAs would run
    // as 1.08 waves of 148; 9 splits = 288 CTAs fill 1.95.
    if (splits > 1) {
      const int sms = sm_count();
      int best = splits;
      double best_eff = 0.0, min_eff = 0.0;
      for (int s = splits; s <= 3 * splits && s <= 64 && num_kb / s >= 8; ++s) {
        const int kps = (num_kb + s - 1) / s;
        const int real = (num_kb + kps - 1) / kps;
        const long long ctas = (long long)tiles * real;
        const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
        if (s == splits) min_eff = eff;
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
      }
      // every extra split costs another atomic epilogue over the whole output: only worth it for a big gain
      if (best_eff > min_eff + 0.15) splits = best;
    }
  }
  p.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  if (splits < 1) splits = 1;
  p.splits = splits;
  if (splits > 1 && !accumulate)
    IPAVSR_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st));
  // Persistent kernel (gemm_f16p.cu) for the big fp16 products: IPAVSR_GEMM_PERSIST = 1 (256-wide tiles, one accumulator),
  // 2 (128-wide tiles, main + cross accumulators), 0 = off; used when the product has more work units than CTA pairs.
  static int persist_env = -1, persist_min = 0;
  if (persist_env < 0) {
    const char* e = getenv("IPAVSR_GEMM_PERSIST");
    persist_env = e ? atoi(e) : 1;
    const char* m = getenv("IPAVSR_GEMM_PERSIST_MIN");
    persist_min = m ? atoi(m) : 2 * (sm_count() / 2);
  }
  // (measured, 13 325 rows: 424-unit products gain 12-18 % timed alone; with 3 units per pair a product timed alone gains
  //  nothing, but inside the step — where the branches' GEMMs of several streams share the machine — 2 units per pair is
  //  the better threshold (512-utterance step 7.47 -> 7.41 ms); k-split products can go to it as well
  //  (IPAVSR_GEMM_PERSIST_SPLIT=1): with their partial tiles leaving as coalesced float4 atomics from the staging tile the
  //  fc1 weight gradient takes 0.153 ms against 0.182 for the tile-per-pair kernel timed alone (with one float4 per row and
  //  lane the atomic epilogue took 21.8 us per unit, longer than the mainloop it hides under: 0.218 ms) — but inside the
  //  step, where the weight gradients overlap the other branches' kernels, it changes nothing (7.58 / 7.60 vs 7.62 / 7.55
  //  ms), so they stay on the kernel with the separate cross-term accumulator and its tighter error bound)
  static int persist_split = -1;
  if (persist_split < 0) {
    const char* e = getenv("IPAVSR_GEMM_PERSIST_SPLIT");
    persist_split = (e && e[0] == '1') ? 1 : 0;
  }
  // small K: the unit is epilogue-bound and the tile-per-pair kernel's lockstep epilogue does as well (K = 150: 41 vs 51 us)
  const bool use_p = f16 && BN == 256 && cg == 2 && persist_env > 0 && tiles / 2 * splits >= persist_min &&
                     (splits == 1 || persist_split) && num_kb / splits >= 8;
  if (use_p) {
    // halves along K per stage of the persistent kernel: 64 = 3 stages of 64 KB, 32 = 6 stages of 32 KB.  Measured at
    // 13 325 rows: no gain from the deeper ring (fc1 forward 0.143 ms either way) and the 64-byte-swizzled K-major tiles
    // cost the product with two K-major operands 27 % (fc2 dgrad 0.129 -> 0.164 ms): 64 stays the default.
    static int persist_kb = -1;
    if (persist_kb < 0) {
      const char* e = getenv("IPAVSR_GEMM_PERSIST_KB");
      persist_kb = (e && atoi(e) == 32) ? 32 : 64;
    }
    const int pbn = persist_env == 2 ? 128 : 256;
    const int pkb = pbn == 256 ? persist_kb : 64;
    if ((rc = make_maps(pkb, pbn / 2))) return rc;
    TcParams pp = p;
    pp.kb_per_split = p.kb_per_split * (bke / pkb);      // counted in blocks of pkb
    rc = gemm_f16p_launch(pbn, pkb, a_mn, b_mn, mA, mAlo, mB, mBlo, pp, st);
    if (rc != -1000) {                  // -1000: declined (no unit counter left for a captured launch)
      if (rc) return rc;
      if (amax != nullptr && splits > 1) return amax_launch(C, ldc, M, N, amax, st);
      return IPAVSR_OK;
    }
  }
  if ((rc = make_maps(bke, BN / cg))) return rc;
#define IPAVSR_TC_DISPATCH(BNV, CGV)                                                                  \
  rc = f16 ? dispatch_major<BNV, 3, true, CGV>(a_mn, b_mn, mA, mAlo, mB, mBlo, p, st)                 \
           : (x3 ? dispatch_major<BNV, 3, false, CGV>(a_mn, b_mn, mA, mAlo, mB, mBlo, p, st)          \
                 : dispatch_major<BNV, 1, false, CGV>(a_mn, b_mn, mA, mAlo, mB, mBlo, p, st))
  if (BN == 64) { IPAVSR_TC_DISPATCH(64, 1); }
  else if (BN == 128) { IPAVSR_TC_DISPATCH(128, 1); }
  else if (cg == 2) { IPAVSR_TC_DISPATCH(256, 2); }
  else { IPAVSR_TC_DISPATCH(256, 1); }
#undef IPAVSR_TC_DISPATCH
  if (rc) return rc;
  // split-K accumulates with atomics, so the kernel cannot produce |C|max itself: one extra pass over C
  if (amax != nullptr && splits > 1) return amax_launch(C, ldc, M, N, amax, st);
  return IPAVSR_OK;
}

int tf32_split_launch(const float* x, float* hi, float* lo, size_t n, cudaStream_t st) {
  const int cap = sm_count() * 8;
  size_t n4 = n / 4;
  if (n4 == 0) return IPAVSR_OK;
  tf32_split_rna_kernel<<<(int)((n4 + 255) / 256 < (size_t)cap ? (n4 + 255) / 256 : cap), 256, 0, st>>>(x, hi, lo, n4);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int gemm_tc(int mode, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
            float* C, int ldc, const float* bias, int act, int accumulate, void* ws, uint64_t ws_bytes,
            cudaStream_t st) {
  const bool x3 = mode == IPAVSR_GEMM_TF32X3;
  const float *Ahi = A, *Alo = A, *Bhi = B, *Blo = B;
  if (x3) {
    const size_t a_n = operand_floats(transA ? K : M, lda), b_n = operand_floats(transB ? N : K, ldb);
    const size_t need = (2 * a_n + 2 * b_n + 16) * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
      set_error("gemm_tc: workspace too small (%llu < %llu bytes)", (unsigned long long)ws_bytes,
                (unsigned long long)need);
      return IPAVSR_ERR_ARG;
    }
    float* w = reinterpret_cast<float*>(ws);
    float *ah = w, *al = w + a_n, *bh = w + 2 * a_n, *bl = w + 2 * a_n + b_n;
    int rc;
    if ((rc = tf32_split_launch(A, ah, al, a_n, st))) return rc;
    if ((rc = tf32_split_launch(B, bh, bl, b_n, st))) return rc;
    Ahi = ah; Alo = al; Bhi = bh; Blo = bl;
  }
  return gemm_tc_core(x3 ? 1 : 0, transA, transB, M, N, K, Ahi, Alo, lda, Bhi, Blo, ldb, C, ldc, bias, act, accumulate,
                      nullptr, nullptr, nullptr, nullptr, nullptr, st);
}

int gemm_tc_presplit(int transA, int transB, int M, int N, int K, const float* Ahi, const float* Alo, int lda,
                     const float* Bhi, const float* Blo, int ldb, float* C, int ldc, const float* bias, int act,
                     int accumulate, float* Chi, float* Clo, cudaStream_t st) {
  return gemm_tc_core(1, transA, transB, M, N, K, Ahi, Alo, lda, Bhi, Blo, ldb, C, ldc, bias, act, accumulate, Chi,
                      Clo, nullptr, nullptr, nullptr, st);
}

// fp16 three-product GEMM on operands split by f16split.cu (hi/lo fp16 arrays, leading dimensions in halves)
int gemm_tc_f16x3(int transA, int transB, int M, int N, int K, const uint16_t* Ahi, const uint16_t* Alo, int lda,
                  const int32_t* expA, const uint16_t* Bhi, const uint16_t* Blo, int ldb, const int32_t* expB, float* C,
                  int ldc, const float* bias, int act, int accumulate, float* amax, uint16_t* C16hi, uint16_t* C16lo,
                  int c16_exp, cudaStream_t st) {
  return gemm_tc_core(2, transA, transB, M, N, K, Ahi, Alo, lda, Bhi, Blo, ldb, C, ldc, bias, act, accumulate, nullptr,
                      nullptr, expA, expB, amax, st, C16hi, C16lo, c16_exp);
}

bool gemm_tc_f16_supported(int M, int N, int K, const void* A, int lda, const void* B, int ldb) {
  if (!aligned16(A) || !aligned16(B) || lda % 8 != 0 || ldb % 8 != 0) return false;
  if (K < 16 || M < 1 || N < 8) return false;
  if ((double)M * N * K < 4.0e6) return false;
  return true;
}

}  // namespace ipavsr

extern "C" int ipavsr_debug_gemm_timestamps(unsigned long long* buf) {
  ipavsr::g_gemm_dbg = buf;
  return IPAVSR_OK;
}
