// Persistent fp16 three-product GEMM for sm_100a (the big products of IPAVSR_GEMM_F16X3: DBNF encoder layers, hoisted LSTM
// input projections and their dgrad / wgrad — reference modelzoo/pretrained_encoder.py:4-9, T.dot behind Lasagne's
// DenseLayer / LSTMLayer).  gemm_tc.cu runs one CTA pair per output tile: its 2.3 us of setup and 6 us of epilogue per
// 18 us mainloop are exposed.  Here ONE CTA pair per TPC stays resident and walks over work units (output tile x k-split)
// drawn from a device-side counter, with the accumulators double-buffered in TMEM so that the epilogue of unit i runs
// under the mainloop of unit i+1:
//
//   warp 0      TMA producer (one lane; both CTAs)   operand stages, ring continues across units
//   warp 1      MMA issuer (one lane; leader CTA)    tcgen05.mma.cta_group::2.kind::f16 into TMEM buffer (unit & 1)
//   warp 2      TMEM allocator
//   warp 3      unit scheduler (one lane; leader)    atomicAdd on the launch's counter, unit ids published to both CTAs
//   warps 4-11  epilogue                             tcgen05.ld -> scale, bias, activation -> coalesced stores (+ fp16 pair)
//
// Tile shapes (both fill the 512 TMEM columns with two buffers):
//   BN = 256, ONEACC: lo*hi + hi*lo + hi*hi accumulate in ONE fp32 accumulator per buffer (possible because the lo halves
//                     are unscaled, common.cuh F16_LO_SCALE); three truncating adds per k-step instead of one.
//   BN = 128        : main (hi*hi) and cross (lo*hi + hi*lo) accumulators per buffer, as in gemm_tc.cu.
// The dynamic scheduler (instead of a static round-robin) matters because the engine runs recurrence kernels on side
// streams: a pair that becomes resident late simply draws fewer units.
#include <stdlib.h>
#include <atomic>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ipavsr {

constexpr int P_THREADS = 384;
constexpr int P_EPI_WARPS = 8;
constexpr int P_NSLOT = 4;                       // unit ids in flight per pair
constexpr int P_STG_BYTES = P_EPI_WARPS * 4096;  // one 32 x 32 float staging tile per epilogue warp
constexpr int P_SCHED_RING = 1024;               // launches in flight that can hold distinct counters
constexpr int P_SCHED_GRAPH = 8192;              // counters owned for good by kernel nodes of captured CUDA graphs

// [2 * i] = next unit, [2 * i + 1] = pairs that have drawn their end marker; the last pair resets both to 0
__device__ unsigned int g_f16p_sched[2 * (P_SCHED_RING + P_SCHED_GRAPH)];

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAITC_DONE;\n"
      "bra WAITC_LOOP;\n"
      "WAITC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// arrive on a barrier of CTA `rank` of the cluster (cluster-scope release: earlier writes of this thread, local or
// st.shared::cluster, are visible to a waiter that acquires at cluster scope)
__device__ __forceinline__ void mbar_arrive_rank(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_shared(smem_u32(bar), rank))
               : "memory");
}
__device__ __forceinline__ void st_shared_rank(int* addr, uint32_t rank, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(mapa_shared(smem_u32(addr), rank)), "r"(v) : "memory");
}

constexpr int IPAVSR_F16P_DECLINED = -1000;      // not an error: no counter left for a captured launch

static std::atomic<uint64_t> g_f16p_launches{0};

struct PUnit {
  int m0, n0, z, kb_begin, nkb;
};

__device__ __forceinline__ PUnit decode_unit(int u, int ncol, int nrp, uint32_t crank, int BN, const TcParams& p,
                                             int num_kb_total) {
  PUnit r;
  const int per = ncol * nrp;
  r.z = u / per;
  const int rem = u - r.z * per;
  const int rp = rem / ncol;
  r.n0 = (rem - rp * ncol) * BN;
  r.m0 = (rp * 2 + (int)crank) * TC_BM;
  r.kb_begin = r.z * p.kb_per_split;
  const int kb_end = min(num_kb_total, r.kb_begin + p.kb_per_split);
  r.nkb = max(kb_end - r.kb_begin, 0);
  return r;
}

// KB = halves along K per operand stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B for a K-major
// operand; an MN-major operand keeps its 128-byte rows of M/N and just takes half as many k-rows).  Half-size stages
// double the ring depth in the same shared memory: a stage can only be refilled once the MMAs that read it have
// completed, so the bytes in flight are (STAGES - 1) / STAGES of the ring — 5/6 instead of 2/3 for the 256-wide tile.
template <int BN, bool A_MN, bool B_MN, bool ONEACC, int KB>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_f16p_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo, TcParams p,
                 int n_units, int ncol, int nrp, unsigned int* sched, int max_units) {
  constexpr int BKE = KB;                             // halves along K per stage
  constexpr int MNBOX = 64;                           // MN-major TMA box: 128 bytes of M/N x BKE k-rows
  constexpr int MNBOX_BYTES = BKE * 128;
  constexpr int A_BYTES = TC_BM * BKE * 2;            // 16 KB at KB = 64
  constexpr int BNL = BN / 2;                         // B-tile columns this CTA loads
  constexpr int B_BYTES = BNL * BKE * 2;
  constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  constexpr int STAGES = (193 * 1024) / STAGE_BYTES < 8 ? (193 * 1024) / STAGE_BYTES : 8;
  static_assert(STAGES >= 3 && STAGES <= 8, "operand ring depth");
  static_assert(KB == 64 || KB == 32, "one or half a 128-byte swizzle row along K");
  // K-major operand tile: rows of KB halves; 8-row swizzle atoms of 8 * 2 * KB bytes (SBO); layout 2 = 128B, 4 = 64B
  constexpr uint32_t KM_SBO = 16 * KB, KM_LT = KB == 64 ? 2 : 4;
  constexpr int ACC_COLS = ONEACC ? BN : 2 * BN;      // TMEM columns of one accumulator buffer
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "two accumulator buffers in a power-of-two TMEM allocation");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* stg_base = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2], sfull_bar[P_NSLOT],
      sempty_bar[P_NSLOT];
  __shared__ int s_unit[P_NSLOT];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();            // 0 = leader of the pair
  if (p.dbg != nullptr && crank == 0 && threadIdx.x == 0) p.dbg[8ull * (blockIdx.x >> 1)] = gtimer();
  const int num_kb_total = (p.K + BKE - 1) / BKE;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 2 * P_EPI_WARPS);       // epilogue warps of both CTAs (leader's copy is the one used)
    }
    for (int s = 0; s < P_NSLOT; ++s) {
      mbar_init(&sfull_bar[s], 1);
      mbar_init(&sempty_bar[s], 2 * (1 + P_EPI_WARPS) + 1);   // producers + epilogue warps of both CTAs + the MMA issuer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // profiling aid (tools/gemm_phases.py): per pair [start, setup done, units, MMA wait for TMEM ns, MMA wait for operands
  // ns, end of the last epilogue, epilogue busy ns, epilogue wait for accumulators ns] (leader CTA, epilogue warp 4)
  unsigned long long* dbg = (p.dbg != nullptr && crank == 0) ? p.dbg + 8ull * (blockIdx.x >> 1) : nullptr;
  if (dbg != nullptr && threadIdx.x == 0) dbg[1] = gtimer();

  if (warp == 3) {
    if (lane == 0 && crank == 0) {
      // ===================== unit scheduler (leader) =====================
      const unsigned int n_pairs = gridDim.x >> 1;
      // a pair retires after max_units units (the grid holds enough pairs for all of them): kernels of other streams —
      // the recurrence kernels the step's critical path waits on — get SMs at that granularity, not at the kernel's end
      for (uint32_t s = 0;; ++s) {
        const uint32_t slot = s % P_NSLOT, ph = (s / P_NSLOT) & 1;
        mbar_wait_cluster(&sempty_bar[slot], ph ^ 1);
        int u = -1;
        if (s < (uint32_t)max_units) {
          const unsigned int got = atomicAdd(&sched[0], 1u);
          if (got < (unsigned int)n_units) u = (int)got;
        }
        s_unit[slot] = u;
        st_shared_rank(&s_unit[slot], 1, u);
        mbar_arrive_rank(&sfull_bar[slot], 0);
        mbar_arrive_rank(&sfull_bar[slot], 1);
        if (u < 0) {
          // every pair draws exactly one end marker; after the last one nobody touches the counters again
          __threadfence();
          if (atomicAdd(&sched[1], 1u) == n_pairs - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
          }
          break;
        }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t fb0 = mapa_shared(smem_u32(&full_bar[0]), 0);
      for (uint32_t s = 0;; ++s) {
        const uint32_t slot = s % P_NSLOT, sph = (s / P_NSLOT) & 1;
        mbar_wait_cluster(&sfull_bar[slot], sph);
        const int u = s_unit[slot];
        mbar_arrive_rank(&sempty_bar[slot], 0);
        if (u < 0) break;
        const PUnit w = decode_unit(u, ncol, nrp, crank, BN, p, num_kb_total);
        const int nl0 = w.n0 + (int)crank * BNL;
        for (int kb = 0; kb < w.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * STAGE_BYTES;
          uint8_t* sB = sA + 2 * A_BYTES;
          const int k0 = (w.kb_begin + kb) * BKE;
          // both CTAs' loads complete on the leader's barrier, which therefore expects the bytes of both
          if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const uint32_t fb = fb0 + (uint32_t)stage * 8u;
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const CUtensorMap* ma = o == 0 ? &mapA : &mapAlo;
            const CUtensorMap* mb = o == 0 ? &mapB : &mapBlo;
            if (!A_MN) {
              tma_load_2d_cg2(sA + o * A_BYTES, ma, fb, k0, w.m0);
            } else {
#pragma unroll
              for (int j = 0; j < TC_BM / MNBOX; ++j)
                tma_load_2d_cg2(sA + o * A_BYTES + j * MNBOX_BYTES, ma, fb, w.m0 + MNBOX * j, k0);
            }
            if (!B_MN) {
              tma_load_2d_cg2(sB + o * B_BYTES, mb, fb, k0, nl0);
            } else {
#pragma unroll
              for (int j = 0; j < BNL / MNBOX; ++j)
                tma_load_2d_cg2(sB + o * B_BYTES + j * MNBOX_BYTES, mb, fb, nl0 + MNBOX * j, k0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN, true, TC_BM * 2);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ul = 0;                                 // units this pair has started
      for (uint32_t s = 0;; ++s) {
        const uint32_t slot = s % P_NSLOT, sph = (s / P_NSLOT) & 1;
        mbar_wait_cluster(&sfull_bar[slot], sph);
        const int u = s_unit[slot];
        mbar_arrive_rank(&sempty_bar[slot], 0);
        if (u < 0) break;
        const PUnit w = decode_unit(u, ncol, nrp, 0, BN, p, num_kb_total);
        const uint32_t buf = ul & 1, uph = (ul >> 1) & 1;
        unsigned long long tw = dbg != nullptr ? gtimer() : 0ull;
        mbar_wait_cluster(&tempty_bar[buf], uph ^ 1);  // both CTAs' epilogues have drained this buffer
        tcgen05_fence_after();
        if (dbg != nullptr) { dbg[3] += gtimer() - tw; dbg[2] += 1; }
        const uint32_t t_main = tmem_base + buf * (uint32_t)ACC_COLS;
        const uint32_t t_cross = ONEACC ? t_main : t_main + (uint32_t)BN;
        for (int kb = 0; kb < w.nkb; ++kb) {
          if (dbg != nullptr) tw = gtimer();
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (dbg != nullptr) dbg[4] += gtimer() - tw;
          const uint32_t sA = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint32_t sB = sA + 2 * A_BYTES;
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            // K-major: advance 32 bytes inside the swizzled row, SBO = 8 rows.
            // MN-major fp16: canonical SWIZZLE_128B atoms of 8 k-rows x 128 bytes; one MMA (K = 16) spans two atoms:
            // advance 2048 bytes, SBO = 1024 (next atom along K), LBO = next TMA box along M/N.
            const uint32_t a_off = A_MN ? k * 2048 : k * 32;
            const uint32_t b_off = B_MN ? k * 2048 : k * 32;
            const uint32_t a_lbo = A_MN ? MNBOX_BYTES : 16, b_lbo = B_MN ? MNBOX_BYTES : 16;
            const uint32_t a_sbo = A_MN ? 1024 : KM_SBO, b_sbo = B_MN ? 1024 : KM_SBO;
            const uint32_t a_lt = A_MN ? 2 : KM_LT, b_lt = B_MN ? 2 : KM_LT;
            const uint64_t a_hi = make_smem_desc(sA + a_off, a_lbo, a_sbo, a_lt);
            const uint64_t b_hi = make_smem_desc(sB + b_off, b_lbo, b_sbo, b_lt);
            const uint64_t a_lo = make_smem_desc(sA + A_BYTES + a_off, a_lbo, a_sbo, a_lt);
            const uint64_t b_lo = make_smem_desc(sB + B_BYTES + b_off, b_lbo, b_sbo, b_lt);
            const uint32_t first = (kb | k) == 0 ? 0u : 1u;
            tcgen05_mma_cg2<true>(t_cross, a_lo, b_hi, idesc, first);
            tcgen05_mma_cg2<true>(t_cross, a_hi, b_lo, idesc, 1u);
            tcgen05_mma_cg2<true>(t_main, a_hi, b_hi, idesc, ONEACC ? 1u : first);
          }
          tcgen05_commit_cg2(&empty_bar[stage]);        // frees this stage in both CTAs once the MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_cg2(&tfull_bar[buf]);            // accumulators of this unit complete (both CTAs' epilogues)
        ++ul;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps =====================
    // Warp w reads TMEM lanes 32 * (w % 4) .. (one output row per thread); warps 4-7 take the even 32-column chunks,
    // warps 8-11 the odd ones.
    const int ew = warp - 4;
    const int q = warp & 3;
    const int half = ew >> 2;
    float* stg = stg_base + ew * 1024;
    const bool vecC = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    const bool split = p.splits > 1;
    const bool vecB = (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
    float ms1, ms2;
    {
      // result = acc * 2^-(eA+eB); one exact factor when it is a normal float, else two
      const int e = -(__ldg(p.expA) + __ldg(p.expB));
      const int e1 = (e > 126 || e < -115) ? e / 2 : e, e2 = e - e1;
      ms1 = __int_as_float((127 + e1) << 23);
      ms2 = __int_as_float((127 + e2) << 23);
    }
    const float cs1 = ms1 * F16_LO_INV;
    float tile_max = 0.f;
    uint32_t ul = 0;
    for (uint32_t s = 0;; ++s) {
      const uint32_t slot = s % P_NSLOT, sph = (s / P_NSLOT) & 1;
      mbar_wait_cluster(&sfull_bar[slot], sph);
      const int u = s_unit[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive_rank(&sempty_bar[slot], 0);
      if (u < 0) break;
      const PUnit w = decode_unit(u, ncol, nrp, crank, BN, p, num_kb_total);
      const uint32_t buf = ul & 1, uph = (ul >> 1) & 1;
      const int m0 = w.m0, n0 = w.n0;
      const int row = m0 + q * 32 + lane;
      unsigned long long te = (dbg != nullptr && threadIdx.x == 128) ? gtimer() : 0ull;
      mbar_wait(&tfull_bar[buf], uph);
      tcgen05_fence_after();
      if (dbg != nullptr && threadIdx.x == 128) { const unsigned long long t1 = gtimer(); dbg[7] += t1 - te; te = t1; }
      const bool fast_ok = vecC;                    // warp-uniform (k-split units: coalesced float4 atomics)
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 64) {
        if (n0 + c0 >= p.N) break;                   // warp-uniform
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)ACC_COLS + (uint32_t)c0;
        float v[32];
        if (ONEACC) {
          tmem_ld32_issue(lane_base, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= ms1;
        } else {
          float x[32];
          tmem_ld32_issue(lane_base + (uint32_t)BN, v);     // cross terms
          tmem_ld32_issue(lane_base, x);                    // main
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], cs1, x[j] * ms1);
        }
        if (ms2 != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= ms2;
        }
        if (fast_ok && n0 + c0 + 32 <= p.N) {
          // ---------------- fast path: a full 32-column chunk, coalesced stores through the staging tile ----------------
          if (row < p.M) {
            if (p.accumulate && !split) {
              const float* cpo = p.C + (size_t)row * p.ldc + n0 + c0;
#pragma unroll
              for (int j4 = 0; j4 < 32; j4 += 4) {
                const float4 old = *reinterpret_cast<const float4*>(cpo + j4);
                v[j4] += old.x; v[j4 + 1] += old.y; v[j4 + 2] += old.z; v[j4 + 3] += old.w;
              }
            }
            if (p.bias != nullptr && (!split || w.z == 0)) {
              if (vecB) {
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j4));
                  v[j4] += b4.x; v[j4 + 1] += b4.y; v[j4 + 2] += b4.z; v[j4 + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + n0 + c0 + j);
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
            if (p.amax != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) tile_max = fmaxf(tile_max, fabsf(v[j]));
            }
          }
          // thread = row writes its 8 float4 at XOR-swizzled positions (conflict-free), then every store instruction
          // covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes
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
              // fp16 hi/lo split of the chunk under the static scale, from the same staging tile
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
          continue;
        }
        // ---------------- general path: ragged edges, split-K atomics, unaligned C ----------------
        if (row < p.M) {
          float* cp = p.C + (size_t)row * p.ldc + n0 + c0;
          if (split && vecC && n0 + c0 + 32 <= p.N) {
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              float4 add = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
              if (p.bias != nullptr && w.z == 0) {
                add.x += __ldg(p.bias + n0 + c0 + j4); add.y += __ldg(p.bias + n0 + c0 + j4 + 1);
                add.z += __ldg(p.bias + n0 + c0 + j4 + 2); add.w += __ldg(p.bias + n0 + c0 + j4 + 3);
              }
              atomicAdd(reinterpret_cast<float4*>(cp + j4), add);
            }
          } else if (split) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < p.N) {
                float add = v[j];
                if (p.bias != nullptr && w.z == 0) add += __ldg(p.bias + n0 + c0 + j);
                atomicAdd(cp + j, add);
              }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + c0 + j < p.N) {
                float o = v[j];
                if (p.accumulate) o += cp[j];
                if (p.bias != nullptr) o += __ldg(p.bias + n0 + c0 + j);
                o = act_epi(o, p.act);
                cp[j] = o;
                tile_max = fmaxf(tile_max, fabsf(o));
                if (p.C16hi != nullptr) {
                  const float xs = o * __int_as_float((127 + p.c16_exp) << 23);
                  const __half h = __float2half_rn(xs);
                  reinterpret_cast<__half*>(p.C16hi)[(size_t)row * p.ldc + n0 + c0 + j] = h;
                  reinterpret_cast<__half*>(p.C16lo)[(size_t)row * p.ldc + n0 + c0 + j] =
                      __float2half_rn((xs - __half2float(h)) * F16_LO_SCALE);
                }
              }
          }
        }
      }
      // this warp's TMEM reads of the unit are complete: hand the buffer back to the MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_rank(&tempty_bar[buf], 0);
      if (dbg != nullptr && threadIdx.x == 128) { const unsigned long long t1 = gtimer(); dbg[6] += t1 - te; dbg[5] = t1; }
      ++ul;
    }
    if (p.amax != nullptr && !split) {
      tile_max = warp_max(tile_max);
      if (lane == 0 && tile_max > 0.f) atomicMax(reinterpret_cast<unsigned int*>(p.amax), __float_as_uint(tile_max));
    }
  }
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();     // the leader's MMAs read the peer's shared memory and write its TMEM
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, bool ONEACC, int KB>
static int launch_f16p(const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB, const CUtensorMap& mBlo,
                       TcParams p, cudaStream_t st) {
  constexpr int A_BYTES = TC_BM * KB * 2, B_BYTES = (BN / 2) * KB * 2;
  constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  constexpr int STAGES = (193 * 1024) / STAGE_BYTES < 8 ? (193 * 1024) / STAGE_BYTES : 8;
  const size_t smem = (size_t)STAGES * STAGE_BYTES + P_STG_BYTES + 1024;
  auto kern = gemm_f16p_kernel<BN, A_MN, B_MN, ONEACC, KB>;
  static int max_pairs = 0;                         // per instantiation
  if (max_pairs == 0) {
    IPAVSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(2 * (sm_count() / 2));
    q.blockDim = dim3(P_THREADS);
    q.dynamicSmemBytes = smem;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = sm_count() / 2;
    }
    max_pairs = n < sm_count() / 2 ? n : sm_count() / 2;
    if (getenv("IPAVSR_GEMM_DEBUG")) fprintf(stderr, "gemm_f16p<%d, kb %d>: %d resident CTA pairs (occupancy query %d)\n", BN, KB, max_pairs, n);
  }
  const int ncol = (p.N + BN - 1) / BN;
  const int nrp = ((p.M + TC_BM - 1) / TC_BM + 1) / 2;
  const int n_units = ncol * nrp * p.splits;
  // units per pair: IPAVSR_GEMM_PERSIST_UNITS (0 = until the work runs out: max_pairs pairs stay for the whole product)
  static int unit_cap = -1;
  if (unit_cap < 0) {
    const char* e = getenv("IPAVSR_GEMM_PERSIST_UNITS");
    unit_cap = e ? atoi(e) : 0;
  }
  int pairs = n_units < max_pairs ? n_units : max_pairs;
  int max_units = n_units;
  if (unit_cap > 0 && (long long)pairs * unit_cap < n_units) {
    max_units = unit_cap;
    pairs = (n_units + unit_cap - 1) / unit_cap;
  }
  // the unit counter of this launch: eager launches rotate over a ring (a counter is back at 0 when its kernel ends); a
  // kernel node of a graph under capture is replayed with the same arguments, so it gets a counter nobody else will use
  static std::atomic<unsigned int> next_slot{0}, next_graph_slot{0};
  static unsigned int* sched_base = nullptr;
  if (sched_base == nullptr) IPAVSR_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&sched_base), g_f16p_sched));
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
    cudaGetLastError();
    cap = cudaStreamCaptureStatusNone;
  }
  unsigned int slot;
  if (cap == cudaStreamCaptureStatusNone) {
    slot = next_slot.fetch_add(1, std::memory_order_relaxed) % P_SCHED_RING;
  } else {
    const unsigned int g = next_graph_slot.fetch_add(1, std::memory_order_relaxed);
    if (g >= (unsigned int)P_SCHED_GRAPH) return IPAVSR_F16P_DECLINED;     // the caller takes the one-tile-per-pair kernel
    slot = P_SCHED_RING + g;
  }
  unsigned int* sched = sched_base + 2 * slot;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  IPAVSR_CUDA(cudaLaunchKernelEx(&cfg, kern, mA, mAlo, mB, mBlo, p, n_units, ncol, nrp, sched, max_units));
  IPAVSR_LAUNCH_CHECK();
  g_f16p_launches.fetch_add(1, std::memory_order_relaxed);
  return IPAVSR_OK;
}

template <int BN, bool ONEACC, int KB>
static int dispatch_f16p(bool a_mn, bool b_mn, const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB,
                         const CUtensorMap& mBlo, const TcParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_f16p<BN, false, false, ONEACC, KB>(mA, mAlo, mB, mBlo, p, st);
  if (!a_mn && b_mn) return launch_f16p<BN, false, true, ONEACC, KB>(mA, mAlo, mB, mBlo, p, st);
  if (a_mn && !b_mn) return launch_f16p<BN, true, false, ONEACC, KB>(mA, mAlo, mB, mBlo, p, st);
  return launch_f16p<BN, true, true, ONEACC, KB>(mA, mAlo, mB, mBlo, p, st);
}

// bn = 256: one accumulator per buffer; bn = 128: main + cross accumulators.  The tensor maps are those of the pair
// kernel of gemm_tc.cu (A box 128 rows, B box bn / 2 columns).
// kb = 64 / 32: halves along K per stage; the maps must have been made for it (K-major box {kb, rows}, 64-byte swizzle
// at kb = 32; MN-major box {64, kb}) and p.kb_per_split counts blocks of kb.
int gemm_f16p_launch(int bn, int kb, bool a_mn, bool b_mn, const CUtensorMap& mA, const CUtensorMap& mAlo,
                     const CUtensorMap& mB, const CUtensorMap& mBlo, const TcParams& p, cudaStream_t st) {
  if (bn == 256)
    return kb == 32 ? dispatch_f16p<256, true, 32>(a_mn, b_mn, mA, mAlo, mB, mBlo, p, st)
                    : dispatch_f16p<256, true, 64>(a_mn, b_mn, mA, mAlo, mB, mBlo, p, st);
  return dispatch_f16p<128, false, 64>(a_mn, b_mn, mA, mAlo, mB, mBlo, p, st);
}

}  // namespace ipavsr

extern "C" uint64_t ipavsr_debug_gemm_persistent_launches(void) {
  return ipavsr::g_f16p_launches.load(std::memory_order_relaxed);
}
