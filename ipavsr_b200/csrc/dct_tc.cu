// compute_dct_features' projection on the tensor cores (reference utils/preprocessing.py:417-462: X.dot(basis) for the
// first K zig-zag coefficients, K <= 32): out (frames, K) = x (frames, D) * basis (D, K) in float32 accuracy.
//
// The product reads 4 D bytes per frame and writes 4 K: with D = 1200, K = 30 it is a stream over x — but at 36 000
// multiply-adds per frame the FFMA kernel (features.cu) is bound by the FP32 pipe at 42 % of HBM.  Here the multiply-adds
// go to tcgen05 (kind::tf32, three-product split for float32 accuracy) and the operand split happens IN the kernel, on
// the shared-memory tiles, so x crosses HBM exactly once:
//
//   warp 0      TMA producer     x tile [128 frames x 32 floats] (K-major, SWIZZLE_128B) and basis tile [32 k-rows x 32
//                                coefficients] (MN-major, SWIZZLE_128B_ATOM_32B) per k-block, 4-stage ring
//   warps 2-5   converters       second tile <- x - trunc_tf32(x) for the x tile (the tensor core reads the upper 19 bits of
//                                a 32-bit operand, so the raw tile is the hi operand), hi / lo by round-to-nearest in place
//                                for the small basis tile; element-wise, so the swizzle never matters; fence.proxy.async
//                                hands the tiles to the tensor core
//   warp 1      MMA issuer       lo*hi + hi*lo into the cross accumulator, hi*hi into the main one (M = 128, N = 32, K = 8)
//   warps 6-9   epilogue         tcgen05.ld of both accumulators (double-buffered in TMEM across row tiles), K floats per row
//
// Persistent: CTA c walks over the 128-frame tiles c, c + gridDim.x, ...  Measured (524 288 frames, D = 1200, K = 30):
// 0.683 ms = 57.7 % of HBM against 0.935 ms = 42 % for the FFMA kernel; the bound is now shared memory — every x tile is
// written once by TMA, read and written once by the converters and read three times by the MMAs (hi twice, lo once).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ipavsr {

int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
             uint32_t box_outer, bool mn_major, bool f16, bool sw64);      // gemm_tc.cu

constexpr int DT_THREADS = 320;
constexpr int DT_STAGES = 4;
constexpr int DT_BN = 32;
constexpr int DT_A_BYTES = TC_BM * TC_BK * 4;        // 16 KB: 128 frames x 32 floats
constexpr int DT_B_BYTES = DT_BN * TC_BK * 4;        // 4 KB: 32 k-rows x 32 coefficients
constexpr int DT_STAGE_BYTES = 2 * (DT_A_BYTES + DT_B_BYTES);
constexpr int DT_SMEM = DT_STAGES * DT_STAGE_BYTES + 1024;
constexpr int DT_TMEM_COLS = 128;                    // 2 buffers x (main 32 + cross 32)

__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(DT_THREADS, 1)
dct_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapB, float* __restrict__ out,
              int ldo, long long frames, int D, int K, int n_tiles) {
  extern __shared__ uint8_t dt_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dt_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[DT_STAGES], conv_bar[DT_STAGES], empty_bar[DT_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (D + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < DT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], 4);           // one arrival per converter warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);         // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"((uint32_t)DT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * DT_STAGE_BYTES;
          uint8_t* sB = sA + 2 * DT_A_BYTES;
          mbar_expect_tx(&full_bar[stage], DT_A_BYTES + DT_B_BYTES);
          tma_load_2d(sA, &mapX, &full_bar[stage], kb * TC_BK, tile * TC_BM);        // box {32 k, 128 frames}
          tma_load_2d(sB, &mapB, &full_bar[stage], 0, kb * TC_BK);                  // box {32 coefficients, 32 k-rows}
          if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc(DT_BN, false, true, false, TC_BM);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1, uph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[buf], uph ^ 1);
        tcgen05_fence_after();
        const uint32_t t_main = tmem_base + buf * 64u, t_cross = t_main + 32u;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&conv_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sA = smem_u32(smem + (size_t)stage * DT_STAGE_BYTES);
          const uint32_t sB = sA + 2 * DT_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            // A K-major SWIZZLE_128B: 32 bytes per k-step, SBO = 1024.  B MN-major tf32 (SWIZZLE_128B_BASE32B, atoms of
            // 4 k-rows x 128 bytes): 8 k-rows = 1024 bytes per k-step, SBO = 512, LBO = one 32-wide box (unused: N = 32).
            const uint64_t a_hi = make_smem_desc(sA + k * 32, 16, 1024, 2);
            const uint64_t a_lo = make_smem_desc(sA + DT_A_BYTES + k * 32, 16, 1024, 2);
            const uint64_t b_hi = make_smem_desc(sB + k * 1024, 4096, 512, 1);
            const uint64_t b_lo = make_smem_desc(sB + DT_B_BYTES + k * 1024, 4096, 512, 1);
            const uint32_t first = (kb | k) == 0 ? 0u : 1u;
            tcgen05_mma_tf32(t_cross, a_lo, b_hi, idesc, first);
            tcgen05_mma_tf32(t_cross, a_hi, b_lo, idesc, 1u);
            tcgen05_mma_tf32(t_main, a_hi, b_hi, idesc, first);
          }
          tcgen05_commit(&empty_bar[stage]);
          if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tfull_bar[buf]);
      }
    }
  } else if (warp < 6) {
    // ===================== converters: 128 threads =====================
    const int tc = threadIdx.x - 64;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        uint8_t* sA = smem + (size_t)stage * DT_STAGE_BYTES;
        float4* a = reinterpret_cast<float4*>(sA);
        float4* alo = reinterpret_cast<float4*>(sA + DT_A_BYTES);
#pragma unroll
        for (int i = 0; i < DT_A_BYTES / 16 / 128; ++i) {
          // the tensor core reads the upper 19 bits of a 32-bit operand, i.e. the raw tile IS the hi operand
          // trunc_tf32(x); only the residual x - trunc_tf32(x) (exact in float32) is written
          const float4 v = a[tc + 128 * i];
          float4 l;
          l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          alo[tc + 128 * i] = l;
        }
        float4* b = reinterpret_cast<float4*>(sA + 2 * DT_A_BYTES);
        float4* blo = reinterpret_cast<float4*>(sA + 2 * DT_A_BYTES + DT_B_BYTES);
#pragma unroll
        for (int i = 0; i < DT_B_BYTES / 16 / 128; ++i) {
          const float4 v = b[tc + 128 * i];
          float4 h, l;
          tf32_hi_lo(v.x, h.x, l.x); tf32_hi_lo(v.y, h.y, l.y); tf32_hi_lo(v.z, h.z, l.z); tf32_hi_lo(v.w, h.w, l.w);
          b[tc + 128 * i] = h;
          blo[tc + 128 * i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive_local(&conv_bar[stage]);
        if (++stage == DT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 6-9, TMEM lane quadrant warp % 4 =====================
    const int q = warp & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1, uph = (it >> 1) & 1;
      mbar_wait(&tfull_bar[buf], uph);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64u;
      float m[32], c[32];
      tmem_ld32_issue(taddr + 32u, c);
      tmem_ld32_issue(taddr, m);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&tempty_bar[buf]);
      const long long row = (long long)tile * TC_BM + q * 32 + lane;
      if (row < frames) {
        float* o = out + row * ldo;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < K) o[j] = m[j] + c[j];
      }
    }
  }
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)DT_TMEM_COLS)
                 : "memory");
  }
}

// returns IPAVSR_OK when the tensor-core kernel took the product, 1 when the shape is not its (caller falls back)
int dct_project_tc(const float* x, int ldx, const float* basis, int ldb, float* out, int ldo, int64_t frames, int D, int K,
                   cudaStream_t st) {
  const bool ok = K <= DT_BN && D >= 64 && frames >= 1024 && D % 4 == 0 && ldx % 4 == 0 && ldb % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(basis) & 15) == 0 &&
                  frames < ((int64_t)1 << 31);
  if (!ok) return 1;
  CUtensorMap mX, mB;
  int rc;
  if ((rc = make_map(&mX, x, D, frames, ldx, TC_BK, TC_BM, false, false, false))) return rc;
  if ((rc = make_map(&mB, basis, K, D, ldb, DT_BN, TC_BK, true, false, false))) return rc;
  static bool attr = false;
  if (!attr) {
    IPAVSR_CUDA(cudaFuncSetAttribute(dct_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
    attr = true;
  }
  const int n_tiles = (int)((frames + TC_BM - 1) / TC_BM);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  dct_tc_kernel<<<grid, DT_THREADS, DT_SMEM, st>>>(mX, mB, out, ldo, (long long)frames, D, K, n_tiles);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // namespace ipavsr
