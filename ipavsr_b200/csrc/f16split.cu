// Operand preparation for the fp16 three-product GEMM mode (IPAVSR_GEMM_F16X3, gemm_tc.cu).
//
// A float32 tensor x is represented as  x * 2^e = hi + lo * 2^-11  with hi, lo in IEEE fp16 and one power-of-two
// scale per tensor:  e = 14 - floor(log2(max|x|))  puts the largest magnitude in [2^14, 2^15) (fp16 overflows at
// 65504), hi = rn_f16(x 2^e), lo = rn_f16((x 2^e - hi) 2^11).  hi carries 11 significant bits and lo the next 11, so
// the pair holds x to ~2^-22 relative — the same as the tf32 hi/lo pair — for every element within 2^29 of the
// tensor's maximum, degrading gracefully (fp16 subnormals) below that.  All scalings are by powers of two: exact.
//
// Two passes over the tensor (max reduction, then split); the reduction is skipped when the producer already
// accumulated |x|max (the GEMM epilogue does, gemm_tc.cu).  The segmented form splits the whole flat parameter arena
// in one go, with one scale per parameter tensor (segment ids per 256-float block, as for adam_vlr).
#include <cuda_fp16.h>
#include "common.cuh"

namespace ipavsr {

__device__ __forceinline__ int f16_scale_exp(float amax) {
  if (!(amax > 0.f) || isinf(amax)) return 0;
  int e = 14 - ilogbf(amax);
  return e < -126 ? -126 : (e > 126 ? 126 : e);
}

__device__ __forceinline__ void f16_scale_factors(int e, float& s1, float& s2) {
  const int e1 = e / 2, e2 = e - e1;
  s1 = __int_as_float((127 + e1) << 23);
  s2 = __int_as_float((127 + e2) << 23);
}

__device__ __forceinline__ void f16_hi_lo(float x, float s1, float s2, __half& hi, __half& lo) {
  const float xs = x * s1 * s2;
  hi = __float2half_rn(xs);
  lo = __float2half_rn((xs - __half2float(hi)) * F16_LO_SCALE);
}

__device__ __forceinline__ void atomic_max_pos(float* addr, float v) {
  if (v > 0.f) atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ---- |x|max over a (rows x cols) matrix, atomically max-combined into amax[0] -------------------------------
__global__ void __launch_bounds__(256) amax2d_kernel(const float* __restrict__ x, int ldx, long long rows, int cols,
                                                     float* __restrict__ amax, int vec) {
  float m = 0.f;
  if (vec) {
    const int c4 = cols >> 2;
    const long long total = rows * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / c4;
      const int c = (int)(i - r * c4) << 2;
      const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / cols;
      m = fmaxf(m, fabsf(x[r * ldx + (i - r * cols)]));
    }
  }
  m = warp_max(m);
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = wm[0];
    for (int i = 1; i < 8; ++i) b = fmaxf(b, wm[i]);
    atomic_max_pos(amax, b);
  }
}

__global__ void __launch_bounds__(256) f16_split2d_kernel(const float* __restrict__ x, int ldx, long long rows, int cols,
                                                          __half* __restrict__ hi, __half* __restrict__ lo, int ldo,
                                                          const float* __restrict__ amax, int32_t* __restrict__ exp_out,
                                                          int vec) {
  const int e = f16_scale_exp(__ldg(amax));
  float s1, s2;
  f16_scale_factors(e, s1, s2);
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e;
  if (vec) {
    const int c4 = cols >> 2;
    const long long total = rows * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / c4;
      const int c = (int)(i - r * c4) << 2;
      const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
      __half h[4], l[4];
      f16_hi_lo(v.x, s1, s2, h[0], l[0]);
      f16_hi_lo(v.y, s1, s2, h[1], l[1]);
      f16_hi_lo(v.z, s1, s2, h[2], l[2]);
      f16_hi_lo(v.w, s1, s2, h[3], l[3]);
      *reinterpret_cast<uint2*>(hi + r * ldo + c) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(lo + r * ldo + c) = *reinterpret_cast<const uint2*>(l);
    }
  } else {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / cols;
      const int c = (int)(i - r * cols);
      __half h, l;
      f16_hi_lo(x[r * ldx + c], s1, s2, h, l);
      hi[r * ldo + c] = h;
      lo[r * ldo + c] = l;
    }
  }
}

// ---- segmented (flat arena): one warp per 256-float block, seg_id[block] selects the tensor ---------------------
__global__ void __launch_bounds__(256) amax_seg_kernel(const float* __restrict__ x, unsigned long long nblk,
                                                       const int32_t* __restrict__ seg_id, float* __restrict__ amax) {
  const int lane = threadIdx.x & 31;
  for (unsigned long long b = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5); b < nblk;
       b += (unsigned long long)gridDim.x * 8) {
    const float4* p = reinterpret_cast<const float4*>(x + b * 256);
    const float4 u = p[lane], v = p[lane + 32];
    float m = fmaxf(fmaxf(fmaxf(fabsf(u.x), fabsf(u.y)), fmaxf(fabsf(u.z), fabsf(u.w))),
                    fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    m = warp_max(m);
    if (lane == 0) atomic_max_pos(amax + (seg_id ? seg_id[b] : 0), m);
  }
}

__global__ void __launch_bounds__(256) f16_split_seg_kernel(const float* __restrict__ x, unsigned long long nblk,
                                                            const int32_t* __restrict__ seg_id,
                                                            const float* __restrict__ amax, __half* __restrict__ hi,
                                                            __half* __restrict__ lo, int32_t* __restrict__ exps) {
  const int lane = threadIdx.x & 31;
  for (unsigned long long b = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5); b < nblk;
       b += (unsigned long long)gridDim.x * 8) {
    const int seg = seg_id ? seg_id[b] : 0;
    const int e = f16_scale_exp(__ldg(amax + seg));
    float s1, s2;
    f16_scale_factors(e, s1, s2);
    if (lane == 0) exps[seg] = e;                 // every block of a segment writes the same value
    const float4* p = reinterpret_cast<const float4*>(x + b * 256);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 v = p[lane + 32 * q];
      __half h[4], l[4];
      f16_hi_lo(v.x, s1, s2, h[0], l[0]);
      f16_hi_lo(v.y, s1, s2, h[1], l[1]);
      f16_hi_lo(v.z, s1, s2, h[2], l[2]);
      f16_hi_lo(v.w, s1, s2, h[3], l[3]);
      const size_t o = (size_t)b * 256 + (size_t)(lane + 32 * q) * 4;
      *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(l);
    }
  }
}

static int grid_for(long long items, int per_block) {
  long long g = (items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int amax_launch(const float* x, int ldx, int rows, int cols, float* amax, cudaStream_t st) {
  const int vec = (cols % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const long long items = (long long)rows * (vec ? cols / 4 : cols);
  amax2d_kernel<<<grid_for(items, 256 * 4), 256, 0, st>>>(x, ldx, rows, cols, amax, vec);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int f16_split_launch(const float* x, int ldx, int rows, int cols, uint16_t* hi, uint16_t* lo, int ldo, float* amax,
                     int32_t* exp_out, int amax_ready, cudaStream_t st) {
  if (!amax_ready) {
    IPAVSR_CUDA(cudaMemsetAsync(amax, 0, sizeof(float), st));
    int rc = amax_launch(x, ldx, rows, cols, amax, st);
    if (rc) return rc;
  }
  const int vec = (cols % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                  (((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0);
  const long long items = (long long)rows * (vec ? cols / 4 : cols);
  f16_split2d_kernel<<<grid_for(items, 256 * 2), 256, 0, st>>>(x, ldx, rows, cols, reinterpret_cast<__half*>(hi),
                                                               reinterpret_cast<__half*>(lo), ldo, amax, exp_out, vec);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // namespace ipavsr

using namespace ipavsr;

extern "C" {

int ipavsr_amax(const float* x, int ldx, int64_t rows, int cols, float* amax, void* stream) {
  IPAVSR_CHECK_ARG(x && amax && rows >= 0 && cols >= 1 && ldx >= cols && rows < (1ll << 31), "bad arguments");
  if (rows == 0) return IPAVSR_OK;
  return amax_launch(x, ldx, (int)rows, cols, amax, reinterpret_cast<cudaStream_t>(stream));
}

int ipavsr_f16_split(const float* x, int ldx, int64_t rows, int cols, uint16_t* hi, uint16_t* lo, int ldo, float* amax,
                     int32_t* exp_out, int amax_ready, void* stream) {
  IPAVSR_CHECK_ARG(x && hi && lo && amax && exp_out, "null pointer");
  IPAVSR_CHECK_ARG(rows >= 0 && rows < (1ll << 31) && cols >= 1 && ldx >= cols && ldo >= cols, "bad sizes");
  if (rows == 0) return IPAVSR_OK;
  return f16_split_launch(x, ldx, (int)rows, cols, hi, lo, ldo, amax, exp_out, amax_ready,
                          reinterpret_cast<cudaStream_t>(stream));
}

int ipavsr_f16_split_segments(const float* x, uint16_t* hi, uint16_t* lo, uint64_t n, const int32_t* seg_id, int nseg,
                              float* amax, int32_t* exps, void* stream) {
  IPAVSR_CHECK_ARG(x && hi && lo && amax && exps && nseg >= 1, "bad arguments");
  IPAVSR_CHECK_ARG(n % 256 == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                       (((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0),
                   "the arena must be a multiple of 256 floats and 16-byte aligned");
  if (n == 0) return IPAVSR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned long long nblk = n / 256;
  IPAVSR_CUDA(cudaMemsetAsync(amax, 0, sizeof(float) * nseg, st));
  const int grid = grid_for((long long)nblk, 8);
  amax_seg_kernel<<<grid, 256, 0, st>>>(x, nblk, seg_id, amax);
  IPAVSR_LAUNCH_CHECK();
  f16_split_seg_kernel<<<grid, 256, 0, st>>>(x, nblk, seg_id, amax, reinterpret_cast<__half*>(hi),
                                             reinterpret_cast<__half*>(lo), exps);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

}  // extern "C"
