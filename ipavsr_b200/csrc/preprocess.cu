// utils/preprocessing.py on the device: sample-wise / feature-wise / sequence-wise
// normalisation, diff images and the feacalc-style FIR deltas.  All streaming, HBM-bound kernels: threads walk the
// contiguous feature axis (coalesced 128-byte rows), frames/utterances are spread over the grid.
#include "common.cuh"

namespace ipavsr {

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- a10 normalize_input (utils/preprocessing.py:218-242): one warp per frame, frame held in registers ----
template <int NV>   // NV float4 per lane  -> D <= 128*NV
__global__ void __launch_bounds__(256) norm_samplewise_kernel(const float* __restrict__ x, int ldx,
                                                              float* __restrict__ y, int ldy, int64_t frames, int D) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int D4 = D >> 2;
  const float invD = 1.0f / (float)D;
  for (int64_t r = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < frames;
       r += (int64_t)gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * ldx);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + 32 * i;
      v[i] = c < D4 ? __ldcs(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * invD;
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + 32 * i;
      if (c < D4) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        s1 += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    // np.std(item) on the centred item re-centres by its (tiny) residual mean
    const float m2 = warp_sum(s1) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + 32 * i;
      if (c < D4) {
        float a = v[i].x - m2, b = v[i].y - m2, cc = v[i].z - m2, d = v[i].w - m2;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float sd = sqrtf(warp_sum(q) * invD);
    float4* yr = reinterpret_cast<float4*>(y + r * ldy);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane + 32 * i;
      if (c < D4) __stcs(yr + c, make_float4(v[i].x / sd, v[i].y / sd, v[i].z / sd, v[i].w / sd));
    }
  }
}

// generic shape (any D, any alignment): one warp per frame, three passes over the (L1/L2-resident) row
__global__ void __launch_bounds__(256) norm_samplewise_generic_kernel(const float* __restrict__ x, int ldx,
                                                                      float* __restrict__ y, int ldy, int64_t frames,
                                                                      int D) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float invD = 1.0f / (float)D;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < frames; r += (int64_t)gridDim.x * wpb) {
    const float* xr = x + r * ldx;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mean = warp_sum(s) * invD;
    float s1 = 0.f;
    for (int c = lane; c < D; c += 32) s1 += xr[c] - mean;
    const float m2 = warp_sum(s1) * invD;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) {
      float a = (xr[c] - mean) - m2;
      q += a * a;
    }
    const float sd = sqrtf(warp_sum(q) * invD);
    for (int c = lane; c < D; c += 32) y[r * ldy + c] = (xr[c] - mean) / sd;
  }
}

// ---- a11 featurewise_normalize_sequence (:245-257) ----
constexpr int FW_ROWS = 256;
// pass = 0: scratch[c] += sum x ; pass = 1: scratch[F+c] += sum (x-mean), scratch[2F+c] += sum (x-mean)^2
__global__ void __launch_bounds__(256) featurewise_sum_kernel(const float* __restrict__ x, int ldx,
                                                              double* __restrict__ scratch, int64_t frames, int F,
                                                              int pass) {
  __shared__ double r1[8][33], r2[8][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int64_t rb = (int64_t)blockIdx.y * FW_ROWS;
  const int64_t re = rb + FW_ROWS < frames ? rb + FW_ROWS : frames;
  double a = 0.0, b = 0.0;
  if (c < F) {
    float mean = 0.f;
    if (pass == 1) mean = (float)(scratch[c] / (double)frames);
    for (int64_t r = rb + rl; r < re; r += 8) {
      float v = x[r * ldx + c];
      if (pass == 0) a += (double)v;
      else {
        float cen = v - mean;
        a += (double)cen;
        b += (double)cen * (double)cen;
      }
    }
  }
  r1[rl][lane] = a;
  r2[rl][lane] = b;
  __syncthreads();
  if (rl == 0 && c < F) {
    double sa = 0, sb = 0;
    for (int i = 0; i < 8; ++i) { sa += r1[i][lane]; sb += r2[i][lane]; }
    if (pass == 0) atomicAdd(scratch + c, sa);
    else {
      atomicAdd(scratch + F + c, sa);
      atomicAdd(scratch + 2 * F + c, sb);
    }
  }
}

__global__ void featurewise_finalize_kernel(const double* __restrict__ scratch, float* __restrict__ mean,
                                            float* __restrict__ std, int64_t frames, int F) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  double n = (double)frames;
  mean[c] = (float)(scratch[c] / n);
  double m1 = scratch[F + c] / n;
  double var = scratch[2 * F + c] / n - m1 * m1;
  std[c] = (float)sqrt(var > 0 ? var : 0.0);
}

// grid (column chunks of 128 floats as float4 lanes, row groups); threads walk rows, lanes walk features
__global__ void __launch_bounds__(256) featurewise_apply_kernel(const float* __restrict__ x, int ldx,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ std, float* __restrict__ y,
                                                                int ldy, int64_t frames, int F, int rows_per_block) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  if (c >= F) return;
  const float m = mean[c], sd = std[c];
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < frames ? r0 + rows_per_block : frames;
  for (int64_t r = r0 + rl; r < r1; r += 8) __stcs(y + r * ldy + c, (__ldcs(x + r * ldx + c) - m) / sd);
}

// vectorised form: a thread owns 4 consecutive features (mean/std in registers) of every row of its row stripe, so a
// warp streams 512 contiguous bytes per row (F % 4 == 0, 16-byte aligned rows)
__global__ void __launch_bounds__(256) featurewise_apply_vec_kernel(const float* __restrict__ x, int ldx,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ std, float* __restrict__ y,
                                                                    int ldy, int64_t frames, int F4) {
  const int64_t total = frames * F4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F4;
    const int c = (int)(i - r * F4) << 2;
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x + r * ldx + c));
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 s = __ldg(reinterpret_cast<const float4*>(std + c));
    __stcs(reinterpret_cast<float4*>(y + r * ldy + c),
           make_float4((v.x - m.x) / s.x, (v.y - m.y) / s.y, (v.z - m.z) / s.z, (v.w - m.w) / s.w));
  }
}

// compute_diff_images, vectorised: block = (utterance, row stripe); every thread produces float4s y[r] = x[r] - x[r-1]
// (the second read of a row hits L1/L2), y[first] = x[first+1] - x[first] (the reference duplicates the first difference)
__global__ void __launch_bounds__(256) diff_image_vec_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                             int ldy, const int64_t* __restrict__ offsets, int D4) {
  const int u = blockIdx.y;
  const int64_t b = offsets[u], e = offsets[u + 1];
  const int64_t len = e - b;
  if (len <= 0) return;
  const int64_t total = len * D4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / D4;
    const int c = (int)(i - t * D4) << 2;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (len > 1) {
      const int64_t r = b + (t == 0 ? 1 : t);
      const float4 cur = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
      const float4 prv = __ldg(reinterpret_cast<const float4*>(x + (r - 1) * ldx + c));
      o = make_float4(cur.x - prv.x, cur.y - prv.y, cur.z - prv.z, cur.w - prv.w);
    }
    __stcs(reinterpret_cast<float4*>(y + (b + t) * ldy + c), o);
  }
}

// ---- a12 sequencewise_mean_image_subtraction (:260-277) and a14 compute_diff_images (:506-517) ----
// grid (column chunks of 128, utterances); a thread owns one pixel column of one utterance and walks its frames in
// order, which reproduces numpy's sequential float32 axis-0 summation exactly.
__global__ void __launch_bounds__(128) seq_mean_sub_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                           int ldy, const int64_t* __restrict__ offsets, int D) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int u = blockIdx.y;
  if (c >= D) return;
  const int64_t b = offsets[u], e = offsets[u + 1];
  if (e <= b) return;
  float s = 0.f;
  for (int64_t r = b; r < e; ++r) s += x[r * ldx + c];
  const float mean = s / (float)(e - b);
  for (int64_t r = b; r < e; ++r) __stcs(y + r * ldy + c, x[r * ldx + c] - mean);
}

__global__ void __launch_bounds__(128) diff_image_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                         int ldy, const int64_t* __restrict__ offsets, int D) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int u = blockIdx.y;
  if (c >= D) return;
  const int64_t b = offsets[u], e = offsets[u + 1];
  if (e <= b) return;
  if (e - b == 1) {   // the reference raises IndexError for a 1-frame utterance; write a zero frame
    y[b * ldy + c] = 0.f;
    return;
  }
  float prev = __ldcs(x + b * ldx + c);
  int64_t r = b + 1;
  // four independent row loads in flight per thread
  for (; r + 3 < e; r += 4) {
    float c0 = __ldcs(x + r * ldx + c), c1 = __ldcs(x + (r + 1) * ldx + c), c2 = __ldcs(x + (r + 2) * ldx + c),
          c3 = __ldcs(x + (r + 3) * ldx + c);
    float d0 = c0 - prev;
    if (r == b + 1) __stcs(y + b * ldy + c, d0);   // frame 0 duplicates the first difference (:514-515)
    __stcs(y + r * ldy + c, d0);
    __stcs(y + (r + 1) * ldy + c, c1 - c0);
    __stcs(y + (r + 2) * ldy + c, c2 - c1);
    __stcs(y + (r + 3) * ldy + c, c3 - c2);
    prev = c3;
  }
  for (; r < e; ++r) {
    float cur = __ldcs(x + r * ldx + c);
    float d = cur - prev;
    if (r == b + 1) __stcs(y + b * ldy + c, d);
    __stcs(y + r * ldy + c, d);
    prev = cur;
  }
}

// ---- a13 deltas (:17-51) + concat_first_second_deltas (:465-489): float64 [x | d1 | d2] ----
// d[t] = sum_{j=-h..h} j * X(t+j),  X(i) = x[1] for i<0 (the reference's left-pad quirk, :43), x[len-1] for i>=len.
// The kernel is bound by the FP64 pipe (two 9-tap float64 passes per output = 18 DFMA per element; measured on the B200 of
// this pool: ~1.6 T DFMA/s whatever the staging — block per utterance through shared memory 0.355 ms per 1 M frames x 30
// columns, one warp per utterance 0.61 ms, one thread per column with everything in registers 0.375 ms), not by HBM.
// OUT = double: the reference's array; OUT = float: the same values rounded once, i.e. `.astype('float32')` of it, which
// is what every runner feeds the network (avletters/bimodal.py:351 `dct_data['dctFeatures'].astype('float32')`).
template <typename OUT>
__global__ void __launch_bounds__(256) deltas_fir_kernel(const float* __restrict__ x, int ldx, OUT* __restrict__ y,
                                                         int ldy, const int64_t* __restrict__ offsets, int F, int h,
                                                         int max_len) {
  extern __shared__ __align__(16) double dsm[];
  double* s0 = dsm;                             // [max_len][32]
  double* s1 = dsm + (size_t)max_len * 32;      // [max_len][32]
  const int u = blockIdx.y;
  const int c0 = blockIdx.x * 32;
  const int64_t b = offsets[u];
  const int len = (int)(offsets[u + 1] - b);
  if (len <= 0) return;
  const int fc = min(32, F - c0);
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  for (int t = rl; t < len; t += 8)
    if (lane < fc) {
      double v = (double)x[(b + t) * ldx + c0 + lane];
      s0[t * 32 + lane] = v;
      y[(b + t) * (int64_t)ldy + c0 + lane] = (OUT)v;
    }
  __syncthreads();
  const int left = len > 1 ? 1 : 0;
  for (int pass = 0; pass < 2; ++pass) {
    const double* src = pass == 0 ? s0 : s1;
    double* dst = pass == 0 ? s1 : s0;
    for (int t = rl; t < len; t += 8)
      if (lane < fc) {
        double acc = 0.0;
        for (int j = -h; j <= h; ++j) {
          int i = t + j;
          i = i < 0 ? left : (i >= len ? len - 1 : i);
          acc += (double)j * src[i * 32 + lane];
        }
        dst[t * 32 + lane] = acc;
        y[(b + t) * (int64_t)ldy + (pass + 1) * F + c0 + lane] = (OUT)acc;
      }
    __syncthreads();
  }
}

static inline int grid_cap(int64_t want, int per_sm) {
  int64_t cap = (int64_t)sm_count() * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace ipavsr

using namespace ipavsr;

template <typename OUT>
static int deltas_fir_launch(const float* x, int ldx, OUT* y, int ldy, const int64_t* offsets, int U, int F, int w,
                             int max_len, void* stream, const char* fn) {
  if (!(x && y && offsets && U >= 0 && F >= 1 && w >= 1 && max_len >= 1)) {
    set_error("%s: bad arguments", fn);
    return IPAVSR_ERR_ARG;
  }
  if (ldy < 3 * F) {
    set_error("%s: ldy too small", fn);
    return IPAVSR_ERR_ARG;
  }
  if (U == 0) return IPAVSR_OK;
  size_t smem = (size_t)2 * max_len * 32 * sizeof(double);
  if (smem > 200 * 1024) {
    set_error("%s: utterance too long for the shared-memory tile (max_len <= 400)", fn);
    return IPAVSR_ERR_ARG;
  }
  if (smem > 48 * 1024)
    IPAVSR_CUDA(cudaFuncSetAttribute(deltas_fir_kernel<OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int u0 = 0; u0 < U; u0 += 65535) {          // gridDim.y limit
    const int n = U - u0 < 65535 ? U - u0 : 65535;
    dim3 grid((F + 31) / 32, n);
    deltas_fir_kernel<OUT><<<grid, 256, smem, S(stream)>>>(x, ldx, y, ldy, offsets + u0, F, w / 2, max_len);
    IPAVSR_LAUNCH_CHECK();
  }
  return IPAVSR_OK;
}

extern "C" {

int ipavsr_norm_samplewise(const float* x, int ldx, float* y, int ldy, int64_t frames, int D, void* stream) {
  IPAVSR_CHECK_ARG(x && y && frames >= 0 && D >= 1 && ldx >= D && ldy >= D, "bad arguments");
  if (frames == 0) return IPAVSR_OK;
  const bool vec = (D % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(x) && aligned16(y) && D <= 2048;
  const int grid = grid_cap((frames + 7) / 8, 8);
  cudaStream_t st = S(stream);
  if (!vec) norm_samplewise_generic_kernel<<<grid, 256, 0, st>>>(x, ldx, y, ldy, frames, D);
  else if (D <= 512) norm_samplewise_kernel<4><<<grid, 256, 0, st>>>(x, ldx, y, ldy, frames, D);
  else if (D <= 1280) norm_samplewise_kernel<10><<<grid, 256, 0, st>>>(x, ldx, y, ldy, frames, D);
  else norm_samplewise_kernel<16><<<grid, 256, 0, st>>>(x, ldx, y, ldy, frames, D);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_norm_featurewise_stats(const float* x, int ldx, float* mean, float* std, double* scratch, int64_t frames,
                                  int F, void* stream) {
  IPAVSR_CHECK_ARG(x && mean && std && scratch && frames >= 1 && F >= 1 && ldx >= F, "bad arguments");
  cudaStream_t st = S(stream);
  IPAVSR_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * F, st));
  dim3 grid((F + 31) / 32, (unsigned)((frames + FW_ROWS - 1) / FW_ROWS));
  featurewise_sum_kernel<<<grid, 256, 0, st>>>(x, ldx, scratch, frames, F, 0);
  IPAVSR_LAUNCH_CHECK();
  featurewise_sum_kernel<<<grid, 256, 0, st>>>(x, ldx, scratch, frames, F, 1);
  IPAVSR_LAUNCH_CHECK();
  featurewise_finalize_kernel<<<(F + 127) / 128, 128, 0, st>>>(scratch, mean, std, frames, F);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_norm_featurewise_apply(const float* x, int ldx, const float* mean, const float* std, float* y, int ldy,
                                  int64_t frames, int F, void* stream) {
  IPAVSR_CHECK_ARG(x && mean && std && y && frames >= 0 && F >= 1, "bad arguments");
  if (frames == 0) return IPAVSR_OK;
  if (F % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(mean) && aligned16(std)) {
    const int64_t items = frames * (F / 4);
    featurewise_apply_vec_kernel<<<grid_cap((items + 255) / 256, 16), 256, 0, S(stream)>>>(x, ldx, mean, std, y, ldy,
                                                                                           frames, F / 4);
    IPAVSR_LAUNCH_CHECK();
    return IPAVSR_OK;
  }
  int rows_per_block = 64;
  while ((frames + rows_per_block - 1) / rows_per_block > 60000) rows_per_block *= 2;
  dim3 grid((F + 31) / 32, (unsigned)((frames + rows_per_block - 1) / rows_per_block));
  featurewise_apply_kernel<<<grid, 256, 0, S(stream)>>>(x, ldx, mean, std, y, ldy, frames, F, rows_per_block);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_seq_mean_sub(const float* x, int ldx, float* y, int ldy, const int64_t* offsets, int U, int D,
                        void* stream) {
  IPAVSR_CHECK_ARG(x && y && offsets && U >= 0 && D >= 1, "bad arguments");
  if (U == 0) return IPAVSR_OK;
  IPAVSR_CHECK_ARG(U <= 65535, "at most 65535 utterances per call");
  dim3 grid((D + 127) / 128, U);
  seq_mean_sub_kernel<<<grid, 128, 0, S(stream)>>>(x, ldx, y, ldy, offsets, D);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_diff_image(const float* x, int ldx, float* y, int ldy, const int64_t* offsets, int U, int D, void* stream) {
  IPAVSR_CHECK_ARG(x && y && offsets && U >= 0 && D >= 1, "bad arguments");
  if (U == 0) return IPAVSR_OK;
  IPAVSR_CHECK_ARG(U <= 65535, "at most 65535 utterances per call");
  if (D % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && aligned16(x) && aligned16(y)) {
    // ~4 blocks per utterance of typical length (40 frames x 300 float4) keep every SM busy with coalesced row streams
    dim3 grid(4, U);
    diff_image_vec_kernel<<<grid, 256, 0, S(stream)>>>(x, ldx, y, ldy, offsets, D / 4);
    IPAVSR_LAUNCH_CHECK();
    return IPAVSR_OK;
  }
  dim3 grid((D + 127) / 128, U);
  diff_image_kernel<<<grid, 128, 0, S(stream)>>>(x, ldx, y, ldy, offsets, D);
  IPAVSR_LAUNCH_CHECK();
  return IPAVSR_OK;
}

int ipavsr_deltas_fir(const float* x, int ldx, double* y, int ldy, const int64_t* offsets, int U, int F, int w,
                      int max_len, void* stream) {
  return deltas_fir_launch<double>(x, ldx, y, ldy, offsets, U, F, w, max_len, stream, "ipavsr_deltas_fir");
}

int ipavsr_deltas_fir_f32(const float* x, int ldx, float* y, int ldy, const int64_t* offsets, int U, int F, int w,
                          int max_len, void* stream) {
  return deltas_fir_launch<float>(x, ldx, y, ldy, offsets, U, F, w, max_len, stream, "ipavsr_deltas_fir_f32");
}

}  // extern "C"
