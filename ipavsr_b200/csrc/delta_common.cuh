// Shared pieces of the DeltaLayer kernels (delta.cu, delta_stream.cu).
#pragma once
#include "common.cuh"

namespace ipavsr {

template <bool EXACT>
__device__ __forceinline__ float delta_step(float acc, float diff, int th) {
  if (EXACT) {
    // The reference evaluates, in float64:  term = (theta*diff) / (2*theta*theta)  [= RN53(diff / (2 theta)), the
    // numerator is exact],  s = RN53(acc + term),  acc = RN24(s)   (utils/signal.py:19-21).  Exact float32 ties of s
    // are common (whenever 2*theta divides diff's mantissa), so the quotient has to be the correctly rounded one:
    // for a power-of-two theta the product diff * (1/(2 theta)) is exact; otherwise one Markstein correction step
    // (q = d*r; rem = fma(-q, c, d) exact; q' = fma(rem, r, q)) gives the correctly rounded quotient for these small
    // integer divisors c.
    // Power-of-two theta: the term diff / (2 theta) is exact in float32 and float32(float64(acc) + term) equals the single
    // rounding of a float32 FMA (a float64 sum of two float32 values only rounds when they are > 2^29 apart, where no
    // float32 rounding boundary is near), so those steps need no float64 arithmetic at all.
    if ((th & (th - 1)) == 0) return fmaf(diff, 1.0f / (2.0f * (float)th), acc);
    const double c = 2.0 * (double)th;
    const double r = 1.0 / c;
    const double d = (double)diff;
    double q = d * r;
    const double rem = fma(-q, c, d);
    q = fma(rem, r, q);
    return (float)((double)acc + q);
  } else {
    return fmaf(diff, 1.0f / (2.0f * (float)th), acc);
  }
}

}  // namespace ipavsr
