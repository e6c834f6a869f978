// Device helpers shared by the persistent grid-wide QRCP kernel (qrcp.cu) and the one-CTA-per-block batched
// kernel (batched.cu): LAPACK's pivot ordering and the LAWN-176 constant.
#pragma once
#include <cstdint>

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// candidate ordering: larger norm wins; ties go to the smaller logical position
// (idamax returns the FIRST maximum).
__device__ __forceinline__ bool cand_better(double v, int lp, double bv, int blp) {
  return (v > bv) || (v == bv && lp < blp);
}

// argmax of (v, lp) over the warp; returns the winning lane
// Non-negative doubles order like their bit patterns, so the argmax runs on the integer pipe with
// four warp collectives (redux.sync) instead of 5 x 3 shuffles + FP64 compares.  v < 0 means "none".
__device__ __forceinline__ int warp_argmax(double v, int lp) {
  const int hi = (v >= 0.0) ? __double2hiint(v) : (int)0x80000000;
  const unsigned lo = (v >= 0.0) ? (unsigned)__double2loint(v) : 0u;
  const int mh = __reduce_max_sync(0xffffffffu, hi);
  const bool c1 = (hi == mh);
  const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
  const bool c2 = c1 && (lo == ml);
  const int mlp = __reduce_min_sync(0xffffffffu, c2 ? lp : 0x7fffffff);
  return __ffs(__ballot_sync(0xffffffffu, c2 && lp == mlp)) - 1;
}

constexpr double TOL3Z = 1.0536712127723509e-08;   // sqrt(2^-53) = sqrt(DLAMCH('Epsilon'))

