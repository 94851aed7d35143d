// fast_math.cuh -- the pair term of the FAST kernels (REBCU_MODE_FAST): FMA + reciprocal square root.
//
// The reference evaluates  prefact = -G/(r*r*r)*m_j  with r = sqrt(r2 + eps^2)  (src/gravity.c:222-230,
// src/tree.c:291-292): one sqrt and one divide, which cost 19 FP64-pipe instructions on this part when
// correctly rounded (strict_math.cuh).  FAST mode computes y = (r2 + eps^2)^(-1/2) directly:
//   seed   rsqrt.approx.ftz.f64 (MUFU.RSQ64H, SFU pipe, relative error ~2^-22)
//   step   one third-order correction  y = y0*(1 + e/2 + 3e^2/8),  e = 1 - a*y0^2   (5 FP64 instructions)
// which leaves a relative error of ~5/16 e^3 < 2^-60 plus the rounding of five operations (a few ulp) -- well
// inside the 1e-12 relative tolerance BASELINE.json states for the direct sum.  CUDA's rsqrt() costs about twice
// as many FP64 instructions for its last-ulp guarantee and carries a range branch.
//   pair term: 3 sub + 3 fma (r2) + 5 (rsqrt) + 3 mul (m*y^3) + 3 fma (accumulate) = 17 FP64 instructions.
// The factor -G is applied once per particle at the end, not per pair.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ double fast_rsqrt(double a) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    const double t = a * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    return fma(ye, p, y0);
}

// m / (r2)^(3/2) for r2 > 0 (r2 already holds the softening).
__device__ __forceinline__ double fast_m_over_r3(double r2, double m) {
    const double y = fast_rsqrt(r2);
    return (m * y) * (y * y);
}
