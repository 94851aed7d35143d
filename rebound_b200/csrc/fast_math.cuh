// fast_math.cuh -- the pair term of the FAST kernels (REBCU_MODE_FAST): FMA + reciprocal square root.
//
// The reference evaluates  prefact = -G/(r*r*r)*m_j  with r = sqrt(r2 + eps^2)  (src/gravity.c:222-230,
// src/tree.c:291-292): one sqrt and one divide, which cost 19 FP64-pipe instructions on this part when
// correctly rounded (strict_math.cuh).  FAST mode computes y = (r2 + eps^2)^(-1/2) directly:
//   seed   rsqrt.approx.ftz.f64 (MUFU.RSQ64H, SFU pipe; relative error measured at about 2^-19.5, see fast_m_over_r3_tree)
//   step   one third-order correction  y = y0*(1 + e/2 + 3e^2/8),  e = 1 - a*y0^2   (5 FP64 instructions)
// which leaves a relative error of ~5/16 e^3 < 1e-18 plus the rounding of five operations (a few ulp) -- well
// inside the 1e-12 relative tolerance BASELINE.json states for the direct sum.  CUDA's rsqrt() costs about twice
// as many FP64 instructions for its last-ulp guarantee and carries a range branch.
//   pair term with fast_rsqrt: 3 sub + 3 fma (r2) + 5 (rsqrt) + 3 mul (m*y^3) + 3 fma (accumulate) = 17 FP64 instructions;
//   with fast_m_over_r3 (below): 3 + 3 + 7 + 3 = 16.
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

// m / a^(3/2) for a > 0 (a = r2 already holds the softening).  The correction is applied to the cube, not to y:
//   m * a^(-3/2) = w * (1 - e)^(-3/2) = w * (1 + e*(3/2 + 15/8 e) + 35/16 e^3 + ...),   e = 1 - a*y0^2,  w = m*y0^3
// truncated after e^2 (remainder 35/16 e^3 < 1e-17): seven FP64 instructions instead of the eight of
// "refine y, then cube" and one rounding less, i.e. 16 per pair term.
__device__ __forceinline__ double fast_m_over_r3(double a, double m) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    const double v = y0 * y0;
    const double e = fma(-a, v, 1.0);
    const double u = m * y0;
    const double w = u * v;
    const double t = fma(1.875, e, 1.5);
    const double we = w * e;
    return fma(we, t, w);
}

// A cheaper variant for the tree's group walk (A/B kernels only, NOT shipped): the series truncated after the linear term,
//   w * (1 + 3e/2),  six FP64 instructions, i.e. 15 per pair term.
// Its error is 15/8 e^2.  Measured on a B200 (gpurun_out/k_tests_m.log): up to 4.5e-12 relative in the theta = 0 tests,
// i.e. the hardware seed is good to about 2^-19.5, not 2^-22 -- outside the 1e-12 those tests ask for, although far inside
// the tree's own error.  It buys 6 % of the walk (53.4 vs 56.7 ms at N = 2^24, profiles/r02_walk_group_ab.txt).
__device__ __forceinline__ double fast_m_over_r3_tree(double a, double m) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    const double v = y0 * y0;
    const double e = fma(-a, v, 1.0);
    const double u = m * y0;
    const double w = u * v;
    const double we = w * e;
    return fma(we, 1.5, w);
}
