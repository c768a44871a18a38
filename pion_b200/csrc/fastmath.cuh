// pion_b200/csrc/fastmath.cuh -- branch-free FP64 reciprocal / square root.
//
// CUDA's `a / b` and `sqrt(x)` expand to a MUFU seed + Newton steps + a range check
// that CALLs a slow path (denormal / zero / huge operands; for `a / b` the check is on
// the NUMERATOR, so an exactly-zero numerator -- a static medium -- takes the slow path
// every time).  The dynamics update only divides by strictly positive, well-scaled
// quantities (densities, wave-speed differences) or guards the result with isfinite
// (HLLD_MHD.cpp:189-224), so the kernels use these sequences instead: the MUFU.RCP64H /
// MUFU.RSQ64H seed (>= 20 good bits) refined by one cubic step (relative error 2^-60 before the final rounding: <= 2 ulp), no branches, no calls.
// Zero / infinite / NaN operands propagate as inf / NaN exactly where the reference's
// IEEE division would produce a non-finite value (what the isfinite guards test).
// tools/micro/fp64_pipe.cu measures the error against IEEE division on the device.
#pragma once
#include <cuda_runtime.h>

namespace pion {

__device__ __forceinline__ double fast_rcp(double x) {
#ifdef PION_STRICT
  return 1.0 / x;
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);        // e + e^2: cubic convergence, 2^-20 -> 2^-60
  return fma(r, e, r);     // <= 1 ulp (measured: tools/micro/fp64_pipe.cu)
#endif
}

// 1/sqrt(x), x > 0
__device__ __forceinline__ double fast_rsqrt(double x) {
#ifdef PION_STRICT
  return 1.0 / sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // y <- y (1 + e/2 + 3 e^2/8), e = 1 - x y^2  (cubic)
  double e = fma(-x * y, y, 1.0);
  double t = fma(0.375, e, 0.5);
  return fma(y * e, t, y);
#endif
}

// sqrt(x), x >= 0 (zero, subnormal and negative x return 0)
__device__ __forceinline__ double fast_sqrt(double x) {
#ifdef PION_STRICT
  return sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);
  double t = fma(0.375, e, 0.5);
  y = fma(y * e, t, y);
  double g = x * y;                       // sqrt(x) to <= 2 ulp
  // the MUFU seed flushes subnormal inputs to zero (seed = inf -> NaN): treat x < DBL_MIN as 0, an
  // absolute error of at most 1.5e-154
  return (x >= 2.2250738585072014e-308) ? g : 0.0;
#endif
}

// sqrt(x) for x known to be >= DBL_MIN (no sub-normal guard: one DSETP + two FSEL fewer).  The wave-speed
// square roots qualify: their arguments are clamped from below (cfast2_ir: pmax(..., MACHINEACCURACY)) or are
// a sum with such a square root.  A negative or NaN argument gives NaN, as the reference's sqrt() does.
__device__ __forceinline__ double fast_sqrt_pos(double x) {
#if defined(PION_STRICT)
  return sqrt(x);
#elif defined(PION_GUARDED_SQRT)
  return fast_sqrt(x);
#else
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);
  double t = fma(0.375, e, 0.5);
  y = fma(y * e, t, y);
  return x * y;
#endif
}

// max / min of two doubles.  CUDA's fmax/fmin carry IEEE NaN handling that costs 7 SASS instructions
// (DSETP.MAX + 2 MOV + FSEL + SEL + LOP3) against 3 for the compare-and-select form; these return the
// SECOND argument when either operand is NaN, so call sites put the safe value second.
__device__ __forceinline__ double pmax(double a, double b) {
#ifdef PION_STRICT
  return fmax(a, b);
#else
  return (a > b) ? a : b;
#endif
}
__device__ __forceinline__ double pmin(double a, double b) {
#ifdef PION_STRICT
  return fmin(a, b);
#else
  return (a < b) ? a : b;
#endif
}

// a / b and sqrt(x) as the kernels use them: the branch-free sequences above unless PION_STRICT
__device__ __forceinline__ double pdiv(double a, double b) {
#ifdef PION_STRICT
  return a / b;
#else
  return a * fast_rcp(b);
#endif
}
__device__ __forceinline__ double psqrt(double x) { return fast_sqrt(x); }

}  // namespace pion
