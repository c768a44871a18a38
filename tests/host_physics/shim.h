// tests/host_physics/shim.h -- TEST INFRASTRUCTURE: lets g++ compile pion_b200/csrc/physics.cuh + fastmath.cuh for the host,
// so that the DEVICE arithmetic (one-sided HLLD, reciprocal-multiply forms, Newton-refined MUFU seeds) can be compared with the
// oracle interface by interface on a machine without a GPU (tests/test_device_physics_on_host.py).  The MUFU.RCP64H / RSQ64H
// seeds are emulated: ~20 good bits, sub-normal inputs flushed to zero, 1/0 = inf -- what matters for the non-finite cases.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline double pion_trunc20(double v) {
  if (!std::isfinite(v) || v == 0.0) return v;
  int64_t b; std::memcpy(&b, &v, 8); b &= ~((int64_t(1) << 32) - 1); std::memcpy(&v, &b, 8); return v;
}
static inline double pion_rcp_approx(double x) { if (std::fabs(x) < DBL_MIN) x = std::copysign(0.0, x); return pion_trunc20(1.0 / x); }
static inline double pion_rsqrt_approx(double x) { if (std::fabs(x) < DBL_MIN) x = std::copysign(0.0, x); return pion_trunc20(1.0 / std::sqrt(x)); }
using std::fabs; using std::fma; using std::fmax; using std::fmin; using std::isfinite; using std::sqrt;
