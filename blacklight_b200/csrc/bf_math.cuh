// Branch-free exp and log for the radiation kernels' per-frequency coefficient code.
//
// CUDA's exp()/log() are accurate but each carries a range-check branch, so a run of independent calls (the
// kappa-distribution fits evaluate ~20 exponentials and ~8 logarithms per frequency, reference
// simulation_coefficients.cpp:608-698) compiles to a chain of small basic blocks that the scheduler cannot
// interleave; with the few resident warps these register-heavy kernels have, the dependent-FMA latency of
// each polynomial is then fully exposed (ncu: stall_wait ~50%).  The versions below are straight-line code
// (clamps are selects, polynomials are evaluated by Estrin's scheme), accurate to < 2 ulp, NaN-propagating,
// and valid for:   exp_bf: any x (clamped to [-745.2, 709.7], i.e. 0 .. 1.6e308)
//                  log_bf: normal positive finite x (the call sites pass 1 + something non-negative)
// They are not used where bit-exactness matters (the geodesic integrator has its own libm restatements).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BF_HD __host__ __device__ __forceinline__
#else
#define BF_HD static inline
#endif

namespace bfm {

BF_HD double make_double(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  __builtin_memcpy(&d, &u, 8);
  return d;
#endif
}
BF_HD int hi_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t u;
  __builtin_memcpy(&u, &x, 8);
  return (int)(u >> 32);
#endif
}
BF_HD int lo_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t u;
  __builtin_memcpy(&u, &x, 8);
  return (int)(uint32_t)u;
#endif
}
// ~20-bit reciprocal seed (MUFU.RCP64H on the device)
BF_HD double rcp_seed(double d) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  return y;
#else
  return (double)(float)(1.0 / d);
#endif
}

BF_HD double exp_bf(double x) {
  // NaN-preserving clamp (both comparisons are false for NaN)
  double xc = x < -745.2 ? -745.2 : (x > 709.7 ? 709.7 : x);
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
  double t = fma(xc, 1.4426950408889634, magic);
  int k = lo_word(t);
  double kd = t - magic;
  double r = fma(kd, -6.93147180369123816490e-01, xc);
  r = fma(kd, -1.90821492927058770002e-10, r);
  // e^r, |r| <= ln2/2, Taylor to r^13 (truncation 4e-18), Estrin
  double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  double a0 = 1.0 + r;
  double a1 = fma(r, 1.0 / 6.0, 0.5);
  double a2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  double a3 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
  double a4 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
  double a5 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
  double a6 = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
  double b0 = fma(a1, r2, a0);
  double b1 = fma(a3, r2, a2);
  double b2 = fma(a5, r2, a4);
  double c0 = fma(b1, r4, b0);
  double c1 = fma(a6, r4, b2);
  double p = fma(c1, r8, c0);
  // 2^k in two normal factors so that results in the subnormal range round once
  int k1 = k >> 1, k2 = k - k1;
  double s1 = make_double((k1 + 1023) << 20, 0), s2 = make_double((k2 + 1023) << 20, 0);
  return (p * s1) * s2;
}

BF_HD double log_bf(double x) {
  int hi = hi_word(x);
  int k = (hi >> 20) - 1023;
  double m = make_double((hi & 0x000fffff) | 0x3ff00000, lo_word(x));  // [1, 2)
  bool big = m > 1.4142135623730951;
  m = big ? 0.5 * m : m;                                             // [sqrt(1/2), sqrt(2)]
  double kd = (double)(k + (big ? 1 : 0));
  // f = (m - 1) / (m + 1) by a Newton-refined reciprocal and one residual correction
  double d = m + 1.0, n = m - 1.0;
  double y = rcp_seed(d);
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  double f = n * y;
  f = fma(fma(-d, f, n), y, f);
  // log m = 2 atanh f = 2 f (1 + s/3 + s^2/5 + ... + s^10/21), s = f^2 <= 0.0295
  double s = f * f, s2 = s * s, s4 = s2 * s2, s8 = s4 * s4;
  double a0 = fma(s, 1.0 / 3.0, 1.0);
  double a1 = fma(s, 1.0 / 7.0, 1.0 / 5.0);
  double a2 = fma(s, 1.0 / 11.0, 1.0 / 9.0);
  double a3 = fma(s, 1.0 / 15.0, 1.0 / 13.0);
  double a4 = fma(s, 1.0 / 19.0, 1.0 / 17.0);
  double b0 = fma(a1, s2, a0);
  double b1 = fma(a3, s2, a2);
  double c0 = fma(b1, s4, b0);
  double c1 = fma(s2, 1.0 / 21.0, a4);
  double p = fma(c1, s8, c0);
  double res = fma(kd, 6.93147180369123816490e-01, fma(kd, 1.90821492927058770002e-10, 2.0 * f * p));
  return x != x ? x : res;
}

// (lo^-x + hi^-x)^(-1/x) from the logarithms a = ln lo, b = ln hi: the bridging form every kappa fit uses
// (simulation_coefficients.cpp:641-698).  Evaluated around the smaller of the two, so no intermediate
// overflows; lo = 0 or hi = 0 (logarithm -inf) gives 0 and NaN propagates, as in the reference.
BF_HD double bridge(double a, double b, double x, double inv_x) {
  double d = a == b ? 0.0 : a - b;
  double m = d < 0.0 ? a : b;
  return exp_bf(m - log_bf(1.0 + exp_bf(-x * fabs(d))) * inv_x);
}

}  // namespace bfm
