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

#include "bf_tables.h"

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

// Tables (tools/bf_tables.py): on the device they sit in global memory behind the read-only path -- the indices of a
// warp diverge, which the constant cache would serialise.
#if defined(__CUDACC__)
static __device__ const double bf_exp_table_dev[64] = BF_EXP_TABLE;
static __device__ const double bf_log_table_dev[256] = BF_LOG_TABLE;
#endif
static const double bf_exp_table_host[64] = BF_EXP_TABLE;
static const double bf_log_table_host[256] = BF_LOG_TABLE;

BF_HD double exp_table(int j) {
#if defined(__CUDA_ARCH__)
  return __ldg(bf_exp_table_dev + j);
#else
  return bf_exp_table_host[j];
#endif
}
BF_HD double log_table(int j) {
#if defined(__CUDA_ARCH__)
  return __ldg(bf_log_table_dev + j);
#else
  return bf_log_table_host[j];
#endif
}

// e^x = 2^k 2^(j/64) e^r with x = (64 k + j) ln2 / 64 + r, |r| <= ln2 / 128: a 64-entry table and a degree-6 polynomial
// (truncation r^7 / 5040 = 3e-20) instead of a degree-13 one -- 15 FP64 instructions.
BF_HD double exp_bf(double x) {
  // NaN-preserving clamp (both comparisons are false for NaN)
  double xc = x < -745.2 ? -745.2 : (x > 709.7 ? 709.7 : x);
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low word
  double t = fma(xc, 92.332482616893657 /* 64 / ln 2 */, magic);
  int ki = lo_word(t);
  double kd = t - magic;
  // ln2 / 64 in two pieces; the high one has 32 significant bits, so kd * hi is exact
  double r = fma(kd, -6.93147180369123816490e-01 / 64.0, xc);
  r = fma(kd, -1.90821492927058770002e-10 / 64.0, r);
  double r2 = r * r;
  double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  double u = fma(r, 1.0 / 6.0, 0.5);
  q = fma(q, r, 1.0 / 24.0);
  double p = fma(fma(q, r2, u), r2, r);   // e^r - 1
  double tj = exp_table(ki & 63);
  double v = fma(tj, p, tj);
  // 2^k in two normal factors so that results in the subnormal range round once
  int k = ki >> 6;
  int k1 = k >> 1, k2 = k - k1;
  double s1 = make_double((k1 + 1023) << 20, 0), s2 = make_double((k2 + 1023) << 20, 0);
  return (v * s1) * s2;
}

// log x = k ln 2 + log c + log1p(r), x = 2^k z, z in [0.6875, 1.375), r = z / c - 1 for the centre c of z's interval (128
// intervals uniform in the bit pattern of z; table of 1 / c and log c), |r| <= 2^-7, degree-8 polynomial for log1p: 14
// FP64 instructions.  Arguments in [1, 1 + 2^-7) have c = 1, r = z - 1 exactly: log_bf(1 + y) keeps full relative accuracy
// for small y (the bridging forms need it); elsewhere the absolute error is ~1e-16 max(1, |log x|).
BF_HD double log_bf(double x) {
  int hi = hi_word(x);
  int tmp = hi - 0x3fe60000;
  int i = (tmp >> 13) & 127;
  int k = tmp >> 20;
  double z = make_double(hi - (tmp & (int)0xfff00000), lo_word(x));
  double invc = log_table(2 * i), logc = log_table(2 * i + 1);
  double r = fma(z, invc, -1.0);
  double kd = (double)k;
  double r2 = r * r, r4 = r2 * r2;
  // log1p(r) - r = r^2 (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7 - r^6/8)
  double a0 = fma(r, 1.0 / 3.0, -0.5);
  double a1 = fma(r, 1.0 / 5.0, -0.25);
  double a2 = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  double b0 = fma(a1, r2, a0);
  double b1 = fma(r2, -0.125, a2);
  double q = fma(b1, r4, b0);
  double head = fma(kd, 6.93147180369123816490e-01, logc);
  double tail = fma(kd, 1.90821492927058770002e-10, fma(q, r2, r));
  double res = head + tail;
  return x != x ? x : res;
}

// log_bf with log 0 = -inf (log_bf itself returns about -744 there: its callers in the frequency loops never pass zero)
BF_HD double log_bf_z(double x) {
  double res = log_bf(x);
  return x == 0.0 ? -1.0 / 0.0 : res;
}

// sinh and cosh of the same argument from two exponentials; below |x| = 0.25 the odd series keeps sinh's
// relative accuracy (selected, not branched).  < 3 ulp.
BF_HD void sinhcosh_bf(double x, double &sh, double &ch) {
  double e1 = exp_bf(x), e2 = exp_bf(-x);
  ch = 0.5 * (e1 + e2);
  double x2 = x * x;
  double series = x * fma(x2, fma(x2, fma(x2, fma(x2, fma(x2, 1.0 / 39916800.0, 1.0 / 362880.0), 1.0 / 5040.0), 1.0 / 120.0),
                                  1.0 / 6.0), 1.0);
  sh = fabs(x) < 0.25 ? series : 0.5 * (e1 - e2);
}

// Modified Bessel functions K_0(x), K_1(x), x > 0 (the reference calls std::cyl_bessel_k,
// simulation_coefficients.cpp:537-539).  Straight-line code, no divisions in loops:
//   x <= 2: ascending series (Abramowitz & Stegun 9.6.11, 9.6.13) as four polynomials of degree 13 in
//           q = x^2/4 (the terms beyond q^13/(13!)^2 = 3e-20 are dropped), coefficients exact to double;
//   x  > 2: K_nu(x) = e^-x x^-1/2 sum_j c_j T_j(4/x - 1), Chebyshev coefficients of degree 22 fitted to 40-digit
//           values (tools/bessel_tables.py), Clenshaw recurrence.
// Relative error < 5e-16 on [1e-3, 150] against 40-digit values (tests/test_cpu_host.py).
BF_HD void bessel_k01(double x, double &k0, double &k1) {
  if (x <= 2.0) {
    const double ci0[14] = {
      1.00000000000000000e+00, 1.00000000000000000e+00, 2.50000000000000000e-01, 2.77777777777777762e-02,
      1.73611111111111101e-03, 6.94444444444444444e-05, 1.92901234567901239e-06, 3.93675988914084175e-08,
      6.15118732678256523e-10, 7.59405842812662392e-12, 7.59405842812662337e-14, 6.27608134555919329e-16,
      4.35838982330499500e-18, 2.57892888952958276e-20};
    const double cs0[14] = {
      0.00000000000000000e+00, 1.00000000000000000e+00, 3.75000000000000000e-01, 5.09259259259259231e-02,
      3.61689814814814816e-03, 1.58564814814814804e-04, 4.72608024691358017e-06, 1.02074559982723252e-07,
      1.67180484131483275e-09, 2.14833502119502765e-11, 2.22427560547629389e-13, 1.89529958700615289e-15,
      1.35250018394848115e-17, 8.20133881368263703e-20};
    const double ci1[14] = {
      1.00000000000000000e+00, 5.00000000000000000e-01, 8.33333333333333287e-02, 6.94444444444444406e-03,
      3.47222222222222235e-04, 1.15740740740740735e-05, 2.75573192239858883e-07, 4.92094986142605219e-09,
      6.83465258531396137e-11, 7.59405842812662312e-13, 6.90368948011511223e-15, 5.23006778796599400e-17,
      3.35260755638845791e-19, 1.84209206394970202e-21};
    const double cs1[14] = {
      -1.54431329803065731e-01, 6.72784335098467134e-01, 1.81575166960855627e-01, 1.91821898393305622e-02,
      1.11535949196652807e-03, 4.14224768927114279e-05, 1.07154591409118091e-06, 2.04528600359387804e-08,
      3.00204874658918806e-10, 3.49592872969288208e-12, 3.30991473525027176e-14, 2.59864113210112861e-16,
      1.71952328269925653e-18, 9.72120751882361755e-21};
    const double euler = 5.77215664901532866e-01;
    double q = 0.25 * x * x, lg = log(0.5 * x);
    double i0 = ci0[13], s0 = cs0[13], i1 = ci1[13], s1 = cs1[13];
    for (int k = 12; k >= 0; k--) {
      i0 = fma(i0, q, ci0[k]);
      s0 = fma(s0, q, cs0[k]);
      i1 = fma(i1, q, ci1[k]);
      s1 = fma(s1, q, cs1[k]);
    }
    k0 = -(lg + euler) * i0 + s0;
    k1 = 1.0 / x + lg * (0.5 * x * i1) - 0.25 * x * s1;
    return;
  }
  const double c0[23] = {
      1.22015154103297774e+00, -3.14481013119645020e-02, 1.56988388573005332e-03, -1.28495495816278017e-04,
      1.39498137188765002e-05, -1.83175552271911953e-06, 2.76681363944501486e-07, -4.66048989768794783e-08,
      8.57403401741422362e-09, -1.69753450938905439e-09, 3.57739728140013962e-10, -7.95748924447235326e-11,
      1.85594911494131028e-11, -4.51459788300317332e-12, 1.14034058718387776e-12, -2.98009689462883721e-13,
      8.03288997119513968e-14, -2.22751103352657940e-14, 6.34001023017130056e-15, -1.84839946836037540e-15,
      5.50630089020145042e-16, -1.66089941374907953e-16, 4.68034840052582055e-17};
  const double c1[23] = {
      1.36031309524222133e+00, 1.03923736576817236e-01, -2.85781685962277921e-03, 1.95215518471351620e-04,
      -1.93619797416608301e-05, 2.40648494783721699e-06, -3.50196060308781256e-07, 5.74108412545004947e-08,
      -1.03457624656780935e-08, 2.01504975519702721e-09, -4.19035475934172483e-10, 9.21831518759994310e-11,
      -2.12996783841327002e-11, 5.13963967308577877e-12, -1.28917395985525712e-12, 3.34841963550466353e-13,
      -8.97670431956193934e-14, 2.47715195964445683e-14, -7.01976576208568517e-15, 2.03849397498716266e-15,
      -6.05082562535566479e-16, 1.81931492334142224e-16, -5.11378840098631785e-17};
  double inv_x = 1.0 / x;
  double u2 = 2.0 * (4.0 * inv_x - 1.0);
  double a1 = 0.0, a2 = 0.0, b1 = 0.0, b2 = 0.0;   // Clenshaw: b_k = 2u b_{k+1} - b_{k+2} + c_k
  for (int k = 22; k >= 1; k--) {
    double a0 = fma(u2, a1, c0[k] - a2), b0 = fma(u2, b1, c1[k] - b2);
    a2 = a1; a1 = a0;
    b2 = b1; b1 = b0;
  }
  double f0 = fma(0.5 * u2, a1, c0[0] - a2), f1 = fma(0.5 * u2, b1, c1[0] - b2);
  double scale = exp(-x) * sqrt(inv_x);
  k0 = f0 * scale;
  k1 = f1 * scale;
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
