// Cartesian Kerr-Schild geometry for the geodesic integrator, evaluated with the exact
// floating-point dataflow of the reference so that trajectories (and therefore termination flags,
// sample counts and sampled cell indices) are reproduced bit for bit.
//
// Mathematics (reference src/geodesic_integrator/geodesic_geometry.cpp:19-276):
//   r^2 = (R^2 - a^2 + hypot(R^2 - a^2, 2 a z)) / 2,  f = 2 r^3 / (r^4 + a^2 z^2)           (M = 1)
//   l = ((r x + a y)/(r^2+a^2), (r y - a x)/(r^2+a^2), z/r),  g_cov = eta + f l l, g^con = eta - f l l
// Instead of materialising 4x4 and 3x4x4 arrays, the Kerr-Schild structure is kept symbolic and only
// the products the reference actually forms are evaluated, in its association order:
//   P_ij = (f l_i) l_j                       -> g_ij = P_ij (+1), g^ij = -P_ij (+1), g_0i = g^0i = f l_i
//   d_a g^00 = -d_a f,  d_a g^0m = d_a g^m0 = (d_a f) l_m + f d_a l_m,
//   d_a g^mn = -(((d_a f) l_m) l_n + (f d_a l_m) l_n + (f l_m) d_a l_n)
// Multiplications by l^0 = -1 / +1 and additions of exact zeros are dropped (they are exact).
// This header must be compiled with -fmad=false: a fused multiply-add would change roundings.
#pragma once
#include "glibc_math.cuh"

namespace ksx {

struct KsPoint {
  double r, r2, f, l1, l2, l3;
};

// Kerr-Schild radius (geodesic_geometry.cpp:19-26)
BL_HD double radius(double a, double x, double y, double z) {
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double r2 = 0.5 * (rr2 - a2 + blmath::hypot_glibc(rr2 - a2, 2.0 * a * z));
  return blmath::sqrt_rn(r2);
}

BL_HD KsPoint ks_point(double a, double x, double y, double z) {
  KsPoint q;
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  q.r2 = 0.5 * (rr2 - a2 + blmath::hypot_glibc(rr2 - a2, 2.0 * a * z));
  q.r = blmath::sqrt_rn(q.r2);
  q.f = blmath::div_rn(2.0 * q.r2 * q.r, q.r2 * q.r2 + a2 * z * z);
  blmath::Recip ra = blmath::recip_of(q.r2 + a2);
  q.l1 = blmath::div_by(q.r * x + a * y, ra);
  q.l2 = blmath::div_by(q.r * y - a * x, ra);
  q.l3 = blmath::div_rn(z, q.r);
  return q;
}

// Solve the null condition g^{mu nu} p_mu p_nu = 0 for a rescaling of the spatial momentum
// (geodesics.cpp:296-309 and :352-371).  p = (p_0, p_1, p_2, p_3) covariant; p[1..3] are scaled.
template <bool flat>
BL_HD_SAMPLE void renormalize_momentum(double a, double x, double y, double z, double p[4]) {
  double g00, g0[3], gs[3][3];
  if (flat) {
    g00 = -1.0;
    for (int i = 0; i < 3; i++) {
      g0[i] = 0.0;
      for (int j = 0; j < 3; j++) gs[i][j] = i == j ? 1.0 : 0.0;
    }
  } else {
    KsPoint q = ks_point(a, x, y, z);
    double l[3] = {q.l1, q.l2, q.l3};
    g00 = -q.f - 1.0;
    for (int i = 0; i < 3; i++) {
      double fl = q.f * l[i];
      g0[i] = fl;
      for (int j = 0; j < 3; j++) gs[i][j] = i == j ? 1.0 - fl * l[j] : -(fl * l[j]);
    }
  }
  double qa = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) qa += gs[i][j] * p[1 + i] * p[1 + j];
  double qb = 0.0;
  for (int i = 0; i < 3; i++) qb += 2.0 * g0[i] * p[0] * p[1 + i];
  double qc = g00 * p[0] * p[0];
  double qd = blmath::sqrt_rn(qb * qb - 4.0 * qa * qc);
  double scale = qb < 0.0 ? blmath::div_rn(qd - qb, 2.0 * qa) : blmath::div_rn(-2.0 * qc, qb + qd);
  for (int i = 0; i < 3; i++) p[1 + i] *= scale;
}

// Hamiltonian right-hand side with proper distance (geodesics.cpp:867-893).
//   in : x,y,z and covariant momentum p[4]
//   out: dx[4] = g^{mu nu} p_nu, dp[3] = -1/2 d_i g^{mu nu} p_mu p_nu, ds = -sqrt(g_ij t^i t^j)
template <bool flat>
BL_HD void rhs(double a, double x, double y, double z, const double p[4], double dx[4],
               double dp[3], double &ds) {
  if (flat) {
    // Minkowski: g = eta, derivatives vanish; sums keep the reference's left-to-right order
    dx[0] = -1.0 * p[0];
    dx[1] = p[1];
    dx[2] = p[2];
    dx[3] = p[3];
    dp[0] = dp[1] = dp[2] = 0.0;
    ds = -blmath::sqrt_rn(p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    return;
  }
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double r2 = 0.5 * (rr2 - a2 + blmath::hypot_glibc(rr2 - a2, 2.0 * a * z));
  double r = blmath::sqrt_rn(r2);
  double r4 = r2 * r2;
  double a2zz = a2 * z * z;
  double f = blmath::div_rn(2.0 * r2 * r, r4 + a2zz);
  // shared denominators: quotients below are the IEEE quotients a / b (blmath::div_by)
  using blmath::div_by;
  const blmath::Recip ra = blmath::recip_of(r2 + a2), rr = blmath::recip_of(r);
  double l[3] = {div_by(r * x + a * y, ra), div_by(r * y - a * x, ra), div_by(z, rr)};
  double fl[3] = {f * l[0], f * l[1], f * l[2]};
  double P[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) P[i][j] = fl[i] * l[j];

  // dx^mu/dlambda = sum_nu g^{mu nu} p_nu (nu ascending)
  double gcon00 = -f - 1.0;
  const blmath::Recip g00 = blmath::recip_of(gcon00);
  dx[0] = gcon00 * p[0] + fl[0] * p[1] + fl[1] * p[2] + fl[2] * p[3];
  for (int i = 0; i < 3; i++) {
    double acc = fl[i] * p[0];
    for (int j = 0; j < 3; j++) {
      double g = i == j ? 1.0 - P[i][j] : -P[i][j];
      acc += g * p[1 + j];
    }
    dx[1 + i] = acc;
  }

  // scalar and vector derivatives (geodesic_geometry.cpp:203-224)
  const blmath::Recip den = blmath::recip_of(2.0 * r2 - rr2 + a2);
  double dr[3] = {div_by(r * x, den), div_by(r * y, den), div_by(r * z + div_by(a2 * z, rr), den)};
  double qn = r4 - 3.0 * a2 * z * z;
  const blmath::Recip w = blmath::recip_of(r * (r4 + a2zz));
  double df[3];
  df[0] = div_by(-qn * dr[0], w) * f;
  df[1] = div_by(-qn * dr[1], w) * f;
  df[2] = div_by(-(qn * dr[2] + 2.0 * a2 * r * z), w) * f;
  double c1 = x - 2.0 * r * l[0];
  double c2 = y - 2.0 * r * l[1];
  double mz = blmath::div_rn(-z, r2);
  double dl[3][3];  // dl[a][m] = d l_m / d x^a
  dl[0][0] = div_by(c1 * dr[0] + r, ra);
  dl[1][0] = div_by(c1 * dr[1] + a, ra);
  dl[2][0] = div_by(c1 * dr[2], ra);
  dl[0][1] = div_by(c2 * dr[0] - a, ra);
  dl[1][1] = div_by(c2 * dr[1] + r, ra);
  dl[2][1] = div_by(c2 * dr[2], ra);
  dl[0][2] = mz * dr[0];
  dl[1][2] = mz * dr[1];
  dl[2][2] = mz * dr[2] + div_by(1.0, rr);

  // dp_a/dlambda = -sum_{mu,nu} (1/2 d_a g^{mu nu}) p_mu p_nu, (mu,nu) row-major, summed one by one
  double hp[4] = {0.5 * p[0], 0.5 * p[1], 0.5 * p[2], 0.5 * p[3]};
  for (int d = 0; d < 3; d++) {
    double e[3], fd[3];
    for (int m = 0; m < 3; m++) {
      fd[m] = f * dl[d][m];
      e[m] = df[d] * l[m] + fd[m];
    }
    double acc = (-df[d]) * hp[0] * p[0];
    for (int n = 0; n < 3; n++) acc += e[n] * hp[0] * p[1 + n];
    for (int m = 0; m < 3; m++) {
      acc += e[m] * hp[1 + m] * p[0];
      double dfl = df[d] * l[m];
      for (int n = 0; n < 3; n++) {
        double g = -(dfl * l[n] + fd[m] * l[n] + fl[m] * dl[d][n]);
        acc += g * hp[1 + m] * p[1 + n];
      }
    }
    dp[d] = -acc;
  }

  // proper-distance rate: t_a = sum_mu (g^{a mu} - g^{0a} g^{0 mu} / g^{00}) p_mu
  double t[3];
  for (int i = 0; i < 3; i++) {
    double acc = (fl[i] - div_by(fl[i] * gcon00, g00)) * p[0];
    for (int j = 0; j < 3; j++) {
      double g = i == j ? 1.0 - P[i][j] : -P[i][j];
      acc += (g - div_by(fl[i] * fl[j], g00)) * p[1 + j];
    }
    t[i] = acc;
  }
  double s2 = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double g = i == j ? P[i][j] + 1.0 : P[i][j];
      s2 += g * t[i] * t[j];
    }
  ds = -blmath::sqrt_rn(s2);
}

}  // namespace ksx
