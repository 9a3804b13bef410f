// Device functions shared by the polarized transfer kernels (the fused single kernel in radiate_pol.cu and the
// three-stage pipeline in radiate_pol_split.cu): Kerr-Schild jet and contracted connection, tetrad legs, the real
// Stokes transport matrix, polarized synchrotron coefficients and the Stokes coupling of one sample.
#pragma once
#include "rad_sample.cuh"
#include "bf_math.cuh"

namespace {

constexpr int kBlock = 128;

// Frequency loop: rolled, with the per-frequency Stokes state in thread-local memory.  The per-frequency
// work is ~1.5e3 instructions (coefficients + 4x4 coupling), so loop and local-memory overhead are nothing,
// while an unrolled body overflows the instruction cache (ncu: no_instruction was the top stall).
#define BL_FREQ_LOOP _Pragma("unroll 1")

// ---------------------------------------------------------------------------------------------------
// Kerr-Schild geometry: g_{mu nu} = eta + f l_mu l_nu, l_mu = (1, l_i), l^mu = (-1, l_i), M = 1.
struct KsJet {
  double f, l[3];      // l_i
  double df[3];        // d_a f
  double dl[3][3];     // dl[i][a] = d_a l_i
};

// r and inv_r = 1/r: the Kerr-Schild radius of the point (the caller has them from the sampling stage).
__device__ __forceinline__ void ks_jet(const RadParams &P, double x, double y, double z, double r, double inv_r, KsJet &J) {
  if (P.ray_flat) {
    J.f = 0.0;
    for (int i = 0; i < 3; i++) {
      J.l[i] = 0.0;
      J.df[i] = 0.0;
      for (int a = 0; a < 3; a++) J.dl[i][a] = 0.0;
    }
    return;
  }
  const double a = P.a, a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double r2 = r * r, r4 = r2 * r2;
  double den_f = r4 + a2 * z * z;
  double inv_den = 1.0 / den_f;
  J.f = 2.0 * r2 * r * inv_den;
  double ra = 1.0 / (r2 + a2);
  J.l[0] = (r * x + a * y) * ra;
  J.l[1] = (r * y - a * x) * ra;
  J.l[2] = z * inv_r;
  // reference geodesic_geometry.cpp:203-224 (derivatives of r, f, l)
  double inv = 1.0 / (2.0 * r2 - rr2 + a2);
  double dr[3] = {r * x * inv, r * y * inv, (r * z + a2 * z * inv_r) * inv};
  double qn = r4 - 3.0 * a2 * z * z;
  double w = J.f * inv_r * inv_den;
  J.df[0] = -qn * dr[0] * w;
  J.df[1] = -qn * dr[1] * w;
  J.df[2] = -(qn * dr[2] + 2.0 * a2 * r * z) * w;
  double c1 = x - 2.0 * r * J.l[0], c2 = y - 2.0 * r * J.l[1], mz = -z * inv_r * inv_r;
  J.dl[0][0] = (c1 * dr[0] + r) * ra;
  J.dl[0][1] = (c1 * dr[1] + a) * ra;
  J.dl[0][2] = c1 * dr[2] * ra;
  J.dl[1][0] = (c2 * dr[0] - a) * ra;
  J.dl[1][1] = (c2 * dr[1] + r) * ra;
  J.dl[1][2] = c2 * dr[2] * ra;
  J.dl[2][0] = mz * dr[0];
  J.dl[2][1] = mz * dr[1];
  J.dl[2][2] = mz * dr[2] + inv_r;
}

// v_mu = g_{mu nu} v^nu and v^mu = g^{mu nu} v_nu
__device__ __forceinline__ void lower(const KsJet &J, const double v[4], double out[4]) {
  double lv = v[0] + J.l[0] * v[1] + J.l[1] * v[2] + J.l[2] * v[3];
  double s = J.f * lv;
  out[0] = -v[0] + s;
  out[1] = v[1] + s * J.l[0];
  out[2] = v[2] + s * J.l[1];
  out[3] = v[3] + s * J.l[2];
}
__device__ __forceinline__ void raise(const KsJet &J, const double v[4], double out[4]) {
  double lv = -v[0] + J.l[0] * v[1] + J.l[1] * v[2] + J.l[2] * v[3];
  double s = J.f * lv;
  out[0] = -v[0] + s;
  out[1] = v[1] - s * J.l[0];
  out[2] = v[2] - s * J.l[1];
  out[3] = v[3] - s * J.l[2];
}

// A^mu_beta = k^alpha Gamma^mu_{alpha beta} for the Kerr-Schild connection (radiation_geometry.cpp:274-410),
// without forming Gamma:  A = 1/2 g^{mu nu} (k.d g_{beta nu} + k^alpha d_beta g_{alpha nu} - k^alpha d_nu g_{alpha beta}).
__device__ __forceinline__ void contracted_connection(const KsJet &J, const double k[4], double A[4][4]) {
  // W_{beta nu} = k.d g_{beta nu} + k^alpha d_beta g_{alpha nu} - k^alpha d_nu g_{alpha beta} written out for
  // g = eta + f l l with l_0 = 1 and nothing depending on time: with
  //   u_i = (k.grad f) l_i + f k.grad l_i,   v_i = d_i f (l.k) + f k^j d_i l_j,
  // W_00 = k.grad f,  W_0j = u_j - v_j,  W_i0 = u_i + v_i,
  // W_ij = W_i0 l_j + l_i (f k.grad l_j - v_j) + f (l.k) (d_i l_j - d_j l_i)
  // (the products with the vanishing time components are dropped instead of being multiplied out).
  const double kf = k[1] * J.df[0] + k[2] * J.df[1] + k[3] * J.df[2];
  const double lk = k[0] + J.l[0] * k[1] + J.l[1] * k[2] + J.l[2] * k[3];
  const double g = J.f * lk;
  double W[4][4];
  double q[3];
  W[0][0] = kf;
  for (int i = 0; i < 3; i++) {
    const double kl = k[1] * J.dl[i][0] + k[2] * J.dl[i][1] + k[3] * J.dl[i][2];   // k.grad l_i
    const double m = k[1] * J.dl[0][i] + k[2] * J.dl[1][i] + k[3] * J.dl[2][i];    // k^j d_i l_j
    const double fkl = J.f * kl;
    const double u = kf * J.l[i] + fkl, v = J.df[i] * lk + J.f * m;
    W[0][1 + i] = u - v;
    W[1 + i][0] = u + v;
    q[i] = fkl - v;
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      W[1 + i][1 + j] = W[1 + i][0] * J.l[j] + J.l[i] * q[j] + g * (J.dl[j][i] - J.dl[i][j]);
  // A^mu_beta = 1/2 g^{mu nu} W_{beta nu},  g^{mu nu} = eta - f l^mu l^nu,  l^mu = (-1, l_i)
  for (int b = 0; b < 4; b++) {
    const double lw = J.l[0] * W[b][1] + J.l[1] * W[b][2] + J.l[2] * W[b][3] - W[b][0];
    const double s = J.f * lw;
    A[0][b] = 0.5 * (s - W[b][0]);
    A[1][b] = 0.5 * (W[b][1] - s * J.l[0]);
    A[2][b] = 0.5 * (W[b][2] - s * J.l[1]);
    A[3][b] = 0.5 * (W[b][3] - s * J.l[2]);
  }
}

// (D v)^mu = -A^mu_beta v^beta : rate of change of a parallel-transported vector's components
__device__ __forceinline__ void transport_rate(const double A[4][4], const double v[4], double out[4]) {
  for (int mu = 0; mu < 4; mu++) out[mu] = -(A[mu][0] * v[0] + A[mu][1] * v[1] + A[mu][2] * v[2] + A[mu][3] * v[3]);
}

__device__ __forceinline__ double dot4(const double a[4], const double b[4]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
}

// Legs 1 and 2 of the orthonormal tetrad of radiation_geometry.cpp:597-658: e_0 = u, e_3 = k/omega - u,
// e_2 = normalised projection of `up` orthogonal to e_0 and e_3, e_1 completes the right-handed frame.
// Outputs contravariant legs e1, e2 and their covariant forms f1, f2.
__device__ __forceinline__ void tetrad_legs(const KsJet &J, const double ucon[4], const double ucov[4],
                                            const double kcon[4], const double kcov[4], const double up[4],
                                            double e1[4], double e2[4], double f1[4], double f2[4]) {
  double inv_omega = -1.0 / dot4(kcov, ucon);
  double k_up = dot4(kcov, up) * inv_omega;
  double u_up = dot4(ucov, up) * inv_omega;
  double e3[4];
  for (int mu = 0; mu < 4; mu++) e3[mu] = kcon[mu] * inv_omega - ucon[mu];
  for (int mu = 0; mu < 4; mu++) e2[mu] = up[mu] - k_up * e3[mu] + u_up * kcon[mu];
  lower(J, e2, f2);
  double inv_norm = rsqrt(dot4(f2, e2));
  for (int mu = 0; mu < 4; mu++) {
    e2[mu] *= inv_norm;
    f2[mu] *= inv_norm;
  }
  const double *t0 = ucon, *t2 = e2, *t3 = e3;
  f1[0] = t0[1] * (t2[3] * t3[2] - t2[2] * t3[3]) + t0[2] * (t2[1] * t3[3] - t2[3] * t3[1]) +
          t0[3] * (t2[2] * t3[1] - t2[1] * t3[2]);
  f1[1] = t0[0] * (t2[2] * t3[3] - t2[3] * t3[2]) + t0[2] * (t2[3] * t3[0] - t2[0] * t3[3]) +
          t0[3] * (t2[0] * t3[2] - t2[2] * t3[0]);
  f1[2] = t0[0] * (t2[3] * t3[1] - t2[1] * t3[3]) + t0[1] * (t2[0] * t3[3] - t2[3] * t3[0]) +
          t0[3] * (t2[1] * t3[0] - t2[0] * t3[1]);
  f1[3] = t0[0] * (t2[1] * t3[2] - t2[2] * t3[1]) + t0[1] * (t2[2] * t3[0] - t2[0] * t3[2]) +
          t0[2] * (t2[0] * t3[1] - t2[1] * t3[0]);
  raise(J, f1, e1);
}

// Transported legs of the previous tetrad and their projections on the new covariant legs.
// T(u (x) v) = u v + h [Da(u) v + u Da(v)] + h h2 [Da Dp(u) v + Dp(u) Da(v) + Da(u) Dp(v) + u Da Dp(v)]
// (predictor with the previous sample's own connection, corrector with the averaged one), h2 < 0 disables
// the corrector (final half step to the camera).
struct LegProj {
  double u[2], up[2], ua[2], uap[2];  // f_a . {e, Dp e, Da e, Da Dp e}
};

__device__ __forceinline__ double pair_proj(const LegProj &c, const LegProj &d, int a, int b, double h, double hh2,
                                            bool corrector) {
  double base = c.u[a] * d.u[b];
  if (!corrector) return base + hh2 * (c.up[a] * d.u[b] + c.u[a] * d.up[b]);
  return base + h * (c.ua[a] * d.u[b] + c.u[a] * d.ua[b]) +
         hh2 * (c.uap[a] * d.u[b] + c.up[a] * d.ua[b] + c.ua[a] * d.up[b] + c.u[a] * d.uap[b]);
}

// M acting on (I,Q,U,V): rows I',Q',U' from the symmetric part, V' from the antisymmetric part.
struct StokesMap {
  double m[3][3];
  double vv;
};

__device__ __forceinline__ void stokes_map(const LegProj L[2], double h, double hh2, bool corrector, StokesMap &M) {
  // P[c][d][a][b] = f_a . T(e_c (x) e_d) . f_b
  double Pm[2][2][2][2];
  for (int c = 0; c < 2; c++)
    for (int d = 0; d < 2; d++)
      for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) Pm[c][d][a][b] = pair_proj(L[c], L[d], a, b, h, hh2, corrector);
  // N = (I+Q) e1e1 + (I-Q) e2e2 + (U - iV) e1e2 + (U + iV) e2e1 ; Stokes' from n'_ab (polarized.cpp:286-292)
  for (int row = 0; row < 3; row++) {
    // combination of n'_ab giving I', Q', U'
    double w11 = row == 0 ? 0.5 : (row == 1 ? 0.5 : 0.0);
    double w22 = row == 0 ? 0.5 : (row == 1 ? -0.5 : 0.0);
    double w12 = row == 2 ? 0.5 : 0.0;
    auto comb = [&](int c, int d) {
      return w11 * Pm[c][d][0][0] + w22 * Pm[c][d][1][1] + w12 * (Pm[c][d][0][1] + Pm[c][d][1][0]);
    };
    double c11 = comb(0, 0), c22 = comb(1, 1), c12 = comb(0, 1) + comb(1, 0);
    M.m[row][0] = c11 + c22;   // I
    M.m[row][1] = c11 - c22;   // Q
    M.m[row][2] = c12;         // U
  }
  M.vv = 0.5 * (Pm[1][0][1][0] - Pm[0][1][1][0] - Pm[1][0][0][1] + Pm[0][1][0][1]);
}

// Stokes transport matrix from the previous sample (jet_p, contravariant momentum k_p, legs e_p, affine step dlam_p) to the
// current one (jet, kcon, covariant legs f1, f2, affine step dlam): predictor with the previous sample's own connection,
// corrector with the average of both samples' connections contracted with the averaged momentum (polarized.cpp:136-192).
__device__ __forceinline__ void transport_map(const KsJet &jet_p, const double k_p[4], const double e_p[2][4], double dlam_p,
                                              const KsJet &jet, const double kcon[4], const double f1[4], const double f2[4],
                                              double dlam, StokesMap &M) {
  const double ks[4] = {k_p[0] + kcon[0], k_p[1] + kcon[1], k_p[2] + kcon[2], k_p[3] + kcon[3]};
  double A_avg[4][4], A_tmp[4][4], A_pp[4][4];
  contracted_connection(jet_p, ks, A_avg);
  contracted_connection(jet, ks, A_tmp);
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) A_avg[a][b] = 0.25 * (A_avg[a][b] + A_tmp[a][b]);
  contracted_connection(jet_p, k_p, A_pp);
  const double h = (dlam_p + dlam) / 2.0, h2 = (dlam_p + dlam) / 4.0;
  LegProj L[2];
  for (int c = 0; c < 2; c++) {
    double vp[4], va[4], vap[4];
    transport_rate(A_pp, e_p[c], vp);
    // corrector derivative acts on the predicted tensor: Da(e + h2 Dp e) = Da e + h2 Da Dp e
    transport_rate(A_avg, e_p[c], va);
    transport_rate(A_avg, vp, vap);
    L[c].u[0] = dot4(f1, e_p[c]);  L[c].u[1] = dot4(f2, e_p[c]);
    L[c].up[0] = dot4(f1, vp);     L[c].up[1] = dot4(f2, vp);
    L[c].ua[0] = dot4(f1, va);     L[c].ua[1] = dot4(f2, va);
    L[c].uap[0] = dot4(f1, vap);   L[c].uap[1] = dot4(f2, vap);
  }
  stokes_map(L, h, h * h2, true, M);
}

// The last half step, from the sample nearest the camera onto the camera tetrad's covariant legs (polarized.cpp:816-833):
// predictor only.
__device__ __forceinline__ void transport_map_final(const KsJet &jet_p, const double k_p[4], const double e_p[2][4], double dlam_p,
                                                    const double f1[4], const double f2[4], StokesMap &M) {
  double A_pp[4][4];
  contracted_connection(jet_p, k_p, A_pp);
  LegProj L[2];
  for (int c = 0; c < 2; c++) {
    double vp[4];
    transport_rate(A_pp, e_p[c], vp);
    L[c].u[0] = dot4(f1, e_p[c]);  L[c].u[1] = dot4(f2, e_p[c]);
    L[c].up[0] = dot4(f1, vp);     L[c].up[1] = dot4(f2, vp);
    L[c].ua[0] = L[c].ua[1] = L[c].uap[0] = L[c].uap[1] = 0.0;
  }
  stokes_map(L, 0.0, dlam_p / 2.0, false, M);
}

struct Coefficients {
  double j[3], a[3], rho[2];  // (I,Q,V), (I,Q,V), (Q,V); Stokes U components vanish in this tetrad
};

// Frequency-independent part of the polarized synchrotron coefficients of one sample.  The reference
// (simulation_coefficients.cpp:458-698) evaluates ~45 std::pow per frequency for the kappa distribution;
// here logarithms of the per-sample quantities are taken once, powers of the pitch angle are hoisted, and
// each remaining power is exp(c * ln x).
struct PolSample {
  double om, inv_om;        // nu = om * image_frequency
  double nu_c, sin_b, cos_b, sin2, n_e, n_nuc;   // n_nuc = n_e e^2 nu_c / c
  double sgn;               // sign of cos(theta_B)
  // thermal
  double inv_nu_s, log_inv_nu_s, h_kt, theta_e, var_d, cos_over_theta;
  double k1_k2, k0, inv_k2; // Bessel ratios (valid iff theta_e >= 0.01)
  // power law / kappa
  double log_om, log_ncs, log_ne, cot;
  double power_vb;          // (3.1 sin^-1.92 - 3.1)^0.512
  double inv_nu_k;          // 1 / nu_kappa
  double lvd_j, lvf_j, lvd_a, lvf_a;   // ln of the pitch-angle factors of j_V and alpha_V (kappa)
};

// DIST: electron distributions compiled into an instantiation -- bit 0 thermal, bit 1 power law, bit 2 kappa; 7 = all,
// chosen at run time from the fractions.  The single-distribution instantiations (thermal only, kappa only) drop
// the other distributions' code and, above all, their live registers from the frequency loop.
template <int DIST> __device__ __forceinline__ bool has_thermal(const RadParams &P) { return DIST == 7 ? P.thermal_frac != 0.0 : (DIST & 1) != 0; }
template <int DIST> __device__ __forceinline__ bool has_power(const RadParams &P) { return DIST == 7 ? P.power_frac != 0.0 : (DIST & 2) != 0; }
template <int DIST> __device__ __forceinline__ bool has_kappa(const RadParams &P) { return DIST == 7 ? P.kappa_frac != 0.0 : (DIST & 4) != 0; }

template <int DIST>
__device__ __forceinline__ void pol_sample(const RadParams &P, const rad::Plasma &s, double om, double sin_b,
                                           double cos_b, const double kk[3], PolSample &q) {
  q.om = om;
  q.inv_om = 1.0 / om;
  q.nu_c = s.bb_cgs * (phys::e / (2.0 * phys::pi * phys::m_e * phys::c));
  q.sin_b = sin_b;
  q.cos_b = cos_b;
  q.sin2 = sin_b * sin_b;
  q.sgn = cos_b >= 0.0 ? 1.0 : -1.0;
  q.n_e = s.n_e_cgs;
  q.n_nuc = s.n_e_cgs * q.nu_c * (phys::e * phys::e / phys::c);
  q.theta_e = s.theta_e;
  q.inv_nu_s = q.log_inv_nu_s = q.h_kt = q.var_d = q.cos_over_theta = 0.0;
  q.k1_k2 = q.k0 = q.inv_k2 = 0.0;
  if (has_thermal<DIST>(P)) {
    q.inv_nu_s = 4.5 * s.inv_theta_e * s.inv_theta_e / (q.nu_c * sin_b);
    q.h_kt = phys::h * s.inv_theta_e * (1.0 / (phys::m_e * phys::c * phys::c));
    double te96 = bfm::exp_bf(0.96 * bfm::log_bf(s.theta_e));
    q.var_d = (7.0 * te96 + 35.0) / (10.0 * te96 + 75.0) * 1.8877486253633870;
    q.cos_over_theta = cos_b * s.inv_theta_e;
    if (s.theta_e >= 0.01) {
      q.log_inv_nu_s = bfm::log_bf(q.inv_nu_s);
      q.inv_k2 = 1.0 / kk[2];
      q.k1_k2 = kk[1] * q.inv_k2;
      q.k0 = kk[0];
    }
  }
  q.log_om = bfm::log_bf(om);
  q.log_ncs = q.log_ne = q.cot = q.power_vb = q.inv_nu_k = 0.0;
  q.lvd_j = q.lvf_j = q.lvd_a = q.lvf_a = 0.0;
  if (has_power<DIST>(P) || has_kappa<DIST>(P)) {
    // table-driven logarithms / exponentials (bf_math.cuh) here too: ten libm calls per sample were an eighth of the
    // coefficient stage.  sin(theta_B) = 0 or 1 gives logarithms of zero, hence the _z variant (-inf, as libm).
    q.log_ncs = bfm::log_bf_z(q.nu_c * sin_b);
    q.log_ne = bfm::log_bf(s.n_e_cgs);
    double log_sin = bfm::log_bf_z(sin_b);
    if (has_power<DIST>(P)) {
      q.cot = cos_b / sin_b;
      q.power_vb = pow(3.1 * exp(-1.92 * log_sin) - 3.1, 0.512);
    }
    if (has_kappa<DIST>(P)) {
      q.inv_nu_k = 1.0 / (q.nu_c * P.plasma_w * P.plasma_w * P.plasma_kappa * P.plasma_kappa * sin_b);
      q.lvd_j = 0.48 * bfm::log_bf_z(bfm::exp_bf(-2.4 * log_sin) - 1.0);
      q.lvf_j = 0.44 * bfm::log_bf_z(bfm::exp_bf(-2.5 * log_sin) - 1.0);
      q.lvd_a = 0.446 * bfm::log_bf_z(bfm::exp_bf(-2.28 * log_sin) - 1.0);
      q.lvf_a = 0.5 * bfm::log_bf_z(bfm::exp_bf(-2.05 * log_sin) - 1.0);
    }
  }
}

// Polarized synchrotron coefficients at image frequency l (simulation_coefficients.cpp:458-698).
template <int DIST>
__device__ __forceinline__ void synchrotron_polarized(const RadParams &P, const PolSample &q, int l, Coefficients &C) {
  const double e2 = phys::e * phys::e;
  double nu_cgs = q.om * P.freqs[l];
  double inv_nu = q.inv_om * P.inv_freqs[l];
  double inv_nu_2 = inv_nu * inv_nu;
  for (int i = 0; i < 3; i++) C.j[i] = C.a[i] = 0.0;
  C.rho[0] = C.rho[1] = 0.0;
  if (has_thermal<DIST>(P)) {
    double xx = nu_cgs * q.inv_nu_s;
    double xx_neg_1_2 = rsqrt(xx);
    double xx_1_2 = xx * xx_neg_1_2, xx_1_3 = cbrt(xx);
    double xx_1_6 = sqrt(xx_1_3);
    double coefficient = P.thermal_frac * q.n_nuc * inv_nu_2 * bfm::exp_bf(-xx_1_3);
    double var_a = phys::sqrt2 * phys::pi / 27.0 * q.sin_b;
    const double var_b = 1.8877486253633870;  // 2^(11/12)
    double var_c = xx_1_2 + var_b * xx_1_6;
    C.j[0] = coefficient * var_a * var_c * var_c;
    double var_e = xx_1_2 + q.var_d * xx_1_6;
    double var_g = phys::pi / 3.0 + phys::pi / 3.0 * xx_1_3 + 2.0 / 300.0 * xx_1_2 + 2.0 / 19.0 * phys::pi * xx_1_3 * xx_1_3;
    C.j[1] = -coefficient * var_a * var_e * var_e;
    C.j[2] = coefficient * q.cos_over_theta * var_g;
    double inv_b_nu = expm1(q.h_kt * nu_cgs) * (phys::c * phys::c / (2.0 * phys::h));
    C.a[0] = C.j[0] * inv_b_nu;
    C.a[1] = C.j[1] * inv_b_nu;
    C.a[2] = C.j[2] * inv_b_nu;
    if (C.a[0] * C.a[0] <= 0x1p-1024) C.a[0] = C.a[1] = C.a[2] = 0.0;
    // Faraday rotation and conversion, with the cold-plasma trap below theta_e = 0.01
    double coefficient_q = -P.thermal_frac * q.n_e * e2 * q.nu_c * q.nu_c * q.sin2 * inv_nu_2 * (1.0 / (phys::m_e * phys::c));
    double coefficient_v = P.thermal_frac * 2.0 * q.n_e * e2 * q.nu_c * q.cos_b * inv_nu * (1.0 / (phys::m_e * phys::c));
    double factor_q = 0.0, factor_v = 1.0;
    if (q.theta_e >= 0.01) {
      double lx = q.log_om + P.log_freqs[l] + q.log_inv_nu_s;  // ln xx
      double va = 2.011 * bfm::exp_bf(-19.78 * bfm::exp_bf(-0.5175 * lx));
      double vb = cos(39.89 * xx_neg_1_2) * bfm::exp_bf(-70.16 * bfm::exp_bf(-0.6 * lx));
      double vc = 0.011 * bfm::exp_bf(-1.69 * xx_neg_1_2);
      double vd = 0.003135 * xx * xx_1_3;
      // 0.5 (1 + tanh y) = 1 - 1/(1 + e^(2y)),  y = 10 ln(0.6648 xx^-1/2),  ln 0.6648 = -0.40826...
      double ve = 1.0 - 1.0 / (1.0 + bfm::exp_bf(20.0 * (-0.4082690354408987 - 0.5 * lx)));
      double f_0 = va - vb - vc;
      double f_m = f_0 + (vc - vd) * ve;
      double delta_jj_5 = 0.4379 * bfm::log_bf(1.0 + 1.3414 * bfm::exp_bf(-0.7515 * lx));
      factor_q = f_m * (q.k1_k2 + 6.0 * q.theta_e);
      factor_v = (q.k0 - delta_jj_5) * q.inv_k2;
      factor_v = (factor_v < 0.0 || factor_v > 1.0) ? 1.0 : factor_v;
    }
    C.rho[0] = coefficient_q * factor_q;
    C.rho[1] = coefficient_v * factor_v;
  }
  if (has_power<DIST>(P) || has_kappa<DIST>(P)) {
    double log_nu = q.log_om + P.log_freqs[l];
    double lr = log_nu - q.log_ncs;  // ln(nu / (nu_c sin(theta_B)))
    if (has_power<DIST>(P)) {
      double e_half = bfm::exp_bf(-0.5 * lr);  // (nu / (nu_c sin))^-1/2
      double coefficient = P.power_frac * q.n_nuc * inv_nu_2 * P.power_jj * q.sin_b * bfm::exp_bf(-(P.plasma_p - 1.0) / 2.0 * lr);
      C.j[0] += coefficient;
      C.j[1] += coefficient * P.power_jj_q;
      C.j[2] += coefficient * P.power_jj_v * q.cot * (1.7320508075688772 * e_half);
      double coefficient_a = P.power_frac * q.n_e * (e2 / (phys::m_e * phys::c)) * P.power_aa * bfm::exp_bf(-(P.plasma_p + 2.0) / 2.0 * lr);
      C.a[0] += coefficient_a;
      C.a[1] += coefficient_a * P.power_aa_q;
      C.a[2] += coefficient_a * P.power_aa_v * q.power_vb * e_half * q.sgn;
      double rb = e_half * e_half;  // nu_c sin / nu
      double ra = q.n_e * (e2 / (phys::m_e * phys::c)) / rb;
      double rc = rb * rb, rd = rc * rb;
      double re = 1.0 - bfm::exp_bf((P.plasma_p / 2.0 - 1.0) * (P.log_power_gmin - lr));
      double coefficient_r = P.power_frac * P.power_rho * ra;
      C.rho[0] += coefficient_r * P.power_rho_q * rd * re;
      C.rho[1] += coefficient_r * P.power_rho_v * rc * q.cot;
    }
    if (has_kappa<DIST>(P)) {
      double lx = lr - P.log_w2k2;   // ln(nu / nu_kappa)
      double xx = nu_cgs * q.inv_nu_k;
      double lm035 = -0.35 * lx, lm12 = -0.5 * lx;
      {
        const double ix_i = P.kappa_inv_x[0], ix_q = P.kappa_inv_x[1], ix_v = P.kappa_inv_x[2];
        double lva = P.log_k_j_pref + q.log_ne + q.log_ncs - 2.0 * log_nu;  // includes the sin(theta_B) factor
        double l_lo = P.log_kjl + lva + lx * (1.0 / 3.0);
        double l_hi = P.log_kjh + lva - (P.plasma_kappa - 2.0) / 2.0 * lx;
        C.j[0] += bfm::bridge(l_lo, l_hi, P.kappa_jj_x_i, ix_i);
        C.j[1] -= bfm::bridge(l_lo + P.log_kj_low_q, l_hi + P.log_kj_high_q, P.kappa_jj_x_q, ix_q);
        C.j[2] += bfm::bridge(l_lo + P.log_kj_low_v + q.lvd_j + lm035, l_hi + P.log_kj_high_v + q.lvf_j + lm12,
                         P.kappa_jj_x_v, ix_v) * q.sgn;
      }
      {
        const double ix_i = P.kappa_inv_x[3], ix_q = P.kappa_inv_x[4], ix_v = P.kappa_inv_x[5];
        double lva = P.log_k_a_pref + q.log_ne;
        double l_lo = P.log_kal + lva - 2.0 / 3.0 * lx;
        double l_hi = P.log_kah_base + lva - (1.0 + P.plasma_kappa) / 2.0 * lx;
        C.a[0] += bfm::bridge(l_lo, l_hi + (P.log_kah - P.log_kah_base), P.kappa_aa_x_i, ix_i);
        C.a[1] -= bfm::bridge(l_lo + P.log_ka_low_q, l_hi + P.log_ka_high_q, P.kappa_aa_x_q, ix_q);
        C.a[2] += bfm::bridge(l_lo + P.log_ka_low_v + q.lvd_a + lm035, l_hi + P.log_ka_high_v + q.lvf_a + lm12,
                         P.kappa_aa_x_v, ix_v) * q.sgn;
      }
      {
        double va = -P.kappa_frac * q.n_e * e2 * q.nu_c * q.nu_c * q.sin2 * inv_nu_2 * (1.0 / (phys::m_e * phys::c));
        double vb = P.kappa_frac * 2.0 * q.n_e * e2 * q.nu_c * q.cos_b * inv_nu * (1.0 / (phys::m_e * phys::c));
        double x084 = bfm::exp_bf(0.84 * lx);
        double xx_m12 = bfm::exp_bf(lm12);
        // The fits are tabulated at kappa = 3.5, 4, 4.5, 5 and blended linearly in between.  On a tabulated value one
        // weight is exactly zero and that entry (4 exponentials, a sine and a logarithm) is not evaluated; its term
        // is 0 x (a finite number) in the reference.
        double q_lo = 0.0, q_hi = 0.0, v_lo = 0.0, v_hi = 0.0;
        if (P.kappa_rho_frac != 1.0) {
          q_lo = va * P.kappa_rho_q_low_a * (1.0 - bfm::exp_bf(P.kappa_rho_q_low_b * x084) -
                 sin(P.kappa_rho_q_low_c * xx) * bfm::exp_bf(P.kappa_rho_q_low_d * bfm::exp_bf(P.kappa_rho_q_low_e * lx)));
          v_lo = P.kappa_rho_v * vb * P.kappa_rho_v_low_a * (1.0 - 0.17 * bfm::log_bf(1.0 + P.kappa_rho_v_low_b * xx_m12));
        }
        if (P.kappa_rho_frac != 0.0) {
          q_hi = va * P.kappa_rho_q_high_a * (1.0 - bfm::exp_bf(P.kappa_rho_q_high_b * x084) -
                 sin(P.kappa_rho_q_high_c * xx) * bfm::exp_bf(P.kappa_rho_q_high_d * bfm::exp_bf(P.kappa_rho_q_high_e * lx)));
          v_hi = P.kappa_rho_v * vb * P.kappa_rho_v_high_a * (1.0 - 0.17 * bfm::log_bf(1.0 + P.kappa_rho_v_high_b * xx_m12));
        }
        C.rho[0] += (1.0 - P.kappa_rho_frac) * q_lo + P.kappa_rho_frac * q_hi;
        C.rho[1] += (1.0 - P.kappa_rho_frac) * v_lo + P.kappa_rho_frac * v_hi;
      }
    }
  }
}

// The polarized kappa coefficients of ALL image frequencies of one sample, term by term (kappa-only plasmas; same
// formulas as synchrotron_polarized<4>, simulation_coefficients.cpp:608-698).  For a bridged coefficient
// (lo^-x + hi^-x)^(-1/x) = exp(m - ln(1 + e^(-x |d|)) / x), d = ln lo - ln hi, m = min(ln lo, ln hi), both logarithms are
// affine in ln nu, so e^(-x d) at frequency l is e^(-x d) at frequency 0 times a host constant: two exponentials per
// sample (e^(-x d_0), e^(+x d_0): the one not selected may overflow, harmlessly) replace one per frequency.  The pure
// powers of nu / nu_kappa inside the Faraday fits are shared the same way.  store(k, l, value): coefficient k = 0..7
// (j_I, j_Q, j_V, alpha_I, alpha_Q, alpha_V, rho_Q, rho_V) at frequency l.
template <class Store>
__device__ __forceinline__ void kappa_polarized_all(const RadParams &P, const PolSample &q, int F, Store store) {
  const double e2 = phys::e * phys::e;
  const double log_nu0 = q.log_om + P.log_freqs[0];
  const double lr0 = log_nu0 - q.log_ncs;
  const double lx0 = lr0 - P.log_w2k2;   // ln(nu_0 / nu_kappa)
  const double lva_j = P.log_k_j_pref + q.log_ne + q.log_ncs - 2.0 * log_nu0;
  const double jlo = P.log_kjl + lva_j + lx0 * (1.0 / 3.0);
  const double jhi = P.log_kjh + lva_j - (P.plasma_kappa - 2.0) / 2.0 * lx0;
  const double lva_a = P.log_k_a_pref + q.log_ne;
  const double alo = P.log_kal + lva_a - 2.0 / 3.0 * lx0;
  const double ahi = P.log_kah_base + lva_a - (1.0 + P.plasma_kappa) / 2.0 * lx0;
  auto bridged = [&](int t, double a0, double b0, double x, double sign) {
    const double inv_x = P.kappa_inv_x[t], s_lo = P.kappa_slope_lo[t], s_hi = P.kappa_slope_hi[t];
    const double d0 = a0 == b0 ? 0.0 : a0 - b0;
    const double e_neg = bfm::exp_bf(-x * d0), e_pos = bfm::exp_bf(x * d0);
BL_FREQ_LOOP
    for (int l = 0; l < F; l++) {
      const double dl = P.dlog_freqs[l];
      const double a = fma(s_lo, dl, a0), b = fma(s_hi, dl, b0);
      const double d = a == b ? 0.0 : a - b;
      const double u = d < 0.0 ? e_pos * P.kappa_kinv[t][l] : e_neg * P.kappa_k[t][l];   // e^(-x |d|)
      const double m = d < 0.0 ? a : b;
      store(t, l, sign * bfm::exp_bf(m - bfm::log_bf(1.0 + u) * inv_x));
    }
  };
  bridged(0, jlo, jhi, P.kappa_jj_x_i, 1.0);
  bridged(1, jlo + P.log_kj_low_q, jhi + P.log_kj_high_q, P.kappa_jj_x_q, -1.0);
  bridged(2, jlo + P.log_kj_low_v + q.lvd_j - 0.35 * lx0, jhi + P.log_kj_high_v + q.lvf_j - 0.5 * lx0, P.kappa_jj_x_v, q.sgn);
  bridged(3, alo, ahi + (P.log_kah - P.log_kah_base), P.kappa_aa_x_i, 1.0);
  bridged(4, alo + P.log_ka_low_q, ahi + P.log_ka_high_q, P.kappa_aa_x_q, -1.0);
  bridged(5, alo + P.log_ka_low_v + q.lvd_a - 0.35 * lx0, ahi + P.log_ka_high_v + q.lvf_a - 0.5 * lx0, P.kappa_aa_x_v, q.sgn);
  // Faraday conversion and rotation: fits tabulated at kappa = 3.5, 4, 4.5, 5, blended (a zero weight skips its entry)
  const double x084_0 = bfm::exp_bf(0.84 * lx0), xm12_0 = bfm::exp_bf(-0.5 * lx0);
  const bool use_lo = P.kappa_rho_frac != 1.0, use_hi = P.kappa_rho_frac != 0.0;
  const double ein_lo0 = use_lo ? bfm::exp_bf(P.kappa_rho_q_low_e * lx0) : 0.0;
  const double ein_hi0 = use_hi ? bfm::exp_bf(P.kappa_rho_q_high_e * lx0) : 0.0;
  const double pref = P.kappa_frac * q.n_e * e2 * q.nu_c * (1.0 / (phys::m_e * phys::c));
BL_FREQ_LOOP
  for (int l = 0; l < F; l++) {
    const double nu_cgs = q.om * P.freqs[l];
    const double inv_nu = q.inv_om * P.inv_freqs[l];
    const double xx = nu_cgs * q.inv_nu_k;
    const double va = -pref * q.nu_c * q.sin2 * inv_nu * inv_nu;
    const double vb = pref * 2.0 * q.cos_b * inv_nu;
    const double x084 = x084_0 * P.rho_c84[l], xx_m12 = xm12_0 * P.rho_cm12[l];
    double q_lo = 0.0, q_hi = 0.0, v_lo = 0.0, v_hi = 0.0;
    if (use_lo) {
      q_lo = va * P.kappa_rho_q_low_a * (1.0 - bfm::exp_bf(P.kappa_rho_q_low_b * x084) -
             sin(P.kappa_rho_q_low_c * xx) * bfm::exp_bf(P.kappa_rho_q_low_d * (ein_lo0 * P.rho_cqe_low[l])));
      v_lo = P.kappa_rho_v * vb * P.kappa_rho_v_low_a * (1.0 - 0.17 * bfm::log_bf(1.0 + P.kappa_rho_v_low_b * xx_m12));
    }
    if (use_hi) {
      q_hi = va * P.kappa_rho_q_high_a * (1.0 - bfm::exp_bf(P.kappa_rho_q_high_b * x084) -
             sin(P.kappa_rho_q_high_c * xx) * bfm::exp_bf(P.kappa_rho_q_high_d * (ein_hi0 * P.rho_cqe_high[l])));
      v_hi = P.kappa_rho_v * vb * P.kappa_rho_v_high_a * (1.0 - 0.17 * bfm::log_bf(1.0 + P.kappa_rho_v_high_b * xx_m12));
    }
    store(6, l, (1.0 - P.kappa_rho_frac) * q_lo + P.kappa_rho_frac * q_hi);
    store(7, l, (1.0 - P.kappa_rho_frac) * v_lo + P.kappa_rho_frac * v_hi);
  }
}

// Clamp to a physically admissible Stokes vector (polarized.cpp:456-466, :782-790)
__device__ __forceinline__ void admissible(double s[4], bool clamp_i) {
  if (clamp_i) s[0] = s[0] < 0.0 ? 0.0 : s[0];
  double pol = s[1] * s[1] + s[2] * s[2] + s[3] * s[3];
  if (pol > s[0] * s[0]) {
    double factor = sqrt(s[0] * s[0] / pol);
    s[1] *= factor;
    s[2] *= factor;
    s[3] *= factor;
  }
}

// Emission + absorption without rotation over a path dl (polarized.cpp:388-453 / :571-653); j_s, a_s indexed
// by Stokes component (U entries zero).
__device__ __forceinline__ void couple_absorb(const double s0[4], const double j[4], const double al[4], double dl,
                                              double delta_tau, bool thin, double out[4]) {
  double alpha_sq = al[1] * al[1] + al[3] * al[3];
  double alpha_p = sqrt(alpha_sq);
  if (al[0] == 0.0) {
    for (int a = 0; a < 4; a++) out[a] = s0[a] + j[a] * dl;
  } else if (alpha_p == 0.0) {
    if (thin) {
      double en = exp(-delta_tau), em = expm1(delta_tau);
      for (int a = 0; a < 4; a++) out[a] = en * (s0[a] + j[a] / al[0] * em);
    } else {
      for (int a = 0; a < 4; a++) out[a] = j[a] / al[0];
    }
  } else if (thin) {
    double exp_neg_i = exp(-delta_tau);
    double xp = alpha_p * dl;
    double exp_neg_p = exp(-xp);
    double sinh_p = sinh(xp), cosh_p = cosh(xp);
    double coshm1_p = 0.5 * (expm1(xp) + exp_neg_p - 1.0);
    double alpha_ss = al[1] * s0[1] + al[3] * s0[3];
    double alpha_j = al[1] * j[1] + al[3] * j[3];
    double fac = 1.0 / (al[0] * al[0] - alpha_sq);
    out[0] = (s0[0] * cosh_p - alpha_ss / alpha_p * sinh_p) * exp_neg_i +
             alpha_j * fac * (-1.0 + (al[0] * sinh_p + alpha_p * cosh_p) / alpha_p * exp_neg_p) +
             al[0] * j[0] * fac * (1.0 - (al[0] * cosh_p + alpha_p * sinh_p) / al[0] * exp_neg_p);
    for (int a = 1; a < 4; a++) {
      double term_1 = (s0[a] + al[a] * alpha_ss / alpha_sq * coshm1_p - s0[0] * al[a] / alpha_p * sinh_p) * exp_neg_i;
      double term_2 = j[a] * (1.0 - exp_neg_i) / al[0];
      double term_3 = alpha_j * al[a] / al[0] * fac *
                      (1.0 - (1.0 - al[0] * al[0] / alpha_sq - al[0] / alpha_sq * (al[0] * cosh_p + alpha_p * sinh_p)) * exp_neg_i);
      double term_4 = j[0] * al[a] / alpha_p * fac * (-alpha_p + (alpha_p * cosh_p + al[0] * sinh_p) * exp_neg_i);
      out[a] = term_1 + term_2 + term_3 + term_4;
    }
  } else {
    double alpha_j = al[1] * j[1] + al[3] * j[3];
    out[0] = (al[0] * j[0] - alpha_j) / (al[0] * al[0] - alpha_sq);
    for (int a = 1; a < 4; a++) out[a] = (j[a] - al[a] * out[0]) / al[0];
  }
}

// Pure Faraday rotation/conversion over dl (polarized.cpp:470-486, :598-612)
__device__ __forceinline__ void couple_rotate(const double s0[4], const double rho[4], double dl, double out[4]) {
  double rho_sq = rho[1] * rho[1] + rho[3] * rho[3];
  double rho_p = sqrt(rho_sq);
  double sin_rho, cos_rho;
  sincos(rho_p * dl, &sin_rho, &cos_rho);
  double sh = sin(rho_p * dl / 2.0);
  double sin_sq = sh * sh;
  double rho_ss = rho[1] * s0[1] + rho[3] * s0[3];
  out[0] = s0[0];
  out[1] = s0[1] * cos_rho + 2.0 * rho[1] * rho_ss / rho_sq * sin_sq - rho[3] * s0[2] / rho_p * sin_rho;
  out[2] = s0[2] * cos_rho + (rho[3] * s0[1] - rho[1] * s0[3]) / rho_p * sin_rho;
  out[3] = s0[3] * cos_rho + 2.0 * rho[3] * rho_ss / rho_sq * sin_sq + rho[1] * s0[2] / rho_p * sin_rho;
}

// One sample's coupling of the Stokes vector to the plasma (polarized.cpp:379-790).
__device__ __forceinline__ void couple(const RadParams &P, const Coefficients &C, double dl, double s[4]) {
  double j[4] = {C.j[0], C.j[1], 0.0, C.j[2]};
  double al[4] = {C.a[0], C.a[1], 0.0, C.a[2]};
  double rho[4] = {0.0, C.rho[0], 0.0, C.rho[1]};
  double delta_tau = al[0] * dl;
  bool thin = delta_tau <= 100.0;
  double alpha_sq = al[1] * al[1] + al[3] * al[3];
  double rho_sq = rho[1] * rho[1] + rho[3] * rho[3];
  // the reference tests rho_p = sqrt(rho_sq) against zero: the same as testing rho_sq (the square root of a positive
  // double is positive)
  const bool no_rotation = rho_sq == 0.0;
  double out[4] = {0.0, 0.0, 0.0, 0.0};
  if (P.rotation_split) {
    // Strang splitting: half absorb/emit, full rotate, half absorb/emit
    couple_absorb(s, j, al, dl / 2.0, delta_tau / 2.0, thin, out);
    admissible(out, true);
    for (int a = 0; a < 4; a++) s[a] = out[a];
    if (!no_rotation) couple_rotate(s, rho, dl, out);
    admissible(out, false);
    for (int a = 0; a < 4; a++) s[a] = out[a];
    couple_absorb(s, j, al, dl / 2.0, delta_tau / 2.0, thin, out);
  } else if (no_rotation) {
    couple_absorb(s, j, al, dl, delta_tau, thin, out);
  } else if (al[0] == 0.0) {
    couple_rotate(s, rho, dl, out);
    for (int a = 0; a < 4; a++) out[a] += j[a] * dl;
  } else {
    // general case: matrix exponential of the 4x4 coupling written with its eigenvalue pair
    // (lambda_1 real, lambda_2 imaginary part) -- polarized.cpp:656-779.  The reference assigns mm_2[1][2]
    // and mm_3[1][2] twice and never sets their [1][3]/[2][3]/[0][2] entries nor mm_4[0][1], [0][3], [1][2],
    // [2][3]; those entries stay zero here as well.
    double alpha_rho = al[1] * rho[1] + al[3] * rho[3];
    double d = alpha_sq - rho_sq;
    double lambda_a = sqrt(d * d / 4.0 + alpha_rho * alpha_rho);
    double lambda_b = d / 2.0;
    double lambda_1 = sqrt(lambda_a + lambda_b);
    double lambda_2 = sqrt(lambda_a - lambda_b);
    double theta = lambda_1 * lambda_1 + lambda_2 * lambda_2;
    double sg = alpha_rho >= 0.0 ? 1.0 : -1.0;
    // Of the sixteen entries of each matrix the reference fills six of mm_2 and mm_3 -- (0,1), (1,0), (0,3), (3,0),
    // (1,2), (2,1) -- and eight of mm_4 -- the diagonal, (0,2), (2,0), (1,3), (3,1); the identity mm_1 has the
    // diagonal.  Only those entries are evaluated: three kinds (diagonal, mm_2/mm_3 entry, off-diagonal mm_4
    // entry), each with the terms of polarized.cpp:735-779 that do not vanish identically.
    const double it = 1.0 / theta, it2 = 2.0 * it;   // 2 / theta, bit for bit: doubling is exact
    const double x01_2 = (lambda_2 * al[1] - sg * lambda_1 * rho[1]) * it;
    const double x03_2 = (lambda_2 * al[3] - sg * lambda_1 * rho[3]) * it;
    const double x12_2 = (sg * lambda_1 * al[1] + lambda_2 * rho[1]) * it;
    const double x01_3 = (lambda_1 * al[1] + sg * lambda_2 * rho[1]) * it;
    const double x03_3 = (lambda_1 * al[3] + sg * lambda_2 * rho[3]) * it;
    const double x12_3 = -(sg * lambda_2 * al[1] - lambda_1 * rho[1]) * it;
    const double half = (alpha_sq + rho_sq) / 2.0;
    const double d0 = half * it2, d1 = (al[1] * al[1] + rho[1] * rho[1] - half) * it2, d2 = -half * it2;
    const double d3 = (al[3] * al[3] + rho[3] * rho[3] - half) * it2;
    const double y02 = (al[1] * rho[3] - al[3] * rho[1]) * it2, y13 = (al[3] * al[1] + rho[3] * rho[1]) * it2;
    double ex = 0.0, sn = 0.0, cs = 0.0, snh = 0.0, csh = 0.0;
    if (thin) {
      ex = bfm::exp_bf(-delta_tau);
      sincos(lambda_2 * dl, &sn, &cs);
      bfm::sinhcosh_bf(lambda_1 * dl, snh, csh);
    }
    const double f_1 = 1.0 / (al[0] * al[0] - lambda_1 * lambda_1);
    const double f_2 = 1.0 / (al[0] * al[0] + lambda_2 * lambda_2);
    const double a1f = al[0] * f_1, a2f = al[0] * f_2, l1f = lambda_1 * f_1, l2f = lambda_2 * f_2;
    // Every entry of polarized.cpp:735-779 is linear in the entry's matrix elements with coefficients that depend only
    // on (alpha_I, lambda_1, lambda_2, dl): with mm_1 +- mm_4 = 2p, 2q and mm_2, mm_3 = x2, x3 the source part of an
    // entry is p Kp + q Kq + x3 L3 + x2 L2 and its propagator part p Gp + q Gq + x2 H2 + x3 H3.  The eight
    // coefficients are formed once per sample instead of inside each of the fourteen entries (thick steps: ex = 0
    // leaves the asymptotic values, exactly the reference's branch).
    const double Kp = a1f - ex * (a1f * csh + l1f * snh), Kq = a2f - ex * (a2f * cs - l2f * sn);
    const double L3 = ex * (l1f * csh + a1f * snh) - l1f, L2 = ex * (l2f * cs + a2f * sn) - l2f;
    const double Gp = ex * csh, Gq = ex * cs, H2 = -(ex * sn), H3 = -(ex * snh);
    const double Ks = 0.5 * (Kp + Kq), Kd = 0.5 * (Kp - Kq), Gs = 0.5 * (Gp + Gq), Gd = 0.5 * (Gp - Gq);
    // diagonal entries: mm_1 = 1, mm_4 = d_k, i.e. p, q = (1 +- d_k) / 2; off-diagonal mm_4 entries y: p, q = +-y / 2
    const double o01_j = x01_3 * L3 + x01_2 * L2, o01_s = x01_2 * H2 + x01_3 * H3;
    const double o03_j = x03_3 * L3 + x03_2 * L2, o03_s = x03_2 * H2 + x03_3 * H3;
    const double o12_j = x12_3 * L3 + x12_2 * L2, o12_s = x12_2 * H2 + x12_3 * H3;
    const double y02_j = y02 * Kd, y02_s = y02 * Gd, y13_j = y13 * Kd, y13_s = y13 * Gd;
    // column 2 multiplies j_U = 0: only the propagator part survives there
    out[0] = (Ks + d0 * Kd) * j[0] + (Gs + d0 * Gd) * s[0] + o01_j * j[1] + o01_s * s[1] + y02_s * s[2] + o03_j * j[3] + o03_s * s[3];
    out[1] = o01_j * j[0] + o01_s * s[0] + (Ks + d1 * Kd) * j[1] + (Gs + d1 * Gd) * s[1] + o12_s * s[2] + y13_j * j[3] + y13_s * s[3];
    out[2] = -(y02_j * j[0] + y02_s * s[0]) - (o12_j * j[1] + o12_s * s[1]) + (Gs + d2 * Gd) * s[2];
    out[3] = o03_j * j[0] + o03_s * s[0] + y13_j * j[1] + y13_s * s[1] + (Ks + d3 * Kd) * j[3] + (Gs + d3 * Gd) * s[3];
  }
  admissible(out, true);
  for (int a = 0; a < 4; a++) s[a] = out[a];
}

}  // namespace
