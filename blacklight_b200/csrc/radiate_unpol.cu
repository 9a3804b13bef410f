// Fused sampling -> coefficients -> unpolarized transfer (-> optional rendering) kernel.
//
// One ray per thread.  A warp walks its 32 rays' step buffers in lock step from the far end towards
// the camera (index n descending == the reference's source->camera order after ReverseGeodesics,
// geodesics.cpp:808-849), so every step-buffer load is a coalesced 256-byte row.  For each sample the
// thread locates the cell, gathers and interpolates the primitives, evaluates the synchrotron (or
// formula) coefficients for every frequency and advances the transfer solution -- the reference's
// sample_*, j_i, alpha_i and cell_values arrays (N x S x ...) are never materialised.
//
// Reference: simulation_sampling.cpp:122-1044, simulation_coefficients.cpp:254-701,
//            formula_coefficients.cpp:25-183, unpolarized.cpp:31-221, rendering.cpp:25-179.
#include "rad_sample.cuh"
#include "bf_math.cuh"

namespace {

constexpr int kBlock = 128;
#ifndef BL_RAD_MINB
#define BL_RAD_MINB 6  // resident blocks per SM the small-bucket kernels are register-capped for
#endif

// Frequency loops: fully unrolled with the per-frequency state in registers for the small buckets
// (FMAX <= 4); a rolled loop over state in thread-local memory for the large one (the per-frequency
// work is ~1e3 instructions, so the loop overhead is nothing and the code stays small).
#if defined(BL_FMAX) && BL_FMAX > 4
#define BL_FREQ_LOOP _Pragma("unroll 1")
#else
#define BL_FREQ_LOOP _Pragma("unroll")
#endif

// Frequency-independent part of the Stokes-I synchrotron coefficients of one sample
// (simulation_coefficients.cpp:458-524 thermal, :559-585 power law, :608-664 kappa; invariant forms
// j/nu^2, alpha*nu).  The reference evaluates every power with std::pow per frequency; here the logarithms
// of the per-sample quantities are taken once and each power is exp(c * ln x).
struct SynchSample {
  double om;          // omega * momentum_factor: nu = om * image_frequency
  double inv_om;      // 1 / om
  double n_nuc;       // n_e e^2 nu_c / c
  double inv_nu_s;    // thermal: 1 / (2/9 nu_c theta_e^2 sin(theta_B))
  double th_shape;    // thermal: thermal_frac * sqrt(2) pi / 27 * sin(theta_B) * n_nuc
  double h_kt;        // thermal: h / (k T_e)
  double log_om, log_ncs, log_ne;  // ln(om), ln(nu_c sin(theta_B)), ln(n_e)  [power law / kappa only]
  double nu_c, sin_theta_b, n_e;
};

// DIST: electron distributions compiled into an instantiation -- bit 0 thermal, bit 1 power law, bit 2 kappa; 7 = all,
// chosen at run time from the fractions.  The thermal-only instantiation of the light-only kernel drops the other
// distributions' code and live registers.
template <int DIST> __device__ __forceinline__ bool has_thermal(const RadParams &P) { return DIST == 7 ? P.thermal_frac != 0.0 : (DIST & 1) != 0; }
template <int DIST> __device__ __forceinline__ bool has_power(const RadParams &P) { return DIST == 7 ? P.power_frac != 0.0 : (DIST & 2) != 0; }
template <int DIST> __device__ __forceinline__ bool has_kappa(const RadParams &P) { return DIST == 7 ? P.kappa_frac != 0.0 : (DIST & 4) != 0; }

template <int DIST>
__device__ __forceinline__ void synch_sample(const RadParams &P, const rad::Plasma &s, double om,
                                             double sin_theta_b, SynchSample &q) {
  q.om = om;
  q.inv_om = 1.0 / om;
  q.nu_c = s.bb_cgs * (phys::e / (2.0 * phys::pi * phys::m_e * phys::c));
  q.sin_theta_b = sin_theta_b;
  q.n_e = s.n_e_cgs;
  q.n_nuc = s.n_e_cgs * q.nu_c * (phys::e * phys::e / phys::c);
  q.inv_nu_s = q.th_shape = q.h_kt = 0.0;
  if (has_thermal<DIST>(P)) {
    // 1/nu_s = 9/2 / (nu_c theta_e^2 sin(theta_B)), with 1/theta_e already known
    q.inv_nu_s = 4.5 * s.inv_theta_e * s.inv_theta_e / (q.nu_c * sin_theta_b);
    q.th_shape = P.thermal_frac * (phys::sqrt2 * phys::pi / 27.0) * sin_theta_b * q.n_nuc;
    q.h_kt = phys::h * s.inv_theta_e * (1.0 / (phys::m_e * phys::c * phys::c));
  }
  q.log_om = q.log_ncs = q.log_ne = 0.0;
  if (has_power<DIST>(P) || has_kappa<DIST>(P)) {
    q.log_om = log(om);
    q.log_ncs = log(q.nu_c * sin_theta_b);
    q.log_ne = log(s.n_e_cgs);
  }
}

// Coefficients at image frequency l.
template <int DIST>
__device__ __forceinline__ void synchrotron_unpolarized(const RadParams &P, const SynchSample &q, int l,
                                                        bool need_j, bool need_a, double &j_out, double &a_out) {
  double nu_cgs = q.om * P.freqs[l];
  double inv_nu = q.inv_om * P.inv_freqs[l];
  double inv_nu_2 = inv_nu * inv_nu;
  double j_val = 0.0, a_val = 0.0;
  if (has_thermal<DIST>(P)) {
    double xx = nu_cgs * q.inv_nu_s;
    double xx_1_2 = sqrt(xx);
    double xx_1_3 = cbrt(xx);
    double xx_1_6 = sqrt(xx_1_3);
    const double var_b = 1.8877486253633870;  // 2^(11/12)
    double var_c = xx_1_2 + var_b * xx_1_6;
    double j_th = q.th_shape * inv_nu_2 * bfm::exp_bf(-xx_1_3) * (var_c * var_c);
    if (need_j) j_val = j_th;
    if (need_a) {
      // Kirchhoff: alpha nu = j/nu^2 / (B_nu/nu^3), B_nu/nu^3 = 2h/c^2 / expm1(h nu / k T_e)
      a_val = j_th * expm1(q.h_kt * nu_cgs) * (phys::c * phys::c / (2.0 * phys::h));
      // absorptivities too small to square are flushed (simulation_coefficients.cpp:513-523:
      // 1/(a*a) == inf  <=>  a*a <= 2^-1024)
      if (a_val * a_val <= 0x1p-1024) a_val = 0.0;
    }
  }
  if (has_power<DIST>(P) || has_kappa<DIST>(P)) {
    double log_nu = q.log_om + P.log_freqs[l];
    double lr = log_nu - q.log_ncs;  // ln(nu / (nu_c sin(theta_B)))
    if (has_power<DIST>(P)) {
      if (need_j)
        j_val += P.power_frac * q.n_nuc * inv_nu_2 * P.power_jj * q.sin_theta_b * bfm::exp_bf(-(P.plasma_p - 1.0) / 2.0 * lr);
      if (need_a)
        a_val += P.power_frac * q.n_e * (phys::e * phys::e / (phys::m_e * phys::c)) * P.power_aa *
                 bfm::exp_bf(-(P.plasma_p + 2.0) / 2.0 * lr);
    }
    if (has_kappa<DIST>(P)) {
      double lx = lr - P.log_w2k2;  // ln(nu / nu_kappa)
      if (need_j) {
        // ln of kappa_frac n_e e^2 nu_c / (c nu^2) * sin(theta_B)
        double lva = P.log_k_j_pref + q.log_ne + q.log_ncs - 2.0 * log_nu;
        double l_lo = P.log_kjl + lva + lx * (1.0 / 3.0);
        double l_hi = P.log_kjh + lva - (P.plasma_kappa - 2.0) / 2.0 * lx;
        j_val += bfm::bridge(l_lo, l_hi, P.kappa_jj_x_i, 1.0 / P.kappa_jj_x_i);
      }
      if (need_a) {
        double lva = P.log_k_a_pref + q.log_ne;
        double l_lo = P.log_kal + lva - 2.0 / 3.0 * lx;
        double l_hi = P.log_kah + lva - (1.0 + P.plasma_kappa) / 2.0 * lx;
        a_val += bfm::bridge(l_lo, l_hi, P.kappa_aa_x_i, 1.0 / P.kappa_aa_x_i);
      }
    }
  }
  j_out = j_val;
  a_out = a_val;
}

// Analytic "formula" plasma of the 2020 ApJ 897 148 code comparison (formula_coefficients.cpp:121-179):
// fluid velocity from a Keplerian-like angular momentum profile in Boyer-Lindquist coordinates,
// Gaussian density, power-law emissivity / absorptivity.  Returns u^mu in CKS and n/n0.
__device__ __forceinline__ void formula_fluid(const RadParams &P, double x, double y, double z, double r,
                                              double ucon[4], double &n_n0) {
  const double a = P.a;
  double rr = sqrt(r * r - z * z);
  double cth = z / r;
  double sth = sqrt(1.0 - cth * cth);
  double ph = atan2(y, x) - atan(a / r);
  double sph, cph;
  sincos(ph, &sph, &cph);
  double delta = r * r - 2.0 * r + a * a;
  double sigma = r * r + a * a * cth * cth;
  double gtt = -(1.0 + 2.0 * r * (r * r + a * a) / (delta * sigma));
  double gtph = -2.0 * a * r / (delta * sigma);
  double gphph = (sigma - 2.0 * r) / (delta * sigma * sth * sth);
  double ll = P.formula_l0 / (1.0 + rr) * pow(rr, 1.0 + P.formula_q);
  double u_norm = 1.0 / sqrt(-gtt + 2.0 * gtph * ll - gphph * ll * ll);
  double u_t = -u_norm, u_ph = u_norm * ll;
  double ut = gtt * u_t + gtph * u_ph;
  double uph = gtph * u_t + gphph * u_ph;
  // u^r = u^theta = 0 in Boyer-Lindquist, so the KS and CKS transformations reduce to the phi column
  ucon[0] = ut;
  ucon[1] = sth * (-r * sph - a * cph) * uph;
  ucon[2] = sth * (r * cph - a * sph) * uph;
  ucon[3] = 0.0;
  n_n0 = exp(-0.5 * (r * r / (P.formula_r0 * P.formula_r0) + P.formula_h * P.formula_h * cth * cth));
}

// LEAN: only the light image is requested (no auxiliary images, no rendering, no inter-block interpolation)
// -- the common case gets a kernel without the dead register state of the rest.
template <int FMAX, bool SIM, bool LEAN, int DIST>
__global__ void __launch_bounds__(kBlock, (FMAX <= 4 ? BL_RAD_MINB : 2))
radiate_unpolarized_kernel(const __grid_constant__ RadArgs A, const __grid_constant__ RadParams P) {
  extern __shared__ double smem_bounds[];
  const GridDev &G = A.grid;
  const double *bounds_s = nullptr;
  if (SIM) {
    // stage block bounds in shared memory when they fit (6 doubles per mesh block)
    int nb6 = G.n_b * 6;
    if ((size_t)nb6 * sizeof(double) <= 48 * 1024) {
      for (int t = threadIdx.x; t < nb6; t += blockDim.x) smem_bounds[t] = G.bounds[t];
      __syncthreads();
      bounds_s = smem_bounds;
    }
  }
  const unsigned full = 0xffffffffu;
  // thread i takes ray order[i] of the wave's list sorted by length (ray_order.cu): a warp's rays end together
  const int64_t i_list = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i_list < A.active;
  int64_t m = valid ? (A.order ? (int64_t)A.order[i_list] : i_list) : 0;
  int num = valid ? A.sample_num[m] : 0;
  bool flagged = valid ? A.sample_flags[m] != 0 : false;
  double mom = valid ? A.mom_factor[m] : 1.0;
  int warp_max = num;
  for (int off = 16; off > 0; off >>= 1) {
    int o = __shfl_xor_sync(full, warp_max, off);
    warp_max = o > warp_max ? o : warp_max;
  }

  const int F = P.num_freq;
  double I[FMAX];
#pragma unroll
  for (int l = 0; l < FMAX; l++) I[l] = 0.0;
  double *img = A.image + m;  // quantity q of this ray at img[q * stride]
  const int64_t stride = A.image_stride;
  const bool aux = !LEAN && (P.image_time || P.image_length || P.image_lambda || P.image_emission || P.image_tau ||
                             P.image_lambda_ave || P.image_emission_ave || P.image_tau_int || P.image_crossings);
  const bool need_j = P.image_light || P.image_emission || P.image_emission_ave;
  const bool need_a = P.image_light || P.image_tau || P.image_tau_int;
  const bool want_coeff = P.image_light || P.image_emission || P.image_tau || P.image_emission_ave || P.image_tau_int;
  const bool do_render = !LEAN && SIM && A.render != nullptr && P.render_num_images > 0;
  bool fill_present = false;
  if (do_render)
    for (int f = 0; f < P.render_feature_start[P.render_num_images]; f++)
      if (P.render_types[f] == 0) fill_present = true;

  // zero the auxiliary slots this thread owns (image[] is accumulated in place for them)
  if (valid && aux)
    for (int q = P.image_light ? F : 0; q < P.num_quantities; q++) img[(size_t)q * stride] = 0.0;
  if (valid && do_render)
    for (int q = 0; q < 3 * P.render_num_images; q++) A.render[m + (size_t)q * stride] = 0.0;

  double int_lambda[FMAX], int_emission[FMAX];
#pragma unroll
  for (int l = 0; l < FMAX; l++) int_lambda[l] = int_emission[l] = 0.0;
  bool plane_sign = false;
  int crossings = 0;
  if (valid && num > 0 && P.image_crossings) {
    // reference looks at sample 0 of the reversed array == our last stored sample
    const double *p0 = A.sb.buf + A.sb.at(num - 1, m);
    plane_sign = P.camera_x[1] * p0[1] + P.camera_x[2] * p0[2] + P.camera_x[3] * p0[3] > 0.0;
  }
  double prev_cv[RAD_NUM_CELL_VALUES];
  for (int q = 0; q < RAD_NUM_CELL_VALUES; q++) prev_cv[q] = nan("");
  rad::CellCache cache = {0, 0, 0, 0};
  rad::SlowLight slow = {0, {0.0, 0.0, 0.0, 0.0}};
  const double inv_mom_x = P.x_unit / mom;  // affine step -> cm per unit image frequency
  unsigned long long processed = 0;
  const double k_t = valid ? A.cam_dir[4 * m] : 0.0;  // conserved covariant time component of the momentum

  for (int n = warp_max - 1; n >= 0; n--) {
    if (n >= num) continue;
    processed++;
    const double2 *src = reinterpret_cast<const double2 *>(A.sb.buf + A.sb.at(n, m));
    rad::prefetch_record(A.sb, n, m, A.prefetch);
    double2 r0 = __ldcs(src), r1 = __ldcs(src + 1), r2 = __ldcs(src + 2), r3 = __ldcs(src + 3);
    double t = r0.x, x = r0.y, y = r1.x, z = r1.y;
    double kc[4] = {k_t, r2.x, r2.y, r3.x};
    double dlam = -r3.y;
    int n_ref = num - 1 - n;  // index in the reference's reversed arrays (taps only)

    double inv_r;
    double r = rad::ks_radius(P.a, x, y, z, inv_r);
    double omega = 0.0;      // -k_mu u^mu
    SynchSample sq;
    bool coupled = false;    // coefficients are nonzero candidates
    bool nan_sample = false;
    rad::Plasma ps;
    double cv[RAD_NUM_CELL_VALUES];
    for (int q = 0; q < RAD_NUM_CELL_VALUES; q++) cv[q] = nan("");
    double n_n0 = 0.0;

    if (SIM) {
      rad::SampleStatus st;
      rad::Prims pr;
      rad::SampleIndex si;
      si.b = si.k = si.j = si.i = -1;
      si.fk = si.fj = si.fi = 0.0;
      if (P.fallback_nan && flagged)
        st = rad::kSampleNan;
      else if (rad::geometric_cut(P, x, y, z, r))
        st = rad::kSampleCut;
      else
        st = rad::sample_grid<!LEAN>(P, G, bounds_s, x, y, z, r, inv_r, t + P.snapshot_time, cache, pr, si, slow);
      if (A.taps.nan_) {
        size_t ti = (size_t)m * A.taps.S + n_ref;
        A.taps.nan_[ti] = st == rad::kSampleNan;
        A.taps.cut[ti] = st == rad::kSampleCut;
        A.taps.fallback[ti] = st == rad::kSampleFallback;
        if (A.taps.inds) {
          A.taps.inds[4 * ti + 0] = si.b; A.taps.inds[4 * ti + 1] = si.k;
          A.taps.inds[4 * ti + 2] = si.j; A.taps.inds[4 * ti + 3] = si.i;
        }
        if (A.taps.fracs) {
          A.taps.fracs[3 * ti + 0] = si.fk; A.taps.fracs[3 * ti + 1] = si.fj; A.taps.fracs[3 * ti + 2] = si.fi;
        }
      }
      if (st == rad::kSampleNan) {
        float qn = nanf("");
        pr.rho = pr.pgas = pr.kappa = pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = qn;
      } else if (st == rad::kSampleFallback) {
        pr.rho = P.fallback_rho; pr.pgas = P.fallback_pgas; pr.kappa = P.fallback_kappa;
        pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = 0.0f;
      }
      if (st != rad::kSampleCut) {
        rad::plasma_state(P, x, y, z, r, inv_r, pr, want_coeff ? 1 : 0, ps);
        if (!ps.value_cut) {
          if (!LEAN && P.need_cell_values) rad::cell_values_of(ps, cv);
          if (want_coeff && !ps.b_zero) {
            omega = -(kc[0] * ps.ucon[0] + kc[1] * ps.ucon[1] + kc[2] * ps.ucon[2] + kc[3] * ps.ucon[3]);
            double kb = kc[0] * ps.bcon[0] + kc[1] * ps.bcon[1] + kc[2] * ps.bcon[2] + kc[3] * ps.bcon[3];
            // fluid-frame pitch angle: |k_spatial| = omega and |b| = sqrt(b^2) in the frame of u
            double c2 = kb * kb / (omega * omega * ps.b_sq);
            c2 = 1.0 < c2 ? 1.0 : c2;
            synch_sample<DIST>(P, ps, omega * mom, sqrt(1.0 - c2), sq);
            coupled = true;
            nan_sample = st == rad::kSampleNan;
          }
        }
      }
    } else {
      // formula model: flagged rays are NaN in frequency slot 0 only (formula_coefficients.cpp:51-59)
      if (P.fallback_nan && flagged) {
        nan_sample = true;
        coupled = true;
      } else if (!rad::geometric_cut(P, x, y, z, r)) {
        double ucon[4];
        formula_fluid(P, x, y, z, r, ucon, n_n0);
        omega = -(ucon[0] * kc[0] + ucon[1] * kc[1] + ucon[2] * kc[2] + ucon[3] * kc[3]);
        coupled = true;
      }
    }
    (void)nan_sample;

    // per-sample auxiliary quantities that do not depend on frequency
    if (aux) {
      if (P.image_time) {
        double t_cgs = t * P.t_unit;
        double cur = img[(size_t)P.off_time * stride];
        img[(size_t)P.off_time * stride] = t_cgs < cur ? t_cgs : cur;
      }
      if (P.image_length)
        img[(size_t)P.off_length * stride] += rad::proper_length_rate(P, x, y, z, kc) * dlam * P.x_unit;
      if (P.image_crossings) {
        bool sign_new = P.camera_x[1] * x + P.camera_x[2] * y + P.camera_x[3] * z > 0.0;
        if (sign_new != plane_sign) crossings++;
        plane_sign = sign_new;
      }
    }
    if (do_render) {
      double dlen = fill_present ? rad::proper_length_rate(P, x, y, z, kc) * dlam * P.x_unit : 0.0;
      rad::render_update(P, A.render + m, stride, prev_cv, cv, dlen);
      for (int q = 0; q < RAD_NUM_CELL_VALUES; q++) prev_cv[q] = cv[q];
    }

    // frequencies
BL_FREQ_LOOP
    for (int l = 0; l < (FMAX > 4 ? F : FMAX); l++) {
      if (l >= F) break;
      double dlam_cgs = dlam * inv_mom_x * P.inv_freqs[l];
      double j = 0.0, alpha = 0.0;
      if (coupled) {
        if (SIM) {
          synchrotron_unpolarized<DIST>(P, sq, l, need_j, need_a, j, alpha);
        } else if (P.fallback_nan && flagged) {
          if (l == 0) j = alpha = nan("");
        } else {
          double nu = omega * P.freqs[l] * mom;
          double jn = P.formula_cn0 * n_n0 * pow(nu / P.formula_nup, -P.formula_alpha);
          j = jn / (nu * nu);
          double an = P.formula_a * P.formula_cn0 * n_n0 * pow(nu / P.formula_nup, -P.formula_beta - P.formula_alpha);
          alpha = an * nu;
        }
      }
      if (!need_j) j = nan("");
      if (!need_a) alpha = nan("");
      double delta_tau = alpha * dlam_cgs;
      bool thin = delta_tau <= 100.0;
      double exp_neg = 0.0, em1 = 0.0;
      if (LEAN) {
        // I <- e^-dtau (I + S expm1(dtau)) = I + (S - I)(1 - e^-dtau)   (unpolarized.cpp:99-110)
        if (alpha > 0.0) {
          double ss = j / alpha;
          I[l] = thin ? I[l] - (ss - I[l]) * expm1(-delta_tau) : ss;
        } else {
          I[l] += j * dlam_cgs;
        }
      } else {
        if (alpha > 0.0 || P.image_tau_int) {
          exp_neg = exp(-delta_tau);
          em1 = expm1(delta_tau);
        }
        if (P.image_light) {
          if (alpha > 0.0) {
            double ss = j / alpha;
            I[l] = thin ? exp_neg * (I[l] + ss * em1) : ss;
          } else {
            I[l] += j * dlam_cgs;
          }
        }
      }
      if (aux) {
        if (P.image_lambda || P.image_lambda_ave) int_lambda[l] += dlam_cgs;
        if (P.image_emission || P.image_emission_ave) int_emission[l] += j * dlam_cgs;
        if (P.image_tau) img[(size_t)(P.off_tau + l) * stride] += delta_tau;
        bool have_cv = SIM && !isnan(cv[0]);
        if (P.image_lambda_ave && have_cv)
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++)
            img[(size_t)(P.off_lambda_ave + l * RAD_NUM_CELL_VALUES + q) * stride] += cv[q] * dlam_cgs;
        if (P.image_emission_ave && have_cv)
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++)
            img[(size_t)(P.off_emission_ave + l * RAD_NUM_CELL_VALUES + q) * stride] += cv[q] * j * dlam_cgs;
        if (P.image_tau_int && have_cv)
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++) {
            double *dst = img + (size_t)(P.off_tau_int + l * RAD_NUM_CELL_VALUES + q) * stride;
            *dst = thin ? exp_neg * (*dst + cv[q] * em1) : cv[q];
          }
      }
    }
  }

  if (valid) {
    if (P.image_light)
BL_FREQ_LOOP
      for (int l = 0; l < (FMAX > 4 ? F : FMAX); l++) {
        if (l >= F) break;
        double f = P.freqs[l];
        img[(size_t)l * stride] = I[l] * (f * f * f);
      }
    if (aux) {
BL_FREQ_LOOP
      for (int l = 0; l < (FMAX > 4 ? F : FMAX); l++) {
        if (l >= F) break;
        if (P.image_lambda) img[(size_t)(P.off_lambda + l) * stride] = int_lambda[l];
        if (P.image_emission) img[(size_t)(P.off_emission + l) * stride] = int_emission[l];
        if (P.image_lambda_ave)
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++)
            img[(size_t)(P.off_lambda_ave + l * RAD_NUM_CELL_VALUES + q) * stride] /= int_lambda[l];
        if (P.image_emission_ave)
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++)
            img[(size_t)(P.off_emission_ave + l * RAD_NUM_CELL_VALUES + q) * stride] /= int_emission[l];
      }
      if (P.image_crossings) img[(size_t)P.off_crossings * stride] = (double)crossings;
    }
  }
  if (!LEAN && SIM) rad::flush_slow_light(A.slow_counters, slow);
  if (A.sample_counter) {
    for (int off = 16; off > 0; off >>= 1) processed += __shfl_down_sync(full, processed, off);
    if ((threadIdx.x & 31) == 0 && processed) atomicAdd(A.sample_counter, processed);
  }
}

template <int FMAX>
cudaError_t launch_fmax(const RadArgs &A, const RadParams &P, cudaStream_t stream) {
  const bool sim = P.model_type == 0;
  const bool lean = P.image_light && !(P.image_time || P.image_length || P.image_lambda || P.image_emission ||
                                       P.image_tau || P.image_lambda_ave || P.image_emission_ave || P.image_tau_int ||
                                       P.image_crossings) &&
                    !(sim && A.render != nullptr && P.render_num_images > 0) && !(sim && (P.block_interp || P.slow_light || P.coord == 2));
  unsigned grid = (unsigned)((A.rays + kBlock - 1) / kBlock);
  size_t smem = 0;
  if (sim && (size_t)A.grid.n_b * 6 * sizeof(double) <= 48 * 1024) smem = (size_t)A.grid.n_b * 6 * sizeof(double);
  const bool thermal_only = P.thermal_frac != 0.0 && P.power_frac == 0.0 && P.kappa_frac == 0.0;
  if (sim && lean && thermal_only)
    radiate_unpolarized_kernel<FMAX, true, true, 1><<<grid, kBlock, smem, stream>>>(A, P);
  else if (sim && lean)
    radiate_unpolarized_kernel<FMAX, true, true, 7><<<grid, kBlock, smem, stream>>>(A, P);
  else if (sim)
    radiate_unpolarized_kernel<FMAX, true, false, 7><<<grid, kBlock, smem, stream>>>(A, P);
  else if (lean)
    radiate_unpolarized_kernel<FMAX, false, true, 7><<<grid, kBlock, 0, stream>>>(A, P);
  else
    radiate_unpolarized_kernel<FMAX, false, false, 7><<<grid, kBlock, 0, stream>>>(A, P);
  return cudaGetLastError();
}

}  // namespace

// One translation unit per frequency-count bucket (BL_FMAX = 1, 4, 32; see the Makefile) so that the
// buckets compile in parallel.
#ifndef BL_FMAX
#define BL_FMAX 1
#endif
#define BL_CAT2(a, b) a##b
#define BL_CAT(a, b) BL_CAT2(a, b)
extern "C" cudaError_t BL_CAT(bl_launch_radiate_unpolarized_f, BL_FMAX)(const RadArgs *args, const RadParams *params,
                                                                        cudaStream_t stream) {
  if (args->rays <= 0) return cudaSuccess;
  return launch_fmax<BL_FMAX>(*args, *params, stream);
}
