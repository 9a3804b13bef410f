// Fused sampling -> coefficients -> polarized (Stokes IQUV) transfer kernel.
//
// Reference: src/radiation_integrator/polarized.cpp:51-973 evolves, per frequency and per ray, a complex
// 4x4 coherency tensor N^{mu nu}: midpoint-rule parallel transport along the geodesic, projection onto a
// fluid-frame tetrad, analytic Stokes coupling, reconstruction of N from the Stokes vector.
// Because N is rebuilt from (I,Q,U,V) and the tetrad legs e_1, e_2 at every sample, and transport is linear,
// everything between two samples collapses to a frequency-INDEPENDENT 4x4 real "Stokes transport matrix"
//     S_start(n) = M(n-1 -> n) S_end(n-1),   M = blockdiag(3x3 on I,Q,U ; 1x1 on V)
// built from the transported legs (vectors, not tensors).  Per sample the thread therefore does the
// geometry once (Kerr-Schild jet, contracted connections, tetrad, M) and per frequency only the synchrotron
// coefficients and the 4x4 coupling -- the reference redoes all of it per frequency with 4x4x4 loops.
// Same one-ray-per-thread, warp-lock-step walk over the SoA step buffer as the unpolarized kernel.
#include "pol_common.cuh"

namespace {

// BI: inter-block interpolation and slow light compiled in (a separate instantiation keeps them out of the
// common kernel)
#ifndef BL_POL_MINB
#define BL_POL_MINB 2  // resident CTAs per SM the kernel is register-capped for
#endif
template <int FMAX, bool BI, int DIST>
__global__ void __launch_bounds__(kBlock, BL_POL_MINB)
radiate_polarized_kernel(const __grid_constant__ RadArgs A, const __grid_constant__ RadParams P) {
  extern __shared__ double smem_bounds[];
  const GridDev &G = A.grid;
  const double *bounds_s = nullptr;
  {
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
  double S[FMAX][4];
#pragma unroll
  for (int l = 0; l < FMAX; l++) S[l][0] = S[l][1] = S[l][2] = S[l][3] = 0.0;
  double *img = A.image + m;
  const int64_t stride = A.image_stride;
  const bool aux = P.image_time || P.image_length || P.image_lambda || P.image_emission || P.image_tau ||
                   P.image_lambda_ave || P.image_emission_ave || P.image_tau_int || P.image_crossings;
  const bool do_render = A.render != nullptr && P.render_num_images > 0;
  bool fill_present = false;
  if (do_render)
    for (int f = 0; f < P.render_feature_start[P.render_num_images]; f++)
      if (P.render_types[f] == 0) fill_present = true;
  if (valid)
    for (int q = 0; q < P.num_quantities; q++) img[(size_t)q * stride] = 0.0;
  if (valid && do_render)
    for (int q = 0; q < 3 * P.render_num_images; q++) A.render[m + (size_t)q * stride] = 0.0;

  double int_lambda[FMAX], int_emission[FMAX];
#pragma unroll
  for (int l = 0; l < FMAX; l++) int_lambda[l] = int_emission[l] = 0.0;
  const double k_t = valid ? A.cam_dir[4 * m] : 0.0;  // conserved covariant time component of the momentum
  bool plane_sign = false;
  int crossings = 0;
  if (valid && num > 0 && P.image_crossings) {
    const double *p0 = A.sb.buf + A.sb.at(num - 1, m);
    plane_sign = P.camera_x[1] * p0[1] + P.camera_x[2] * p0[2] + P.camera_x[3] * p0[3] > 0.0;
  }
  double prev_cv[RAD_NUM_CELL_VALUES];
  for (int q = 0; q < RAD_NUM_CELL_VALUES; q++) prev_cv[q] = nan("");
  rad::CellCache cache = {0, 0, 0, 0};
  rad::SlowLight slow = {0, {0.0, 0.0, 0.0, 0.0}};
  const double inv_mom_x = P.x_unit / mom;  // affine step -> cm per unit image frequency
  unsigned long long processed = 0;

  // frequency-independent state carried from the previous sample
  KsJet jet_p;
  double k_p[4] = {0, 0, 0, 0}, e_p[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  double dlam_p = 0.0;
  bool have_prev = false;

  for (int n = warp_max - 1; n >= 0; n--) {
    if (n >= num) continue;
    processed++;
    const double2 *src = reinterpret_cast<const double2 *>(A.sb.buf + A.sb.at(n, m));
    rad::prefetch_record(A.sb, n, m, A.prefetch);
    double2 r0 = __ldcs(src), r1 = __ldcs(src + 1), r2 = __ldcs(src + 2), r3 = __ldcs(src + 3);
    double t = r0.x, x = r0.y, y = r1.x, z = r1.y;
    double kc[4] = {k_t, r2.x, r2.y, r3.x};
    double dlam = -r3.y;
    double inv_r;
    double r = rad::ks_radius(P.a, x, y, z, inv_r);

    // ---- sample the plasma ----
    rad::SampleStatus st;
    rad::Prims pr;
    rad::SampleIndex si;
    pr.rho = pr.pgas = pr.kappa = pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = 0.0f;
    if (P.fallback_nan && flagged)
      st = rad::kSampleNan;
    else if (rad::geometric_cut(P, x, y, z, r))
      st = rad::kSampleCut;
    else
      st = rad::sample_grid<BI>(P, G, bounds_s, x, y, z, r, inv_r, t + P.snapshot_time, cache, pr, si, slow);
    if (st == rad::kSampleNan) {
      float qn = nanf("");
      pr.rho = pr.pgas = pr.kappa = pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = qn;
    } else if (st == rad::kSampleFallback) {
      pr.rho = P.fallback_rho; pr.pgas = P.fallback_pgas; pr.kappa = P.fallback_kappa;
      pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = 0.0f;
    }
    rad::Plasma ps;
    rad::plasma_state(P, x, y, z, r, inv_r, pr, 2, ps);
    bool coupled = st != rad::kSampleCut && !ps.value_cut && !ps.b_zero;
    double cv[RAD_NUM_CELL_VALUES];
    for (int q = 0; q < RAD_NUM_CELL_VALUES; q++) cv[q] = nan("");
    if (st != rad::kSampleCut && !ps.value_cut && P.need_cell_values) rad::cell_values_of(ps, cv);

    // ---- geometry: metric jet, momenta, fluid-frame tetrad ----
    KsJet jet;
    ks_jet(P, x, y, z, r, inv_r, jet);
    double kcon[4], ucov[4];
    raise(jet, kc, kcon);
    lower(jet, ps.ucon, ucov);
    double up[4] = {0.0, 0.0, 0.0, 1.0};
    if (!ps.b_zero)
      for (int mu = 0; mu < 4; mu++) up[mu] = ps.bcon[mu];
    double e1[4], e2[4], f1[4], f2[4];
    tetrad_legs(jet, ps.ucon, ucov, kcon, kc, up, e1, e2, f1, f2);

    // ---- Stokes transport matrix from the previous sample to this one ----
    StokesMap M;
    if (have_prev) transport_map(jet_p, k_p, e_p, dlam_p, jet, kcon, f1, f2, dlam, M);

    // pitch angle from invariants (see radiate_unpol.cu)
    double omega = -dot4(kc, ps.ucon);
    PolSample sq;
    if (coupled) {
      double kk[3] = {0.0, 0.0, 0.0};
      double kb = dot4(kc, ps.bcon);
      double c2 = kb * kb / (omega * omega * ps.b_sq);
      c2 = 1.0 < c2 ? 1.0 : c2;
      double sin_theta_b = sqrt(1.0 - c2);
      double cos_theta_b = sqrt(c2) * (kb >= 0.0 ? 1.0 : -1.0);
      if (has_thermal<DIST>(P) && ps.theta_e >= 0.01) {
        bfm::bessel_k01(ps.inv_theta_e, kk[0], kk[1]);
        kk[2] = kk[0] + 2.0 * ps.theta_e * kk[1];
      }
      pol_sample<DIST>(P, ps, omega * mom, sin_theta_b, cos_theta_b, kk, sq);
    }

    // ---- per-sample auxiliary quantities ----
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

    // ---- frequencies: rotate Stokes into the new frame, couple to the plasma ----
BL_FREQ_LOOP
    for (int l = 0; l < F; l++) {
      if (l >= F) break;
      double dl_cgs = dlam * inv_mom_x * P.inv_freqs[l];
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      if (have_prev) {
        s[0] = M.m[0][0] * S[l][0] + M.m[0][1] * S[l][1] + M.m[0][2] * S[l][2];
        s[1] = M.m[1][0] * S[l][0] + M.m[1][1] * S[l][1] + M.m[1][2] * S[l][2];
        s[2] = M.m[2][0] * S[l][0] + M.m[2][1] * S[l][1] + M.m[2][2] * S[l][2];
        s[3] = M.vv * S[l][3];
      }
      Coefficients C;
      for (int q = 0; q < 3; q++) C.j[q] = C.a[q] = 0.0;
      C.rho[0] = C.rho[1] = 0.0;
      if (coupled) synchrotron_polarized<DIST>(P, sq, l, C);
      double delta_tau = C.a[0] * dl_cgs;
      if (aux) {
        bool thin = delta_tau <= 100.0;
        if (P.image_lambda || P.image_lambda_ave) int_lambda[l] += dl_cgs;
        if (P.image_emission || P.image_emission_ave) int_emission[l] += C.j[0] * dl_cgs;
        if (P.image_tau) img[(size_t)(P.off_tau + l) * stride] += delta_tau;
        bool have_cv = !isnan(cv[0]);
        if (P.image_lambda_ave && have_cv)
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++)
            img[(size_t)(P.off_lambda_ave + l * RAD_NUM_CELL_VALUES + q) * stride] += cv[q] * dl_cgs;
        if (P.image_emission_ave && have_cv)
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++)
            img[(size_t)(P.off_emission_ave + l * RAD_NUM_CELL_VALUES + q) * stride] += cv[q] * C.j[0] * dl_cgs;
        if (P.image_tau_int && have_cv) {
          double en = thin ? exp(-delta_tau) : 0.0, em = thin ? expm1(delta_tau) : 0.0;
          for (int q = 0; q < RAD_NUM_CELL_VALUES; q++) {
            double *dst = img + (size_t)(P.off_tau_int + l * RAD_NUM_CELL_VALUES + q) * stride;
            *dst = thin ? en * (*dst + cv[q] * em) : cv[q];
          }
        }
      }
      couple(P, C, dl_cgs, s);
      S[l][0] = s[0]; S[l][1] = s[1]; S[l][2] = s[2]; S[l][3] = s[3];
    }

    // ---- carry state ----
    jet_p = jet;
    for (int mu = 0; mu < 4; mu++) {
      k_p[mu] = kcon[mu];
      e_p[0][mu] = e1[mu];
      e_p[1][mu] = e2[mu];
    }
    dlam_p = dlam;
    have_prev = true;
  }

  if (valid) {
    if (num > 0) {
      // last half step of transport, then projection on the camera tetrad (polarized.cpp:816-833, :875-939)
      const double *cp = A.cam_pos + 4 * m, *cd = A.cam_dir + 4 * m;
      KsJet jc;
      double inv_rc;
      double rc = rad::ks_radius(P.a, cp[1], cp[2], cp[3], inv_rc);
      ks_jet(P, cp[1], cp[2], cp[3], rc, inv_rc, jc);
      double kcov[4] = {cd[0], cd[1], cd[2], cd[3]}, kcon[4];
      raise(jc, kcov, kcon);
      const double *uc = P.camera_u_con, *ul = P.camera_u_cov, *vc = P.camera_vert_con_c;
      double up[4];
      up[0] = uc[0] * vc[0] - (ul[1] * vc[1] + ul[2] * vc[2] + ul[3] * vc[3]) / ul[0];
      up[1] = vc[1] + uc[1] * vc[0];
      up[2] = vc[2] + uc[2] * vc[0];
      up[3] = vc[3] + uc[3] * vc[0];
      double e1[4], e2[4], f1[4], f2[4];
      tetrad_legs(jc, uc, ul, kcon, kcov, up, e1, e2, f1, f2);
      StokesMap M;
      transport_map_final(jet_p, k_p, e_p, dlam_p, f1, f2, M);
      if (P.image_light)
BL_FREQ_LOOP
        for (int l = 0; l < F; l++) {
          if (l >= F) break;
          double f = P.freqs[l], nu_cu = f * f * f;
          img[(size_t)(4 * l + 0) * stride] = (M.m[0][0] * S[l][0] + M.m[0][1] * S[l][1] + M.m[0][2] * S[l][2]) * nu_cu;
          img[(size_t)(4 * l + 1) * stride] = (M.m[1][0] * S[l][0] + M.m[1][1] * S[l][1] + M.m[1][2] * S[l][2]) * nu_cu;
          img[(size_t)(4 * l + 2) * stride] = (M.m[2][0] * S[l][0] + M.m[2][1] * S[l][1] + M.m[2][2] * S[l][2]) * nu_cu;
          img[(size_t)(4 * l + 3) * stride] = M.vv * S[l][3] * nu_cu;
        }
      if (aux) {
BL_FREQ_LOOP
        for (int l = 0; l < F; l++) {
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
  }
  if (BI) rad::flush_slow_light(A.slow_counters, slow);
  if (A.sample_counter) {
    for (int off = 16; off > 0; off >>= 1) processed += __shfl_down_sync(full, processed, off);
    if ((threadIdx.x & 31) == 0 && processed) atomicAdd(A.sample_counter, processed);
  }
}

template <int FMAX>
cudaError_t launch_fmax(const RadArgs &A, const RadParams &P, cudaStream_t stream) {
  unsigned grid = (unsigned)((A.rays + kBlock - 1) / kBlock);
  size_t smem = 0;
  if ((size_t)A.grid.n_b * 6 * sizeof(double) <= 48 * 1024) smem = (size_t)A.grid.n_b * 6 * sizeof(double);
  const bool thermal_only = P.thermal_frac != 0.0 && P.power_frac == 0.0 && P.kappa_frac == 0.0;
  const bool kappa_only = P.kappa_frac != 0.0 && P.power_frac == 0.0 && P.thermal_frac == 0.0;
  if (P.block_interp || P.slow_light || P.coord == 2)
    radiate_polarized_kernel<FMAX, true, 7><<<grid, kBlock, smem, stream>>>(A, P);
  else if (thermal_only)
    radiate_polarized_kernel<FMAX, false, 1><<<grid, kBlock, smem, stream>>>(A, P);
  else if (kappa_only)
    radiate_polarized_kernel<FMAX, false, 4><<<grid, kBlock, smem, stream>>>(A, P);
  else
    radiate_polarized_kernel<FMAX, false, 7><<<grid, kBlock, smem, stream>>>(A, P);
  return cudaGetLastError();
}

}  // namespace

extern "C" cudaError_t bl_launch_radiate_polarized(const RadArgs *args, const RadParams *params, cudaStream_t stream) {
  if (args->rays <= 0) return cudaSuccess;
  return launch_fmax<RAD_MAX_FREQ>(*args, *params, stream);
}
