// Polarized (Stokes IQUV) transfer as a pipeline of four kernels over slabs of the step buffer.
//
// The fused kernel of radiate_pol.cu walks a ray with everything in one thread: geometry (Kerr-Schild jet,
// tetrad, Stokes transport matrix), per-frequency synchrotron coefficients and the Stokes coupling.  Its
// register state (255 + 2 KB of stack) allows two warps per scheduler, and ncu shows it waiting on
// dependent FP64 latencies most of the time.  Only the last few dozen operations per sample and frequency
// are really sequential along the ray (s <- M s; s <- O(coefficients) s + c), so the work is split by what
// it depends on:
//   pol_sampling_kernel      one thread per ray, a slab of `slab` consecutive samples: sampling of the grid and
//                            plasma state; writes the fluid frame (u^mu, the tetrad's "up" vector: 8 doubles)
//                            and the 7 scalars the coefficient stage needs.  Carried state: the cell hint.
//   pol_geometry_kernel      one thread per ray over the same slab: metric jet, tetrad legs, and the
//                            frequency-independent transport matrix M from the previous sample
//                            (polarized.cpp:136-292, :816-833); the only carried state is the previous sample's
//                            frame, re-derived from one halo sample at the top of the slab.  Writes 11 doubles.
//                            (Round 2 started with these two as one kernel: 168 registers, 12 warps per SM, and
//                            more than half of its time in the sampling code's gathers.)
//   pol_coefficient_kernel   one thread per (ray, sample): the eight polarized synchrotron coefficients of
//                            every frequency (simulation_coefficients.cpp:458-698).  No carried state at all,
//                            so it runs at whatever occupancy its registers allow.  Writes 8 F doubles.
//   pol_transfer_kernel      one thread per (ray, frequency): s <- M s, Stokes coupling (polarized.cpp:379-790),
//                            sequential over the slab; the Stokes state between slabs lives in the image
//                            itself.  The last slab applies the half step to the camera and the projection on
//                            the camera tetrad (:816-939).
// Slabs run from the far end of the rays (large n) to the camera (n = 0), four launches each, on one stream.
// The scratch between the stages is field-major, scratch[(field * slab + j) * rays + i]: every access of a warp
// is a contiguous 256-byte row.  i is the ray's position in the wave's list of rays sorted by length (ray_order.cu;
// the ray itself without a list): a slab is launched over the rays still alive in it, a prefix of that list.
#include <cstdio>
#include <cstdlib>

#include "pol_common.cuh"

namespace {

enum : int {
  kFieldM = 0,        // 9 entries of the 3x3 block of M, then vv
  kFieldDlam = 10,
  kFieldOm = 11,      // omega * momentum factor (nu = om * image frequency); 0 marks an uncoupled sample
  kFieldSin = 12,
  kFieldCos = 13,
  kFieldBb = 14,      // |B| in gauss
  kFieldNe = 15,      // electron number density
  kFieldTheta = 16,
  kFieldInvTheta = 17,
  kFieldCoef = 18     // 8 per frequency: j_I, j_Q, j_V, alpha_I, alpha_Q, alpha_V, rho_Q, rho_V
};

struct SplitArgs {
  double *frame;       // (8, slab + 1, rays): u^mu and the tetrad's "up" vector of every sample of the slab + its halo
  double *scratch;     // (18 + 8 F, slab, rays)
  double *cam_map;     // (10, rays): the half step to the camera, filled by the slab that holds n = 0
  int32_t slab;        // samples per slab
  int32_t n_lo, n_hi;  // this launch covers samples n_lo <= n < n_hi of every ray
};

__device__ __forceinline__ double *field_ptr(const SplitArgs &X, int64_t rays, int field, int j, int64_t i) {
  return X.scratch + ((size_t)field * X.slab + j) * (size_t)rays + i;
}

__device__ __forceinline__ double *frame_ptr(const SplitArgs &X, int64_t rays, int field, int j, int64_t i) {
  return X.frame + ((size_t)field * (X.slab + 1) + j) * (size_t)rays + i;
}

// ---- stage 1a: sampling -------------------------------------------------------------------------------------
// The plasma at every sample of the slab: grid sampling (cell search from the previous sample's cell, trilinear
// gather), plasma state, and from it what the later stages need -- the fluid four-velocity and the "up" vector of the
// tetrad (the magnetic field) for the frame stage, the seven scalars of the coefficient stage.  Half of the old
// single geometry kernel's time was spent here, waiting on gathers with the 12 warps per SM its transport-matrix
// arithmetic left; on its own the sampling code needs 88 registers and runs with 20.  When the slab does not start at
// the ray's far end, the sample before it (the frame stage's halo) is sampled again into the extra slot j = n_hi - n_lo.
template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
pol_sampling_kernel(const __grid_constant__ RadArgs A, const __grid_constant__ RadParams P, const SplitArgs X) {
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
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // position in the list: addresses the scratch
  const bool valid = i < A.active;
  const int64_t m = valid ? (A.order ? (int64_t)A.order[i] : i) : 0;   // the ray
  const int num = valid ? A.sample_num[m] : 0;
  unsigned long long processed = 0;
  if (num > X.n_lo) {
    const bool flagged = A.sample_flags[m] != 0;
    const double mom = A.mom_factor[m];
    const double k_t = A.cam_dir[4 * m];
    const int top = (X.n_hi < num ? X.n_hi : num) - 1;   // first sample of this slab in walking order
    const bool halo = X.n_hi < num;                        // the sample before it belongs to the previous slab
    rad::CellCache cache = {0, 0, 0, 0};
    rad::SlowLight slow = {0, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll 1
    for (int n = halo ? top + 1 : top; n >= X.n_lo; n--) {
      const bool store = n <= top;
      const int j = n - X.n_lo;
      const double2 *src = reinterpret_cast<const double2 *>(A.sb.buf + A.sb.at(n, m));
      rad::prefetch_record(A.sb, n, m, A.prefetch);
      double2 r0 = __ldcs(src), r1 = __ldcs(src + 1), r2 = __ldcs(src + 2), r3 = __ldcs(src + 3);
      double t = r0.x, x = r0.y, y = r1.x, z = r1.y;
      double kc[4] = {k_t, r2.x, r2.y, r3.x};
      double inv_r;
      double r = rad::ks_radius(P.a, x, y, z, inv_r);

      rad::SampleStatus st;
      rad::Prims pr;
      rad::SampleIndex si;
      pr.rho = pr.pgas = pr.kappa = pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = 0.0f;
      if (P.fallback_nan && flagged)
        st = rad::kSampleNan;
      else if (rad::geometric_cut(P, x, y, z, r))
        st = rad::kSampleCut;
      else
        st = rad::sample_grid<false>(P, G, bounds_s, x, y, z, r, inv_r, t + P.snapshot_time, cache, pr, si, slow);
      if (st == rad::kSampleNan) {
        float qn = nanf("");
        pr.rho = pr.pgas = pr.kappa = pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = qn;
      } else if (st == rad::kSampleFallback) {
        pr.rho = P.fallback_rho; pr.pgas = P.fallback_pgas; pr.kappa = P.fallback_kappa;
        pr.uu1 = pr.uu2 = pr.uu3 = pr.bb1 = pr.bb2 = pr.bb3 = 0.0f;
      }
      rad::Plasma ps;
      rad::plasma_state(P, x, y, z, r, inv_r, pr, 2, ps);
      const bool coupled = st != rad::kSampleCut && !ps.value_cut && !ps.b_zero;

      // ---- the fluid frame for the frame stage ----
      for (int mu = 0; mu < 4; mu++) {
        __stcs(frame_ptr(X, A.rays, mu, j, i), ps.ucon[mu]);
        __stcs(frame_ptr(X, A.rays, 4 + mu, j, i), ps.b_zero ? (mu == 3 ? 1.0 : 0.0) : ps.bcon[mu]);
      }
      if (store) {
        processed++;
        // ---- what the coefficient stage needs: frequency scale, pitch angle, field strength, n_e, theta_e ----
        double om = 0.0, sin_theta_b = 0.0, cos_theta_b = 0.0;
        if (coupled) {
          double omega = -dot4(kc, ps.ucon);
          double kb = dot4(kc, ps.bcon);
          double c2 = kb * kb / (omega * omega * ps.b_sq);
          c2 = 1.0 < c2 ? 1.0 : c2;
          sin_theta_b = sqrt(1.0 - c2);
          cos_theta_b = sqrt(c2) * (kb >= 0.0 ? 1.0 : -1.0);
          om = omega * mom;
        }
        __stcs(field_ptr(X, A.rays, kFieldOm, j, i), om);
        if (coupled) {
          __stcs(field_ptr(X, A.rays, kFieldSin, j, i), sin_theta_b);
          __stcs(field_ptr(X, A.rays, kFieldCos, j, i), cos_theta_b);
          __stcs(field_ptr(X, A.rays, kFieldBb, j, i), ps.bb_cgs);
          __stcs(field_ptr(X, A.rays, kFieldNe, j, i), ps.n_e_cgs);
          __stcs(field_ptr(X, A.rays, kFieldTheta, j, i), ps.theta_e);
          __stcs(field_ptr(X, A.rays, kFieldInvTheta, j, i), ps.inv_theta_e);
        }
      }
    }
  }
  if (A.sample_counter) {
    for (int off = 16; off > 0; off >>= 1) processed += __shfl_down_sync(full, processed, off);
    if ((threadIdx.x & 31) == 0 && processed) atomicAdd(A.sample_counter, processed);
  }
}

// ---- stage 1b: frames and the Stokes transport matrix ------------------------------------------------------------
// Metric jet, photon momentum, fluid-frame tetrad at every sample (from the sampling stage's u^mu and "up"), and the
// frequency-independent transport matrix M from the previous sample (polarized.cpp:136-292, :816-833).  The carried
// state is the previous sample's frame, re-derived at a slab's top from the halo sample.
template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
pol_geometry_kernel(const __grid_constant__ RadArgs A, const __grid_constant__ RadParams P, const SplitArgs X) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.active) return;
  const int64_t m = A.order ? (int64_t)A.order[i] : i;
  const int num = A.sample_num[m];
  if (num <= X.n_lo) return;
  const double k_t = A.cam_dir[4 * m];
  const int top = (X.n_hi < num ? X.n_hi : num) - 1;
  const bool halo = X.n_hi < num;
  KsJet jet_p;
  double k_p[4] = {0, 0, 0, 0}, e_p[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  double dlam_p = 0.0;
  bool have_prev = false;

#pragma unroll 1
  for (int n = halo ? top + 1 : top; n >= X.n_lo; n--) {
    const bool store = n <= top;
    const int j = n - X.n_lo;
    const double2 *src = reinterpret_cast<const double2 *>(A.sb.buf + A.sb.at(n, m));
    rad::prefetch_record(A.sb, n, m, A.prefetch);
    double2 r0 = __ldcs(src), r1 = __ldcs(src + 1), r2 = __ldcs(src + 2), r3 = __ldcs(src + 3);
    double ucon[4], up[4];
    for (int mu = 0; mu < 4; mu++) {
      ucon[mu] = __ldcs(frame_ptr(X, A.rays, mu, j, i));
      up[mu] = __ldcs(frame_ptr(X, A.rays, 4 + mu, j, i));
    }
    double x = r0.y, y = r1.x, z = r1.y;
    double kc[4] = {k_t, r2.x, r2.y, r3.x};
    double dlam = -r3.y;
    double inv_r;
    double r = rad::ks_radius(P.a, x, y, z, inv_r);

    // ---- metric jet, momenta, fluid-frame tetrad ----
    KsJet jet;
    ks_jet(P, x, y, z, r, inv_r, jet);
    double kcon[4], ucov[4];
    raise(jet, kc, kcon);
    lower(jet, ucon, ucov);
    double e1[4], e2[4], f1[4], f2[4];
    tetrad_legs(jet, ucon, ucov, kcon, kc, up, e1, e2, f1, f2);

    if (store) {
      // ---- Stokes transport matrix from the previous sample to this one ----
      StokesMap M;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) M.m[a][b] = 0.0;
      M.vv = 0.0;
      if (have_prev) transport_map(jet_p, k_p, e_p, dlam_p, jet, kcon, f1, f2, dlam, M);
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) __stcs(field_ptr(X, A.rays, kFieldM + 3 * a + b, j, i), M.m[a][b]);
      __stcs(field_ptr(X, A.rays, kFieldM + 9, j, i), M.vv);
      __stcs(field_ptr(X, A.rays, kFieldDlam, j, i), dlam);
    }

    // ---- carry the frame to the next sample ----
    jet_p = jet;
    for (int mu = 0; mu < 4; mu++) {
      k_p[mu] = kcon[mu];
      e_p[0][mu] = e1[mu];
      e_p[1][mu] = e2[mu];
    }
    dlam_p = dlam;
    have_prev = true;
  }

  if (X.n_lo == 0) {
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
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) X.cam_map[(size_t)(3 * a + b) * A.rays + m] = M.m[a][b];
    X.cam_map[(size_t)9 * A.rays + m] = M.vv;
  }
}

// ---- stage 2: coefficients ----------------------------------------------------------------------------------
template <int DIST, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
pol_coefficient_kernel(const __grid_constant__ RadArgs A, const __grid_constant__ RadParams P, const SplitArgs X) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= A.active) return;
  const int64_t m = A.order ? (int64_t)A.order[i] : i;
  const int n = X.n_lo + j;
  if (n >= X.n_hi || n >= A.sample_num[m]) return;
  const int F = P.num_freq;
  const double om = __ldcs(field_ptr(X, A.rays, kFieldOm, j, i));
  if (om == 0.0) {
    for (int q = 0; q < 8 * F; q++) __stcs(field_ptr(X, A.rays, kFieldCoef + q, j, i), 0.0);
    return;
  }
  rad::Plasma s;
  const double sin_b = __ldcs(field_ptr(X, A.rays, kFieldSin, j, i));
  const double cos_b = __ldcs(field_ptr(X, A.rays, kFieldCos, j, i));
  s.bb_cgs = __ldcs(field_ptr(X, A.rays, kFieldBb, j, i));
  s.n_e_cgs = __ldcs(field_ptr(X, A.rays, kFieldNe, j, i));
  s.theta_e = __ldcs(field_ptr(X, A.rays, kFieldTheta, j, i));
  s.inv_theta_e = __ldcs(field_ptr(X, A.rays, kFieldInvTheta, j, i));
  double kk[3] = {0.0, 0.0, 0.0};
  if (has_thermal<DIST>(P) && s.theta_e >= 0.01) {
    bfm::bessel_k01(s.inv_theta_e, kk[0], kk[1]);
    kk[2] = kk[0] + 2.0 * s.theta_e * kk[1];
  }
  PolSample sq;
  pol_sample<DIST>(P, s, om, sin_b, cos_b, kk, sq);
  const size_t fs = (size_t)X.slab * (size_t)A.rays;   // distance between consecutive fields
  if constexpr (DIST == 4) {
    // kappa only: term-major over the frequencies, values go straight to the scratch
    double *dst = field_ptr(X, A.rays, kFieldCoef, j, i);
    kappa_polarized_all(P, sq, F, [&](int k, int l, double v) { __stcs(dst + (size_t)(8 * l + k) * fs, v); });
  } else {
BL_FREQ_LOOP
    for (int l = 0; l < F; l++) {
      Coefficients C;
      synchrotron_polarized<DIST>(P, sq, l, C);
      double *dst = field_ptr(X, A.rays, kFieldCoef + 8 * l, j, i);
      __stcs(dst, C.j[0]); __stcs(dst + fs, C.j[1]); __stcs(dst + 2 * fs, C.j[2]);
      __stcs(dst + 3 * fs, C.a[0]); __stcs(dst + 4 * fs, C.a[1]); __stcs(dst + 5 * fs, C.a[2]);
      __stcs(dst + 6 * fs, C.rho[0]); __stcs(dst + 7 * fs, C.rho[1]);
    }
  }
}

// ---- stage 3: transfer --------------------------------------------------------------------------------------
// FW: frequencies per CTA (1, 2 or 4); the CTA's 128 threads are 128/FW adjacent rays x FW frequencies, so that a
// warp is 32 adjacent rays at one frequency and the FW warps of a ray group share M through L1.
template <int FW, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
pol_transfer_kernel(const __grid_constant__ RadArgs A, const __grid_constant__ RadParams P, const SplitArgs X) {
  constexpr int kRays = kBlock / FW;
  const int64_t i = (int64_t)blockIdx.x * kRays + (threadIdx.x % kRays);
  const int l = blockIdx.y * FW + threadIdx.x / kRays;
  if (i >= A.active || l >= P.num_freq) return;
  const int64_t m = A.order ? (int64_t)A.order[i] : i;
  const int num = A.sample_num[m];
  if (num <= X.n_lo) return;
  const int top = (X.n_hi < num ? X.n_hi : num) - 1;
  const int64_t stride = A.image_stride;
  double *img = A.image + m;
  // the Stokes state between slabs lives in the image's own slots (zeroed before the first slab)
  double s[4];
  for (int a = 0; a < 4; a++) s[a] = img[(size_t)(4 * l + a) * stride];
  const double dl_factor = P.x_unit / A.mom_factor[m] * P.inv_freqs[l];
  const size_t fs = (size_t)X.slab * (size_t)A.rays;
  double tau = 0.0, lam = 0.0, emi = 0.0;
  if (P.image_tau) tau = img[(size_t)(P.off_tau + l) * stride];
  if (P.image_lambda) lam = img[(size_t)(P.off_lambda + l) * stride];
  if (P.image_emission) emi = img[(size_t)(P.off_emission + l) * stride];

#pragma unroll 1
  for (int n = top; n >= X.n_lo; n--) {
    const int j = n - X.n_lo;
    const double *lk = field_ptr(X, A.rays, kFieldM, j, i);
    const double *cf = field_ptr(X, A.rays, kFieldCoef + 8 * l, j, i);
    double mm[10];
#pragma unroll
    for (int q = 0; q < 10; q++) mm[q] = __ldg(lk + q * fs);
    const double dlam = __ldg(lk + 10 * fs);
    Coefficients C;
    C.j[0] = __ldcs(cf); C.j[1] = __ldcs(cf + fs); C.j[2] = __ldcs(cf + 2 * fs);
    C.a[0] = __ldcs(cf + 3 * fs); C.a[1] = __ldcs(cf + 4 * fs); C.a[2] = __ldcs(cf + 5 * fs);
    C.rho[0] = __ldcs(cf + 6 * fs); C.rho[1] = __ldcs(cf + 7 * fs);
    const double dl_cgs = dlam * dl_factor;
    double t0 = mm[0] * s[0] + mm[1] * s[1] + mm[2] * s[2];
    double t1 = mm[3] * s[0] + mm[4] * s[1] + mm[5] * s[2];
    double t2 = mm[6] * s[0] + mm[7] * s[1] + mm[8] * s[2];
    double t3 = mm[9] * s[3];
    s[0] = t0; s[1] = t1; s[2] = t2; s[3] = t3;
    if (P.image_tau) tau += C.a[0] * dl_cgs;
    if (P.image_lambda) lam += dl_cgs;
    if (P.image_emission) emi += C.j[0] * dl_cgs;
    couple(P, C, dl_cgs, s);
  }

  if (X.n_lo == 0) {
    double mm[10];
    for (int q = 0; q < 10; q++) mm[q] = X.cam_map[(size_t)q * A.rays + m];
    double f = P.freqs[l], nu_cu = f * f * f;
    img[(size_t)(4 * l + 0) * stride] = (mm[0] * s[0] + mm[1] * s[1] + mm[2] * s[2]) * nu_cu;
    img[(size_t)(4 * l + 1) * stride] = (mm[3] * s[0] + mm[4] * s[1] + mm[5] * s[2]) * nu_cu;
    img[(size_t)(4 * l + 2) * stride] = (mm[6] * s[0] + mm[7] * s[1] + mm[8] * s[2]) * nu_cu;
    img[(size_t)(4 * l + 3) * stride] = mm[9] * s[3] * nu_cu;
  } else {
    for (int a = 0; a < 4; a++) img[(size_t)(4 * l + a) * stride] = s[a];
  }
  if (P.image_tau) img[(size_t)(P.off_tau + l) * stride] = tau;
  if (P.image_lambda) img[(size_t)(P.off_lambda + l) * stride] = lam;
  if (P.image_emission) img[(size_t)(P.off_emission + l) * stride] = emi;
}

template <int MINB>
void launch_coefficients(int dist, dim3 grid, cudaStream_t stream, const RadArgs &A, const RadParams &P, const SplitArgs &X) {
  if (dist == 1) pol_coefficient_kernel<1, MINB><<<grid, kBlock, 0, stream>>>(A, P, X);
  else if (dist == 4) pol_coefficient_kernel<4, MINB><<<grid, kBlock, 0, stream>>>(A, P, X);
  else pol_coefficient_kernel<7, MINB><<<grid, kBlock, 0, stream>>>(A, P, X);
}

template <int MINB>
void launch_transfer(int fw, dim3 grid, cudaStream_t stream, const RadArgs &A, const RadParams &P, const SplitArgs &X) {
  if (fw == 4) pol_transfer_kernel<4, MINB><<<grid, kBlock, 0, stream>>>(A, P, X);
  else if (fw == 2) pol_transfer_kernel<2, MINB><<<grid, kBlock, 0, stream>>>(A, P, X);
  else pol_transfer_kernel<1, MINB><<<grid, kBlock, 0, stream>>>(A, P, X);
}

// Resident CTAs per SM each stage is register-capped for.  The defaults are the measured best on B200; the
// environment variables exist for re-tuning (BL_POL_OCC="g,c,t").
struct Occupancy { int g, c, t, s; };
Occupancy stage_occupancy() {
  static Occupancy occ = [] {
    Occupancy o = {3, 0, 5, 5};   // coefficient stage: 0 = by electron distribution
    if (const char *e = getenv("BL_POL_OCC")) sscanf(e, "%d,%d,%d,%d", &o.g, &o.c, &o.t, &o.s);
    return o;
  }();
  return occ;
}

}  // namespace

extern "C" int bl_polarized_split_fields(int num_freq) { return kFieldCoef + 8 * num_freq; }

// One pass of the three stages over all slabs of a wave: samples [0, s_top) of args->rays rays.  scratch holds
// bl_polarized_split_fields(F) * slab * rays doubles, cam_map 10 * rays; the wave's image columns must be zero.
// frame holds 8 * (slab + 1) * rays doubles.  events (or nullptr): 4 * slabs + 1 events recorded on `stream` around
// every launch (sampling, geometry, coefficients, transfer), so that the caller can attribute device time to the stages
// after synchronising; *launches is advanced by the kernels launched.
extern "C" int bl_polarized_split_slabs(int slab, int s_top) { return s_top <= 0 ? 0 : (s_top + slab - 1) / slab; }

extern "C" cudaError_t bl_launch_radiate_polarized_split(const RadArgs *args, const RadParams *params, double *scratch,
                                                         double *frame, double *cam_map, int slab, int s_top, cudaStream_t stream,
                                                         cudaEvent_t *events, long long *launches,
                                                         const int64_t *alive, int num_alive) {
  RadArgs A = *args;
  const RadParams &P = *params;
  if (A.rays <= 0 || s_top <= 0) return cudaSuccess;
  size_t smem = 0;
  if ((size_t)A.grid.n_b * 6 * sizeof(double) <= 48 * 1024) smem = (size_t)A.grid.n_b * 6 * sizeof(double);
  const bool thermal_only = P.thermal_frac != 0.0 && P.power_frac == 0.0 && P.kappa_frac == 0.0;
  const bool kappa_only = P.kappa_frac != 0.0 && P.power_frac == 0.0 && P.thermal_frac == 0.0;
  const int dist = thermal_only ? 1 : (kappa_only ? 4 : 7);
  const Occupancy occ = stage_occupancy();
  const int F = P.num_freq;
  const int fw = F >= 4 ? 4 : (F >= 2 ? 2 : 1);
  int ev = 0;
  if (events) cudaEventRecord(events[ev++], stream);
  for (int n_hi = (s_top + slab - 1) / slab * slab; n_hi > 0; n_hi -= slab) {
    SplitArgs X;
    X.scratch = scratch; X.frame = frame; X.cam_map = cam_map; X.slab = slab;
    X.n_lo = n_hi - slab; X.n_hi = n_hi < s_top ? n_hi : s_top;
    // the rays alive in this slab: with the sorted list a prefix of it (alive[s]: rays longer than s slabs)
    const int slab_index = X.n_lo / slab;
    if (alive && A.order) A.active = slab_index < num_alive ? alive[slab_index] : 0;
    const unsigned ray_blocks = (unsigned)((A.active + kBlock - 1) / kBlock);
    if (ray_blocks == 0) {
      if (events) for (int q = 0; q < 4; q++) cudaEventRecord(events[ev++], stream);
      continue;
    }
    if (occ.s == 4) pol_sampling_kernel<4><<<ray_blocks, kBlock, smem, stream>>>(A, P, X);
    else if (occ.s == 6) pol_sampling_kernel<6><<<ray_blocks, kBlock, smem, stream>>>(A, P, X);
    else pol_sampling_kernel<5><<<ray_blocks, kBlock, smem, stream>>>(A, P, X);
    if (events) cudaEventRecord(events[ev++], stream);
    if (occ.g == 2) pol_geometry_kernel<2><<<ray_blocks, kBlock, 0, stream>>>(A, P, X);
    else if (occ.g == 4) pol_geometry_kernel<4><<<ray_blocks, kBlock, 0, stream>>>(A, P, X);
    else if (occ.g == 5) pol_geometry_kernel<5><<<ray_blocks, kBlock, 0, stream>>>(A, P, X);
    else pol_geometry_kernel<3><<<ray_blocks, kBlock, 0, stream>>>(A, P, X);
    if (events) cudaEventRecord(events[ev++], stream);
    dim3 cgrid(ray_blocks, (unsigned)(X.n_hi - X.n_lo));
    // the thermal-only coefficient code is light enough for five CTAs per SM (58 -> 51 ms per 1024^2 frame), the
    // term-major kappa code for six (155 -> 141 -> 134 ms at four, five, six)
    const int occ_c = occ.c > 0 ? occ.c : (dist == 1 ? 5 : (dist == 4 ? 6 : 4));
    if (occ_c == 3) launch_coefficients<3>(dist, cgrid, stream, A, P, X);
    else if (occ_c == 5) launch_coefficients<5>(dist, cgrid, stream, A, P, X);
    else if (occ_c == 6) launch_coefficients<6>(dist, cgrid, stream, A, P, X);
    else if (occ_c == 7) launch_coefficients<7>(dist, cgrid, stream, A, P, X);
    else if (occ_c == 8) launch_coefficients<8>(dist, cgrid, stream, A, P, X);
    else launch_coefficients<4>(dist, cgrid, stream, A, P, X);
    if (events) cudaEventRecord(events[ev++], stream);
    dim3 tgrid((unsigned)((A.active + kBlock / fw - 1) / (kBlock / fw)), (unsigned)((F + fw - 1) / fw));
    if (occ.t == 3) launch_transfer<3>(fw, tgrid, stream, A, P, X);
    else if (occ.t == 4) launch_transfer<4>(fw, tgrid, stream, A, P, X);
    else if (occ.t == 6) launch_transfer<6>(fw, tgrid, stream, A, P, X);
    else launch_transfer<5>(fw, tgrid, stream, A, P, X);
    if (events) cudaEventRecord(events[ev++], stream);
    if (launches) *launches += 4;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
