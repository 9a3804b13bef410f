// C-ABI glue: context, HBM residency, wave scheduling and the parity taps.
// See include/blacklight_b200.h for the contract of each entry point.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/blacklight_b200.h"
#include "rad_types.cuh"

extern "C" cudaError_t bl_launch_geodesic_dp(const GeoArgs *args, int flat, int sm_count, int min_blocks, cudaStream_t stream);
extern "C" cudaError_t bl_launch_geodesic_rk(const GeoArgs *args, int flat, int order, int sm_count, cudaStream_t stream);
#define BL_DECL_RAD(name) extern "C" cudaError_t name(const RadArgs *args, const RadParams *params, cudaStream_t stream)
BL_DECL_RAD(bl_launch_radiate_unpolarized_f1); BL_DECL_RAD(bl_launch_radiate_unpolarized_f4);
BL_DECL_RAD(bl_launch_radiate_unpolarized_f32);
BL_DECL_RAD(bl_launch_radiate_polarized);
extern "C" int bl_polarized_split_fields(int num_freq);
extern "C" int bl_polarized_split_slabs(int slab, int s_top);
extern "C" cudaError_t bl_launch_radiate_polarized_split(const RadArgs *args, const RadParams *params, double *scratch,
                                                         double *frame, double *cam_map, int slab, int s_top, cudaStream_t stream,
                                                         cudaEvent_t *events, long long *launches,
                                                         const int64_t *alive, int num_alive);
extern "C" cudaError_t bl_launch_relayout_grid(const float *prim, int n_var, const int *var_index, size_t cells,
                                               float4 *out, float *kappa_out, cudaStream_t stream);
extern "C" cudaError_t bl_launch_unpack_samples(const StepBuffer *sb, const int32_t *num, const double *cam_dir,
                                                int64_t rays, int S, double *pos, double *dir, double *len,
                                                cudaStream_t stream);
extern "C" cudaError_t bl_launch_pack_samples(const StepBuffer *sb, const int32_t *num, int64_t ray0, int64_t count, int S,
                                              const double *pos, const double *dir, const double *len,
                                              cudaStream_t stream);
extern "C" cudaError_t bl_launch_refine(const double *image, int64_t stride, int level, const int32_t *block_locs,
                                        int64_t num_blocks, const bl_params *params_dev, uint8_t *flags,
                                        cudaStream_t stream);
extern "C" cudaError_t bl_launch_camera_pixels(const CameraDev *cam, int kind, const int32_t *units, int eff_res, int block_size,
                                               int64_t num_pixels, double *cam_pos, double *cam_dir, double *mom_factor,
                                               cudaStream_t stream);
extern "C" int bl_ray_order_max_buckets(void);
extern "C" size_t bl_ray_order_workspace(int64_t rays, int buckets);
extern "C" cudaError_t bl_launch_ray_order(const int32_t *num, int64_t rays, int unit, int buckets, int32_t *workspace,
                                           int32_t *order, cudaStream_t stream);
extern "C" cudaError_t bl_launch_impact_order(const double *cam_pos, const double *cam_dir, int64_t rays, int32_t *keys,
                                              unsigned int *b_max_bits, int buckets, int32_t *workspace, int32_t *order,
                                              cudaStream_t stream);
extern "C" cudaError_t bl_launch_fp64_peak(double *out, int blocks, int iters, cudaStream_t stream);
extern "C" cudaError_t bl_launch_division_selftest(unsigned long long seed, int blocks, int iters,
                                                   unsigned long long *mismatches, cudaStream_t stream);

namespace {

constexpr int kImpactBuckets = 130;   // impact-parameter buckets of the integrator's queue order

struct Level {
  int64_t rays = 0;
  double *cam_pos = nullptr, *cam_dir = nullptr, *mom = nullptr;  // device (rays,4),(rays,4),(rays)
  int32_t *num = nullptr;     // device (rays)
  uint8_t *flags = nullptr;   // device (rays)
  double *step = nullptr;     // device step buffer of one wave
  int64_t wave_rays = 0;      // rays per wave
  bool resident = false;      // whole level traced and kept in `step`
  bool traced = false;
  double *image = nullptr;    // device (Q, rays)
  double *render = nullptr;   // device (R,3,rays)
  // polarized pipeline (radiate_pol_split.cu): slab scratch and the camera half-step map of one wave
  // rays of one wave sorted by length, longest first (ray_order.cu): the radiation kernels take their rays from it
  int32_t *order = nullptr;        // device (wave_rays)
  int32_t *order_ws = nullptr;     // device workspace of the sort; its tail holds the bucket totals
  int32_t order_unit = 0, order_buckets = 0;   // bucket = ceil(num / unit); 0 buckets: rays are taken in index order
  int64_t order_alive0 = -1;       // rays with at least one sample in the last radiated wave (-1: no list was used)
  double *scratch = nullptr;  // device (fields, slab, wave_rays)
  double *frame = nullptr;    // device (8, slab + 1, wave_rays): fluid frame between the sampling and the geometry stage
  double *cam_map = nullptr;  // device (10, wave_rays)
  int32_t slab = 0;           // samples per slab; 0 = the level uses the fused kernel
  double ms_stage[4] = {0.0, 0.0, 0.0, 0.0};  // sampling, geometry, coefficients, transfer: device time of the last radiate call
  bl_level_stats stats{};
  bl_slow_stats slow{};
  // taps (allocated on demand)
  int32_t *tap_inds = nullptr; double *tap_fracs = nullptr; uint8_t *tap_nan = nullptr, *tap_cut = nullptr, *tap_fb = nullptr;
  int32_t tap_S = 0;
};

}  // namespace

struct bl_ctx {
  bl_params params;
  RadParams rad;            // passed by value (constant bank) with every radiation launch
  bl_params *params_dev = nullptr;
  GridDev grid{};
  std::vector<void *> grid_allocs;
  bool have_grid = false;
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  GeoCounters *counters = nullptr;          // device
  unsigned long long *rad_counter = nullptr;  // device
  unsigned long long *slow_counters = nullptr;  // device, 8 words (see RadArgs::slow_counters)
  std::vector<Level> levels;
  std::string error;
  bool taps_enabled = false;
  long long launches = 0;   // kernels of ours launched so far
  int geo_min_blocks = 0;   // occupancy variant of the DP kernel: 0 = by rays per thread (trace_wave); BL_GEO_BLOCKS overrides
  std::vector<cudaEvent_t> stage_events;   // per-launch events of the polarized pipeline
  bool have_camera = false; // bl_set_camera was called
  CameraDev camera;         // device-side camera description (camera_kernel.cu)
  int32_t *units_dev = nullptr;   // unit list (rows / block locations) of the last bl_trace_level_pixels
  size_t units_cap = 0;           // its capacity in int32
  int32_t *refine_locs = nullptr; // bl_refine_level: block locations and flags on the device, kept between calls
  uint8_t *refine_flags = nullptr;
  size_t refine_cap = 0;          // blocks they hold
  int rad_prefetch = 2;     // BL_RAD_PREFETCH: samples ahead the radiation kernels prefetch step-buffer records into L2 (0 = off)
  int pol_slab = 0;         // BL_POL_SLAB: samples per slab of that pipeline (0 = chosen from the HBM budget)
  int geo_order = 1;        // BL_GEO_ORDER=0: the integrator's queue hands out the rays in index order
  int ray_order = 1;        // BL_RAY_ORDER=0: radiation kernels take the rays in index order (A/B comparisons)
  bool pol_fused = false;   // BL_POL_FUSED=1: keep the single fused polarized kernel (A/B comparisons, parity cross-check)
};

namespace {

std::string g_create_error;

int bl_fail(bl_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->error = buf; else g_create_error = buf;
  return code;
}

int bl_fail_cuda(bl_ctx *ctx, cudaError_t err, const char *what, const char *file, int line) {
  return bl_fail(ctx, BL_ERR_CUDA, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(err), cudaGetErrorString(err),
                 file, line, what);
}

template <typename T>
cudaError_t dev_alloc(T **p, size_t count) {
  return cudaMalloc((void **)p, count * sizeof(T) > 0 ? count * sizeof(T) : 1);
}

void free_level(Level &L) {
  cudaFree(L.cam_pos); cudaFree(L.cam_dir); cudaFree(L.mom); cudaFree(L.num); cudaFree(L.flags);
  cudaFree(L.step); cudaFree(L.image); cudaFree(L.render); cudaFree(L.scratch); cudaFree(L.frame); cudaFree(L.cam_map);
  cudaFree(L.order); cudaFree(L.order_ws);
  cudaFree(L.tap_inds); cudaFree(L.tap_fracs); cudaFree(L.tap_nan); cudaFree(L.tap_cut); cudaFree(L.tap_fb);
  L = Level();
}

// 2F1 via the Pfaff transformation and a 10-term series (reference simulation_coefficients.cpp:740-773)
double hypergeometric(double alpha, double beta, double gamma, double z) {
  double a = alpha, b = gamma - beta, c = gamma, x = z / (z - 1.0);
  double result = 1.0, a_k = 1.0, b_k = 1.0, c_k = 1.0, xk = 1.0, kf = 1.0;
  for (int k = 1; k <= 10; k++) {
    a_k *= a + k - 1.0; b_k *= b + k - 1.0; c_k *= c + k - 1.0; xk *= x; kf *= k;
    result += a_k * b_k * xk / (c_k * kf);
  }
  return result * std::pow(1.0 - z, -alpha);
}

// Slot layout of image[level] (reference radiation_integrator.cpp:436-520)
void image_layout(const bl_params &p, RadParams &r) {
  int q = 0;
  bool pol = p.model_type == BL_MODEL_SIMULATION && p.image_polarization;
  int F = p.image_num_frequencies;
  if (p.image_light) q += F * (pol ? 4 : 1);
  r.off_time = q; if (p.image_time) q += 1;
  r.off_length = q; if (p.image_length) q += 1;
  r.off_lambda = q; if (p.image_lambda) q += F;
  r.off_emission = q; if (p.image_emission) q += F;
  r.off_tau = q; if (p.image_tau) q += F;
  r.off_lambda_ave = q; if (p.image_lambda_ave) q += F * BL_NUM_CELL_VALUES;
  r.off_emission_ave = q; if (p.image_emission_ave) q += F * BL_NUM_CELL_VALUES;
  r.off_tau_int = q; if (p.image_tau_int) q += F * BL_NUM_CELL_VALUES;
  r.off_crossings = q; if (p.image_crossings) q += 1;
  r.num_quantities = q;
}

// Distribution-function constants (reference simulation_coefficients.cpp:53-193)
void plasma_constants(const bl_params &p, RadParams &r) {
  const double pi = phys::pi;
  bool pol = p.image_light && p.image_polarization;
  if (p.plasma_power_frac != 0.0) {
    double pp = p.plasma_p;
    double var_a = std::pow(3.0, pp / 2.0) * (pp - 1.0);
    double var_b = 2.0 * (pp + 1.0);
    double var_c = std::pow(p.plasma_gamma_min, 1.0 - pp) - std::pow(p.plasma_gamma_max, 1.0 - pp);
    double var_d = std::tgamma((3.0 * pp - 1.0) / 12.0);
    double var_e = std::tgamma((3.0 * pp + 19.0) / 12.0);
    double var_f = std::pow(3.0, (pp + 1.0) / 2.0) * (pp - 1.0) / 4.0;
    double var_g = std::tgamma((3.0 * pp + 2.0) / 12.0);
    double var_h = std::tgamma((3.0 * pp + 22.0) / 12.0);
    r.power_jj = var_a / var_b / var_c * var_d * var_e;
    r.power_aa = var_f / var_c * var_g * var_h;
    if (pol) {
      double var_i = 2.0 * (pp + 2.0) / (pp + 1.0);
      double var_j = std::pow(p.plasma_gamma_min, -(pp + 1.0));
      double var_k = std::log(p.plasma_gamma_min);
      r.power_jj_q = -(pp + 1.0) / (pp + 7.0 / 3.0);
      r.power_jj_v = 0.684 * std::pow(pp, 0.49);
      r.power_aa_q = -std::pow(0.034 * pp - 0.0344, 0.086);
      r.power_aa_v = std::pow(0.71 * pp + 0.0352, 0.394);
      r.power_rho = (pp - 1.0) / var_c;
      r.power_rho_q = -std::pow(p.plasma_gamma_min, 2.0 - pp) / (pp / 2.0 - 1.0);
      r.power_rho_v = var_i * var_j * var_k;
    }
  }
  if (p.plasma_kappa_frac != 0.0) {
    double kk = p.plasma_kappa, w = p.plasma_w;
    double var_a = 4.0 * pi * std::tgamma(kk - 4.0 / 3.0);
    double var_b = std::pow(3.0, 7.0 / 3.0) * std::tgamma(kk - 2.0);
    double var_c = std::pow(3.0, (kk - 1.0) / 2.0);
    double var_d = (kk - 2.0) * (kk - 1.0) / 4.0;
    double var_e = std::tgamma(kk / 4.0 - 1.0 / 3.0);
    double var_f = std::tgamma(kk / 4.0 + 4.0 / 3.0);
    double var_g = std::pow(3.0, 1.0 / 6.0) * 10.0 / 41.0;
    double var_h = w * kk;
    double var_i = 2.0 * pi * std::pow(var_h, kk - 10.0 / 3.0);
    double var_j = (kk - 2.0) * (kk - 1.0) * kk;
    double var_k = 3.0 * kk - 1.0;
    double var_l = std::tgamma(5.0 / 3.0);
    double var_m = hypergeometric(kk - 1.0 / 3.0, kk + 1.0, kk + 2.0 / 3.0, -var_h);
    double var_n = std::pow(pi, 1.5) / 3.0;
    double var_o = var_j / (var_h * var_h * var_h);
    double var_p = 2.0 * std::tgamma(2.0 + kk / 2.0) / (2.0 + kk) - 1.0;
    r.kappa_jj_low = var_a / var_b;
    r.kappa_jj_high = var_c * var_d * var_e * var_f;
    r.kappa_jj_x_i = 3.0 * std::pow(kk, -1.5);
    r.kappa_aa_low = var_g * var_i * var_j / var_k * var_l * var_m;
    r.kappa_aa_high = var_n * var_o * var_p;
    r.kappa_aa_x_i = std::pow(-1.75 + 1.6 * kk, -0.86);
    // Stokes-I absorptivity blends with this factor in every mode (simulation_coefficients.cpp:652), but the
    // reference only assigns it for polarized runs (:121); in an unpolarized run it reads the never-written
    // member of a freshly allocated object, which is zero in practice, and the kappa absorptivity bridges to
    // (lo^-x + 0^-x)^(-1/x) = 0.  Reproduced, not fixed (verified against the reference
    // binary by tests/test_gpu_parity.py::test_live_reference_unpolarized).
    r.kappa_aa_high_i = pol ? std::pow(3.0 / kk, 4.75) + 0.6 : 0.0;
    if (pol) {
      double var_q = 14.3 * std::pow(w, -0.928);
      double var_r = 169.0 * std::pow(kk, -8.0) + 0.0052 * kk - 0.0526 + 47.0 / (200.0 * kk);
      r.kappa_jj_low_q = 0.5;
      r.kappa_jj_low_v = 0.5625 * std::pow(kk, -0.528) / w;
      r.kappa_jj_high_q = 0.64 + 0.02 * kk;
      r.kappa_jj_high_v = 0.765625 * std::pow(kk, -0.44) / w;
      r.kappa_jj_x_q = 3.7 * std::pow(kk, -1.6);
      r.kappa_jj_x_v = r.kappa_jj_x_i;
      r.kappa_aa_low_q = 25.0 / 48.0;
      r.kappa_aa_low_v = 77.0 / (100.0 * w) * std::pow(kk, -0.7);
      r.kappa_aa_high_q = 441.0 * std::pow(kk, -5.76) + 0.55;
      r.kappa_aa_high_v = var_q * var_r;
      r.kappa_aa_x_q = 1.4 * std::pow(kk, -1.15);
      r.kappa_aa_x_v = 1.22 * std::pow(kk, -1.136) + 0.007;
      r.kappa_rho_v = std::cyl_bessel_k(0.0, 1.0 / w) / std::cyl_bessel_k(2.0, 1.0 / w);
      // Faraday-rotation fits are tabulated at kappa = 3.5, 4, 4.5, 5 and blended linearly in between
      struct Fit { double qa, qb, qc, qd, qe, va, vb; };
      double sw = std::sqrt(w), e5 = std::exp(-5.0 * w);
      Fit f35 = {17.0 * w + sw * (-3.0 + 7.0 * e5), -1.0 / 30.0, 0.1, -1.5, 0.471,
                 (w * w + 2.0 * w + 1.0) / (3.125 * w * w + 4.0 * w + 1.0), 0.447};
      Fit f40 = {46.0 / 3.0 * w + sw * (-5.0 / 3.0 + 17.0 / 3.0 * e5), -1.0 / 18.0, 1.0 / 6.0, -1.75, 0.5,
                 (w * w + 54.0 * w + 50.0) / (30.0 / 11.0 * w * w + 134.0 * w + 50.0), 0.391};
      Fit f45 = {14.0 * w + sw * (-1.625 + 4.5 * e5), -1.0 / 12.0, 0.25, -2.0, 0.525,
                 (w * w + 43.0 * w + 38.0) / (7.0 / 3.0 * w * w + 92.5 * w + 38.0), 0.348};
      Fit f50 = {12.5 * w + sw * (-1.0 + 5.0 * e5), -0.125, 0.375, -2.25, 0.541,
                 (w + 13.0 / 14.0) / (2.0 * w + 13.0 / 14.0), 0.313};
      Fit lo, hi;
      if (kk < 4.0) { r.kappa_rho_frac = (kk - 3.5) / (4.0 - 3.5); lo = f35; hi = f40; }
      else if (kk < 4.5) { r.kappa_rho_frac = (kk - 4.0) / (4.5 - 4.0); lo = f40; hi = f45; }
      else { r.kappa_rho_frac = (kk - 4.5) / (5.0 - 4.5); lo = f45; hi = f50; }
      r.kappa_rho_q_low_a = lo.qa; r.kappa_rho_q_low_b = lo.qb; r.kappa_rho_q_low_c = lo.qc;
      r.kappa_rho_q_low_d = lo.qd; r.kappa_rho_q_low_e = lo.qe;
      r.kappa_rho_q_high_a = hi.qa; r.kappa_rho_q_high_b = hi.qb; r.kappa_rho_q_high_c = hi.qc;
      r.kappa_rho_q_high_d = hi.qd; r.kappa_rho_q_high_e = hi.qe;
      r.kappa_rho_v_low_a = lo.va; r.kappa_rho_v_low_b = lo.vb;
      r.kappa_rho_v_high_a = hi.va; r.kappa_rho_v_high_b = hi.vb;
    }
  }
}

void fill_rad_params(const bl_params &p, RadParams &r) {
  std::memset(&r, 0, sizeof r);
  r.model_type = p.model_type; r.ray_flat = p.ray_flat; r.coord = p.simulation_coord; r.interp = p.simulation_interp;
  r.block_interp = p.simulation_interp && p.simulation_block_interp;
  r.slow_light = p.model_type == BL_MODEL_SIMULATION && p.slow_light_on;
  r.slow_interp = r.slow_light && p.slow_interp;
  r.slow_count = 0;   // set by bl_set_time_window
  r.extrap_tol = p.extrapolation_tolerance;
  r.a = p.bh_a; r.camera_r = p.camera_r;
  for (int i = 0; i < 4; i++) {
    r.camera_x[i] = p.camera_x[i]; r.camera_u_con[i] = p.camera_u_con[i];
    r.camera_u_cov[i] = p.camera_u_cov[i]; r.camera_vert_con_c[i] = p.camera_vert_con_c[i];
  }
  r.num_freq = p.image_num_frequencies;
  for (int l = 0; l < p.image_num_frequencies; l++) {
    r.freqs[l] = p.image_frequencies[l];
    r.inv_freqs[l] = 1.0 / p.image_frequencies[l];
    r.log_freqs[l] = std::log(p.image_frequencies[l]);
  }
  r.x_unit = phys::gg_msun * p.mass_msun / (phys::c * phys::c);
  r.t_unit = r.x_unit / phys::c;
  r.image_light = p.image_light; r.image_time = p.image_time; r.image_length = p.image_length;
  r.image_lambda = p.image_lambda; r.image_emission = p.image_emission; r.image_tau = p.image_tau;
  bool sim = p.model_type == BL_MODEL_SIMULATION;
  r.image_lambda_ave = sim && p.image_lambda_ave; r.image_emission_ave = sim && p.image_emission_ave;
  r.image_tau_int = sim && p.image_tau_int; r.image_crossings = p.image_crossings;
  r.polarization = sim && p.image_light && p.image_polarization;
  r.rotation_split = p.image_rotation_split;
  bl_params q = p;
  q.image_lambda_ave = r.image_lambda_ave; q.image_emission_ave = r.image_emission_ave; q.image_tau_int = r.image_tau_int;
  image_layout(q, r);
  r.render_num_images = sim ? p.render_num_images : 0;
  r.need_cell_values = r.image_lambda_ave || r.image_emission_ave || r.image_tau_int || r.render_num_images > 0;
  r.d_unit = p.simulation_rho_cgs;
  r.e_unit = r.d_unit * phys::c * phys::c;
  r.b_unit = std::sqrt(4.0 * phys::pi * r.e_unit);
  r.plasma_mu = p.plasma_mu; r.plasma_ne_ni = p.plasma_ne_ni; r.plasma_model = p.plasma_model; r.plasma_use_p = p.plasma_use_p;
  r.plasma_gamma = p.plasma_gamma; r.plasma_gamma_i = p.plasma_gamma_i; r.plasma_gamma_e = p.plasma_gamma_e;
  r.plasma_rat_low = p.plasma_rat_low; r.plasma_rat_high = p.plasma_rat_high;
  r.power_frac = p.plasma_power_frac; r.kappa_frac = p.plasma_kappa_frac;
  r.thermal_frac = 1.0 - (p.plasma_power_frac + p.plasma_kappa_frac);
  r.plasma_p = p.plasma_p; r.plasma_gamma_min = p.plasma_gamma_min; r.plasma_gamma_max = p.plasma_gamma_max;
  r.plasma_kappa = p.plasma_kappa; r.plasma_w = p.plasma_w;
  if (sim) plasma_constants(p, r);
  r.formula_r0 = p.formula_r0; r.formula_h = p.formula_h; r.formula_l0 = p.formula_l0; r.formula_q = p.formula_q;
  r.formula_nup = p.formula_nup; r.formula_cn0 = p.formula_cn0; r.formula_alpha = p.formula_alpha;
  r.formula_a = p.formula_a; r.formula_beta = p.formula_beta;
  r.cut_rho_min = p.cut_rho_min; r.cut_rho_max = p.cut_rho_max; r.cut_n_e_min = p.cut_n_e_min; r.cut_n_e_max = p.cut_n_e_max;
  r.cut_p_gas_min = p.cut_p_gas_min; r.cut_p_gas_max = p.cut_p_gas_max; r.cut_theta_e_min = p.cut_theta_e_min;
  r.cut_theta_e_max = p.cut_theta_e_max; r.cut_b_min = p.cut_b_min; r.cut_b_max = p.cut_b_max;
  r.cut_sigma_min = p.cut_sigma_min; r.cut_sigma_max = p.cut_sigma_max;
  r.cut_beta_inverse_min = p.cut_beta_inverse_min; r.cut_beta_inverse_max = p.cut_beta_inverse_max;
  r.cut_omit_near = p.cut_omit_near; r.cut_omit_far = p.cut_omit_far; r.cut_plane = p.cut_plane;
  r.cut_omit_in = p.cut_omit_in; r.cut_omit_out = p.cut_omit_out;
  r.cut_midplane_theta = p.cut_midplane_theta; r.cut_midplane_z = p.cut_midplane_z;
  for (int i = 0; i < 3; i++) { r.cut_plane_origin[i] = p.cut_plane_origin[i]; r.cut_plane_normal[i] = p.cut_plane_normal[i]; }
  r.n_e_factor = 1.0 / (p.plasma_mu * phys::m_p) / (1.0 + 1.0 / p.plasma_ne_ni);
  r.any_value_cut = p.cut_rho_min >= 0.0 || p.cut_rho_max >= 0.0 || p.cut_n_e_min >= 0.0 || p.cut_n_e_max >= 0.0 ||
                    p.cut_p_gas_min >= 0.0 || p.cut_p_gas_max >= 0.0 || p.cut_theta_e_min >= 0.0 || p.cut_theta_e_max >= 0.0 ||
                    p.cut_b_min >= 0.0 || p.cut_b_max >= 0.0 || p.cut_sigma_min >= 0.0 || p.cut_sigma_max >= 0.0 ||
                    p.cut_beta_inverse_min >= 0.0 || p.cut_beta_inverse_max >= 0.0;
  r.need_sigma_beta = r.need_cell_values;
  if (sim && p.plasma_kappa_frac != 0.0) {
    r.log_w2k2 = std::log(p.plasma_w * p.plasma_w * p.plasma_kappa * p.plasma_kappa);
    r.log_kjl = std::log(r.kappa_jj_low); r.log_kjh = std::log(r.kappa_jj_high);
    r.log_kal = std::log(r.kappa_aa_low); r.log_kah = std::log(r.kappa_aa_high * r.kappa_aa_high_i);
    r.log_k_j_pref = std::log(p.plasma_kappa_frac * phys::e * phys::e / phys::c);
    r.log_k_a_pref = std::log(p.plasma_kappa_frac * phys::e * phys::e / (phys::m_e * phys::c));
    r.log_kah_base = std::log(r.kappa_aa_high);
    if (r.polarization) {
      r.log_kj_low_q = std::log(r.kappa_jj_low_q); r.log_kj_low_v = std::log(r.kappa_jj_low_v);
      r.log_kj_high_q = std::log(r.kappa_jj_high_q); r.log_kj_high_v = std::log(r.kappa_jj_high_v);
      r.log_ka_low_q = std::log(r.kappa_aa_low_q); r.log_ka_low_v = std::log(r.kappa_aa_low_v);
      r.log_ka_high_q = std::log(r.kappa_aa_high_q); r.log_ka_high_v = std::log(r.kappa_aa_high_v);
    }
  }
  if (sim && p.plasma_kappa_frac != 0.0 && r.polarization) {
    const double kk = p.plasma_kappa;
    const double x[6] = {r.kappa_jj_x_i, r.kappa_jj_x_q, r.kappa_jj_x_v, r.kappa_aa_x_i, r.kappa_aa_x_q, r.kappa_aa_x_v};
    // d ln(lo) / d ln nu and d ln(hi) / d ln nu of the six bridged coefficients (pol_common.cuh: synchrotron_polarized)
    const double j_lo = -2.0 + 1.0 / 3.0, j_hi = -2.0 - (kk - 2.0) / 2.0, a_lo = -2.0 / 3.0, a_hi = -(1.0 + kk) / 2.0;
    const double s_lo[6] = {j_lo, j_lo, j_lo - 0.35, a_lo, a_lo, a_lo - 0.35};
    const double s_hi[6] = {j_hi, j_hi, j_hi - 0.5, a_hi, a_hi, a_hi - 0.5};
    for (int l = 0; l < p.image_num_frequencies; l++) {
      const double d = std::log(p.image_frequencies[l]) - std::log(p.image_frequencies[0]);
      r.dlog_freqs[l] = d;
      for (int t = 0; t < 6; t++) {
        r.kappa_k[t][l] = std::exp(-x[t] * (s_lo[t] - s_hi[t]) * d);
        r.kappa_kinv[t][l] = std::exp(x[t] * (s_lo[t] - s_hi[t]) * d);
      }
      r.rho_c84[l] = std::exp(0.84 * d);
      r.rho_cm12[l] = std::exp(-0.5 * d);
      r.rho_cqe_low[l] = std::exp(r.kappa_rho_q_low_e * d);
      r.rho_cqe_high[l] = std::exp(r.kappa_rho_q_high_e * d);
    }
    for (int t = 0; t < 6; t++) {
      r.kappa_inv_x[t] = 1.0 / x[t];
      r.kappa_slope_lo[t] = s_lo[t];
      r.kappa_slope_hi[t] = s_hi[t];
    }
  }
  if (sim && p.plasma_power_frac != 0.0) r.log_power_gmin = std::log(2.0 * p.plasma_gamma_min * p.plasma_gamma_min / 3.0);
  r.fallback_nan = p.fallback_nan; r.fallback_rho = p.fallback_rho; r.fallback_pgas = p.fallback_pgas;
  r.fallback_kappa = p.fallback_kappa;
  for (int i = 0; i <= BL_MAX_RENDER_FEATURES; i++) r.render_feature_start[i] = p.render_feature_start[i];
  for (int i = 0; i < BL_MAX_RENDER_FEATURES; i++) {
    r.render_quantities[i] = p.render_quantities[i]; r.render_types[i] = p.render_types[i];
    r.render_min_vals[i] = p.render_min_vals[i]; r.render_max_vals[i] = p.render_max_vals[i];
    r.render_thresh_vals[i] = p.render_thresh_vals[i]; r.render_tau_scales[i] = p.render_tau_scales[i];
    r.render_opacities[i] = p.render_opacities[i];
    r.render_x_vals[i] = p.render_x_vals[i]; r.render_y_vals[i] = p.render_y_vals[i]; r.render_z_vals[i] = p.render_z_vals[i];
  }
}

int validate_params(const bl_params &p) {
  if (p.abi_version != BL_ABI_VERSION) return bl_fail(nullptr, BL_ERR_ARG, "bl_params.abi_version %d != %d", p.abi_version, BL_ABI_VERSION);
  if (p.model_type != BL_MODEL_SIMULATION && p.model_type != BL_MODEL_FORMULA) return bl_fail(nullptr, BL_ERR_ARG, "unknown model_type");
  // messages below are the reference's own (geodesic_integrator.cpp:39-104, radiation_integrator.cpp:201)
  if (p.ray_max_steps <= 0) return bl_fail(nullptr, BL_ERR_ARG, "Must have positive ray_max_steps.");
  if (p.ray_integrator == BL_INTEGRATOR_DP && p.ray_max_retries <= 0) return bl_fail(nullptr, BL_ERR_ARG, "Must have nonnegative ray_max_retries.");
  if (p.image_num_frequencies < 1) return bl_fail(nullptr, BL_ERR_ARG, "Must have positive image_num_frequencies.");
  if (p.image_num_frequencies > BL_MAX_FREQ) return bl_fail(nullptr, BL_ERR_UNSUPPORTED, "image_num_frequencies > %d not supported", BL_MAX_FREQ);
  for (int l = 0; l < p.image_num_frequencies; l++)
    if (!(p.image_frequencies[l] > 0.0)) return bl_fail(nullptr, BL_ERR_ARG, "Must choose positive image_frequency.");
  bool sim = p.model_type == BL_MODEL_SIMULATION;
  if (sim && p.slow_light_on && (p.slow_chunk_size < 2 || p.slow_chunk_size > BL_MAX_SLICES))
    return bl_fail(nullptr, p.slow_chunk_size < 2 ? BL_ERR_ARG : BL_ERR_UNSUPPORTED,
                   p.slow_chunk_size < 2 ? "Must have slow_chunk_size be at least 2." : "slow_chunk_size > %d not supported", BL_MAX_SLICES);
  if (!(p.image_light || p.image_time || p.image_length || p.image_lambda || p.image_emission || p.image_tau ||
        (sim && (p.image_lambda_ave || p.image_emission_ave || p.image_tau_int)) || p.image_crossings ||
        (sim && p.render_num_images > 0)))
    return bl_fail(nullptr, BL_ERR_ARG, "No image or rendering selected.");
  if (sim && p.render_num_images > 0 && p.render_feature_start[p.render_num_images] > BL_MAX_RENDER_FEATURES)
    return bl_fail(nullptr, BL_ERR_UNSUPPORTED, "more than %d render features", BL_MAX_RENDER_FEATURES);
  if (sim && p.image_light && p.image_polarization && p.plasma_kappa_frac != 0.0 &&
      (p.plasma_kappa < 3.5 || p.plasma_kappa > 5.0))
    return bl_fail(nullptr, BL_ERR_ARG, "Polarized transport only supports kappa in [3.5, 5].");
  if (p.adaptive_max_level > 0) {
    if (!p.image_light) return bl_fail(nullptr, BL_ERR_ARG, "Adaptive ray tracing requires image_light.");
    if (p.adaptive_block_size <= 0) return bl_fail(nullptr, BL_ERR_ARG, "Must have positive adaptive_block_size.");
    if (p.camera_resolution % p.adaptive_block_size != 0) return bl_fail(nullptr, BL_ERR_ARG, "Must have adaptive_block_size divide camera_resolution.");
    if (p.adaptive_num_regions > BL_MAX_REGIONS) return bl_fail(nullptr, BL_ERR_UNSUPPORTED, "more than %d adaptive regions", BL_MAX_REGIONS);
  }
  return BL_OK;
}

}  // namespace

extern "C" {

const char *bl_last_error(const bl_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int bl_create(const bl_params *params, bl_ctx **out) {
  if (!params || !out) return bl_fail(nullptr, BL_ERR_ARG, "bl_create: null argument");
  *out = nullptr;
  int rc = validate_params(*params);
  if (rc) return rc;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count <= 0)
    return bl_fail(nullptr, BL_ERR_CUDA, "no usable CUDA device (%s); blacklight_b200 has no CPU fallback",
                   err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
  if (params->device < 0 || params->device >= count) return bl_fail(nullptr, BL_ERR_ARG, "device %d out of range (%d devices)", params->device, count);
  bl_ctx *ctx = new (std::nothrow) bl_ctx();
  if (!ctx) return bl_fail(nullptr, BL_ERR_NOMEM, "out of host memory");
  ctx->params = *params;
  ctx->device = params->device;
  fill_rad_params(*params, ctx->rad);
  ctx->levels.resize((size_t)params->adaptive_max_level + 1);
  if (const char *e = getenv("BL_GEO_BLOCKS")) ctx->geo_min_blocks = atoi(e);
  if (const char *e = getenv("BL_POL_SLAB")) ctx->pol_slab = atoi(e);
  if (const char *e = getenv("BL_RAD_PREFETCH")) ctx->rad_prefetch = atoi(e);
  if (const char *e = getenv("BL_RAY_ORDER")) ctx->ray_order = atoi(e);
  if (const char *e = getenv("BL_GEO_ORDER")) ctx->geo_order = atoi(e);
  if (const char *e = getenv("BL_POL_FUSED")) ctx->pol_fused = atoi(e) != 0;
#define CREATE_CHECK(call)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      bl_fail(nullptr, BL_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      bl_destroy(ctx);                                                                       \
      return BL_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)
  CREATE_CHECK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CREATE_CHECK(cudaGetDeviceProperties(&prop, ctx->device));
  ctx->sm_count = prop.multiProcessorCount;
  CREATE_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CREATE_CHECK(cudaEventCreate(&ctx->ev0));
  CREATE_CHECK(cudaEventCreate(&ctx->ev1));
  CREATE_CHECK(dev_alloc(&ctx->params_dev, 1));
  CREATE_CHECK(dev_alloc(&ctx->counters, 1));
  CREATE_CHECK(dev_alloc(&ctx->rad_counter, 1));
  CREATE_CHECK(dev_alloc(&ctx->slow_counters, 8));
  CREATE_CHECK(cudaMemcpy(ctx->params_dev, &ctx->params, sizeof(bl_params), cudaMemcpyHostToDevice));
#undef CREATE_CHECK
  *out = ctx;
  return BL_OK;
}

void bl_destroy(bl_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto &L : ctx->levels) free_level(L);
  for (void *p : ctx->grid_allocs) cudaFree(p);
  cudaFree(ctx->units_dev); cudaFree(ctx->refine_locs); cudaFree(ctx->refine_flags); cudaFree(ctx->params_dev); cudaFree(ctx->counters); cudaFree(ctx->rad_counter); cudaFree(ctx->slow_counters);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (cudaEvent_t e : ctx->stage_events) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int bl_device_count(void) {
  int count = 0;
  return cudaGetDeviceCount(&count) == cudaSuccess ? count : 0;
}

int bl_image_num_quantities(const bl_ctx *ctx) { return ctx ? ctx->rad.num_quantities : -1; }

int bl_set_taps(bl_ctx *ctx, int enabled) {
  if (!ctx) return BL_ERR_ARG;
  ctx->taps_enabled = enabled != 0;
  return BL_OK;
}

int bl_device_info(bl_ctx *ctx, char *name, int name_len, int *sm_count, double *hbm_free_gb) {
  if (!ctx) return BL_ERR_ARG;
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  BL_CUDA_CHECK(cudaGetDeviceProperties(&prop, ctx->device));
  if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", prop.name);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  size_t fr = 0, tot = 0;
  BL_CUDA_CHECK(cudaMemGetInfo(&fr, &tot));
  if (hbm_free_gb) *hbm_free_gb = (double)fr / 1e9;
  return BL_OK;
}

int bl_measure_fp64_peak(bl_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return BL_ERR_ARG;
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  double *sink = nullptr;
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  BL_CUDA_CHECK(dev_alloc(&sink, (size_t)blocks * threads));
  BL_CUDA_CHECK(bl_launch_fp64_peak(sink, blocks, iters, ctx->stream));  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    BL_CUDA_CHECK(bl_launch_fp64_peak(sink, blocks, iters, ctx->stream));
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    BL_CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    BL_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    // 16 independent FMA chains per thread, 2 flop per FMA
    double flop = (double)blocks * threads * (double)iters * 16.0 * 2.0;
    double tf = flop / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaFree(sink);
  *tflops = best;
  return BL_OK;
}

int bl_selftest_division(bl_ctx *ctx, uint64_t seed, int64_t num_pairs, int64_t *mismatches) {
  if (!ctx || !mismatches || num_pairs <= 0) return BL_ERR_ARG;
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  unsigned long long *dev = nullptr;
  BL_CUDA_CHECK(dev_alloc(&dev, 1));
  const int blocks = ctx->sm_count * 8, threads = 256;
  int iters = (int)((num_pairs + (int64_t)blocks * threads - 1) / ((int64_t)blocks * threads));
  cudaError_t e = cudaMemsetAsync(dev, 0, sizeof *dev, ctx->stream);
  if (e == cudaSuccess) e = bl_launch_division_selftest(seed, blocks, iters, dev, ctx->stream);
  unsigned long long host = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&host, dev, sizeof host, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dev);
  if (e != cudaSuccess) return bl_fail_cuda(ctx, e, "division self-test", __FILE__, __LINE__);
  *mismatches = (int64_t)host;
  ctx->launches++;
  return BL_OK;
}

static int upload_grid_slot(bl_ctx *ctx, const bl_grid_view *gv, int slot);

int bl_upload_grid(bl_ctx *ctx, const bl_grid_view *gv) { return upload_grid_slot(ctx, gv, 0); }

int bl_upload_grid_slice(bl_ctx *ctx, const bl_grid_view *gv, int slot) {
  if (!ctx) return BL_ERR_ARG;
  if (!ctx->rad.slow_light) return bl_fail(ctx, BL_ERR_STATE, "bl_upload_grid_slice: slow_light_on is false");
  if (slot < 0 || slot >= ctx->params.slow_chunk_size) return bl_fail(ctx, BL_ERR_ARG, "bl_upload_grid_slice: slot %d outside [0, slow_chunk_size)", slot);
  return upload_grid_slot(ctx, gv, slot);
}

int bl_set_time_window(bl_ctx *ctx, int count, const int32_t *slots, const double *times, double snapshot_time) {
  if (!ctx || !slots || !times) return BL_ERR_ARG;
  if (!ctx->rad.slow_light) return bl_fail(ctx, BL_ERR_STATE, "bl_set_time_window: slow_light_on is false");
  if (count != ctx->params.slow_chunk_size) return bl_fail(ctx, BL_ERR_ARG, "bl_set_time_window: %d entries, slow_chunk_size is %d", count, ctx->params.slow_chunk_size);
  for (int t = 0; t < count; t++) {
    if (slots[t] < 0 || slots[t] >= count) return bl_fail(ctx, BL_ERR_ARG, "bl_set_time_window: slot %d out of range", slots[t]);
    if (t > 0 && !(times[t] < times[t - 1])) return bl_fail(ctx, BL_ERR_ARG, "bl_set_time_window: times must decrease (entry 0 is the latest snapshot)");
    ctx->rad.slow_slot[t] = slots[t];
    ctx->rad.slow_time[t] = times[t];
  }
  ctx->rad.slow_count = count;
  ctx->rad.snapshot_time = snapshot_time;
  return BL_OK;
}

int bl_slow_light_stats(bl_ctx *ctx, int level, bl_slow_stats *out) {
  if (!ctx || !out) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_slow_light_stats: level %d out of range", level);
  *out = ctx->levels[level].slow;
  return BL_OK;
}

static int upload_grid_slot(bl_ctx *ctx, const bl_grid_view *gv, int slot) {
  if (!ctx || !gv) return BL_ERR_ARG;
  if (gv->n_b <= 0 || gv->n_i <= 0 || gv->n_j <= 0 || gv->n_k <= 0 || !gv->prim || !gv->x1f || !gv->x2f ||
      !gv->x3f || !gv->x1v || !gv->x2v || !gv->x3v)
    return bl_fail(ctx, BL_ERR_ARG, "bl_upload_grid: incomplete grid view");
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  bool same_shape = ctx->have_grid && ctx->grid.n_b == gv->n_b && ctx->grid.n_k == gv->n_k &&
                    ctx->grid.n_j == gv->n_j && ctx->grid.n_i == gv->n_i;
  bool want_kappa = ctx->params.plasma_model == BL_PLASMA_CODE_KAPPA;
  if (want_kappa && (gv->ind_kappa < 0 || gv->ind_kappa >= gv->n_var))
    return bl_fail(ctx, BL_ERR_ARG, "plasma_model = code_kappa needs an electron entropy variable");
  const bool fmks = ctx->params.simulation_coord == BL_COORD_FMKS;
  if (fmks && (gv->n_b != 1 || !gv->sks_map || gv->sks_map_n1 < 2 || gv->sks_map_n2 < 2 || !(gv->sks_map_dr > 0.0) ||
               !(gv->sks_map_dtheta > 0.0) || gv->n_i < 2 || gv->n_j < 2))
    return bl_fail(ctx, BL_ERR_ARG, "simulation_coord = fmks needs a single block and the reader's sks_map in the grid view");
  // all argument validation happens before anything is freed or overwritten
  {
    const int vidx[8] = {gv->ind_rho, gv->ind_pgas, gv->ind_uu1, gv->ind_uu2, gv->ind_uu3, gv->ind_bb1, gv->ind_bb2, gv->ind_bb3};
    for (int q = 0; q < 8; q++)
      if (vidx[q] < 0 || vidx[q] >= gv->n_var) return bl_fail(ctx, BL_ERR_ARG, "bl_upload_grid: variable index %d out of range", q);
    if (ctx->rad.block_interp) {
      if (!gv->levels || !gv->locations)
        return bl_fail(ctx, BL_ERR_ARG, "simulation_block_interp needs the blocks' levels and logical locations");
      if (ctx->params.simulation_coord == BL_COORD_SKS && (gv->n_3_root <= 0 || gv->n_3_root % gv->n_k != 0))
        return bl_fail(ctx, BL_ERR_ARG, "simulation_block_interp needs RootGridSize[2] (n_3_root) as a multiple of the block size");
    }
    if (fmks && same_shape && (ctx->grid.map_n1 != gv->sks_map_n1 || ctx->grid.map_n2 != gv->sks_map_n2))
      return bl_fail(ctx, BL_ERR_ARG, "bl_upload_grid: sks_map changed shape between snapshots");
  }
  // from here on the grid counts as absent until the upload has completed: a CUDA failure below must not leave a
  // half-built grid usable (slices of a slow-light window are added to a grid that stays valid meanwhile)
  if (slot == 0) ctx->have_grid = false;
  size_t cells = (size_t)gv->n_b * gv->n_k * gv->n_j * gv->n_i;
  GridDev &g = ctx->grid;
  if (!same_shape) {
    for (void *p : ctx->grid_allocs) cudaFree(p);
    ctx->grid_allocs.clear();
    g = GridDev();
    g.n_b = gv->n_b; g.n_k = gv->n_k; g.n_j = gv->n_j; g.n_i = gv->n_i;
    double *d = nullptr;
    auto alloc_d = [&](const double **dst, size_t n) -> cudaError_t {
      cudaError_t e = dev_alloc(&d, n);
      if (e == cudaSuccess) { *dst = d; ctx->grid_allocs.push_back(d); }
      return e;
    };
    BL_CUDA_CHECK(alloc_d(&g.x1f, (size_t)g.n_b * (g.n_i + 1)));
    BL_CUDA_CHECK(alloc_d(&g.x2f, (size_t)g.n_b * (g.n_j + 1)));
    BL_CUDA_CHECK(alloc_d(&g.x3f, (size_t)g.n_b * (g.n_k + 1)));
    // one element of padding: the reference's inter-block fractions read x?v(b, n) -- the first centre of the
    // next block, and one element past the array for the last block (simulation_sampling.cpp:519-521)
    BL_CUDA_CHECK(alloc_d(&g.x1v, (size_t)g.n_b * g.n_i + 1));
    BL_CUDA_CHECK(alloc_d(&g.x2v, (size_t)g.n_b * g.n_j + 1));
    BL_CUDA_CHECK(alloc_d(&g.x3v, (size_t)g.n_b * g.n_k + 1));
    BL_CUDA_CHECK(cudaMemsetAsync((void *)(g.x1v + (size_t)g.n_b * g.n_i), 0, sizeof(double), ctx->stream));
    BL_CUDA_CHECK(cudaMemsetAsync((void *)(g.x2v + (size_t)g.n_b * g.n_j), 0, sizeof(double), ctx->stream));
    BL_CUDA_CHECK(cudaMemsetAsync((void *)(g.x3v + (size_t)g.n_b * g.n_k), 0, sizeof(double), ctx->stream));
    if (ctx->rad.block_interp) {
      uint32_t size = 16;
      while (size < 2u * (uint32_t)g.n_b) size <<= 1;
      g.hash_mask = size - 1;
      int32_t *ip = nullptr;
      unsigned long long *kp = nullptr;
      BL_CUDA_CHECK(dev_alloc(&ip, (size_t)g.n_b)); g.levels = ip; ctx->grid_allocs.push_back(ip);
      BL_CUDA_CHECK(dev_alloc(&ip, (size_t)g.n_b * 3)); g.locs = ip; ctx->grid_allocs.push_back(ip);
      BL_CUDA_CHECK(dev_alloc(&ip, (size_t)size)); g.hash_vals = ip; ctx->grid_allocs.push_back(ip);
      BL_CUDA_CHECK(dev_alloc(&kp, (size_t)size)); g.hash_keys = kp; ctx->grid_allocs.push_back(kp);
    }
    BL_CUDA_CHECK(alloc_d(&g.bounds, (size_t)g.n_b * 6));
    if (fmks) {
      BL_CUDA_CHECK(alloc_d(&g.sks_map, (size_t)2 * gv->sks_map_n2 * gv->sks_map_n1));
      g.map_n1 = gv->sks_map_n1;
      g.map_n2 = gv->sks_map_n2;
    }
    BL_CUDA_CHECK(alloc_d(&g.x1d, (size_t)g.n_b * g.n_i));
    BL_CUDA_CHECK(alloc_d(&g.x2d, (size_t)g.n_b * g.n_j));
    BL_CUDA_CHECK(alloc_d(&g.x3d, (size_t)g.n_b * g.n_k));
    // slow light keeps slow_chunk_size snapshots resident, back to back
    const size_t slices = ctx->rad.slow_light ? (size_t)ctx->params.slow_chunk_size : 1;
    g.slice_cells = cells;
    float4 *c4 = nullptr;
    BL_CUDA_CHECK(dev_alloc(&c4, cells * 2 * slices));
    g.cells = c4; ctx->grid_allocs.push_back(c4);
    if (want_kappa) {
      float *kp = nullptr;
      BL_CUDA_CHECK(dev_alloc(&kp, cells * slices));
      g.kappa = kp; ctx->grid_allocs.push_back(kp);
    }
  }
  auto h2d = [&](const double *dst, const double *src, size_t n) {
    return cudaMemcpyAsync((void *)dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  };
  BL_CUDA_CHECK(h2d(g.x1f, gv->x1f, (size_t)g.n_b * (g.n_i + 1)));
  BL_CUDA_CHECK(h2d(g.x2f, gv->x2f, (size_t)g.n_b * (g.n_j + 1)));
  BL_CUDA_CHECK(h2d(g.x3f, gv->x3f, (size_t)g.n_b * (g.n_k + 1)));
  BL_CUDA_CHECK(h2d(g.x1v, gv->x1v, (size_t)g.n_b * g.n_i));
  BL_CUDA_CHECK(h2d(g.x2v, gv->x2v, (size_t)g.n_b * g.n_j));
  BL_CUDA_CHECK(h2d(g.x3v, gv->x3v, (size_t)g.n_b * g.n_k));
  std::vector<double> bounds((size_t)g.n_b * 6);
  for (int b = 0; b < g.n_b; b++) {
    bounds[6 * b + 0] = gv->x1f[(size_t)b * (g.n_i + 1)];
    bounds[6 * b + 1] = gv->x1f[(size_t)b * (g.n_i + 1) + g.n_i];
    bounds[6 * b + 2] = gv->x2f[(size_t)b * (g.n_j + 1)];
    bounds[6 * b + 3] = gv->x2f[(size_t)b * (g.n_j + 1) + g.n_j];
    bounds[6 * b + 4] = gv->x3f[(size_t)b * (g.n_k + 1)];
    bounds[6 * b + 5] = gv->x3f[(size_t)b * (g.n_k + 1) + g.n_k];
  }
  if (fmks) {
    // the grid's extent in (r, theta, phi) stands in for the native face positions (simulation_sampling.cpp:190-198)
    if (g.map_n1 != gv->sks_map_n1 || g.map_n2 != gv->sks_map_n2)
      return bl_fail(ctx, BL_ERR_ARG, "bl_upload_grid: sks_map changed shape between snapshots");
    for (int d = 0; d < 6; d++) bounds[(size_t)d] = gv->simulation_bounds[d];
    g.map_r_in = gv->sks_map_r_in;
    g.map_dr = gv->sks_map_dr;
    g.map_dtheta = gv->sks_map_dtheta;
    BL_CUDA_CHECK(h2d(g.sks_map, gv->sks_map, (size_t)2 * g.map_n2 * g.map_n1));
  }
  BL_CUDA_CHECK(cudaMemcpyAsync((void *)g.bounds, bounds.data(), bounds.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // mesh topology: levels, logical locations and the (level, location) -> block hash
  std::vector<unsigned long long> hkeys;
  std::vector<int32_t> hvals;
  if (ctx->rad.block_interp) {
    if (!gv->levels || !gv->locations)
      return bl_fail(ctx, BL_ERR_ARG, "simulation_block_interp needs the blocks' levels and logical locations");
    // blocks around x3 at the root level: only the azimuthal wrap of spherical grids looks at it
    const bool periodic_x3 = ctx->params.simulation_coord == BL_COORD_SKS;
    if (periodic_x3 && (gv->n_3_root <= 0 || gv->n_3_root % g.n_k != 0))
      return bl_fail(ctx, BL_ERR_ARG, "simulation_block_interp needs RootGridSize[2] (n_3_root) as a multiple of the block size");
    g.n3_root = periodic_x3 ? gv->n_3_root / g.n_k : 0;
    g.max_level = 0;
    hkeys.assign((size_t)g.hash_mask + 1, ~0ull);
    hvals.assign((size_t)g.hash_mask + 1, -1);
    for (int b = 0; b < g.n_b; b++) {
      int lev = gv->levels[b];
      const int32_t *loc = gv->locations + 3 * (size_t)b;
      if (lev < 0 || lev > 40 || loc[0] < 0 || loc[1] < 0 || loc[2] < 0 || loc[0] >= (1 << 19) || loc[1] >= (1 << 19) || loc[2] >= (1 << 19))
        return bl_fail(ctx, BL_ERR_ARG, "block %d: level/location outside the supported range", b);
      if (lev > g.max_level) g.max_level = lev;
      unsigned long long key = block_key(lev, loc[0], loc[1], loc[2]);
      uint32_t h = block_hash(key) & g.hash_mask;
      while (hkeys[h] != ~0ull) h = (h + 1) & g.hash_mask;
      hkeys[h] = key;
      hvals[h] = b;
    }
    BL_CUDA_CHECK(cudaMemcpyAsync((void *)g.levels, gv->levels, (size_t)g.n_b * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    BL_CUDA_CHECK(cudaMemcpyAsync((void *)g.locs, gv->locations, (size_t)g.n_b * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    BL_CUDA_CHECK(cudaMemcpyAsync((void *)g.hash_keys, hkeys.data(), hkeys.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    BL_CUDA_CHECK(cudaMemcpyAsync((void *)g.hash_vals, hvals.data(), hvals.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  // reciprocal cell-centre spacings, so that interpolation fractions need no division on the device
  std::vector<double> inv1((size_t)g.n_b * g.n_i), inv2((size_t)g.n_b * g.n_j), inv3((size_t)g.n_b * g.n_k);
  auto fill_inv = [&](std::vector<double> &out, const double *xv, int n) {
    for (int b = 0; b < g.n_b; b++)
      for (int i = 0; i < n; i++)
        out[(size_t)b * n + i] = i + 1 < n ? 1.0 / (xv[(size_t)b * n + i + 1] - xv[(size_t)b * n + i]) : 0.0;
  };
  fill_inv(inv1, gv->x1v, g.n_i);
  fill_inv(inv2, gv->x2v, g.n_j);
  fill_inv(inv3, gv->x3v, g.n_k);
  BL_CUDA_CHECK(h2d(g.x1d, inv1.data(), inv1.size()));
  BL_CUDA_CHECK(h2d(g.x2d, inv2.data(), inv2.size()));
  BL_CUDA_CHECK(h2d(g.x3d, inv3.data(), inv3.size()));
  BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the staging vectors above go out of scope
  // primitives: stage the reader's (var, b, k, j, i) planes, then re-lay out to one record per cell
  float *stage = nullptr;
  BL_CUDA_CHECK(dev_alloc(&stage, (size_t)gv->n_var * cells));
  cudaError_t e = cudaMemcpyAsync(stage, gv->prim, (size_t)gv->n_var * cells * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  int idx[9] = {gv->ind_rho, gv->ind_pgas, gv->ind_uu1, gv->ind_uu2, gv->ind_uu3, gv->ind_bb1, gv->ind_bb2, gv->ind_bb3,
                want_kappa ? gv->ind_kappa : -1};
  for (int q = 0; q < 8 && e == cudaSuccess; q++)
    if (idx[q] < 0 || idx[q] >= gv->n_var) { cudaFree(stage); return bl_fail(ctx, BL_ERR_ARG, "bl_upload_grid: variable index %d out of range", q); }
  for (int q = 0; q < 9; q++) {
    g.next_slot[q] = -1;
    for (int u = 0; u < 9; u++)
      if (idx[q] >= 0 && idx[u] == idx[q] + 1) g.next_slot[q] = (int8_t)u;
  }
  int *idx_dev = nullptr;
  if (e == cudaSuccess) e = dev_alloc(&idx_dev, 9);
  if (e == cudaSuccess) e = cudaMemcpyAsync(idx_dev, idx, sizeof idx, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
    e = bl_launch_relayout_grid(stage, gv->n_var, idx_dev, cells, const_cast<float4 *>(g.cells) + 2 * cells * (size_t)slot,
                                g.kappa ? const_cast<float *>(g.kappa) + cells * (size_t)slot : nullptr, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(stage);
  cudaFree(idx_dev);
  if (e != cudaSuccess) return bl_fail_cuda(ctx, e, "grid upload", __FILE__, __LINE__);
  ctx->have_grid = true;
  return BL_OK;
}

}  // extern "C"

namespace {

// The polarized pipeline (radiate_pol_split.cu) covers the light image and the per-frequency sums
// (tau, lambda, emission); everything that needs per-sample side outputs stays on the fused kernel.
bool pol_split_eligible(const bl_ctx *ctx) {
  const RadParams &r = ctx->rad;
  return r.polarization && !ctx->pol_fused && !r.block_interp && !r.slow_light && r.coord != 2 &&
         !(r.image_time || r.image_length || r.image_lambda_ave || r.image_emission_ave || r.image_tau_int ||
           r.image_crossings) && r.render_num_images == 0;
}

// Slab length and scratch bytes per ray of that pipeline for a level of num_rays rays under `budget` bytes.
int pol_split_slab(const bl_ctx *ctx, int64_t num_rays, size_t budget, size_t *bytes_per_ray) {
  *bytes_per_ray = 0;
  if (!pol_split_eligible(ctx)) return 0;
  const size_t nf = (size_t)bl_polarized_split_fields(ctx->rad.num_freq);
  int slab = ctx->pol_slab > 0 ? ctx->pol_slab : 64;
  // small levels: fewer, longer slabs while the scratch stays under a tenth of the budget
  if (ctx->pol_slab <= 0)
    while (slab < 256 && (double)num_rays * (2.0 * slab) * (double)nf * 8.0 <= 0.1 * (double)budget) slab *= 2;
  if (slab > ctx->params.ray_max_steps) slab = ctx->params.ray_max_steps;
  *bytes_per_ray = ((size_t)slab * nf + (size_t)(slab + 1) * 8 + 10) * sizeof(double) + sizeof(int32_t);   // scratch, frame, camera map, list
  return slab;
}

// Trace rays [first, first+count) of a level into L.step (one wave).
int trace_wave(bl_ctx *ctx, Level &L, int64_t first, int64_t count) {
  // the integrator's queue hands out the rays by increasing impact parameter: longest first (ray_order.cu)
  const bool queue_order = ctx->params.ray_integrator == BL_INTEGRATOR_DP && L.order && L.order_buckets > 0 && ctx->geo_order;
  if (queue_order) {
    BL_CUDA_CHECK(cudaMemsetAsync(&ctx->counters->b_max_bits, 0, sizeof(unsigned int), ctx->stream));
    BL_CUDA_CHECK(bl_launch_impact_order(L.cam_pos + 4 * first, L.cam_dir + 4 * first, count, L.num + first,
                                         &ctx->counters->b_max_bits, kImpactBuckets, L.order_ws, L.order, ctx->stream));
    ctx->launches += 5;
  }
  const bl_params &p = ctx->params;
  GeoArgs g{};
  g.cam_pos = L.cam_pos + 4 * first;
  g.cam_dir = L.cam_dir + 4 * first;
  g.rays = count;
  g.a = p.bh_a; g.camera_r = p.camera_r; g.r_terminate = p.r_terminate; g.r_horizon = p.r_horizon;
  g.ray_step = p.ray_step; g.tol_abs = p.ray_tol_abs; g.tol_rel = p.ray_tol_rel;
  g.max_steps = p.ray_max_steps; g.max_retries = p.ray_max_retries;
  g.sb.buf = L.step; g.sb.rays = count; g.sb.cap = p.ray_max_steps;
  g.sample_num = L.num + first;
  g.sample_flags = L.flags + first;
  g.counters = ctx->counters;
  g.order = queue_order ? L.order : nullptr;
  BL_CUDA_CHECK(cudaMemsetAsync(&ctx->counters->next_ray, 0, sizeof(unsigned long long), ctx->stream));
  // Three CTAs per SM give the best throughput (56 ms against 62 at two, 1024^2 rays of the simulation camera); with
  // fewer than eight rays per thread the kernel is bound by the serial chain of its longest rays instead, and those run
  // faster with two (example_formula at 512^2: 113 -> 103 ms)
  int min_blocks = ctx->geo_min_blocks;
  if (min_blocks <= 0) min_blocks = count < (int64_t)8 * ctx->sm_count * 3 * 128 ? 2 : 3;
  if (p.ray_integrator == BL_INTEGRATOR_DP)
    BL_CUDA_CHECK(bl_launch_geodesic_dp(&g, p.ray_flat, ctx->sm_count, min_blocks, ctx->stream));
  else
    BL_CUDA_CHECK(bl_launch_geodesic_rk(&g, p.ray_flat, p.ray_integrator == BL_INTEGRATOR_RK4 ? 4 : 2, ctx->sm_count, ctx->stream));
  ctx->launches++;
  return BL_OK;
}

int read_geo_counters(bl_ctx *ctx, Level &L) {
  GeoCounters c;
  BL_CUDA_CHECK(cudaMemcpyAsync(&c, ctx->counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
  BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  L.stats.num_rays = L.rays;
  L.stats.geodesic_num_steps = c.max_samples;
  L.stats.num_bad_geodesics = (int64_t)c.bad;
  L.stats.num_samples = (int64_t)c.samples;
  L.stats.num_attempts = (int64_t)c.attempts;
  L.stats.num_accepted = (int64_t)c.accepted;
  return BL_OK;
}

// s_top: upper bound of the sample counts of these rays (the slabs of the polarized pipeline start there).
// *split_slabs: slabs launched by the polarized pipeline (0: the fused kernels ran).
int radiate_wave(bl_ctx *ctx, Level &L, int64_t first, int64_t count, int s_top, int *split_slabs) {
  *split_slabs = 0;
  RadArgs A{};
  A.grid = ctx->grid;
  A.sb.buf = L.step; A.sb.rays = count; A.sb.cap = ctx->params.ray_max_steps;
  A.sample_num = L.num + first;
  A.sample_flags = L.flags + first;
  A.mom_factor = L.mom + first;
  A.cam_pos = L.cam_pos + 4 * first;
  A.cam_dir = L.cam_dir + 4 * first;
  A.rays = count;
  A.image = L.image + first;
  A.image_stride = L.rays;
  A.render = L.render ? L.render + first : nullptr;
  A.sample_counter = ctx->rad_counter;
  A.prefetch = ctx->rad_prefetch;
  A.slow_counters = ctx->rad.slow_light ? ctx->slow_counters : nullptr;
  if (L.tap_nan) {
    size_t o = (size_t)first * L.tap_S;
    A.taps.S = L.tap_S;
    A.taps.nan_ = L.tap_nan + o; A.taps.cut = L.tap_cut + o; A.taps.fallback = L.tap_fb + o;
    A.taps.inds = L.tap_inds ? L.tap_inds + 4 * o : nullptr;
    A.taps.fracs = L.tap_fracs ? L.tap_fracs + 3 * o : nullptr;
  }
  // parameters travel by value in the kernel's constant bank (no device copy, no per-sample loads)
  // the unpolarized kernel is instantiated for frequency-count buckets of 1, 4 and 32 (one object file each)
  const int F = ctx->rad.num_freq;
  cudaError_t le;
  // rays sorted by length, longest first; alive[s] = rays with more than s * unit samples (a prefix of the list)
  std::vector<int64_t> alive;
  A.order = nullptr;
  A.active = count;
  if (L.order_buckets > 0 && L.order) {
    const int nb = L.order_buckets;
    BL_CUDA_CHECK(bl_launch_ray_order(L.num + first, count, L.order_unit, nb, L.order_ws, L.order, ctx->stream));
    ctx->launches += 3;
    std::vector<int32_t> totals((size_t)nb);
    const int64_t chunks = (count + 1023) / 1024;
    BL_CUDA_CHECK(cudaMemcpyAsync(totals.data(), L.order_ws + (size_t)nb * (size_t)chunks, (size_t)nb * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    // bucket b holds the rays with ceil(num / unit) == nb - 1 - b
    alive.assign((size_t)nb, 0);
    int64_t run = 0;
    for (int key = nb - 1; key >= 1; key--) {
      run += totals[(size_t)(nb - 1 - key)];
      alive[(size_t)(key - 1)] = run;     // rays with ceil(num / unit) > key - 1
    }
    A.order = L.order;                    // the fused kernels walk the whole list (rays without samples still write
                                          // their pixels); the pipeline's slabs take the prefixes alive[s]
  }
  L.order_alive0 = A.order ? alive[0] : -1;
  if (ctx->rad.polarization && L.slab > 0 && L.scratch && !L.tap_nan) {
    // the Stokes state of the pipeline starts (and stays between slabs) in the image columns of these rays
    BL_CUDA_CHECK(cudaMemset2DAsync(L.image + first, (size_t)L.rays * sizeof(double), 0, (size_t)count * sizeof(double),
                                    (size_t)ctx->rad.num_quantities, ctx->stream));
    const int slabs = bl_polarized_split_slabs(L.slab, s_top);
    while ((int)ctx->stage_events.size() < 4 * slabs + 1) {
      cudaEvent_t e;
      BL_CUDA_CHECK(cudaEventCreate(&e));
      ctx->stage_events.push_back(e);
    }
    BL_CUDA_CHECK(bl_launch_radiate_polarized_split(&A, &ctx->rad, L.scratch, L.frame, L.cam_map, L.slab, s_top, ctx->stream,
                                                    ctx->stage_events.data(), &ctx->launches,
                                                    alive.empty() ? nullptr : alive.data(), (int)alive.size()));
    *split_slabs = slabs;
    return BL_OK;
  }
  if (ctx->rad.polarization)
    le = bl_launch_radiate_polarized(&A, &ctx->rad, ctx->stream);
  else
    le = F <= 1 ? bl_launch_radiate_unpolarized_f1(&A, &ctx->rad, ctx->stream)
       : F <= 4 ? bl_launch_radiate_unpolarized_f4(&A, &ctx->rad, ctx->stream)
                 : bl_launch_radiate_unpolarized_f32(&A, &ctx->rad, ctx->stream);
  BL_CUDA_CHECK(le);
  ctx->launches++;
  return BL_OK;
}

// After the stream has been synchronised: add the device time of each stage of the last radiate_wave.
int collect_stage_times(bl_ctx *ctx, Level &L, int slabs) {
  for (int k = 0; k < 4 * slabs; k++) {
    float ms = 0.f;
    BL_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->stage_events[k], ctx->stage_events[k + 1]));
    L.ms_stage[k % 4] += ms;
  }
  return BL_OK;
}

int alloc_split_scratch(bl_ctx *ctx, Level &L, int slab) {
  L.slab = slab;
  if (slab <= 0) return BL_OK;
  const size_t nf = (size_t)bl_polarized_split_fields(ctx->rad.num_freq);
  BL_CUDA_CHECK(dev_alloc(&L.scratch, (size_t)L.wave_rays * (size_t)slab * nf));
  BL_CUDA_CHECK(dev_alloc(&L.frame, (size_t)L.wave_rays * (size_t)(slab + 1) * 8));
  BL_CUDA_CHECK(dev_alloc(&L.cam_map, (size_t)L.wave_rays * 10));
  return BL_OK;
}

}  // namespace

extern "C" {

}  // extern "C"

namespace {

// Common part of bl_trace_level and bl_trace_level_pixels: (re)allocate the level for num_rays rays, let `fill` put the
// camera arrays into L.cam_pos / L.cam_dir / L.mom (enqueued on the context's stream), size the waves, trace if resident.
template <class Fill>
int trace_level_impl(bl_ctx *ctx, int level, int64_t num_rays, bl_level_stats *stats, Fill fill) {
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  Level &L = ctx->levels[level];
  // any failure below leaves the level empty rather than half built
  struct Guard { Level &L; bool ok = false; ~Guard() { if (!ok) free_level(L); } } guard{L};
  // Device buffers are kept when the ray count is unchanged (time series, repeated renders): a
  // cudaFree/cudaMalloc pair of a ~100 GB step buffer costs far more than the kernels themselves.
  const bool reuse = L.rays == num_rays && num_rays > 0 && L.cam_pos && L.step;
  if (reuse) {
    cudaFree(L.tap_inds); cudaFree(L.tap_fracs); cudaFree(L.tap_nan); cudaFree(L.tap_cut); cudaFree(L.tap_fb);
    L.tap_inds = nullptr; L.tap_fracs = nullptr; L.tap_nan = L.tap_cut = L.tap_fb = nullptr; L.tap_S = 0;
    L.traced = false;
  } else {
    free_level(L);
  }
  L.rays = num_rays;
  if (num_rays == 0) { L.traced = true; L.resident = true; guard.ok = true; if (stats) *stats = L.stats; return BL_OK; }
  const int Q = ctx->rad.num_quantities, R = ctx->rad.render_num_images;
  if (!reuse) {
    BL_CUDA_CHECK(dev_alloc(&L.cam_pos, (size_t)num_rays * 4));
    BL_CUDA_CHECK(dev_alloc(&L.cam_dir, (size_t)num_rays * 4));
    BL_CUDA_CHECK(dev_alloc(&L.mom, (size_t)num_rays));
    BL_CUDA_CHECK(dev_alloc(&L.num, (size_t)num_rays));
    BL_CUDA_CHECK(dev_alloc(&L.flags, (size_t)num_rays));
    BL_CUDA_CHECK(dev_alloc(&L.image, (size_t)num_rays * (size_t)(Q > 0 ? Q : 1)));
    if (R > 0) BL_CUDA_CHECK(dev_alloc(&L.render, (size_t)num_rays * 3 * R));
  }
  {
    int rc = fill(L);
    if (rc) return rc;
  }

  if (!reuse) {
    // wave size from the HBM budget: 64 bytes per sample slot, ray_max_steps slots per ray
    size_t fr = 0, tot = 0;
    BL_CUDA_CHECK(cudaMemGetInfo(&fr, &tot));
    size_t per_ray = (size_t)ctx->params.ray_max_steps * StepBuffer::kRecord * sizeof(double);
    size_t budget = (size_t)((double)fr * 0.80);
    size_t per_ray_split = 0;
    const int slab = pol_split_slab(ctx, num_rays, budget, &per_ray_split);
    int64_t fit = (int64_t)(budget / (per_ray + per_ray_split));
    if (fit < 128) {
      free_level(L);
      return bl_fail(ctx, BL_ERR_NOMEM, "not enough free HBM for a 128-ray wave (%zu bytes per ray)", per_ray + per_ray_split);
    }
    // a requested wave size is honoured in whole 128-ray groups (at least one)
    if (ctx->params.tile_rays > 0 && ctx->params.tile_rays < fit) fit = ctx->params.tile_rays < 128 ? 128 : ctx->params.tile_rays;
    if (fit >= num_rays) {
      L.wave_rays = num_rays;
      L.resident = true;
    } else {
      // split into equal waves of whole 128-ray groups that never exceed the budget
      fit = fit / 128 * 128;
      int64_t waves = (num_rays + fit - 1) / fit;
      int64_t w = (num_rays + waves - 1) / waves;
      w = (w + 127) / 128 * 128;
      L.wave_rays = w < fit ? w : fit;
      L.resident = false;
    }
    BL_CUDA_CHECK(dev_alloc(&L.step, (size_t)L.wave_rays * per_ray / sizeof(double)));
    int rc = alloc_split_scratch(ctx, L, slab);
    if (rc) return rc;
    // length buckets of the ray ordering: the pipeline's slabs, else 64 groups over the step capacity
    L.order_unit = slab > 0 ? slab : (ctx->params.ray_max_steps + 63) / 64;
    L.order_buckets = ctx->params.ray_max_steps / L.order_unit + 2;
    if (!ctx->ray_order || L.order_buckets > bl_ray_order_max_buckets()) L.order_buckets = 0;
    if (L.order_buckets > 0) {
      BL_CUDA_CHECK(dev_alloc(&L.order, (size_t)L.wave_rays));
      const int ws_buckets = L.order_buckets > kImpactBuckets ? L.order_buckets : kImpactBuckets;
      BL_CUDA_CHECK(dev_alloc(&L.order_ws, bl_ray_order_workspace(L.wave_rays, ws_buckets)));
    }
  }
  BL_CUDA_CHECK(cudaMemsetAsync(ctx->counters, 0, sizeof(GeoCounters), ctx->stream));
  L.stats = bl_level_stats();
  if (L.resident) {
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = trace_wave(ctx, L, 0, num_rays);
    if (rc) return rc;
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = read_geo_counters(ctx, L);
    if (rc) return rc;
    float ms = 0.f;
    BL_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    L.stats.ms_geodesic = ms;
    L.traced = true;
  } else {
    // count-only information is produced by the first radiate pass; report what is known
    L.stats.num_rays = num_rays;
    L.traced = false;
    BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  guard.ok = true;
  if (stats) *stats = L.stats;
  return BL_OK;
}

}  // namespace

extern "C" {

int bl_trace_level(bl_ctx *ctx, int level, const double *cam_pos, const double *cam_dir, const double *mom_factor,
                   int64_t num_rays, bl_level_stats *stats) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level: level %d out of range", level);
  if (!cam_pos || !cam_dir || !mom_factor || num_rays < 0) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level: null camera arrays");
  return trace_level_impl(ctx, level, num_rays, stats, [&](Level &L) -> int {
    BL_CUDA_CHECK(cudaMemcpyAsync(L.cam_pos, cam_pos, (size_t)num_rays * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BL_CUDA_CHECK(cudaMemcpyAsync(L.cam_dir, cam_dir, (size_t)num_rays * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BL_CUDA_CHECK(cudaMemcpyAsync(L.mom, mom_factor, (size_t)num_rays * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return BL_OK;
  });
}

int bl_set_camera(bl_ctx *ctx, const bl_camera *camera) {
  if (!ctx || !camera) return BL_ERR_ARG;
  if (camera->type != BL_CAMERA_PLANE && camera->type != BL_CAMERA_PINHOLE) return bl_fail(ctx, BL_ERR_ARG, "bl_set_camera: unknown camera type %d", camera->type);
  if (camera->normalization != BL_NORM_CAMERA && camera->normalization != BL_NORM_INFINITY)
    return bl_fail(ctx, BL_ERR_ARG, "bl_set_camera: unknown image normalization %d", camera->normalization);
  CameraDev &c = ctx->camera;
  c.type = camera->type;
  c.normalization = camera->normalization;
  c.flat = ctx->params.ray_flat ? 1 : 0;
  c.pad = 0;
  c.a = ctx->params.bh_a;
  c.width = camera->width;
  c.r = camera->r;
  for (int m = 0; m < 4; m++) {
    c.x[m] = camera->x[m];
    c.u_con[m] = camera->u_con[m];
    c.u_cov[m] = camera->u_cov[m];
    c.norm_con[m] = camera->norm_con[m];
    c.norm_con_c[m] = camera->norm_con_c[m];
    c.hor_con_c[m] = camera->hor_con_c[m];
    c.vert_con_c[m] = camera->vert_con_c[m];
  }
  ctx->have_camera = true;
  return BL_OK;
}

int bl_trace_level_pixels(bl_ctx *ctx, int level, int kind, const int32_t *units, int64_t num_units, bl_level_stats *stats) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level_pixels: level %d out of range", level);
  if (!ctx->have_camera) return bl_fail(ctx, BL_ERR_STATE, "bl_trace_level_pixels: no camera (bl_set_camera)");
  if (kind != BL_PIXELS_ROWS && kind != BL_PIXELS_BLOCKS) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level_pixels: unknown unit kind %d", kind);
  if (num_units < 0 || (kind == BL_PIXELS_BLOCKS && !units && num_units > 0)) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level_pixels: null unit list");
  const int64_t eff_res64 = (int64_t)ctx->params.camera_resolution << level;
  if (ctx->params.camera_resolution <= 0 || eff_res64 > 0x7fffffff) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level_pixels: effective resolution out of range");
  const int eff_res = (int)eff_res64;
  const int bs = ctx->params.adaptive_block_size;
  if (kind == BL_PIXELS_BLOCKS && (bs <= 0 || eff_res % bs != 0)) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level_pixels: adaptive_block_size must divide the effective resolution");
  // the unit list is validated here, on the host: a bad entry would only move a pixel, never fault, but it is a caller bug
  const int64_t per_unit = kind == BL_PIXELS_ROWS ? eff_res : (int64_t)bs * bs;
  const int64_t limit = kind == BL_PIXELS_ROWS ? eff_res : eff_res / bs;
  const int64_t entries = kind == BL_PIXELS_ROWS ? num_units : 2 * num_units;
  if (!units && num_units > eff_res) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level_pixels: %lld rows of a %d-row raster", (long long)num_units, eff_res);
  if (units)
    for (int64_t q = 0; q < entries; q++)
      if (units[q] < 0 || units[q] >= limit) return bl_fail(ctx, BL_ERR_ARG, "bl_trace_level_pixels: unit entry %d outside [0, %lld)", units[q], (long long)limit);
  const int64_t num_rays = num_units * per_unit;
  return trace_level_impl(ctx, level, num_rays, stats, [&](Level &L) -> int {
    const int32_t *units_dev = nullptr;
    if (units && entries > 0) {
      if ((size_t)entries > ctx->units_cap) {
        cudaFree(ctx->units_dev);
        ctx->units_dev = nullptr;
        ctx->units_cap = 0;
        BL_CUDA_CHECK(dev_alloc(&ctx->units_dev, (size_t)entries));
        ctx->units_cap = (size_t)entries;
      }
      BL_CUDA_CHECK(cudaMemcpyAsync(ctx->units_dev, units, (size_t)entries * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      units_dev = ctx->units_dev;
    }
    BL_CUDA_CHECK(bl_launch_camera_pixels(&ctx->camera, kind, units_dev, eff_res, bs, num_rays, L.cam_pos, L.cam_dir, L.mom, ctx->stream));
    ctx->launches++;
    return BL_OK;
  });
}

int bl_download_camera(bl_ctx *ctx, int level, double *cam_pos, double *cam_dir, double *mom_factor) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_download_camera: level %d out of range", level);
  Level &L = ctx->levels[level];
  if (L.rays > 0 && !L.cam_pos) return bl_fail(ctx, BL_ERR_STATE, "bl_download_camera: level %d has no camera arrays", level);
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (L.rays > 0) {
    if (cam_pos) BL_CUDA_CHECK(cudaMemcpyAsync(cam_pos, L.cam_pos, (size_t)L.rays * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (cam_dir) BL_CUDA_CHECK(cudaMemcpyAsync(cam_dir, L.cam_dir, (size_t)L.rays * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (mom_factor) BL_CUDA_CHECK(cudaMemcpyAsync(mom_factor, L.mom, (size_t)L.rays * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  }
  return BL_OK;
}

int bl_upload_samples(bl_ctx *ctx, int level, const double *cam_pos, const double *cam_dir, const double *mom_factor,
                      int64_t num_rays, int32_t S, const uint8_t *flags, const int32_t *num, const double *pos,
                      const double *dir, const double *len, bl_level_stats *stats) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_upload_samples: level %d out of range", level);
  if (!cam_pos || !cam_dir || !mom_factor || !flags || !num || !pos || !dir || !len || num_rays <= 0 || S <= 0)
    return bl_fail(ctx, BL_ERR_ARG, "bl_upload_samples: null or empty argument");
  if (S > ctx->params.ray_max_steps) return bl_fail(ctx, BL_ERR_ARG, "bl_upload_samples: %d samples per ray exceed ray_max_steps = %d", S, ctx->params.ray_max_steps);
  for (int64_t m = 0; m < num_rays; m++)
    if (num[m] < 0 || num[m] > S) return bl_fail(ctx, BL_ERR_ARG, "bl_upload_samples: sample count %d of ray %lld outside [0, %d]", num[m], (long long)m, S);
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  Level &L = ctx->levels[level];
  free_level(L);
  L.rays = num_rays;
  const int Q = ctx->rad.num_quantities, R = ctx->rad.render_num_images;
  BL_CUDA_CHECK(dev_alloc(&L.cam_pos, (size_t)num_rays * 4));
  BL_CUDA_CHECK(dev_alloc(&L.cam_dir, (size_t)num_rays * 4));
  BL_CUDA_CHECK(dev_alloc(&L.mom, (size_t)num_rays));
  BL_CUDA_CHECK(dev_alloc(&L.num, (size_t)num_rays));
  BL_CUDA_CHECK(dev_alloc(&L.flags, (size_t)num_rays));
  BL_CUDA_CHECK(dev_alloc(&L.image, (size_t)num_rays * (size_t)(Q > 0 ? Q : 1)));
  if (R > 0) BL_CUDA_CHECK(dev_alloc(&L.render, (size_t)num_rays * 3 * R));
  size_t per_ray = (size_t)ctx->params.ray_max_steps * StepBuffer::kRecord * sizeof(double);
  size_t fr = 0, tot = 0;
  BL_CUDA_CHECK(cudaMemGetInfo(&fr, &tot));
  size_t per_ray_split = 0;
  const int slab = pol_split_slab(ctx, num_rays, (size_t)(0.80 * (double)fr), &per_ray_split);
  if ((double)(per_ray + per_ray_split) * (double)num_rays > 0.80 * (double)fr)
    return bl_fail(ctx, BL_ERR_NOMEM, "bl_upload_samples: the level's step buffer (%zu bytes per ray) does not fit in HBM; uploaded geodesics must be resident", per_ray);
  L.wave_rays = num_rays;
  L.resident = true;
  BL_CUDA_CHECK(dev_alloc(&L.step, (size_t)num_rays * per_ray / sizeof(double)));
  {
    int rc = alloc_split_scratch(ctx, L, slab);
    if (rc) return rc;
  }
  BL_CUDA_CHECK(cudaMemcpyAsync(L.cam_pos, cam_pos, (size_t)num_rays * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  BL_CUDA_CHECK(cudaMemcpyAsync(L.cam_dir, cam_dir, (size_t)num_rays * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  BL_CUDA_CHECK(cudaMemcpyAsync(L.mom, mom_factor, (size_t)num_rays * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  BL_CUDA_CHECK(cudaMemcpyAsync(L.num, num, (size_t)num_rays * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  BL_CUDA_CHECK(cudaMemcpyAsync(L.flags, flags, (size_t)num_rays, cudaMemcpyHostToDevice, ctx->stream));
  // stage the (N, S, .) host arrays through a bounded device buffer, a chunk of rays at a time
  int64_t chunk = (int64_t)((size_t)64 << 20) / ((size_t)S * 9 * sizeof(double));
  if (chunk < 1) chunk = 1;
  if (chunk > num_rays) chunk = num_rays;
  double *spos = nullptr, *sdir = nullptr, *slen = nullptr;
  BL_CUDA_CHECK(dev_alloc(&spos, (size_t)chunk * S * 4));
  BL_CUDA_CHECK(dev_alloc(&sdir, (size_t)chunk * S * 4));
  BL_CUDA_CHECK(dev_alloc(&slen, (size_t)chunk * S));
  StepBuffer sb; sb.buf = L.step; sb.rays = num_rays; sb.cap = ctx->params.ray_max_steps;
  cudaError_t e = cudaSuccess;
  for (int64_t m0 = 0; m0 < num_rays && e == cudaSuccess; m0 += chunk) {
    int64_t cnt = num_rays - m0 < chunk ? num_rays - m0 : chunk;
    size_t o = (size_t)m0 * S;
    e = cudaMemcpyAsync(spos, pos + 4 * o, (size_t)cnt * S * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sdir, dir + 4 * o, (size_t)cnt * S * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(slen, len + o, (size_t)cnt * S * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = bl_launch_pack_samples(&sb, L.num, m0, cnt, S, spos, sdir, slen, ctx->stream);
    ctx->launches++;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(spos); cudaFree(sdir); cudaFree(slen);
  if (e != cudaSuccess) return bl_fail_cuda(ctx, e, "sample upload", __FILE__, __LINE__);
  L.stats = bl_level_stats();
  L.stats.num_rays = num_rays;
  L.stats.geodesic_num_steps = S;
  for (int64_t m = 0; m < num_rays; m++) {
    L.stats.num_samples += num[m];
    L.stats.num_bad_geodesics += flags[m] ? 1 : 0;
  }
  L.traced = true;
  if (stats) *stats = L.stats;
  return BL_OK;
}

long long bl_launch_count(const bl_ctx *ctx) { return ctx ? ctx->launches : -1; }

int bl_device_image(bl_ctx *ctx, int level, void **image, int64_t *num_rays) {
  if (!ctx || !image) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_device_image: level %d out of range", level);
  const Level &L = ctx->levels[level];
  if (!L.image) return bl_fail(ctx, BL_ERR_STATE, "bl_device_image: level %d has no image", level);
  *image = L.image;
  if (num_rays) *num_rays = L.rays;
  return BL_OK;
}

int bl_download_polarized_scratch(bl_ctx *ctx, int level, double *out, double *cam_map, int64_t *num_fields, int64_t *slab,
                                  int64_t *num_rays) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_download_polarized_scratch: level %d out of range", level);
  Level &L = ctx->levels[level];
  if (!L.scratch || L.slab <= 0 || !L.resident)
    return bl_fail(ctx, BL_ERR_STATE, "bl_download_polarized_scratch: level %d was not rendered by the slab pipeline as one wave", level);
  const int64_t nf = bl_polarized_split_fields(ctx->rad.num_freq);
  if (num_fields) *num_fields = nf;
  if (slab) *slab = L.slab;
  if (num_rays) *num_rays = L.wave_rays;
  if (out) {
    BL_CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t rows = (size_t)nf * (size_t)L.slab, rays = (size_t)L.wave_rays;
    if (L.order_alive0 < 0) {
      BL_CUDA_CHECK(cudaMemcpy(out, L.scratch, rows * rays * sizeof(double), cudaMemcpyDeviceToHost));
    } else {
      // the pipeline addressed the scratch by position in the sorted ray list: hand it out by ray
      std::vector<double> tmp(rows * rays);
      std::vector<int32_t> order(rays);
      BL_CUDA_CHECK(cudaMemcpy(tmp.data(), L.scratch, rows * rays * sizeof(double), cudaMemcpyDeviceToHost));
      BL_CUDA_CHECK(cudaMemcpy(order.data(), L.order, rays * sizeof(int32_t), cudaMemcpyDeviceToHost));
      std::memset(out, 0, rows * rays * sizeof(double));
      for (size_t r = 0; r < rows; r++)
        for (size_t i = 0; i < (size_t)L.order_alive0; i++) out[r * rays + (size_t)order[i]] = tmp[r * rays + i];
    }
  }
  if (cam_map) {
    BL_CUDA_CHECK(cudaSetDevice(ctx->device));
    BL_CUDA_CHECK(cudaMemcpy(cam_map, L.cam_map, (size_t)10 * (size_t)L.wave_rays * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return BL_OK;
}

int bl_polarized_stage_ms(bl_ctx *ctx, int level, double *ms3, int32_t *slab) {
  if (!ctx || !ms3) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_polarized_stage_ms: level %d out of range", level);
  const Level &L = ctx->levels[level];
  for (int k = 0; k < 3; k++) ms3[k] = L.ms_stage[k + 1];
  if (slab) *slab = (L.slab > 0 && L.scratch && !L.tap_nan) ? L.slab : 0;
  return BL_OK;
}

int bl_polarized_sampling_ms(bl_ctx *ctx, int level, double *ms) {
  if (!ctx || !ms) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_polarized_sampling_ms: level %d out of range", level);
  *ms = ctx->levels[level].ms_stage[0];
  return BL_OK;
}

void *bl_cuda_stream(const bl_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int bl_retrace_level(bl_ctx *ctx, int level, bl_level_stats *stats) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_retrace_level: level %d out of range", level);
  Level &L = ctx->levels[level];
  if (!L.cam_pos) return bl_fail(ctx, BL_ERR_STATE, "bl_retrace_level: level %d has no camera arrays", level);
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (L.resident) {
    BL_CUDA_CHECK(cudaMemsetAsync(ctx->counters, 0, sizeof(GeoCounters), ctx->stream));
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = trace_wave(ctx, L, 0, L.rays);
    if (rc) return rc;
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = read_geo_counters(ctx, L);
    if (rc) return rc;
    float ms = 0.f;
    BL_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    L.stats.ms_geodesic = ms;
    L.traced = true;
  }
  if (stats) *stats = L.stats;
  return BL_OK;
}

int bl_radiate_level(bl_ctx *ctx, int level, int snapshot, double *image, double *render, bl_level_stats *stats) {
  (void)snapshot;
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_radiate_level: level %d out of range", level);
  Level &L = ctx->levels[level];
  if (L.rays > 0 && (!L.cam_pos || !L.step || L.wave_rays <= 0))
    return bl_fail(ctx, BL_ERR_STATE, "bl_radiate_level: level %d has not been traced", level);
  bool sim = ctx->params.model_type == BL_MODEL_SIMULATION;
  if (sim && !ctx->have_grid) return bl_fail(ctx, BL_ERR_STATE, "bl_radiate_level: no grid uploaded");
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  const int Q = ctx->rad.num_quantities, R = ctx->rad.render_num_images;
  if (L.rays == 0) { if (stats) *stats = L.stats; return BL_OK; }
  if (ctx->taps_enabled && sim && !L.tap_nan) {
    if (!L.resident) return bl_fail(ctx, BL_ERR_STATE, "parity taps need a resident level (reduce rays or raise tile_rays)");
    L.tap_S = L.stats.geodesic_num_steps > 0 ? L.stats.geodesic_num_steps : 1;
    size_t ns = (size_t)L.rays * L.tap_S;
    BL_CUDA_CHECK(dev_alloc(&L.tap_inds, ns * 4));
    BL_CUDA_CHECK(cudaMemsetAsync(L.tap_inds, 0xff, ns * 4 * sizeof(int32_t), ctx->stream));
    if (ctx->params.simulation_interp) {
      BL_CUDA_CHECK(dev_alloc(&L.tap_fracs, ns * 3));
      BL_CUDA_CHECK(cudaMemsetAsync(L.tap_fracs, 0, ns * 3 * sizeof(double), ctx->stream));
    }
    BL_CUDA_CHECK(dev_alloc(&L.tap_nan, ns)); BL_CUDA_CHECK(dev_alloc(&L.tap_cut, ns)); BL_CUDA_CHECK(dev_alloc(&L.tap_fb, ns));
    BL_CUDA_CHECK(cudaMemsetAsync(L.tap_nan, 0, ns, ctx->stream));
    BL_CUDA_CHECK(cudaMemsetAsync(L.tap_cut, 0, ns, ctx->stream));
    BL_CUDA_CHECK(cudaMemsetAsync(L.tap_fb, 0, ns, ctx->stream));
  }
  BL_CUDA_CHECK(cudaMemsetAsync(ctx->rad_counter, 0, sizeof(unsigned long long), ctx->stream));
  if (ctx->rad.slow_light) {
    if (ctx->rad.slow_count != ctx->params.slow_chunk_size)
      return bl_fail(ctx, BL_ERR_STATE, "bl_radiate_level: slow light needs bl_set_time_window before radiating");
    BL_CUDA_CHECK(cudaMemsetAsync(ctx->slow_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
  }
  double ms_geo = 0.0, ms_rad = 0.0;
  float ms = 0.f;
  int split_slabs = 0;
  L.ms_stage[0] = L.ms_stage[1] = L.ms_stage[2] = L.ms_stage[3] = 0.0;
  if (L.resident) {
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = radiate_wave(ctx, L, 0, L.rays, L.stats.geodesic_num_steps, &split_slabs);
    if (rc) return rc;
    BL_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    BL_CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
    BL_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ms_rad = ms;
    rc = collect_stage_times(ctx, L, split_slabs);
    if (rc) return rc;
  } else {
    // waves: trace then radiate, reusing one step buffer; geodesic statistics accumulate over waves
    BL_CUDA_CHECK(cudaMemsetAsync(ctx->counters, 0, sizeof(GeoCounters), ctx->stream));
    for (int64_t first = 0; first < L.rays; first += L.wave_rays) {
      int64_t count = L.rays - first < L.wave_rays ? L.rays - first : L.wave_rays;
      BL_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
      int rc = trace_wave(ctx, L, first, count);
      if (rc) return rc;
      BL_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
      BL_CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
      BL_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
      ms_geo += ms;
      // sample counts so far bound this wave's (the maximum only grows): where the polarized pipeline's slabs start
      int s_top = ctx->params.ray_max_steps;
      if (L.slab > 0) {
        BL_CUDA_CHECK(cudaMemcpyAsync(&s_top, &ctx->counters->max_samples, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
      }
      BL_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
      rc = radiate_wave(ctx, L, first, count, s_top, &split_slabs);
      if (rc) return rc;
      BL_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
      BL_CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
      BL_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
      ms_rad += ms;
      rc = collect_stage_times(ctx, L, split_slabs);
      if (rc) return rc;
    }
    int rc = read_geo_counters(ctx, L);
    if (rc) return rc;
    L.stats.ms_geodesic = ms_geo;
    L.traced = true;
  }
  L.stats.ms_radiation = ms_rad;
  if (image && Q > 0)
    BL_CUDA_CHECK(cudaMemcpyAsync(image, L.image, (size_t)L.rays * Q * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (render && R > 0 && L.render)
    BL_CUDA_CHECK(cudaMemcpyAsync(render, L.render, (size_t)L.rays * 3 * R * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long slow_raw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ctx->rad.slow_light)
    BL_CUDA_CHECK(cudaMemcpyAsync(slow_raw, ctx->slow_counters, sizeof slow_raw, cudaMemcpyDeviceToHost, ctx->stream));
  BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  L.slow = bl_slow_stats();
  if (ctx->rad.slow_light)
    for (int side = 0; side < 2; side++) {
      L.slow.num_small[side] = (int64_t)slow_raw[2 * side];
      L.slow.num_large[side] = (int64_t)slow_raw[2 * side + 1];
      std::memcpy(&L.slow.val_small[side], &slow_raw[4 + 2 * side], sizeof(double));
      std::memcpy(&L.slow.val_large[side], &slow_raw[4 + 2 * side + 1], sizeof(double));
    }
  if (stats) *stats = L.stats;
  return BL_OK;
}

int bl_refine_level(bl_ctx *ctx, int level, const int32_t *block_locs, int64_t num_blocks, uint8_t *flags, int64_t *n_refined) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_refine_level: level %d out of range", level);
  if (!block_locs || !flags || num_blocks < 0) return bl_fail(ctx, BL_ERR_ARG, "bl_refine_level: null argument");
  Level &L = ctx->levels[level];
  const bl_params &p = ctx->params;
  if (p.adaptive_max_level <= 0) return bl_fail(ctx, BL_ERR_STATE, "bl_refine_level: adaptive_max_level is 0");
  int64_t bs2 = (int64_t)p.adaptive_block_size * p.adaptive_block_size;
  if (num_blocks * bs2 != L.rays || !L.image) return bl_fail(ctx, BL_ERR_STATE, "bl_refine_level: level %d image has %lld rays, expected %lld blocks x %lld", level, (long long)L.rays, (long long)num_blocks, (long long)bs2);
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (n_refined) *n_refined = 0;
  if (num_blocks == 0) return BL_OK;
  if ((size_t)num_blocks > ctx->refine_cap) {   // grow-only buffers: no allocation per call
    cudaFree(ctx->refine_locs); cudaFree(ctx->refine_flags);
    ctx->refine_locs = nullptr; ctx->refine_flags = nullptr; ctx->refine_cap = 0;
    BL_CUDA_CHECK(dev_alloc(&ctx->refine_locs, (size_t)num_blocks * 2));
    BL_CUDA_CHECK(dev_alloc(&ctx->refine_flags, (size_t)num_blocks));
    ctx->refine_cap = (size_t)num_blocks;
  }
  int32_t *locs_dev = ctx->refine_locs;
  uint8_t *flags_dev = ctx->refine_flags;
  cudaError_t e = cudaMemcpyAsync(locs_dev, block_locs, (size_t)num_blocks * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->ev0, ctx->stream);
  if (e == cudaSuccess) e = bl_launch_refine(L.image, L.rays, level, locs_dev, num_blocks, ctx->params_dev, flags_dev, ctx->stream);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->ev1, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(flags, flags_dev, (size_t)num_blocks, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  if (e != cudaSuccess) return bl_fail_cuda(ctx, e, "refine", __FILE__, __LINE__);
  L.stats.ms_refine = ms;
  int64_t cnt = 0;
  for (int64_t b = 0; b < num_blocks; b++) cnt += flags[b] ? 1 : 0;
  if (n_refined) *n_refined = cnt;
  return BL_OK;
}

int bl_download_samples(bl_ctx *ctx, int level, uint8_t *flags, int32_t *num, double *pos, double *dir, double *len) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_download_samples: level %d out of range", level);
  Level &L = ctx->levels[level];
  if (!L.traced) return bl_fail(ctx, BL_ERR_STATE, "bl_download_samples: level %d not traced", level);
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (L.rays == 0) return BL_OK;
  if (flags) BL_CUDA_CHECK(cudaMemcpyAsync(flags, L.flags, (size_t)L.rays, cudaMemcpyDeviceToHost, ctx->stream));
  if (num) BL_CUDA_CHECK(cudaMemcpyAsync(num, L.num, (size_t)L.rays * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (pos || dir || len) {
    if (!L.resident) return bl_fail(ctx, BL_ERR_STATE, "bl_download_samples: level %d is traced in waves; samples are not resident", level);
    int S = L.stats.geodesic_num_steps;
    size_t ns = (size_t)L.rays * (size_t)(S > 0 ? S : 1);
    double *dpos = nullptr, *ddir = nullptr, *dlen = nullptr;
    if (pos) BL_CUDA_CHECK(dev_alloc(&dpos, ns * 4));
    if (dir) BL_CUDA_CHECK(dev_alloc(&ddir, ns * 4));
    if (len) BL_CUDA_CHECK(dev_alloc(&dlen, ns));
    StepBuffer sb; sb.buf = L.step; sb.rays = L.rays; sb.cap = ctx->params.ray_max_steps;
    cudaError_t e = bl_launch_unpack_samples(&sb, L.num, L.cam_dir, L.rays, S, dpos, ddir, dlen, ctx->stream);
    if (e == cudaSuccess && pos) e = cudaMemcpyAsync(pos, dpos, ns * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && dir) e = cudaMemcpyAsync(dir, ddir, ns * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && len) e = cudaMemcpyAsync(len, dlen, ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dpos); cudaFree(ddir); cudaFree(dlen);
    if (e != cudaSuccess) return bl_fail_cuda(ctx, e, "sample download", __FILE__, __LINE__);
  }
  BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return BL_OK;
}

int bl_download_sample_inds(bl_ctx *ctx, int level, int32_t *inds, double *fracs, uint8_t *nan_, uint8_t *cut, uint8_t *fallback) {
  if (!ctx) return BL_ERR_ARG;
  if (level < 0 || level >= (int)ctx->levels.size()) return bl_fail(ctx, BL_ERR_ARG, "bl_download_sample_inds: level %d out of range", level);
  Level &L = ctx->levels[level];
  if (!L.tap_nan) return bl_fail(ctx, BL_ERR_STATE, "bl_download_sample_inds: call bl_set_taps(ctx, 1) before bl_radiate_level");
  BL_CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t ns = (size_t)L.rays * L.tap_S;
  if (inds) BL_CUDA_CHECK(cudaMemcpyAsync(inds, L.tap_inds, ns * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (fracs && L.tap_fracs) BL_CUDA_CHECK(cudaMemcpyAsync(fracs, L.tap_fracs, ns * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (nan_) BL_CUDA_CHECK(cudaMemcpyAsync(nan_, L.tap_nan, ns, cudaMemcpyDeviceToHost, ctx->stream));
  if (cut) BL_CUDA_CHECK(cudaMemcpyAsync(cut, L.tap_cut, ns, cudaMemcpyDeviceToHost, ctx->stream));
  if (fallback) BL_CUDA_CHECK(cudaMemcpyAsync(fallback, L.tap_fb, ns, cudaMemcpyDeviceToHost, ctx->stream));
  BL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return BL_OK;
}

}  // extern "C"
