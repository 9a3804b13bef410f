// Device-side parameter blocks of the radiation kernels.
#pragma once
#include <stdint.h>
#include "device_types.cuh"

#define RAD_MAX_FREQ 32
#define RAD_MAX_SLICES 64
#define RAD_MAX_FEATURES 64
#define RAD_NUM_CELL_VALUES 7

// CGS constants (reference src/blacklight.hpp:18-27)
namespace phys {
constexpr double pi = 3.141592653589793;
constexpr double sqrt2 = 1.4142135623730951;
constexpr double c = 2.99792458e10;
constexpr double h = 6.62607015e-27;
constexpr double k_b = 1.380649e-16;
constexpr double m_p = 1.67262192369e-24;
constexpr double m_e = 9.1093837015e-28;
constexpr double e = 4.80320425e-10;
constexpr double gg_msun = 1.32712440018e26;
}  // namespace phys

// Grid resident in HBM.  Cell primitives are re-laid out at upload from the reader's
// (var, b, k, j, i) planes into one 32-byte record per cell, so a trilinear gather touches
// 8 fully used 32-byte sectors instead of 32 quarter-used ones:
//   cells[2*c + 0] = (rho, pgas, uu1, uu2), cells[2*c + 1] = (uu3, bb1, bb2, bb3),
//   c = ((b*n_k + k)*n_j + j)*n_i + i
struct GridDev {
  int32_t n_b, n_k, n_j, n_i;
  const double *x1f, *x2f, *x3f;  // (n_b, n+1)
  const double *x1v, *x2v, *x3v;  // (n_b, n)
  const double *bounds;           // (n_b, 6): x1min, x1max, x2min, x2max, x3min, x3max
  const double *x1d, *x2d, *x3d;  // (n_b, n): 1 / (xv[i+1] - xv[i]) for i < n-1 (interpolation weights)
  const float4 *cells;
  const float *kappa;             // (n_b, n_k, n_j, n_i) electron entropy, or nullptr
  // slow light: `cells` / `kappa` hold several snapshots back to back, slot s at cells + s * slice_cells * 2
  // and kappa + s * slice_cells
  size_t slice_cells;
  // mesh topology for inter-block interpolation (simulation_block_interp): refinement level and logical
  // location of every block, and an open-addressing hash (level, location) -> block built at upload, which
  // replaces the reference's linear scans over all blocks (simulation_sampling.cpp:1068-1321)
  const int32_t *levels;          // (n_b)
  const int32_t *locs;            // (n_b, 3): x1, x2, x3 logical locations
  const unsigned long long *hash_keys;
  const int32_t *hash_vals;
  uint32_t hash_mask;             // table size - 1 (power of two)
  int32_t max_level;
  int32_t n3_root;                // blocks around x3 at level 0 (RootGridSize[2] / n_k)
  // simulation_coord = fmks: x1f, x2f stay native; (r, theta) -> native (x1, x2) through the reader's table
  // sks_map(0|1, j, i) sampled at r = map_r_in + i map_dr, theta = j map_dtheta (simulation_geometry.cpp:330-413)
  const double *sks_map;          // (2, map_n2, map_n1) or nullptr
  int32_t map_n1, map_n2;
  double map_r_in, map_dr, map_dtheta;
  // The reference indexes one zone past the last row / plane there; in its (variable, cell) array that addresses the
  // following cells and, past a variable's last cell, the first cells of the variable stored after it.  next_slot[q]:
  // record slot (0-7; 8 = kappa) holding the variable that follows slot q in the reader's array, -1 = none (read as 0).
  int8_t next_slot[9];
};

// key of a block in the topology hash; locations are < 2^19 on every level that can occur
__host__ __device__ inline unsigned long long block_key(int level, int li, int lj, int lk) {
  return ((unsigned long long)level << 57) | ((unsigned long long)li << 38) | ((unsigned long long)lj << 19) |
         (unsigned long long)lk;
}
__host__ __device__ inline uint32_t block_hash(unsigned long long key) {
  key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ull; key ^= key >> 33;
  return (uint32_t)key;
}

struct RadParams {
  int32_t model_type, ray_flat, coord, interp, block_interp;
  double a, camera_r;
  double camera_x[4];
  int32_t num_freq;
  double freqs[RAD_MAX_FREQ];
  double x_unit, t_unit;
  // image selection and slot offsets (reference radiation_integrator.cpp:436-520)
  int32_t image_light, image_time, image_length, image_lambda, image_emission, image_tau;
  int32_t image_lambda_ave, image_emission_ave, image_tau_int, image_crossings, polarization;
  int32_t off_time, off_length, off_lambda, off_emission, off_tau, off_lambda_ave;
  int32_t off_emission_ave, off_tau_int, off_crossings, num_quantities;
  int32_t need_cell_values;
  // units and plasma model
  double d_unit, e_unit, b_unit;
  double plasma_mu, plasma_ne_ni;
  int32_t plasma_model, plasma_use_p;
  double plasma_gamma, plasma_gamma_i, plasma_gamma_e, plasma_rat_low, plasma_rat_high;
  double thermal_frac, power_frac, kappa_frac;
  double plasma_p, plasma_gamma_min, plasma_gamma_max, plasma_kappa, plasma_w;
  // precomputed distribution constants (reference simulation_coefficients.cpp:53-193)
  double power_jj, power_aa, power_jj_q, power_jj_v, power_aa_q, power_aa_v;
  double power_rho, power_rho_q, power_rho_v;
  double kappa_jj_low, kappa_jj_high, kappa_jj_x_i, kappa_aa_low, kappa_aa_high, kappa_aa_x_i;
  double kappa_jj_low_q, kappa_jj_low_v, kappa_jj_high_q, kappa_jj_high_v, kappa_jj_x_q, kappa_jj_x_v;
  double kappa_aa_low_q, kappa_aa_low_v, kappa_aa_high_i, kappa_aa_high_q, kappa_aa_high_v;
  double kappa_aa_x_q, kappa_aa_x_v, kappa_rho_v, kappa_rho_frac;
  double kappa_rho_q_low_a, kappa_rho_q_low_b, kappa_rho_q_low_c, kappa_rho_q_low_d, kappa_rho_q_low_e;
  double kappa_rho_q_high_a, kappa_rho_q_high_b, kappa_rho_q_high_c, kappa_rho_q_high_d, kappa_rho_q_high_e;
  double kappa_rho_v_low_a, kappa_rho_v_low_b, kappa_rho_v_high_a, kappa_rho_v_high_b;
  // formula model
  double formula_r0, formula_h, formula_l0, formula_q, formula_nup, formula_cn0;
  double formula_alpha, formula_a, formula_beta;
  // cuts
  double cut_rho_min, cut_rho_max, cut_n_e_min, cut_n_e_max, cut_p_gas_min, cut_p_gas_max;
  double cut_theta_e_min, cut_theta_e_max, cut_b_min, cut_b_max, cut_sigma_min, cut_sigma_max;
  double cut_beta_inverse_min, cut_beta_inverse_max;
  int32_t cut_omit_near, cut_omit_far, cut_plane;
  double cut_omit_in, cut_omit_out, cut_midplane_theta, cut_midplane_z;
  double cut_plane_origin[3], cut_plane_normal[3];
  // fallback
  int32_t fallback_nan;
  float fallback_rho, fallback_pgas, fallback_kappa;
  // polarized extras
  int32_t rotation_split;
  double camera_u_con[4], camera_u_cov[4], camera_vert_con_c[4];
  // derived on the host (abi.cu fill_rad_params): reciprocals and logarithms hoisted out of the kernels
  double inv_freqs[RAD_MAX_FREQ];   // 1 / image_frequencies[l]
  double log_freqs[RAD_MAX_FREQ];   // ln image_frequencies[l]
  double n_e_factor;                // n_e_cgs = rho_cgs * n_e_factor
  int32_t any_value_cut;            // any of the cut_{rho,...,beta_inverse}_{min,max} enabled
  // slow light: the resident time window (index 0 = latest snapshot), see bl_set_time_window
  int32_t slow_light, slow_interp, slow_count;
  int32_t slow_slot[RAD_MAX_SLICES];
  double slow_time[RAD_MAX_SLICES];
  double snapshot_time, extrap_tol;
  // logarithms of distribution constants: powers of per-sample quantities are evaluated as exp(c * ln x)
  // with the logarithms shared between all exponents and frequencies
  double log_w2k2;                  // ln(w^2 kappa^2)
  double log_kjl, log_kjh, log_kal, log_kah;   // ln of kappa_jj_low, _high, kappa_aa_low, kappa_aa_high*kappa_aa_high_i
  double log_k_j_pref, log_k_a_pref;           // ln(kappa_frac e^2 / c), ln(kappa_frac e^2 / (m_e c))
  double log_kah_base;                         // ln kappa_aa_high (without the Stokes-I factor)
  double log_kj_low_q, log_kj_low_v, log_kj_high_q, log_kj_high_v;   // ln kappa_jj_{low,high}_{q,v}
  double log_ka_low_q, log_ka_low_v, log_ka_high_q, log_ka_high_v;   // ln kappa_aa_{low,high}_{q,v}
  double log_power_gmin;                       // ln(2 gamma_min^2 / 3)
  // Term-major evaluation of the polarized kappa coefficients over the image frequencies (pol_common.cuh:
  // kappa_polarized_all).  Every bridged coefficient t = j_I, j_Q, j_V, alpha_I, alpha_Q, alpha_V is
  // (lo^-x + hi^-x)^(-1/x) with ln lo, ln hi affine in ln nu: slopes kappa_slope_lo/hi[t], so that between image
  // frequencies only the host constants kappa_k[t][l] = exp(-x_t (slope_lo - slope_hi) ln(nu_l / nu_0)) and their
  // reciprocals change; likewise the pure powers of nu / nu_kappa inside the Faraday fits.
  double dlog_freqs[RAD_MAX_FREQ];             // ln(image_frequencies[l] / image_frequencies[0])
  double kappa_inv_x[6], kappa_slope_lo[6], kappa_slope_hi[6];
  double kappa_k[6][RAD_MAX_FREQ], kappa_kinv[6][RAD_MAX_FREQ];
  double rho_c84[RAD_MAX_FREQ], rho_cm12[RAD_MAX_FREQ];          // (nu_l / nu_0)^0.84, (nu_l / nu_0)^-0.5
  double rho_cqe_low[RAD_MAX_FREQ], rho_cqe_high[RAD_MAX_FREQ];  // (nu_l / nu_0)^kappa_rho_q_{low,high}_e
  int32_t need_sigma_beta;          // sigma / beta_inverse needed as values (cell values or cuts)
  // rendering
  int32_t render_num_images;
  int32_t render_feature_start[RAD_MAX_FEATURES + 1];
  int32_t render_quantities[RAD_MAX_FEATURES], render_types[RAD_MAX_FEATURES];
  double render_min_vals[RAD_MAX_FEATURES], render_max_vals[RAD_MAX_FEATURES];
  double render_thresh_vals[RAD_MAX_FEATURES], render_tau_scales[RAD_MAX_FEATURES];
  double render_opacities[RAD_MAX_FEATURES];
  double render_x_vals[RAD_MAX_FEATURES], render_y_vals[RAD_MAX_FEATURES], render_z_vals[RAD_MAX_FEATURES];
};

// Optional parity taps, filled in the reference's host layout (source->camera order)
struct SampleTaps {
  int32_t *inds;     // (rays_level, S, 4) or nullptr
  double *fracs;     // (rays_level, S, 3) or nullptr
  uint8_t *nan_, *cut, *fallback;  // (rays_level, S)
  int32_t S;
};

struct RadArgs {
  GridDev grid;
  StepBuffer sb;
  const int32_t *sample_num;    // (wave rays)
  const uint8_t *sample_flags;  // (wave rays)
  const double *mom_factor;     // (wave rays)
  const double *cam_pos;        // (wave rays,4) camera position / covariant momentum of each ray:
  const double *cam_dir;        //   the polarized kernel projects onto the camera tetrad at the end
  int64_t rays;                 // rays in this wave
  const int32_t *order;         // rays of the wave sorted by length, longest first (ray_order.cu), or nullptr: thread i
  int64_t active;               //   takes ray order[i] (i itself without a list), i < active
  double *image;                // (Q, level_rays) device, already offset to this wave's first ray
  int64_t image_stride;         // level_rays
  double *render;               // (R, 3, level_rays) or nullptr, offset likewise
  SampleTaps taps;              // pointers already offset to this wave's first ray
  unsigned long long *sample_counter;  // processed (ray, sample) pairs, for roofline accounting
  int32_t prefetch;             // step-buffer records are prefetched this many samples ahead into L2 (0 = off)
  unsigned long long *slow_counters;   // [0..3] pixels extrapolating (camera small, camera large, source small,
                                       // source large), [4..7] the largest extrapolations as double bits; or nullptr
};
