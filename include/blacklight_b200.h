/* blacklight_b200 -- C ABI of the B200-native ray-tracing hot path.
 *
 * The reference (c-white/blacklight) has no plugin/FFI layer: its `main` wires three objects
 * together and calls three hot methods (reference src/blacklight.cpp:94,204,221):
 *     double GeodesicIntegrator::Integrate()                      geodesic_integrator.cpp:194
 *     double GeodesicIntegrator::AddGeodesics(const RadiationIntegrator*)          ...:236
 *     bool   RadiationIntegrator::Integrate(int snapshot, double*, double*, double*)
 *                                                                 radiation_integrator.cpp:676
 * with the grid handed over by RadiationIntegrator::ObtainGridData (simulation_sampling.cpp:26).
 * Each entry point below replaces one of those seams; the reference-side stub a maintainer
 * would add is shown in INTEGRATION.md.  Plain pointers and sizes only; all host buffers are
 * caller-owned; nothing but the context needs freeing.  Every function returns 0 on success
 * and a nonzero bl_status otherwise, with a message retrievable by bl_last_error().
 * There is no CPU fallback: bl_create fails if no CUDA device is usable.
 */
#ifndef BLACKLIGHT_B200_H_
#define BLACKLIGHT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BL_ABI_VERSION 1
#define BL_MAX_FREQ 32          /* image_num_frequencies upper bound held in constant memory */
#define BL_MAX_RENDER_FEATURES 64
#define BL_MAX_REGIONS 32
#define BL_NUM_CELL_VALUES 7    /* rho, n_e, p_gas, theta_e, bb, sigma, beta_inv (blacklight.hpp:30-33) */

typedef enum bl_status {
  BL_OK = 0,
  BL_ERR_ARG = 1,       /* invalid argument / inconsistent parameters */
  BL_ERR_CUDA = 2,      /* CUDA runtime failure (message carries cudaGetErrorString) */
  BL_ERR_STATE = 3,     /* call order violated (e.g. radiate before trace) */
  BL_ERR_NOMEM = 4,
  BL_ERR_UNSUPPORTED = 5
} bl_status;

/* enum values follow the declaration order in the reference's blacklight.hpp:36-46 */
enum { BL_MODEL_SIMULATION = 0, BL_MODEL_FORMULA = 1 };
enum { BL_COORD_CKS = 0, BL_COORD_SKS = 1, BL_COORD_FMKS = 2 };
enum { BL_CAMERA_PLANE = 0, BL_CAMERA_PINHOLE = 1 };
enum { BL_INTEGRATOR_DP = 0, BL_INTEGRATOR_RK4 = 1, BL_INTEGRATOR_RK2 = 2 };
enum { BL_NORM_CAMERA = 0, BL_NORM_INFINITY = 1 };
enum { BL_RENDER_FILL = 0, BL_RENDER_THRESH = 1, BL_RENDER_RISE = 2, BL_RENDER_FALL = 3 };
enum { BL_PLASMA_TI_TE_BETA = 0, BL_PLASMA_CODE_KAPPA = 1 };

/* POD mirror of exactly the fields the two integrators copy in their constructors
 * (geodesic_integrator.cpp:26-123, radiation_integrator.cpp:31-433).  Angles in radians. */
typedef struct bl_params {
  int32_t abi_version;            /* must be BL_ABI_VERSION */
  int32_t model_type;             /* BL_MODEL_* */
  /* geometry */
  double bh_a;                    /* simulation_a or formula_spin; bh_m is 1 */
  double mass_msun;               /* simulation_m_msun, or formula_mass*c^2/gg_msun */
  /* ray tracing (geodesic_integrator.cpp:57-75) */
  int32_t ray_flat;
  int32_t ray_integrator;         /* BL_INTEGRATOR_* */
  double ray_step;
  int32_t ray_max_steps;
  int32_t ray_max_retries;
  double ray_tol_abs, ray_tol_rel;
  double r_terminate;             /* geodesic_integrator.cpp:117-123 */
  double r_horizon;
  /* camera frame (camera.cpp:53-380), needed by cuts, crossings, polarized projection */
  double camera_r;
  double camera_x[4];
  double camera_u_con[4], camera_u_cov[4];
  double camera_vert_con_c[4];
  /* image */
  int32_t image_light;
  int32_t image_num_frequencies;
  double image_frequencies[BL_MAX_FREQ];
  int32_t image_polarization, image_rotation_split;
  int32_t image_time, image_length, image_lambda, image_emission, image_tau;
  int32_t image_lambda_ave, image_emission_ave, image_tau_int, image_crossings;
  /* simulation sampling */
  int32_t simulation_coord;       /* BL_COORD_* */
  int32_t simulation_interp, simulation_block_interp;
  double simulation_rho_cgs;
  /* plasma (radiation_integrator.cpp:273-314; gammas from the reader) */
  double plasma_mu, plasma_ne_ni;
  int32_t plasma_model, plasma_use_p;
  double plasma_gamma, plasma_gamma_i, plasma_gamma_e;
  double plasma_rat_low, plasma_rat_high;
  double plasma_power_frac, plasma_p, plasma_gamma_min, plasma_gamma_max;
  double plasma_kappa_frac, plasma_kappa, plasma_w;
  /* formula model (radiation_integrator.cpp:70-83) */
  double formula_r0, formula_h, formula_l0, formula_q, formula_nup, formula_cn0;
  double formula_alpha, formula_a, formula_beta;
  /* cuts (negative / zero = disabled exactly as in the reference) */
  double cut_rho_min, cut_rho_max, cut_n_e_min, cut_n_e_max, cut_p_gas_min, cut_p_gas_max;
  double cut_theta_e_min, cut_theta_e_max, cut_b_min, cut_b_max, cut_sigma_min, cut_sigma_max;
  double cut_beta_inverse_min, cut_beta_inverse_max;
  int32_t cut_omit_near, cut_omit_far;
  double cut_omit_in, cut_omit_out, cut_midplane_theta, cut_midplane_z;
  int32_t cut_plane;
  double cut_plane_origin[3], cut_plane_normal[3];
  /* fallback */
  int32_t fallback_nan;
  float fallback_rho, fallback_pgas, fallback_kappa;
  /* rendering (rendering.cpp:25-179): features of all images flattened, image r owns
   * features [render_feature_start[r], render_feature_start[r+1]) */
  int32_t render_num_images;
  int32_t render_feature_start[BL_MAX_RENDER_FEATURES + 1];
  int32_t render_quantities[BL_MAX_RENDER_FEATURES];
  int32_t render_types[BL_MAX_RENDER_FEATURES];
  double render_min_vals[BL_MAX_RENDER_FEATURES], render_max_vals[BL_MAX_RENDER_FEATURES];
  double render_thresh_vals[BL_MAX_RENDER_FEATURES], render_tau_scales[BL_MAX_RENDER_FEATURES];
  double render_opacities[BL_MAX_RENDER_FEATURES];
  double render_x_vals[BL_MAX_RENDER_FEATURES], render_y_vals[BL_MAX_RENDER_FEATURES];
  double render_z_vals[BL_MAX_RENDER_FEATURES];
  /* adaptive refinement (radiation_adaptive.cpp) */
  int32_t adaptive_max_level, adaptive_block_size, adaptive_frequency_num;
  int32_t camera_resolution;
  double camera_width;
  double adaptive_val_cut, adaptive_val_frac, adaptive_abs_grad_cut, adaptive_abs_grad_frac;
  double adaptive_rel_grad_cut, adaptive_rel_grad_frac, adaptive_abs_lapl_cut, adaptive_abs_lapl_frac;
  double adaptive_rel_lapl_cut, adaptive_rel_lapl_frac;
  int32_t adaptive_num_regions;
  int32_t adaptive_region_levels[BL_MAX_REGIONS];
  double adaptive_region_x_min[BL_MAX_REGIONS], adaptive_region_x_max[BL_MAX_REGIONS];
  double adaptive_region_y_min[BL_MAX_REGIONS], adaptive_region_y_max[BL_MAX_REGIONS];
  /* B200 knobs (new, optional; 0 = default) */
  int32_t device;                 /* CUDA device ordinal */
  int64_t tile_rays;              /* rays traced per wave; 0 = sized from free HBM */
  /* slow light (radiation_integrator.cpp:203-214, simulation_reader.hpp:99) */
  int32_t slow_light_on, slow_interp, slow_chunk_size;
  double extrapolation_tolerance; /* the reference's constant 1.0 */
  int32_t level0_block_major;     /* 1: level-0 rays are given block by block (m = block*bs^2 + row*bs + col, like the
                                   * refined levels) instead of as the full raster -- lets the root blocks of an
                                   * adaptive image be sharded over GPUs; affects only
                                   * bl_refine_level's addressing of level 0 */
} bl_params;

/* Host view of one snapshot exactly as SimulationReader leaves it
 * (simulation_reader.hpp:114-126; Athena++ branch simulation_reader.cpp:591-621,762-781). */
typedef struct bl_grid_view {
  int32_t n_b, n_k, n_j, n_i;     /* blocks, cells per block in x3, x2, x1 */
  int32_t n_var;                  /* leading dimension of prim */
  const int32_t *levels;          /* (n_b) or NULL */
  const int32_t *locations;       /* (n_b,3) or NULL */
  const double *x1f, *x2f, *x3f;  /* (n_b, n+1) */
  const double *x1v, *x2v, *x3v;  /* (n_b, n) */
  const float *prim;              /* (n_var, n_b, n_k, n_j, n_i) */
  int32_t ind_rho, ind_pgas, ind_kappa, ind_uu1, ind_uu2, ind_uu3, ind_bb1, ind_bb2, ind_bb3;
  int32_t n_3_root;               /* RootGridSize[2], inter-block interpolation only */
  /* simulation_coord = fmks only (simulation_reader.hpp:103-112, simulation_geometry.cpp:330-413): x1f, x2f, x?v stay in
   * the native coordinates; sks_map(0|1, j, i) = native x1 | x2 at r = r_in + i dr, theta = j dtheta; the block test uses
   * simulation_bounds = (r_min, r_max, theta_min, theta_max, phi_min, phi_max) (simulation_sampling.cpp:190-198,397-413) */
  const double *sks_map;          /* (2, sks_map_n2, sks_map_n1) or NULL */
  int32_t sks_map_n1, sks_map_n2;
  double sks_map_r_in, sks_map_dr, sks_map_dtheta;
  double simulation_bounds[6];
} bl_grid_view;

typedef struct bl_level_stats {
  int64_t num_rays;
  int32_t geodesic_num_steps;     /* max sample_num over rays (geodesics.cpp:374-386) */
  int64_t num_bad_geodesics;      /* rays with sample_flags set (geodesics.cpp:379-394) */
  int64_t num_samples;            /* sum of sample_num */
  int64_t num_attempts;           /* DP step attempts (accepted + rejected), for roofline accounting */
  int64_t num_accepted;
  double ms_geodesic;             /* CUDA-event times of the kernels run by this call */
  double ms_radiation;
  double ms_refine;
} bl_level_stats;

typedef struct bl_ctx bl_ctx;

/* Replaces the two integrator constructors.  Also reports image_num_quantities
 * (radiation_integrator.cpp:436-520) through bl_image_num_quantities(). */
int bl_create(const bl_params *params, bl_ctx **out);
void bl_destroy(bl_ctx *ctx);
const char *bl_last_error(const bl_ctx *ctx);   /* ctx may be NULL: error of a failed bl_create */
int bl_image_num_quantities(const bl_ctx *ctx);
int bl_device_count(void);                       /* usable CUDA devices (0 if none): one bl_ctx drives one of them */

/* Replaces RadiationIntegrator::ObtainGridData: one H2D copy of the whole grid (per snapshot). */
int bl_upload_grid(bl_ctx *ctx, const bl_grid_view *grid);

/* Replaces GeodesicIntegrator::Integrate (level 0) / AddGeodesics (level > 0) after the host
 * has built the camera arrays (camera.cpp:390-413 / 426-504).  cam_pos, cam_dir: (N,4) row-major
 * f64 (dir = covariant momentum), mom_factor: (N).  Rays are traced immediately if the level's
 * step buffer fits the HBM budget, otherwise tile by tile inside bl_radiate_level. */
int bl_trace_level(bl_ctx *ctx, int level, const double *cam_pos, const double *cam_dir,
                   const double *mom_factor, int64_t num_rays, bl_level_stats *stats);

/* Camera pixels generated on the device (reference camera.cpp:390-413 root raster, :471-499 refined blocks, :528-671
 * SetPixelPlane / SetPixelPinhole): bl_set_camera hands over what InitializeCamera leaves (camera.cpp:53-380; the host
 * layer's blh_camera_frame), bl_trace_level_pixels then replaces "build the camera arrays on the host + bl_trace_level":
 * the level's rays are the pixels of the listed units at effective resolution camera_resolution * 2^level, computed in
 * HBM bit for bit as the host code computes them (csrc/camera_kernel.cu), so only the unit list crosses PCIe.
 *   BL_PIXELS_ROWS:   units = image rows (num_units int32), ray m = unit_index * eff_res + column; units == NULL
 *                     means rows 0 .. num_units-1 (num_units = eff_res: the whole raster)
 *   BL_PIXELS_BLOCKS: units = (num_units, 2) int32 (v, u) block locations, adaptive_block_size^2 rays per block,
 *                     block-major like the reference's refined levels
 * bl_download_camera returns the level's camera arrays (any pointer may be NULL), e.g. for output_camera. */
typedef struct bl_camera {
  int32_t type;                   /* BL_CAMERA_* */
  int32_t normalization;          /* BL_NORM_* (image_normalization) */
  double width, r;                /* camera_width, camera_r */
  double x[4], u_con[4], u_cov[4];
  double norm_con[4], norm_con_c[4], hor_con_c[4], vert_con_c[4];
} bl_camera;
enum { BL_PIXELS_ROWS = 0, BL_PIXELS_BLOCKS = 1 };
int bl_set_camera(bl_ctx *ctx, const bl_camera *camera);
int bl_trace_level_pixels(bl_ctx *ctx, int level, int kind, const int32_t *units, int64_t num_units,
                          bl_level_stats *stats);
int bl_download_camera(bl_ctx *ctx, int level, double *cam_pos, double *cam_dir, double *mom_factor);

/* Slow light (reference simulation_reader.cpp:211-303, simulation_sampling.cpp:298-349): the context keeps
 * slow_chunk_size snapshots resident in HBM.  bl_upload_grid_slice puts one snapshot into slot `slot`
 * (0 <= slot < slow_chunk_size; slot 0 is what bl_upload_grid fills); bl_set_time_window then declares, for the
 * next bl_radiate_level calls, which slot holds window entry t (t = 0 the latest ... count-1 the earliest, as in
 * the reader's prim[t]/time[t]), the entries' simulation times, and the camera time of the image
 * (slow_t_start + slow_dt * snapshot).  Shifting the window is a permutation of `slots` on the host; no device
 * data moves. */
#define BL_MAX_SLICES 64
int bl_upload_grid_slice(bl_ctx *ctx, const bl_grid_view *grid, int slot);
int bl_set_time_window(bl_ctx *ctx, int count, const int32_t *slots, const double *times, double snapshot_time);

/* Extrapolation accounting of the last bl_radiate_level on a level (simulation_sampling.cpp:556-618): pixels that
 * needed time slices beyond the window, by less / more than extrapolation_tolerance, and by how much at most.
 * Index 0: forward in time (camera side), 1: backward (source side). */
typedef struct bl_slow_stats {
  int64_t num_small[2], num_large[2];
  double val_small[2], val_large[2];
} bl_slow_stats;
int bl_slow_light_stats(bl_ctx *ctx, int level, bl_slow_stats *out);

/* Load a level whose geodesics were integrated elsewhere (the reference's checkpoint_geodesic_load,
 * geodesic_checkpoint.cpp:77-108) instead of tracing it: camera arrays as for bl_trace_level plus the
 * reference's sample arrays in their host layouts and source->camera order -- flags, num: (N); pos, dir:
 * (N,S,4) f64; len: (N,S) f64 > 0; S = geodesic_num_steps <= ray_max_steps.  The level must fit in HBM. */
int bl_upload_samples(bl_ctx *ctx, int level, const double *cam_pos, const double *cam_dir,
                      const double *mom_factor, int64_t num_rays, int32_t S, const uint8_t *flags,
                      const int32_t *num, const double *pos, const double *dir, const double *len,
                      bl_level_stats *stats);

/* Re-integrate the geodesics of a level from the camera arrays already resident in HBM (no host
 * transfer).  A no-op for levels traced wave by wave inside bl_radiate_level. */
int bl_retrace_level(bl_ctx *ctx, int level, bl_level_stats *stats);

/* Number of CUDA kernels of this library launched by the context so far (bench accounting). */
long long bl_launch_count(const bl_ctx *ctx);

/* The level's image where bl_radiate_level leaves it in HBM, (image_num_quantities, num_rays) f64 -- for a host that
 * gathers the images of several GPUs device to device (peer copies, NCCL) instead of bouncing them through host memory.
 * Valid until the level is traced again with a different ray count or the context is destroyed; complete when
 * bl_radiate_level has returned. */
int bl_device_image(bl_ctx *ctx, int level, void **image, int64_t *num_rays);

/* Polarized levels without per-sample side outputs are rendered by a pipeline of four kernels over slabs of the step
 * buffer (grid sampling + plasma state | metric jet, tetrad, Stokes transport matrix | synchrotron coefficients | Stokes
 * coupling; csrc/radiate_pol_split.cu) in place of the single fused kernel.  ms3: device time of the last three stages
 * during the last bl_radiate_level of the level (CUDA events around every launch), bl_polarized_sampling_ms: of the
 * first; *slab: samples per slab, 0 if the fused kernel ran.  Environment, tuning only: BL_POL_FUSED=1 keeps the fused
 * kernel, BL_POL_SLAB=n fixes the slab length, BL_RAY_ORDER=0 takes the rays in index order instead of by length. */
int bl_polarized_stage_ms(bl_ctx *ctx, int level, double *ms3, int32_t *slab);
int bl_polarized_sampling_ms(bl_ctx *ctx, int level, double *ms);
/* Parity tap of that pipeline: the scratch of the LAST slab it processed (samples 0 <= n < slab of every ray; with
 * BL_POL_SLAB >= ray_max_steps the whole level), (num_fields, slab, num_rays) f64 -- per sample the 3x3 + 1 entries of
 * the Stokes transport matrix, the affine step, seven plasma scalars and 8 synchrotron coefficients per frequency
 * (field order: csrc/radiate_pol_split.cu); cam_map: (10, num_rays), the half step from the last sample onto the camera
 * tetrad.  out / cam_map may be NULL (e.g. to query the extents first).  Resident levels only. */
int bl_download_polarized_scratch(bl_ctx *ctx, int level, double *out, double *cam_map, int64_t *num_fields, int64_t *slab,
                                  int64_t *num_rays);

/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on, so that a host
 * can bracket calls with its own events or order other work against them. */
void *bl_cuda_stream(const bl_ctx *ctx);

/* Replaces RadiationIntegrator::Integrate's sampling + coefficient + transfer (+ render) stages
 * for one level.  image: (image_num_quantities, N) f64; render: (R,3,N) f64 or NULL. */
int bl_radiate_level(bl_ctx *ctx, int level, int snapshot, double *image, double *render,
                     bl_level_stats *stats);

/* Replaces RadiationIntegrator::CheckAdaptiveRefinement (radiation_adaptive.cpp:19-139) for the
 * image last produced at this level.  block_locs: (B,2) int32 (v,u) block coordinates.
 * flags: (B) bytes out; n_refined: number of flags set. */
int bl_refine_level(bl_ctx *ctx, int level, const int32_t *block_locs, int64_t num_blocks,
                    uint8_t *flags, int64_t *n_refined);

/* Enable (1) / disable (0) recording of the sampling taps during bl_radiate_level (parity tests only;
 * costs N*S*(4*4+3*8+3) bytes of HBM per level). */
int bl_set_taps(bl_ctx *ctx, int enabled);

/* Parity taps mirroring the reference's checkpoint dumps (geodesic_checkpoint.cpp:36-57,
 * sample_checkpoint.cpp:31-35), in the reference's own host layouts and source->camera order.
 * Any pointer may be NULL.  flags: (N) bytes; num: (N) int32; pos, dir: (N,S,4) f64; len: (N,S),
 * S = geodesic_num_steps of the level; entries with n >= num[m] are zero. */
int bl_download_samples(bl_ctx *ctx, int level, uint8_t *flags, int32_t *num, double *pos,
                        double *dir, double *len);
/* inds: (N,S,4) int32 (b,k,j,i); fracs: (N,S,3) f64 (f_k,f_j,f_i) or NULL when not interpolating;
 * nan_, cut, fallback: (N,S) bytes.  Entries the reference leaves unset are -1 / 0. */
int bl_download_sample_inds(bl_ctx *ctx, int level, int32_t *inds, double *fracs, uint8_t *nan_,
                            uint8_t *cut, uint8_t *fallback);

/* Device properties and a measured FP64 FMA peak (TFLOP/s) for roofline denominators. */
int bl_device_info(bl_ctx *ctx, char *name, int name_len, int *sm_count, double *hbm_free_gb);
int bl_measure_fp64_peak(bl_ctx *ctx, double *tflops);


/* Self-test of the geodesic kernel's branch-free division and square root (csrc/glibc_math.cuh: div_by,
 * sqrt_rn) against the hardware IEEE operations on num_pairs pseudo-random operand pairs; *mismatches =
 * differing results (must be 0: the integrator's bit-for-bit parity with the reference rests on it). */
int bl_selftest_division(bl_ctx *ctx, uint64_t seed, int64_t num_pairs, int64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* BLACKLIGHT_B200_H_ */
