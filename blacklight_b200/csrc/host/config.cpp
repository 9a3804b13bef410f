#include "config.hpp"

#include <omp.h>

#include <cmath>
#include <cstring>

namespace blh {

namespace {
constexpr double kPi = 3.141592653589793;
constexpr double kC = 2.99792458e10;
constexpr double kGgMsun = 1.32712440018e26;

int cell_quantity(const std::string &v) {
  static const char *names[] = {"rho", "n_e", "p_gas", "Theta_e", "B", "sigma", "beta_inverse"};
  for (int i = 0; i < 7; i++)
    if (v == names[i]) return i;
  throw Error("Invalid render quantity (" + v + ") in input file.");
}

// sRGB (0-255) -> CIE XYZ (reference utils/colors.cpp:24-36)
void rgb_to_xyz(double r, double g, double b, double *x, double *y, double *z) {
  auto lin = [](double c) {
    double c1 = c / 255.0;
    return c1 <= 0.040449936 ? c1 / 12.92 : std::pow((c1 + 0.055) / 1.055, 2.4);
  };
  double lr = lin(r), lg = lin(g), lb = lin(b);
  *x = 0.4123955889674142 * lr + 0.3575834307637148 * lg + 0.18049264738170154 * lb;
  *y = 0.21258623078559552 * lr + 0.715170303703411 * lg + 0.0722004986433362 * lb;
  *z = 0.019297215491746938 * lr + 0.11918386458084851 * lg + 0.9504971251315798 * lb;
}
}  // namespace

RunConfig make_config(const InputFile &in) {
  RunConfig c;
  bl_params &p = c.params;
  std::memset(&p, 0, sizeof p);
  p.abi_version = BL_ABI_VERSION;
  p.model_type = in.choice("model_type", {"simulation", "formula"}, "ModelType");
  const bool sim = p.model_type == BL_MODEL_SIMULATION;
  c.num_runs = in.num_runs();
  // host-side parallel loops (camera pixels, reader conversions) use the input file's thread count, as the
  // reference's main does (blacklight.cpp:77)
  c.num_threads = in.integer("num_threads");
  if (c.num_threads > 0) omp_set_num_threads(c.num_threads);

  // output
  c.output_format = in.choice("output_format", {"npz", "npy", "raw"}, "OutputFormat");
  c.output_file = in.str("output_file");
  if (c.output_format == 0) c.output_camera = in.flag("output_camera");

  // checkpoints (geodesic_integrator.cpp:29-34, radiation_integrator.cpp:35-52); sample checkpoints cannot
  // be loaded (sampling is fused into the transfer kernels, there is nothing to load them into)
  c.checkpoint_geodesic_save = in.flag("checkpoint_geodesic_save");
  c.checkpoint_geodesic_load = in.flag("checkpoint_geodesic_load");
  if (c.checkpoint_geodesic_save && c.checkpoint_geodesic_load) throw Error("Cannot both save and load a geodesic checkpoint.");
  if (c.checkpoint_geodesic_save || c.checkpoint_geodesic_load) c.checkpoint_geodesic_file = in.str("checkpoint_geodesic_file");
  if (sim) {
    c.checkpoint_sample_save = in.flag("checkpoint_sample_save");
    bool sample_load = in.flag("checkpoint_sample_load");
    if (c.checkpoint_sample_save && sample_load) throw Error("Cannot both save and load a sample checkpoint.");
    if (sample_load) throw Error("checkpoint_sample_load is outside the B200 hot-path scope.");
    if (c.checkpoint_sample_save) c.checkpoint_sample_file = in.str("checkpoint_sample_file");
  } else {
    if (in.has("checkpoint_sample_save") && in.flag("checkpoint_sample_save")) warning("Ignoring checkpoint_sample_save selection.");
    if (in.has("checkpoint_sample_load") && in.flag("checkpoint_sample_load")) warning("Ignoring checkpoint_sample_load selection.");
  }

  // camera
  CameraSetup &cam = c.camera;
  cam.type = in.choice("camera_type", {"plane", "pinhole"}, "Camera");
  cam.r = in.real("camera_r");
  cam.th = in.real("camera_th") * kPi / 180.0;
  cam.ph = in.real("camera_ph") * kPi / 180.0;
  cam.urn = in.real("camera_urn");
  cam.uthn = in.real("camera_uthn");
  cam.uphn = in.real("camera_uphn");
  cam.k_r = in.real("camera_k_r");
  cam.k_th = in.real("camera_k_th");
  cam.k_ph = in.real("camera_k_ph");
  cam.rotation = in.real("camera_rotation") * kPi / 180.0;
  cam.width = in.real("camera_width");
  cam.resolution = in.integer("camera_resolution");
  if (cam.resolution <= 0) throw Error("Must have positive camera_resolution.");
  cam.pole = in.camera_pole();

  // ray tracing
  p.ray_flat = in.flag("ray_flat");
  cam.flat = p.ray_flat;
  int terminate = in.choice("ray_terminate", {"photon", "multiplicative", "additive"}, "RayTerminate");
  double ray_factor = terminate != 0 ? in.real("ray_factor") : 0.0;
  p.ray_integrator = in.choice("ray_integrator", {"dp", "rk4", "rk2"}, "RayIntegrator");
  p.ray_step = in.real("ray_step");
  p.ray_max_steps = in.integer("ray_max_steps");
  if (p.ray_max_steps <= 0) throw Error("Must have positive ray_max_steps.");
  if (p.ray_integrator == BL_INTEGRATOR_DP) {
    p.ray_max_retries = in.integer("ray_max_retries");
    if (p.ray_max_retries <= 0) throw Error("Must have nonnegative ray_max_retries.");
    p.ray_tol_abs = in.real("ray_tol_abs");
    p.ray_tol_rel = in.real("ray_tol_rel");
  }

  // image frequencies
  p.image_num_frequencies = in.integer("image_num_frequencies");
  double f_single = 0.0, f_start = 0.0, f_end = 0.0;
  int spacing = 0;
  if (p.image_num_frequencies == 1) {
    f_single = in.real("image_frequency");
    if (f_single <= 0.0) throw Error("Must choose positive image_frequency.");
  } else if (p.image_num_frequencies > 1) {
    f_start = in.real("image_frequency_start");
    if (f_start <= 0.0) throw Error("Must choose positive image_frequency_start.");
    f_end = in.real("image_frequency_end");
    if (f_end <= 0.0) throw Error("Must choose positive image_frequency_end.");
    spacing = in.choice("image_frequency_spacing", {"lin_freq", "lin_wave", "log"}, "FrequencySpacing");
  } else {
    throw Error("Must have positive image_num_frequencies.");
  }
  if (p.image_num_frequencies > BL_MAX_FREQ) throw Error("image_num_frequencies exceeds the B200 build limit.");
  cam.normalization = in.choice("image_normalization", {"camera", "infinity"}, "FrequencyNormalization");
  c.frequencies = image_frequencies(p.image_num_frequencies, f_single, f_start, f_end, spacing);
  for (int l = 0; l < p.image_num_frequencies; l++) p.image_frequencies[l] = c.frequencies[(size_t)l];

  // adaptive (geodesic side)
  p.adaptive_max_level = in.integer("adaptive_max_level");
  if (p.adaptive_max_level > 0) {
    p.adaptive_block_size = in.integer("adaptive_block_size");
    if (p.adaptive_block_size <= 0) throw Error("Must have positive adaptive_block_size.");
    if (cam.resolution % p.adaptive_block_size != 0) throw Error("Must have adaptive_block_size divide camera_resolution.");
  }

  // geometry
  p.bh_a = sim ? in.real("simulation_a") : in.real("formula_spin");
  cam.a = p.bh_a;
  p.r_horizon = 1.0 + std::sqrt(1.0 * 1.0 - p.bh_a * p.bh_a);
  if (terminate == 0)
    p.r_terminate = 2.0 * 1.0 * (1.0 + std::cos(2.0 / 3.0 * std::acos(-std::abs(p.bh_a) / 1.0)));
  else if (terminate == 1)
    p.r_terminate = p.r_horizon * ray_factor;
  else
    p.r_terminate = p.r_horizon + ray_factor;

  // ---- radiation side ----
  if (sim) {
    c.simulation_format = in.choice("simulation_format", {"athena", "athenak", "iharm3d", "harm3d"}, "SimulationFormat");
    c.simulation_file = in.str("simulation_file");
    c.simulation_multiple = in.flag("simulation_multiple");
    if (c.simulation_multiple) {
      c.simulation_start = in.integer("simulation_start");
      if (c.simulation_start < 0) throw Error("Must have nonnegative index simulation_start.");
      c.simulation_end = in.integer("simulation_end");
      if (c.simulation_end < c.simulation_start) throw Error("Must have simulation_end at least as large as simulation_start.");
    }
    const std::string &coord = in.str("simulation_coord");
    if (coord == "cks") p.simulation_coord = BL_COORD_CKS;
    else if (coord == "sks" || coord == "mks") p.simulation_coord = BL_COORD_SKS;
    else if (coord == "fmks") p.simulation_coord = BL_COORD_FMKS;
    else throw Error("Unknown string used for Coordinates value.");
    p.mass_msun = in.real("simulation_m_msun");
    p.simulation_rho_cgs = in.real("simulation_rho_cgs");
    p.simulation_interp = in.flag("simulation_interp");
    if ((c.simulation_format == 0 || c.simulation_format == 1) && p.simulation_interp)
      p.simulation_block_interp = in.flag("simulation_block_interp");
    else if (in.has("simulation_block_interp"))
      warning("Ignoring simulation_block_interp selection.");
    // slow light (simulation_reader.cpp:64-82, radiation_integrator.cpp:203-214)
    p.slow_light_on = in.flag("slow_light_on");
    p.extrapolation_tolerance = 1.0;   // simulation_reader.hpp:99
    if (p.slow_light_on) {
      if (!c.simulation_multiple) throw Error("Must enable simulation_multiple to use slow light.");
      if (c.checkpoint_sample_save) throw Error("Cannot use sample checkpoints with slow light.");
      p.slow_interp = in.flag("slow_interp");
      p.slow_chunk_size = in.integer("slow_chunk_size");
      if (p.slow_chunk_size < 2) throw Error("Must have slow_chunk_size be at least 2.");
      if (p.slow_chunk_size > c.simulation_end - c.simulation_start + 1) throw Error("Not enough simulation files for given slow_chunk_size.");
      c.slow_t_start = in.real("slow_t_start");
      c.slow_dt = in.real("slow_dt");
      if (c.slow_dt <= 0.0) throw Error("Must have positive time interval slow_dt.");
      c.slow_offset = in.integer("slow_offset");
    }
  } else {
    double formula_mass = in.real("formula_mass");
    p.mass_msun = formula_mass * kC * kC / kGgMsun;
    p.formula_r0 = in.real("formula_r0");
    p.formula_h = in.real("formula_h");
    p.formula_l0 = in.real("formula_l0");
    p.formula_q = in.real("formula_q");
    p.formula_nup = in.real("formula_nup");
    p.formula_cn0 = in.real("formula_cn0");
    p.formula_alpha = in.real("formula_alpha");
    p.formula_a = in.real("formula_a");
    p.formula_beta = in.real("formula_beta");
    if (in.has("slow_light_on") && in.flag("slow_light_on")) throw Error("Can only use slow light with simulation data.");
  }
  p.camera_r = cam.r;
  p.camera_width = cam.width;
  p.camera_resolution = cam.resolution;

  // images
  p.image_light = in.flag("image_light");
  if (p.image_light) {
    if (sim) p.image_polarization = in.flag("image_polarization");
    else if (in.has("image_polarization") && in.flag("image_polarization")) warning("Ignoring image_polarization selection.");
    if (p.image_polarization) p.image_rotation_split = in.flag("image_rotation_split");
  } else if (in.has("image_polarization") && in.flag("image_polarization")) {
    warning("Ignoring image_polarization selection.");
  }
  p.image_time = in.flag("image_time");
  p.image_length = in.flag("image_length");
  p.image_lambda = in.flag("image_lambda");
  p.image_emission = in.flag("image_emission");
  p.image_tau = in.flag("image_tau");
  if (sim) {
    p.image_lambda_ave = in.flag("image_lambda_ave");
    p.image_emission_ave = in.flag("image_emission_ave");
    p.image_tau_int = in.flag("image_tau_int");
  } else {
    if (in.has("image_lambda_ave") && in.flag("image_lambda_ave")) warning("Ignoring image_lambda_ave selection.");
    if (in.has("image_emission_ave") && in.flag("image_emission_ave")) warning("Ignoring image_emission_ave selection.");
    if (in.has("image_tau_int") && in.flag("image_tau_int")) warning("Ignoring image_tau_int selection.");
  }
  p.image_crossings = in.flag("image_crossings");

  // rendering
  if (sim) {
    p.render_num_images = in.integer("render_num_images");
  } else {
    if (in.has("render_num_images") && in.integer("render_num_images") > 0) warning("Ignoring request for rendering.");
    p.render_num_images = 0;
  }
  int feature = 0;
  p.render_feature_start[0] = 0;
  for (int im = 0; im < p.render_num_images; im++) {
    std::string base = "render_" + std::to_string(im + 1) + "_";
    int nf = in.integer(base + "num_features");
    if (nf <= 0) throw Error("Must have positive number of features for each rendered image.");
    for (int f = 0; f < nf; f++, feature++) {
      if (feature >= BL_MAX_RENDER_FEATURES) throw Error("Too many render features for the B200 build limit.");
      std::string fb = base + std::to_string(f + 1) + "_";
      p.render_quantities[feature] = cell_quantity(in.str(fb + "quantity"));
      const std::string &type = in.str(fb + "type");
      if (type == "fill") p.render_types[feature] = BL_RENDER_FILL;
      else if (type == "thresh") p.render_types[feature] = BL_RENDER_THRESH;
      else if (type == "rise") p.render_types[feature] = BL_RENDER_RISE;
      else if (type == "fall") p.render_types[feature] = BL_RENDER_FALL;
      else throw Error("Invalid render type (" + type + ") in input file.");
      if (p.render_types[feature] == BL_RENDER_FILL) {
        p.render_min_vals[feature] = in.real(fb + "min");
        p.render_max_vals[feature] = in.real(fb + "max");
        p.render_tau_scales[feature] = in.real(fb + "tau_scale");
      } else {
        p.render_thresh_vals[feature] = in.real(fb + "thresh");
        p.render_opacities[feature] = in.real(fb + "opacity");
      }
      // colour: sRGB triple converted to XYZ, or XYZ given directly (render_reader.cpp:183-210)
      if (in.has(fb + "xyz")) {
        double xyz[3];
        in.triple(fb + "xyz", xyz);
        p.render_x_vals[feature] = xyz[0]; p.render_y_vals[feature] = xyz[1]; p.render_z_vals[feature] = xyz[2];
      } else {
        double rgb[3];
        in.triple(fb + "rgb", rgb);
        rgb_to_xyz(rgb[0], rgb[1], rgb[2], &p.render_x_vals[feature], &p.render_y_vals[feature], &p.render_z_vals[feature]);
      }
    }
    p.render_feature_start[im + 1] = feature;
  }
  for (int im = p.render_num_images; im < BL_MAX_RENDER_FEATURES; im++) p.render_feature_start[im + 1] = feature;
  if (!(p.image_light || p.image_time || p.image_length || p.image_lambda || p.image_emission || p.image_tau ||
        p.image_lambda_ave || p.image_emission_ave || p.image_tau_int || p.image_crossings || p.render_num_images > 0))
    throw Error("No image or rendering selected.");

  // adaptive (radiation side)
  if (p.adaptive_max_level > 0) {
    if (!p.image_light) throw Error("Adaptive ray tracing requires image_light.");
    if (p.image_num_frequencies > 1) {
      p.adaptive_frequency_num = in.integer("adaptive_frequency_num") - 1;
      if (p.adaptive_frequency_num < 0 || p.adaptive_frequency_num >= p.image_num_frequencies)
        throw Error("Must choose adaptive_frequency_num from 1 to image_num_frequencies.");
    }
    p.adaptive_val_frac = in.real("adaptive_val_frac");
    if (p.adaptive_val_frac >= 0.0) p.adaptive_val_cut = in.real("adaptive_val_cut");
    p.adaptive_abs_grad_frac = in.real("adaptive_abs_grad_frac");
    if (p.adaptive_abs_grad_frac >= 0.0) p.adaptive_abs_grad_cut = in.real("adaptive_abs_grad_cut");
    p.adaptive_rel_grad_frac = in.real("adaptive_rel_grad_frac");
    if (p.adaptive_rel_grad_frac >= 0.0) p.adaptive_rel_grad_cut = in.real("adaptive_rel_grad_cut");
    p.adaptive_abs_lapl_frac = in.real("adaptive_abs_lapl_frac");
    if (p.adaptive_abs_lapl_frac >= 0.0) p.adaptive_abs_lapl_cut = in.real("adaptive_abs_lapl_cut");
    p.adaptive_rel_lapl_frac = in.real("adaptive_rel_lapl_frac");
    if (p.adaptive_rel_lapl_frac >= 0.0) p.adaptive_rel_lapl_cut = in.real("adaptive_rel_lapl_cut");
    p.adaptive_num_regions = in.integer("adaptive_num_regions");
    if (p.adaptive_num_regions > BL_MAX_REGIONS) throw Error("Too many adaptive regions for the B200 build limit.");
    for (int r = 0; r < p.adaptive_num_regions; r++) {
      std::string base = "adaptive_region_" + std::to_string(r + 1) + "_";
      p.adaptive_region_levels[r] = in.integer(base + "level");
      p.adaptive_region_x_min[r] = in.real(base + "x_min");
      p.adaptive_region_x_max[r] = in.real(base + "x_max");
      p.adaptive_region_y_min[r] = in.real(base + "y_min");
      p.adaptive_region_y_max[r] = in.real(base + "y_max");
    }
  }

  // plasma (radiation_integrator.cpp:273-314) and the adiabatic indices the reader owns (simulation_reader.cpp:93-150)
  if (sim) {
    p.plasma_mu = in.real("plasma_mu");
    p.plasma_ne_ni = in.real("plasma_ne_ni");
    p.plasma_model = in.choice("plasma_model", {"ti_te_beta", "code_kappa"}, "PlasmaModel");
    if (p.plasma_model == BL_PLASMA_TI_TE_BETA) {
      p.plasma_use_p = in.flag("plasma_use_p");
      p.plasma_rat_low = in.real("plasma_rat_low");
      p.plasma_rat_high = in.real("plasma_rat_high");
      if (p.plasma_use_p) {
        if (in.has("plasma_gamma")) { p.plasma_gamma = in.real("plasma_gamma"); c.gamma_set = true; }
        if (in.has("plasma_gamma_i")) warning("Ignoring plasma_gamma_i selection.");
        if (in.has("plasma_gamma_e")) warning("Ignoring plasma_gamma_e selection.");
      } else {
        if (c.simulation_format == 0 || in.has("plasma_gamma")) { p.plasma_gamma = in.real("plasma_gamma"); c.gamma_set = true; }
        if (c.simulation_format != 2) {
          p.plasma_gamma_i = in.real("plasma_gamma_i");
          p.plasma_gamma_e = in.real("plasma_gamma_e");
        } else {   // iharm3d: the dump's header/gam_p, gam_e stand in (simulation_reader.cpp:113-125)
          if (in.has("plasma_gamma_i")) { p.plasma_gamma_i = in.real("plasma_gamma_i"); c.gamma_i_set = true; }
          if (in.has("plasma_gamma_e")) { p.plasma_gamma_e = in.real("plasma_gamma_e"); c.gamma_e_set = true; }
        }
      }
    } else {
      c.simulation_kappa_name = in.str("simulation_kappa_name");
      if (in.has("plasma_gamma")) { p.plasma_gamma = in.real("plasma_gamma"); c.gamma_set = true; }
      if (in.has("plasma_gamma_i")) warning("Ignoring plasma_gamma_i selection.");
      if (in.has("plasma_gamma_e")) warning("Ignoring plasma_gamma_e selection.");
    }
    p.plasma_power_frac = in.real("plasma_power_frac");
    if (p.plasma_power_frac < 0.0 || p.plasma_power_frac > 1.0) warning("Fraction of power-law electrons outside [0, 1].");
    if (p.plasma_power_frac != 0.0) {
      p.plasma_p = in.real("plasma_p");
      p.plasma_gamma_min = in.real("plasma_gamma_min");
      p.plasma_gamma_max = in.real("plasma_gamma_max");
    }
    p.plasma_kappa_frac = in.real("plasma_kappa_frac");
    if (p.plasma_kappa_frac < 0.0 || p.plasma_kappa_frac > 1.0) warning("Fraction of kappa-distribution electrons outside [0, 1].");
    if (p.plasma_kappa_frac != 0.0) {
      p.plasma_kappa = in.real("plasma_kappa");
      if (p.image_light && p.image_polarization) {
        if (p.plasma_kappa < 3.5 || p.plasma_kappa > 5.0) throw Error("Polarized transport only supports kappa in [3.5, 5].");
        else if (p.plasma_kappa != 3.5 && p.plasma_kappa != 4.0 && p.plasma_kappa != 4.5 && p.plasma_kappa != 5.0)
          warning("Polarized transport will interpolate formulas based on kappa.");
      }
      p.plasma_w = in.real("plasma_w");
    }
    double thermal = 1.0 - (p.plasma_power_frac + p.plasma_kappa_frac);
    if (thermal < 0.0 || thermal > 1.0) warning("Fraction of thermal electrons outside [0, 1].");
  }

  // cuts
  if (sim) {
    p.cut_rho_min = in.real("cut_rho_min"); p.cut_rho_max = in.real("cut_rho_max");
    p.cut_n_e_min = in.real("cut_n_e_min"); p.cut_n_e_max = in.real("cut_n_e_max");
    p.cut_p_gas_min = in.real("cut_p_gas_min"); p.cut_p_gas_max = in.real("cut_p_gas_max");
    p.cut_theta_e_min = in.real("cut_theta_e_min"); p.cut_theta_e_max = in.real("cut_theta_e_max");
    p.cut_b_min = in.real("cut_b_min"); p.cut_b_max = in.real("cut_b_max");
    p.cut_sigma_min = in.real("cut_sigma_min"); p.cut_sigma_max = in.real("cut_sigma_max");
    p.cut_beta_inverse_min = in.real("cut_beta_inverse_min"); p.cut_beta_inverse_max = in.real("cut_beta_inverse_max");
  }
  p.cut_omit_near = in.flag("cut_omit_near");
  p.cut_omit_far = in.flag("cut_omit_far");
  p.cut_omit_in = in.real("cut_omit_in");
  p.cut_omit_out = in.real("cut_omit_out");
  p.cut_midplane_theta = in.real("cut_midplane_theta") * kPi / 180.0;
  p.cut_midplane_z = in.real("cut_midplane_z");
  p.cut_plane = in.flag("cut_plane");
  if (p.cut_plane) {
    in.triple("cut_plane_origin", p.cut_plane_origin);
    in.triple("cut_plane_normal", p.cut_plane_normal);
  }

  // fallback
  p.fallback_nan = in.flag("fallback_nan");
  if (sim && !p.fallback_nan) {
    p.fallback_rho = in.real32("fallback_rho");
    p.fallback_pgas = in.real32("fallback_pgas");
    if (p.plasma_model == BL_PLASMA_CODE_KAPPA) p.fallback_kappa = in.real32("fallback_kappa");
  }

  // camera frame
  c.frame = build_camera_frame(cam);
  for (int m = 0; m < 4; m++) {
    p.camera_x[m] = c.frame.x[m];
    p.camera_u_con[m] = c.frame.u_con[m];
    p.camera_u_cov[m] = c.frame.u_cov[m];
    p.camera_vert_con_c[m] = c.frame.vert_con_c[m];
  }
  return c;
}

bl_camera make_bl_camera(const RunConfig &cfg) {
  bl_camera c{};
  c.type = cfg.camera.type == 0 ? BL_CAMERA_PLANE : BL_CAMERA_PINHOLE;
  c.normalization = cfg.camera.normalization == 0 ? BL_NORM_CAMERA : BL_NORM_INFINITY;
  c.width = cfg.camera.width;
  c.r = cfg.camera.r;
  for (int m = 0; m < 4; m++) {
    c.x[m] = cfg.frame.x[m];
    c.u_con[m] = cfg.frame.u_con[m];
    c.u_cov[m] = cfg.frame.u_cov[m];
    c.norm_con[m] = cfg.frame.norm_con[m];
    c.norm_con_c[m] = cfg.frame.norm_con_c[m];
    c.hor_con_c[m] = cfg.frame.hor_con_c[m];
    c.vert_con_c[m] = cfg.frame.vert_con_c[m];
  }
  return c;
}

}  // namespace blh
