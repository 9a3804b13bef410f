#include "input_file.hpp"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <fstream>
#include <set>
#include <sstream>

namespace blh {

void warning(const std::string &message) { std::fprintf(stderr, "Warning: %s\n", message.c_str()); }

namespace {

// Every fixed key of the reference's schema (input/example.input is the master list).
const std::set<std::string> &fixed_keys() {
  static const std::set<std::string> keys = {
      "model_type", "num_threads", "output_format", "output_file", "output_camera",
      "checkpoint_geodesic_save", "checkpoint_geodesic_load", "checkpoint_geodesic_file",
      "checkpoint_sample_save", "checkpoint_sample_load", "checkpoint_sample_file",
      "simulation_format", "simulation_file", "simulation_multiple", "simulation_start", "simulation_end",
      "simulation_coord", "simulation_a", "simulation_m_msun", "simulation_rho_cgs", "simulation_kappa_name",
      "simulation_interp", "simulation_block_interp",
      "formula_mass", "formula_spin", "formula_r0", "formula_h", "formula_l0", "formula_q", "formula_nup",
      "formula_cn0", "formula_alpha", "formula_a", "formula_beta",
      "camera_type", "camera_r", "camera_th", "camera_ph", "camera_urn", "camera_uthn", "camera_uphn",
      "camera_k_r", "camera_k_th", "camera_k_ph", "camera_rotation", "camera_width", "camera_resolution",
      "ray_flat", "ray_terminate", "ray_factor", "ray_integrator", "ray_step", "ray_max_steps",
      "ray_max_retries", "ray_tol_abs", "ray_tol_rel",
      "image_light", "image_num_frequencies", "image_frequency", "image_frequency_start", "image_frequency_end",
      "image_frequency_spacing", "image_normalization", "image_polarization", "image_rotation_split",
      "image_time", "image_length", "image_lambda", "image_emission", "image_tau", "image_lambda_ave",
      "image_emission_ave", "image_tau_int", "image_crossings",
      "slow_light_on", "slow_interp", "slow_chunk_size", "slow_t_start", "slow_dt", "slow_num_images", "slow_offset",
      "adaptive_max_level", "adaptive_block_size", "adaptive_frequency_num", "adaptive_val_cut", "adaptive_val_frac",
      "adaptive_abs_grad_cut", "adaptive_abs_grad_frac", "adaptive_rel_grad_cut", "adaptive_rel_grad_frac",
      "adaptive_abs_lapl_cut", "adaptive_abs_lapl_frac", "adaptive_rel_lapl_cut", "adaptive_rel_lapl_frac",
      "adaptive_num_regions",
      "plasma_mu", "plasma_ne_ni", "plasma_model", "plasma_use_p", "plasma_gamma", "plasma_gamma_i", "plasma_gamma_e",
      "plasma_rat_low", "plasma_rat_high", "plasma_power_frac", "plasma_p", "plasma_gamma_min", "plasma_gamma_max",
      "plasma_kappa_frac", "plasma_kappa", "plasma_w",
      "cut_rho_min", "cut_rho_max", "cut_n_e_min", "cut_n_e_max", "cut_p_gas_min", "cut_p_gas_max",
      "cut_theta_e_min", "cut_theta_e_max", "cut_b_min", "cut_b_max", "cut_sigma_min", "cut_sigma_max",
      "cut_beta_inverse_min", "cut_beta_inverse_max", "cut_omit_near", "cut_omit_far", "cut_omit_in", "cut_omit_out",
      "cut_midplane_theta", "cut_midplane_z", "cut_plane", "cut_plane_origin", "cut_plane_normal",
      "fallback_nan", "fallback_rho", "fallback_pgas", "fallback_kappa"};
  return keys;
}

bool ends_with(const std::string &s, const char *suffix) {
  std::string t(suffix);
  return s.size() >= t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0;
}

// render_<i>_num_features, render_<i>_<f>_{quantity,type,min,max,thresh,tau_scale,opacity,rgb,xyz}, render_num_images
bool render_key_ok(const std::string &rest) {
  if (rest == "num_images") return true;
  static const char *suffixes[] = {"_num_features", "_quantity", "_type", "_min", "_max", "_thresh",
                                   "_tau_scale", "_opacity", "_rgb", "_xyz"};
  for (const char *s : suffixes)
    if (ends_with(rest, s) && rest.size() > std::string(s).size()) return true;
  return false;
}

bool region_key_ok(const std::string &rest) {
  static const char *suffixes[] = {"_level", "_x_min", "_x_max", "_y_min", "_y_max"};
  for (const char *s : suffixes)
    if (ends_with(rest, s) && rest.size() > std::string(s).size()) return true;
  return false;
}

}  // namespace

InputFile::InputFile(const std::string &path) {
  std::ifstream in(path);
  if (!in.is_open()) throw Error("Could not open input file.");
  for (std::string line; std::getline(in, line);) {
    line.erase(std::remove_if(line.begin(), line.end(), [](unsigned char c) { return std::isspace(c) != 0; }), line.end());
    std::string::size_type pos = line.find('#');
    if (pos != std::string::npos) line.erase(pos);
    if (line.empty()) continue;
    pos = line.find('=');
    if (pos == std::string::npos) throw Error("Invalid assignment in input file.");
    std::string key = line.substr(0, pos), val = line.substr(pos + 1);
    bool known = fixed_keys().count(key) != 0;
    if (!known && key.compare(0, 7, "render_") == 0) {
      if (!render_key_ok(key.substr(7))) throw Error("Unknown key (render_" + key.substr(7) + ") in input file.");
      known = true;
    }
    if (!known && key.compare(0, 16, "adaptive_region_") == 0) {
      if (!region_key_ok(key.substr(16))) throw Error("Unknown key (" + key + ") in input file.");
      known = true;
    }
    if (!known) throw Error("Unknown key (" + key + ") in input file.");
    values_[key] = val;
    if (key == "camera_th") {
      double th = std::stod(val);
      camera_pole_ = th == 0.0 || th == 180.0;
    }
  }
}

const std::string &InputFile::str(const std::string &key) const {
  auto it = values_.find(key);
  if (it == values_.end()) throw Error("Missing input parameter: " + key + ".");
  return it->second;
}

bool InputFile::flag(const std::string &key) const {
  const std::string &v = str(key);
  if (v == "true") return true;
  if (v == "false") return false;
  throw Error("Unknown string used for boolean value.");
}

int InputFile::integer(const std::string &key) const { return std::stoi(str(key)); }
double InputFile::real(const std::string &key) const { return std::stod(str(key)); }
float InputFile::real32(const std::string &key) const { return std::stof(str(key)); }

void InputFile::triple(const std::string &key, double out[3]) const {
  const std::string &s = str(key);
  std::size_t p1 = 0, p2 = 0;
  out[0] = std::stod(s, &p1);
  out[1] = std::stod(s.substr(p1 + 1), &p2);
  out[2] = std::stod(s.substr(p1 + p2 + 2));
  if (s[p1] != ',' || s[p1 + p2 + 1] != ',') throw Error("Invalid triple (" + s + ") in input file.");
}

int InputFile::choice(const std::string &key, const std::vector<std::string> &names, const char *type_name) const {
  const std::string &v = str(key);
  for (std::size_t i = 0; i < names.size(); i++)
    if (names[i] == v) return (int)i;
  throw Error(std::string("Unknown string used for ") + type_name + " value.");
}

int InputFile::num_runs() const {
  if (str("model_type") == "simulation" && flag("simulation_multiple")) {
    if (flag("slow_light_on")) return integer("slow_num_images");
    return integer("simulation_end") - integer("simulation_start") + 1;
  }
  return 1;
}

}  // namespace blh
