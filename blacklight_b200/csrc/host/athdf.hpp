// Minimal reader for Athena++ .athdf snapshots (HDF5 superblock v0/v1, symbol-table groups, v1 object
// headers, contiguous datasets) -- the subset the reference's hand-written parser accepts
// (reference src/simulation_reader/hdf5_format_*.cpp, simulation_reader.cpp:304-351,591-621,762-781,
// 1141-1216).  Produces the host arrays the C ABI's bl_grid_view points at.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/blacklight_b200.h"

namespace blh {

// What the first file of a series fixes and later files of the series (reuse_layout) rely on.  It belongs to the grid
// that was read, so two open snapshots -- or two threads -- never see each other's.
struct ReaderLayout {
  // athenak: position in the file of each internal variable, bytes per block record, and what the layout was taken from
  int file_ind[9] = {0, 0, 0, 0, 0, 0, 0, 0, -1};
  long block_bytes = 0;
  int num_file_variables = 0, location_size = 0, variable_size = 0;
  // harm3d / iharm3d: x2 cell centres in the modified coordinates (for the Jacobian) and the coordinate parameters
  // (simulation_reader.cpp:362-428)
  std::vector<double> x2v_mod;
  bool fmks = false;
  double a = 0.0, h = 1.0, r_in = 0.0, poly_xt = 0.0, poly_alpha = 0.0, mks_smooth = 0.0, poly_norm = 0.0;
};

struct AthenaGrid {
  ReaderLayout layout;
  int n_b = 0, n_k = 0, n_j = 0, n_i = 0, n_var = 0;
  std::vector<int32_t> levels, locations;
  std::vector<double> x1f, x2f, x3f, x1v, x2v, x3v;  // float32 file values widened to double
  std::vector<float> prim;                           // (n_var, n_b, n_k, n_j, n_i): "prim" then "B"
  int ind_rho = -1, ind_pgas = -1, ind_kappa = -1, ind_uu1 = -1, ind_uu2 = -1, ind_uu3 = -1;
  int ind_bb1 = -1, ind_bb2 = -1, ind_bb3 = -1;
  int n_3_root = 0;
  double time = 0.0;
  // simulation_coord = fmks only: map from spherical Kerr-Schild (r, theta) to the native (x1, x2), its sampling
  // and the grid's extent in (r, theta, phi) (simulation_reader.hpp:103-112)
  std::vector<double> sks_map;                       // (2, sks_map_n2, sks_map_n1)
  int sks_map_n1 = 0, sks_map_n2 = 0;
  double sks_map_r_in = 0.0, sks_map_dr = 0.0, sks_map_dtheta = 0.0;
  double simulation_bounds[6] = {0, 0, 0, 0, 0, 0};
  bl_grid_view view() const;
};

// kappa_name: electron-entropy variable to locate when plasma_model = code_kappa ("" = none).
// reuse_layout: keep coordinates of `grid` (already read from the first snapshot) and only refresh cell data.
void read_athdf(const std::string &path, const std::string &kappa_name, bool reuse_layout, AthenaGrid &grid);
// only the file's Time attribute (slow light looks ahead for the first snapshot late enough)
double read_athdf_time(const std::string &path);

// File name of snapshot `number` from a pattern holding one `{Nd}` field (simulation_reader.cpp:870-904)
std::string format_numbered(const std::string &pattern, int number, const char *what);

}  // namespace blh
