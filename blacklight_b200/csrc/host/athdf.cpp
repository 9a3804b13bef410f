#include "athdf.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>

#include "h5file.hpp"
#include "input_file.hpp"

namespace blh {

namespace {

using namespace h5;

void coordinate_dataset(const H5File &f, const std::string &name, int n_b, std::vector<double> &out, int &n) {
  Datatype dt;
  std::vector<uint64_t> dims;
  const uint8_t *d = f.dataset(name, dt, dims);
  if (dt.cls != 1 || dt.size != 4 || dims.size() != 2 || (int)dims[0] != n_b) throw Error("Unexpected layout of dataset " + name + ".");
  n = (int)dims[1];
  out.resize((size_t)n_b * n);
  for (size_t i = 0; i < out.size(); i++) {
    float v;
    std::memcpy(&v, d + 4 * i, 4);
    out[i] = static_cast<double>(v);
  }
}

int find_name(const std::vector<std::string> &names, int begin, int end, const std::string &want) {
  for (int i = begin; i < end; i++)
    if (names[(size_t)i] == want) return i;
  return -1;
}

}  // namespace

bl_grid_view AthenaGrid::view() const {
  bl_grid_view v{};
  v.n_b = n_b; v.n_k = n_k; v.n_j = n_j; v.n_i = n_i; v.n_var = n_var;
  v.levels = levels.data(); v.locations = locations.data();
  v.x1f = x1f.data(); v.x2f = x2f.data(); v.x3f = x3f.data();
  v.x1v = x1v.data(); v.x2v = x2v.data(); v.x3v = x3v.data();
  v.prim = prim.data();
  v.ind_rho = ind_rho; v.ind_pgas = ind_pgas; v.ind_kappa = ind_kappa;
  v.ind_uu1 = ind_uu1; v.ind_uu2 = ind_uu2; v.ind_uu3 = ind_uu3;
  v.ind_bb1 = ind_bb1; v.ind_bb2 = ind_bb2; v.ind_bb3 = ind_bb3;
  v.n_3_root = n_3_root;
  v.sks_map = sks_map.empty() ? nullptr : sks_map.data();
  v.sks_map_n1 = sks_map_n1; v.sks_map_n2 = sks_map_n2;
  v.sks_map_r_in = sks_map_r_in; v.sks_map_dr = sks_map_dr; v.sks_map_dtheta = sks_map_dtheta;
  for (int d = 0; d < 6; d++) v.simulation_bounds[d] = simulation_bounds[d];
  return v;
}

double read_athdf_time(const std::string &path) {
  H5File f(path);
  Datatype dt;
  size_t count = 0;
  const uint8_t *d = f.attribute("Time", dt, count);
  if (dt.cls != 1 || dt.size != 4) throw Error("Unexpected HDF5 datatype for attribute Time.");
  float t;
  std::memcpy(&t, d, 4);
  return t;
}

void read_athdf(const std::string &path, const std::string &kappa_name, bool reuse_layout, AthenaGrid &g) {
  H5File f(path);
  {
    Datatype dt;
    size_t count = 0;
    const uint8_t *d = f.attribute("Time", dt, count);
    if (dt.cls != 1 || dt.size != 4) throw Error("Unexpected HDF5 datatype for attribute Time.");
    float t;
    std::memcpy(&t, d, 4);
    g.time = t;
  }
  if (!reuse_layout) {
    std::vector<int32_t> root = int_attribute(f, "RootGridSize");
    if (root.size() != 3) throw Error("Unexpected RootGridSize in data file.");
    g.n_3_root = root[2];
    Datatype dt;
    std::vector<uint64_t> dims;
    const uint8_t *d = f.dataset("Levels", dt, dims);
    if (dims.size() != 1) throw Error("Unexpected layout of dataset Levels.");
    g.n_b = (int)dims[0];
    g.levels = int_values(d, dt, (size_t)g.n_b, "Levels");
    d = f.dataset("LogicalLocations", dt, dims);
    if (dims.size() != 2 || (int)dims[0] != g.n_b || dims[1] != 3) throw Error("Unexpected layout of dataset LogicalLocations.");
    g.locations = int_values(d, dt, (size_t)g.n_b * 3, "LogicalLocations");
    int n1f, n2f, n3f;
    coordinate_dataset(f, "x1f", g.n_b, g.x1f, n1f);
    coordinate_dataset(f, "x2f", g.n_b, g.x2f, n2f);
    coordinate_dataset(f, "x3f", g.n_b, g.x3f, n3f);
    coordinate_dataset(f, "x1v", g.n_b, g.x1v, g.n_i);
    coordinate_dataset(f, "x2v", g.n_b, g.x2v, g.n_j);
    coordinate_dataset(f, "x3v", g.n_b, g.x3v, g.n_k);
    if (n1f != g.n_i + 1 || n2f != g.n_j + 1 || n3f != g.n_k + 1) throw Error("Inconsistent face and cell coordinate arrays.");

    // variable bookkeeping: datasets are stacked "prim" then "B"; indices refer to the stacked array
    std::vector<std::string> dataset_names = string_attribute(f, "DatasetNames");
    std::vector<std::string> variable_names = string_attribute(f, "VariableNames");
    std::vector<int32_t> num_variables = int_attribute(f, "NumVariables");
    if (num_variables.size() != dataset_names.size()) throw Error("Inconsistent dataset metadata in data file.");
    int ind_hydro = -1, ind_bb = -1, prim_off = 0, bb_off = 0, running = 0;
    for (size_t i = 0; i < dataset_names.size(); i++) {
      if (dataset_names[i] == "prim" && ind_hydro < 0) { ind_hydro = (int)i; prim_off = running; }
      if (dataset_names[i] == "B" && ind_bb < 0) { ind_bb = (int)i; bb_off = running; }
      running += num_variables[i];
    }
    if (ind_hydro < 0) throw Error("Unable to locate array \"prim\" in data file.");
    if (ind_bb < 0) throw Error("Unable to locate array \"B\" in data file.");
    int n_hydro = num_variables[(size_t)ind_hydro], n_bb = num_variables[(size_t)ind_bb];
    auto hydro = [&](const char *nm, const char *msg) {
      int i = find_name(variable_names, prim_off, prim_off + n_hydro, nm);
      if (i < 0) throw Error(msg);
      return i - prim_off;
    };
    g.ind_rho = hydro("rho", "Unable to locate \"rho\" slice of \"prim\" in data file.");
    g.ind_pgas = hydro("press", "Unable to locate \"press\" slice of \"prim\" in data file.");
    if (!kappa_name.empty()) g.ind_kappa = hydro(kappa_name.c_str(), "Unable to locate electron entropy slice of \"prim\" in data file.");
    g.ind_uu1 = hydro("vel1", "Unable to locate \"vel1\" slice of \"prim\" in data file.");
    g.ind_uu2 = hydro("vel2", "Unable to locate \"vel2\" slice of \"prim\" in data file.");
    g.ind_uu3 = hydro("vel3", "Unable to locate \"vel3\" slice of \"prim\" in data file.");
    auto field = [&](const char *nm, const char *msg) {
      int i = find_name(variable_names, bb_off, bb_off + n_bb, nm);
      if (i < 0) throw Error(msg);
      return n_hydro + (i - bb_off);
    };
    g.ind_bb1 = field("Bcc1", "Unable to locate \"Bcc1\" slice of \"prim\" in data file.");
    g.ind_bb2 = field("Bcc2", "Unable to locate \"Bcc2\" slice of \"prim\" in data file.");
    g.ind_bb3 = field("Bcc3", "Unable to locate \"Bcc3\" slice of \"prim\" in data file.");
    g.n_var = n_hydro + n_bb;
    g.prim.resize((size_t)g.n_var * g.n_b * g.n_k * g.n_j * g.n_i);
  }
  size_t cells = (size_t)g.n_b * g.n_k * g.n_j * g.n_i;
  size_t filled = 0;
  for (const char *name : {"prim", "B"}) {
    Datatype dt;
    std::vector<uint64_t> dims;
    const uint8_t *d = f.dataset(name, dt, dims);
    if (dt.cls != 1 || dt.size != 4 || dims.size() != 5 || (int)dims[1] != g.n_b || (int)dims[2] != g.n_k ||
        (int)dims[3] != g.n_j || (int)dims[4] != g.n_i)
      throw Error(std::string("Unexpected layout of dataset ") + name + ".");
    size_t n = (size_t)dims[0] * cells;
    if (filled + n > g.prim.size()) throw Error("Cell data larger than declared number of variables.");
    std::memcpy(g.prim.data() + filled, d, n * sizeof(float));
    filled += n;
  }
  if (filled != g.prim.size()) throw Error("Cell data smaller than declared number of variables.");
}

std::string format_numbered(const std::string &pattern, int number, const char *what) {
  std::string err = std::string("Invalid ") + what + " for multiple runs.";
  std::string::size_type open = pattern.find_first_of('{');
  if (open == std::string::npos) throw Error(err);
  std::string::size_type close = pattern.find_first_of('}', open);
  if (close == std::string::npos) throw Error(err);
  if (pattern[close - 1] != 'd') throw Error(err);
  int width = 0;
  if (close - open > 2) width = std::stoi(pattern.substr(open + 1, close - open - 2));
  char digits[32];
  int len = std::snprintf(digits, sizeof digits, "%d", number);
  std::string out = pattern.substr(0, open);
  for (int i = len; i < width; i++) out += '0';
  out += digits;
  out += pattern.substr(close + 1);
  return out;
}

}  // namespace blh
