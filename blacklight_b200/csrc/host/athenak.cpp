#include "athenak.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

#include "input_file.hpp"

namespace blh {

namespace {

struct Header {
  double time = 0.0;
  int location_size = 0, variable_size = 0;
  std::vector<std::string> names;
  std::streamoff params_begin = 0, data_begin = 0;
};

// "  key=value" line of the pre-header (simulation_reader.cpp:924-1010)
std::string keyed(std::ifstream &in, const char *key) {
  std::string line;
  std::getline(in, line);
  size_t n = std::strlen(key);
  if (!in || line.compare(0, n, key) != 0) throw Error("Invalid AthenaK file header.");
  return line.substr(n);
}

Header read_header(std::ifstream &in) {
  Header hd;
  std::string line;
  std::getline(in, line);
  if (!in || line != "Athena binary output version=1.1") throw Error("Unknown AthenaK file format.");
  std::getline(in, line);                                  // size of preheader
  hd.time = std::strtod(keyed(in, "  time=").c_str(), nullptr);
  std::getline(in, line);                                  // cycle
  hd.location_size = std::atoi(keyed(in, "  size of location=").c_str());
  if (hd.location_size != 4 && hd.location_size != 8) throw Error("Unsupported size of location.");
  hd.variable_size = std::atoi(keyed(in, "  size of variable=").c_str());
  if (hd.variable_size != 4 && hd.variable_size != 8) throw Error("Unsupported size of variables.");
  int num = std::atoi(keyed(in, "  number of variables=").c_str());
  std::istringstream names(keyed(in, "  variables:"));
  std::string name;
  while ((int)hd.names.size() < num && names >> name) hd.names.push_back(name);
  if ((int)hd.names.size() != num || num <= 0) throw Error("Invalid AthenaK file header.");
  long offset = std::atol(keyed(in, "  header offset=").c_str());
  hd.params_begin = in.tellg();
  hd.data_begin = hd.params_begin + offset;
  return hd;
}

// The dump's copy of the simulation's parameter file: <section> lines and `name = value` lines
// (simulation_reader.cpp:1024-1131).  Returns whether <mhd> gamma was present.
bool read_parameters(std::ifstream &in, const Header &hd, AthenaKExpect *expect, double *gamma_out) {
  in.seekg(hd.params_begin);
  std::string line, section;
  bool found = false;
  auto mismatch = [](const char *what, double given, double file) {
    std::ostringstream msg;
    msg << "Given " << what << " of " << given << " does not match file value of " << file << "; ignoring the latter.";
    warning(msg.str());
  };
  while (in.tellg() < hd.data_begin) {
    if (!std::getline(in, line)) break;
    if (line.empty() || line[0] == '#') continue;
    if (line.front() == '<' && line.back() == '>') {
      section = line.substr(1, line.size() - 2);
      continue;
    }
    size_t eq = line.find('=');
    if (eq == std::string::npos) throw Error("Error parsing inputs in AthenaK file.");
    std::string key = line.substr(0, eq);
    key.erase(std::remove(key.begin(), key.end(), ' '), key.end());
    auto value = [&]() { return std::stod(line.substr(eq + 1)); };
    if (expect && section == "coord" && key == "a" && value() != expect->simulation_a)
      mismatch("spin", expect->simulation_a, value());
    if (expect && section == "units" && key == "bhmass_msun" && value() != expect->simulation_m_msun)
      mismatch("mass", expect->simulation_m_msun, value());
    if (expect && section == "units" && key == "density_cgs" && value() != expect->simulation_rho_cgs)
      mismatch("density scale", expect->simulation_rho_cgs, value());
    if (expect && section == "units" && key == "mu" && value() != expect->plasma_mu)
      mismatch("density scale", expect->plasma_mu, value());   // the reference's wording (:1104-1107)
    if (section == "mhd" && key == "gamma") {
      double g = value();
      if (expect && expect->gamma_set && expect->plasma_gamma != g)
        mismatch("total adiabatic index", expect->plasma_gamma, g);
      else if (expect && !expect->gamma_set)
        expect->plasma_gamma = g;
      if (gamma_out) *gamma_out = g;
      found = true;
    }
  }
  if (!found) throw Error("Missing adiabatic index.");
  return found;
}

int locate(const Header &hd, const std::string &name, const char *message) {
  for (size_t n = 0; n < hd.names.size(); n++)
    if (hd.names[n] == name) return (int)n;
  throw Error(message);
}

// n equal cells between the two block edges; centres are face averages (simulation_reader.cpp:508-529)
void block_axis(double lo, double hi, int n, double *f, double *v) {
  f[0] = lo;
  f[n] = hi;
  double d = (hi - lo) / n;
  for (int i = 1; i < n; i++) f[i] = lo + i * d;
  for (int i = 0; i < n; i++) v[i] = 0.5 * (f[i] + f[i + 1]);
}

}  // namespace

void read_athenak_header(const std::string &path, double *time, double *gamma_adi) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw Error("Could not open file for reading.");
  Header hd = read_header(in);
  if (time) *time = hd.time;
  if (gamma_adi) read_parameters(in, hd, nullptr, gamma_adi);
}

void read_athenak(const std::string &path, const std::string &kappa_name, bool reuse_layout, AthenaKExpect &expect,
                  AthenaGrid &g) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw Error("Could not open file for reading.");
  Header hd = read_header(in);
  g.time = hd.time;
  const bool want_kappa = !kappa_name.empty();
  ReaderLayout &layout = g.layout;   // found in the first file of the series, kept with the grid
  if (reuse_layout) {
    // a later file must have the records the stored layout describes
    if (g.n_b <= 0 || layout.block_bytes <= 0) throw Error("AthenaK series: no first snapshot to take the layout from.");
    if ((int)hd.names.size() != layout.num_file_variables || hd.location_size != layout.location_size ||
        hd.variable_size != layout.variable_size)
      throw Error("AthenaK file does not match the layout of the first snapshot of the series.");
    in.seekg(0, std::ios::end);
    if (((long)in.tellg() - (long)hd.data_begin) / layout.block_bytes != g.n_b)
      throw Error("AthenaK file does not match the layout of the first snapshot of the series.");
  }
  if (!reuse_layout) {
    layout = ReaderLayout();
    layout.num_file_variables = (int)hd.names.size();
    layout.location_size = hd.location_size;
    layout.variable_size = hd.variable_size;
    // internal order rho, uu1, uu2, uu3, pgas, bb1, bb2, bb3, [kappa] (simulation_reader.cpp:1279-1288)
    layout.file_ind[0] = locate(hd, "dens", "Unable to locate \"dens\" values in data file.");
    layout.file_ind[4] = locate(hd, "eint", "Unable to locate \"eint\" values in data file.");
    if (want_kappa) layout.file_ind[8] = locate(hd, kappa_name, "Unable to locate electron entropy values in data file.");
    layout.file_ind[1] = locate(hd, "velx", "Unable to locate \"velx\" values in data file.");
    layout.file_ind[2] = locate(hd, "vely", "Unable to locate \"vely\" values in data file.");
    layout.file_ind[3] = locate(hd, "velz", "Unable to locate \"velz\" values in data file.");
    layout.file_ind[5] = locate(hd, "bcc1", "Unable to locate \"bcc1\" values in data file.");
    layout.file_ind[6] = locate(hd, "bcc2", "Unable to locate \"bcc2\" values in data file.");
    layout.file_ind[7] = locate(hd, "bcc3", "Unable to locate \"bcc3\" values in data file.");
    read_parameters(in, hd, &expect, nullptr);
    expect.gamma_set = true;

    in.seekg(hd.data_begin);
    int32_t bounds[6];
    in.read(reinterpret_cast<char *>(bounds), sizeof bounds);
    if (!in) throw Error("Unexpected end of AthenaK file.");
    g.n_i = bounds[1] - bounds[0] + 1;
    g.n_j = bounds[3] - bounds[2] + 1;
    g.n_k = bounds[5] - bounds[4] + 1;
    if (g.n_i <= 0 || g.n_j <= 0 || g.n_k <= 0) throw Error("Invalid AthenaK block size.");
    const long cells = (long)g.n_k * g.n_j * g.n_i;
    layout.block_bytes = 24 + 16 + 6L * hd.location_size + (long)hd.names.size() * cells * hd.variable_size;
    in.seekg(0, std::ios::end);
    // Complete block records only.  (With libstdc++ >= 11 the reference's counting loop, :449-454, runs once more
    // than there are records and carries a last block of unread memory; it repeats the coordinates of the block
    // before it, so the first-match block search never selects it.)
    g.n_b = (int)(((long)in.tellg() - (long)hd.data_begin) / layout.block_bytes);
    if (g.n_b <= 0) throw Error("Unexpected end of AthenaK file.");
    g.levels.assign((size_t)g.n_b, 0);
    g.locations.assign((size_t)g.n_b * 3, 0);
    g.x1f.assign((size_t)g.n_b * (g.n_i + 1), 0.0);
    g.x2f.assign((size_t)g.n_b * (g.n_j + 1), 0.0);
    g.x3f.assign((size_t)g.n_b * (g.n_k + 1), 0.0);
    g.x1v.assign((size_t)g.n_b * g.n_i, 0.0);
    g.x2v.assign((size_t)g.n_b * g.n_j, 0.0);
    g.x3v.assign((size_t)g.n_b * g.n_k, 0.0);
    g.n_var = want_kappa ? 9 : 8;
    g.ind_rho = 0; g.ind_uu1 = 1; g.ind_uu2 = 2; g.ind_uu3 = 3; g.ind_pgas = 4;
    g.ind_bb1 = 5; g.ind_bb2 = 6; g.ind_bb3 = 7; g.ind_kappa = want_kappa ? 8 : -1;
    g.n_3_root = 0;
  }
  const long cells = (long)g.n_k * g.n_j * g.n_i;
  const size_t plane = (size_t)g.n_b * cells;
  g.prim.assign((size_t)g.n_var * plane, 0.0f);
  std::vector<double> wide(hd.variable_size == 8 ? (size_t)cells : 0);
  for (int b = 0; b < g.n_b; b++) {
    std::streamoff at = hd.data_begin + (std::streamoff)b * layout.block_bytes + 24;
    in.seekg(at);
    if (!reuse_layout) {
      int32_t where[4];
      in.read(reinterpret_cast<char *>(where), sizeof where);
      for (int d = 0; d < 3; d++) g.locations[(size_t)b * 3 + d] = where[d];
      g.levels[(size_t)b] = where[3];
      double face[6];
      if (hd.location_size == 4) {
        float single[6];
        in.read(reinterpret_cast<char *>(single), sizeof single);
        for (int d = 0; d < 6; d++) face[d] = single[d];
      } else {
        in.read(reinterpret_cast<char *>(face), sizeof face);
      }
      if (!in) throw Error("Unexpected end of AthenaK file.");
      block_axis(face[0], face[1], g.n_i, &g.x1f[(size_t)b * (g.n_i + 1)], &g.x1v[(size_t)b * g.n_i]);
      block_axis(face[2], face[3], g.n_j, &g.x2f[(size_t)b * (g.n_j + 1)], &g.x2v[(size_t)b * g.n_j]);
      block_axis(face[4], face[5], g.n_k, &g.x3f[(size_t)b * (g.n_k + 1)], &g.x3v[(size_t)b * g.n_k]);
    }
    std::streamoff cell_data = at + 16 + 6 * hd.location_size;
    for (int v = 0; v < g.n_var; v++) {
      float *dst = &g.prim[(size_t)v * plane + (size_t)b * cells];
      in.seekg(cell_data + (std::streamoff)layout.file_ind[v] * cells * hd.variable_size);
      if (hd.variable_size == 4) {
        in.read(reinterpret_cast<char *>(dst), cells * 4);
      } else {
        in.read(reinterpret_cast<char *>(wide.data()), cells * 8);
        for (long c = 0; c < cells; c++) dst[c] = static_cast<float>(wide[(size_t)c]);
      }
      if (!in) throw Error("Unexpected end of AthenaK file.");
    }
  }
  // internal energy -> pressure, in single precision (simulation_reader.cpp:581-587)
  const float gm1 = static_cast<float>(expect.plasma_gamma - 1.0);
  float *pgas = &g.prim[(size_t)g.ind_pgas * plane];
  for (size_t c = 0; c < plane; c++) pgas[c] *= gm1;
}

}  // namespace blh
