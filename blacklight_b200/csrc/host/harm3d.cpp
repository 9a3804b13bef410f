#include "harm3d.hpp"

#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <vector>

#include "input_file.hpp"

namespace blh {

namespace {

constexpr double kPi = 3.141592653589793;
constexpr double kAngularDomainTolerance = 0.1;   // simulation_reader.hpp:100

struct Header {
  double time = 0.0, x_start[3] = {0, 0, 0}, dx[3] = {0, 0, 0}, a = 0.0, gamma = 0.0, h = 1.0;
  int n[3] = {0, 0, 0};
  std::streampos cell_data;
};

Header read_header(std::ifstream &in) {
  Header hd;
  double skip;
  in >> hd.time;
  in >> hd.n[0] >> hd.n[1] >> hd.n[2];
  in >> hd.x_start[0] >> hd.x_start[1] >> hd.x_start[2];
  in >> hd.dx[0] >> hd.dx[1] >> hd.dx[2];
  in >> hd.a >> hd.gamma >> skip >> hd.h >> skip;
  if (!in || hd.n[0] <= 0 || hd.n[1] <= 0 || hd.n[2] <= 0) throw Error("Could not read harm3d header.");
  in.seekg(1, std::ios_base::cur);
  hd.cell_data = in.tellg();
  return hd;
}

// faces x_start + (i+1) dx, centres as face averages (simulation_reader.cpp:669-693)
void uniform_axis(double start, double dx, int n, std::vector<double> &f, std::vector<double> &v) {
  f.assign((size_t)n + 1, 0.0);
  v.assign((size_t)n, 0.0);
  f[0] = start;
  for (int i = 0; i < n; i++) {
    f[(size_t)i + 1] = start + (i + 1) * dx;
    v[(size_t)i] = 0.5 * (f[(size_t)i] + f[(size_t)i + 1]);
  }
}

}  // namespace

void read_harm3d_header(const std::string &path, double *time, double *gamma_adi) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw Error("Could not open file for reading.");
  Header hd = read_header(in);
  if (time) *time = hd.time;
  if (gamma_adi) *gamma_adi = hd.gamma;
}

void read_harm3d(const std::string &path, bool want_kappa, bool gamma_set, double *plasma_gamma, double simulation_a,
                 bool reuse_layout, AthenaGrid &g) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw Error("Could not open file for reading.");
  Header hd = read_header(in);
  g.time = hd.time;
  std::vector<double> &x2v_mod = g.layout.x2v_mod;   // x2 centres in modified coordinates (for the Jacobian)
  double &metric_h = g.layout.h;
  if (reuse_layout && (g.n_b != 1 || g.n_i != hd.n[0] || g.n_j != hd.n[1] || g.n_k != hd.n[2] || (int)x2v_mod.size() != g.n_j))
    throw Error("harm3d file does not match the layout of the first snapshot of the series.");
  if (!reuse_layout) {
    g.n_b = 1;
    g.n_i = hd.n[0]; g.n_j = hd.n[1]; g.n_k = hd.n[2];
    g.levels.assign(1, 0);
    g.locations.assign(3, 0);
    g.n_3_root = g.n_k;
    uniform_axis(hd.x_start[0], hd.dx[0], g.n_i, g.x1f, g.x1v);
    uniform_axis(hd.x_start[1], hd.dx[1], g.n_j, g.x2f, g.x2v);
    uniform_axis(hd.x_start[2], hd.dx[2], g.n_k, g.x3f, g.x3v);
    if (hd.a != simulation_a) {
      std::ostringstream msg;
      msg << "Given spin of " << simulation_a << " does not match file value of " << hd.a << "; ignoring the latter.";
      warning(msg.str());
    }
    if (!gamma_set) {
      *plasma_gamma = hd.gamma;
    } else if (*plasma_gamma != hd.gamma) {
      std::ostringstream msg;
      msg << "Given total adiabatic index of " << *plasma_gamma << " does not match file value of " << hd.gamma
          << "; ignoring the latter.";
      warning(msg.str());
    }
    metric_h = hd.h;
    // modified -> spherical Kerr-Schild coordinates (simulation_geometry.cpp:62-90)
    x2v_mod = g.x2v;
    for (double &x : g.x1f) x = std::exp(x);
    for (double &x : g.x1v) x = std::exp(x);
    for (double &x : g.x2f) x = kPi * x + (1.0 - metric_h) / 2.0 * std::sin(2.0 * kPi * x);
    for (double &x : g.x2v) x = kPi * x + (1.0 - metric_h) / 2.0 * std::sin(2.0 * kPi * x);
    // snap the angular ranges to [0, pi] x [0, 2 pi] (simulation_reader.cpp:722-757)
    auto snap = [](std::vector<double> &f, double upper, const char *name, const char *range) {
      size_t n = f.size();
      bool low = std::abs(f[0]) > (f[1] - f[0]) * kAngularDomainTolerance;
      bool high = std::abs(f[n - 1] - upper) > (f[n - 1] - f[n - 2]) * kAngularDomainTolerance;
      if (low || high) {
        std::ostringstream msg;
        msg << std::scientific << std::setprecision(16) << "Changing " << name << " range from [" << f[0] << ", " << f[n - 1]
            << "] to " << range << ".";
        warning(msg.str());
        f[0] = 0.0;
        f[n - 1] = upper;
      }
    };
    snap(g.x2f, kPi, "theta", "[0, pi]");
    snap(g.x3f, 2.0 * kPi, "phi", "[0, 2*pi]");
    g.n_var = want_kappa ? 11 : 10;
    g.ind_rho = 0; g.ind_pgas = 1; g.ind_kappa = want_kappa ? 10 : -1;
    g.ind_uu1 = 3; g.ind_uu2 = 4; g.ind_uu3 = 5;
    g.ind_bb1 = 7; g.ind_bb2 = 8; g.ind_bb3 = 9;
  }
  const int n1 = g.n_i, n2 = g.n_j, n3 = g.n_k, nv = g.n_var;
  const size_t cells = (size_t)n1 * n2 * n3;
  std::vector<float> record((size_t)(nv + 6) * cells);
  in.seekg(hd.cell_data);
  in.read(reinterpret_cast<char *>(record.data()), (std::streamsize)(record.size() * sizeof(float)));
  if (!in) throw Error("Unexpected end of harm3d file.");
  g.prim.assign((size_t)nv * cells, 0.0f);
  // file: variable fastest, then x3, x2, x1 slowest; ours: (var, k, j, i)
#pragma omp parallel for schedule(static) collapse(2)
  for (int v = 0; v < nv; v++)
    for (int k = 0; k < n3; k++)
      for (int j = 0; j < n2; j++)
        for (int i = 0; i < n1; i++)
          g.prim[(((size_t)v * n3 + k) * n2 + j) * n1 + i] = record[(((size_t)i * n2 + j) * n3 + k) * (nv + 6) + v + 6];
  const float gm1 = static_cast<float>(*plasma_gamma - 1.0);
  for (size_t c = 0; c < cells; c++) g.prim[(size_t)g.ind_pgas * cells + c] *= gm1;
  // coordinate-frame four-vectors -> normal-frame velocity and coordinate-frame field in standard coordinates
  // (ConvertPrimitives4, simulation_geometry.cpp:242-327)
  const double a = simulation_a;
  auto at = [&](int v, int k, int j, int i) -> float & { return g.prim[(((size_t)v * n3 + k) * n2 + j) * n1 + i]; };
  const std::vector<double> &x2_mod = x2v_mod;
  const double h_slope = metric_h;
#pragma omp parallel for schedule(static) collapse(2)
  for (int k = 0; k < n3; k++)
    for (int j = 0; j < n2; j++)
      for (int i = 0; i < n1; i++) {
        double r = g.x1v[(size_t)i], th = g.x2v[(size_t)j], cth = std::cos(th), x2 = x2_mod[(size_t)j];
        double u0 = at(2, k, j, i), u1 = at(3, k, j, i), u2 = at(4, k, j, i), u3 = at(5, k, j, i);
        double b0 = at(6, k, j, i), b1 = at(7, k, j, i), b2 = at(8, k, j, i), b3 = at(9, k, j, i);
        double dr_dx1 = r;
        double dth_dx2 = kPi + (1.0 - h_slope) * kPi * std::cos(2.0 * kPi * x2);
        double sigma = r * r + a * a * cth * cth;
        double f = 2.0 * r / sigma;
        double gtt = -(1.0 + f), gtr = f, gtth = 0.0, gtph = 0.0;
        double alpha = 1.0 / std::sqrt(-gtt);
        double ut = u0, ur = dr_dx1 * u1, uth = dth_dx2 * u2, uph = u3;
        double uur = ur + alpha * alpha * gtr * ut;
        double uuth = uth + alpha * alpha * gtth * ut;
        double uuph = uph + alpha * alpha * gtph * ut;
        double bt = b0, br = dr_dx1 * b1, bth = dth_dx2 * b2, bph = b3;
        double bbr = br * ut - bt * ur, bbth = bth * ut - bt * uth, bbph = bph * ut - bt * uph;
        at(3, k, j, i) = static_cast<float>(uur);
        at(4, k, j, i) = static_cast<float>(uuth);
        at(5, k, j, i) = static_cast<float>(uuph);
        at(7, k, j, i) = static_cast<float>(bbr);
        at(8, k, j, i) = static_cast<float>(bbth);
        at(9, k, j, i) = static_cast<float>(bbph);
      }
}

}  // namespace blh
