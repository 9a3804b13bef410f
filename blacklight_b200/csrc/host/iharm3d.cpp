#include "iharm3d.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <iomanip>
#include <sstream>

#include "h5file.hpp"
#include "input_file.hpp"

namespace blh {

namespace {

using namespace h5;

constexpr double kPi = 3.141592653589793;
constexpr double kAngularDomainTolerance = 0.1;   // simulation_reader.hpp:100
constexpr int kMapN1 = 2048, kMapN2 = 2048, kMapMaxIter = 1000;   // :109-112
constexpr double kMapTol = 1.0e-8;

// parameters of the modified coordinates (simulation_reader.cpp:362-428): kept with the grid they were read for
using Metric = ReaderLayout;

double scalar(const H5File &f, const std::string &path) {
  std::vector<double> v = double_dataset(f, path);
  if (v.size() != 1) throw Error("Unexpected HDF5 floating-point array size.");
  return v[0];
}

// GetSKSCoordinates (simulation_geometry.cpp:416-434), theta only
double fmks_theta(const Metric &metric, double x1, double x2) {
  double y = 2.0 * x2 - 1.0;
  double theta_g = kPi * x2 + (1.0 - metric.h) / 2.0 * std::sin(2.0 * kPi * x2);
  double theta_j = 0.5 * kPi + metric.poly_norm * y * (1.0 + std::pow(y / metric.poly_xt, metric.poly_alpha) / (metric.poly_alpha + 1.0));
  return theta_g + std::exp(metric.mks_smooth * (std::log(metric.r_in) - x1)) * (theta_j - theta_g);
}

// Jacobian of (r, theta) with respect to the native (x1, x2): dr/dx1 = r, and the two theta derivatives -- for MKS the
// h-slope map only, for FMKS the blend theta_g + s (theta_j - theta_g), s = exp(smooth (ln r_in - x1)), of the h-slope
// map theta_g and the polynomial map theta_j (values as simulation_geometry.cpp:440-471 evaluates them).
struct Jacobian {
  double dr_dx1, dth_dx1, dth_dx2;
};

Jacobian jacobian(const Metric &metric, double x1, double x2) {
  Jacobian J;
  J.dr_dx1 = std::exp(x1);
  const double slope_cos = (1.0 - metric.h) * kPi * std::cos(2.0 * kPi * x2);   // d theta_g / d x2 - pi
  if (!metric.fmks) {
    J.dth_dx1 = 0.0;
    J.dth_dx2 = kPi + slope_cos;
    return J;
  }
  const double y = 2.0 * x2 - 1.0;
  const double blend = std::exp(metric.mks_smooth * (std::log(metric.r_in) - x1));
  const double y_pow = std::pow(y / metric.poly_xt, metric.poly_alpha);
  const double order = 1.0 + metric.poly_alpha;
  const double poly = metric.poly_norm * (1.0 + y_pow / order);          // (theta_j - pi/2) / y
  // theta_j - theta_g = pi (1/2 - x2) + poly y - (1 - h)/2 sin(2 pi x2)
  const double gap = kPi * (0.5 - x2) + poly * y + -0.5 * (1.0 - metric.h) * std::sin(2.0 * kPi * x2);
  J.dth_dx1 = -metric.mks_smooth * blend * gap;
  // d(theta_j - theta_g)/dx2 = -pi + 2 poly + 2 norm alpha y^alpha / (1 + alpha) - (1 - h) pi cos(2 pi x2)
  const double gap_dx2 = (-kPi + 2.0 * poly) + 2.0 * metric.poly_norm * metric.poly_alpha * y_pow / order + -slope_cos;
  J.dth_dx2 = (kPi + slope_cos) + blend * gap_dx2;
  return J;
}

// GenerateSKSMap (simulation_geometry.cpp:330-413): x2(r, theta) by bisection on a uniform (r, theta) lattice.  The
// lattice points are independent, so the rows are spread over the host threads (the reference runs them serially).
void generate_sks_map(const Metric &m, double r_in, double r_out, AthenaGrid &g) {
  g.sks_map_n1 = kMapN1;
  g.sks_map_n2 = kMapN2;
  g.sks_map.assign((size_t)2 * kMapN2 * kMapN1, 0.0);
  double dr = (r_out - r_in) / (kMapN1 - 1);
  double dtheta = kPi / (kMapN2 - 1);
  g.sks_map_r_in = r_in;
  g.sks_map_dr = dr;
  g.sks_map_dtheta = dtheta;
  double *map_x1 = g.sks_map.data(), *map_x2 = g.sks_map.data() + (size_t)kMapN2 * kMapN1;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < kMapN1; ++i) {
    double r = r_in + i * dr;
    double x1 = std::log(r);
    for (int j = 0; j < kMapN2; ++j) {
      double theta = std::min(j * dtheta, kPi);
      double x2 = 0.5;
      if (theta > kMapTol && std::abs(kPi - theta) > kMapTol) {
        double x2_a = 0.0, x2_b = 1.0;
        x2 = (x2_b + x2_a) / 2.0;
        double theta_b = fmks_theta(m, x1, x2_b), theta_c = kPi / 2.0;
        for (int n = 0; n < kMapMaxIter; n++) {
          theta_c = fmks_theta(m, x1, x2);
          if ((theta_c - theta) * (theta_b - theta) < 0.0) {
            x2_a = x2;
          } else {
            theta_b = theta_c;
            x2_b = x2;
          }
          x2 = (x2_a + x2_b) / 2.0;
          if (std::abs(theta - theta_c) < kMapTol) break;
        }
      } else if (theta < kMapTol) {
        x2 = 0.0;
      } else if (theta > kPi - kMapTol) {
        x2 = 1.0;
      }
      map_x1[(size_t)j * kMapN1 + i] = x1;
      map_x2[(size_t)j * kMapN1 + i] = x2;
    }
  }
}

void uniform_axis(int n, double start, double dx, std::vector<double> &f, std::vector<double> &v) {
  f.assign((size_t)n + 1, 0.0);
  v.assign((size_t)n, 0.0);
  f[0] = start;
  for (int i = 0; i < n; i++) {
    f[(size_t)i + 1] = start + (i + 1) * dx;
    v[(size_t)i] = 0.5 * (f[(size_t)i] + f[(size_t)i + 1]);
  }
}

void snap(std::vector<double> &f, double upper, const char *name, const char *range) {
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
}

// header/gam, gam_p, gam_e against the input file (VerifyVariablesHarm, simulation_reader.cpp:1366-1422)
void settle_gamma(const H5File &f, const char *path, bool is_set, double *value, const char *what, const char *missing, bool warn) {
  if (f.has_dataset(path)) {
    double file = scalar(f, path);
    if (!is_set) {
      *value = file;
    } else if (*value != file && warn) {
      std::ostringstream msg;
      msg << "Given " << what << " adiabatic index of " << *value << " does not match file value of " << file
          << "; ignoring the latter.";
      warning(msg.str());
    }
  } else if (!is_set) {
    throw Error(missing);
  }
}

void settle_gammas(const H5File &f, Iharm3dExpect &e, bool warn) {
  settle_gamma(f, "header/gam", e.gamma_set, &e.plasma_gamma, "total", "Could not find total adiabatic index in input or data file.", warn);
  if (e.need_gamma_ie) {
    settle_gamma(f, "header/gam_p", e.gamma_i_set, &e.plasma_gamma_i, "ion", "Could not find ion adiabatic index in input or data file.", warn);
    settle_gamma(f, "header/gam_e", e.gamma_e_set, &e.plasma_gamma_e, "electron", "Could not find electron adiabatic index in input or data file.", warn);
  }
}

int locate(const std::vector<std::string> &names, const std::string &want, const std::string &message) {
  for (size_t n = 0; n < names.size(); n++)
    if (names[n] == want) return (int)n;
  throw Error(message);
}

}  // namespace

double read_iharm3d_time(const std::string &path) {
  H5File f(path);
  return scalar(f, "t");
}

void read_iharm3d_gammas(const std::string &path, Iharm3dExpect &expect) {
  H5File f(path);
  settle_gammas(f, expect, false);
}

void read_iharm3d(const std::string &path, const std::string &kappa_name, bool reuse_layout, Iharm3dExpect &expect,
                  AthenaGrid &g) {
  H5File f(path);
  g.time = scalar(f, "t");
  Metric &metric = g.layout;
  std::vector<double> &x2v_mod = g.layout.x2v_mod;   // x2 centres in modified coordinates (for the Jacobian)
  if (reuse_layout && (g.n_b != 1 || g.n_var <= 0 || (!metric.fmks && (int)x2v_mod.size() != g.n_j)))
    throw Error("iharm3d series: no first snapshot to take the layout from.");
  if (!reuse_layout) {
    // metric (simulation_reader.cpp:362-432)
    std::vector<std::string> name = string_dataset(f, "header/metric");
    if (name.empty()) throw Error("Unexpected HDF5 string array size.");
    std::string lower = name[0];
    for (char &c : lower) c = (char)std::tolower((unsigned char)c);
    if (name[0] != "MKS" && name[0] != "MMKS" && name[0] != "FMKS")
      warning("Given metric mks does not match file value of " + name[0] + "; ignoring the latter.");
    metric = Metric();
    metric.fmks = expect.fmks;
    metric.a = scalar(f, "header/geom/" + lower + "/a");
    metric.h = scalar(f, "header/geom/" + lower + "/hslope");
    if (metric.a != expect.simulation_a) {
      std::ostringstream msg;
      msg << "Given spin of " << expect.simulation_a << " does not match file value of " << metric.a << "; ignoring the latter.";
      warning(msg.str());
    }
    if (name[0] == "MMKS" || name[0] == "FMKS") {
      if (f.has_dataset("header/geom/" + lower + "/r_in")) metric.r_in = scalar(f, "header/geom/" + lower + "/r_in");
      else if (f.has_dataset("header/geom/" + lower + "/Rin")) metric.r_in = scalar(f, "header/geom/" + lower + "/Rin");
      else throw Error("Unable to identify r_in parameter for iharm3d-format file.");
      metric.poly_xt = scalar(f, "header/geom/" + lower + "/poly_xt");
      metric.poly_alpha = scalar(f, "header/geom/" + lower + "/poly_alpha");
      metric.mks_smooth = scalar(f, "header/geom/" + lower + "/mks_smooth");
      metric.poly_norm = (metric.poly_alpha + 1.0) * std::pow(metric.poly_xt, metric.poly_alpha);
      metric.poly_norm = 0.5 * kPi * metric.poly_norm / (metric.poly_norm + 1.0);
    }
    // one block (:598-607); uniform native coordinates (:622-660)
    g.n_b = 1;
    g.levels.assign(1, 0);
    g.locations.assign(3, 0);
    auto axis = [&](const char *n, const char *start, const char *dx, int &count, std::vector<double> &faces, std::vector<double> &centres) {
      std::vector<int32_t> cells = int_dataset(f, std::string("header/") + n);
      if (cells.size() != 1 || cells[0] <= 0) throw Error("Unexpected HDF5 integer array size.");
      count = cells[0];
      uniform_axis(count, scalar(f, std::string("header/geom/") + start), scalar(f, std::string("header/geom/") + dx), faces, centres);
    };
    axis("n1", "startx1", "dx1", g.n_i, g.x1f, g.x1v);
    axis("n2", "startx2", "dx2", g.n_j, g.x2f, g.x2v);
    axis("n3", "startx3", "dx3", g.n_k, g.x3f, g.x3v);
    g.n_3_root = g.n_k;
    g.sks_map.clear();
    // ConvertCoordinates (simulation_geometry.cpp:29-90)
    if (expect.fmks) {
      generate_sks_map(metric, std::exp(g.x1f[0]), std::exp(g.x1f[(size_t)g.n_i]), g);
      g.simulation_bounds[0] = std::exp(g.x1f[0]);
      g.simulation_bounds[2] = fmks_theta(metric, g.x1f[0], 0.0);
      g.simulation_bounds[4] = 0.0;
      g.simulation_bounds[1] = std::exp(g.x1f[(size_t)g.n_i]);
      g.simulation_bounds[3] = fmks_theta(metric, g.x1f[(size_t)g.n_i], 1.0);
      g.simulation_bounds[5] = 2.0 * kPi;
    } else {
      x2v_mod = g.x2v;
      for (double &x : g.x1f) x = std::exp(x);
      for (double &x : g.x1v) x = std::exp(x);
      for (double &x : g.x2f) x = kPi * x + (1.0 - metric.h) / 2.0 * std::sin(2.0 * kPi * x);
      for (double &x : g.x2v) x = kPi * x + (1.0 - metric.h) / 2.0 * std::sin(2.0 * kPi * x);
      snap(g.x2f, kPi, "theta", "[0, pi]");
    }
    snap(g.x3f, 2.0 * kPi, "phi", "[0, 2*pi]");
    // variables by name (VerifyVariablesHarm, :1303-1364)
    std::vector<int32_t> n_prim = int_dataset(f, "header/n_prim");
    std::vector<std::string> names = string_dataset(f, "header/prim_names");
    if (n_prim.size() != 1 || n_prim[0] != (int)names.size()) throw Error("Inconsistency in number of primitive variables.");
    g.n_var = n_prim[0];
    g.ind_rho = locate(names, "RHO", "Unable to locate \"RHO\" slice of \"prims\" in data file.");
    g.ind_pgas = locate(names, "UU", "Unable to locate \"UU\" slice of \"prims\" in data file.");
    g.ind_kappa = kappa_name.empty() ? -1 : locate(names, kappa_name, "Unable to locate electron entropy slice of \"prims\" in data file.");
    g.ind_uu1 = locate(names, "U1", "Unable to locate \"U1\" slice of \"prims\" in data file.");
    g.ind_uu2 = locate(names, "U2", "Unable to locate \"U2\" slice of \"prims\" in data file.");
    g.ind_uu3 = locate(names, "U3", "Unable to locate \"U3\" slice of \"prims\" in data file.");
    g.ind_bb1 = locate(names, "B1", "Unable to locate \"B1\" slice of \"prims\" in data file.");
    g.ind_bb2 = locate(names, "B2", "Unable to locate \"B2\" slice of \"prims\" in data file.");
    g.ind_bb3 = locate(names, "B3", "Unable to locate \"B3\" slice of \"prims\" in data file.");
    settle_gammas(f, expect, true);
    expect.gamma_set = true;
    if (expect.need_gamma_ie) expect.gamma_i_set = expect.gamma_e_set = true;
  }
  const int n1 = g.n_i, n2 = g.n_j, n3 = g.n_k, nv = g.n_var;
  const size_t cells = (size_t)n1 * n2 * n3;
  Datatype dt;
  std::vector<uint64_t> dims;
  const uint8_t *d = f.dataset("prims", dt, dims);
  if (dt.cls != 1 || dt.size != 4 || dims.size() != 4 || (int)dims[0] != n1 || (int)dims[1] != n2 || (int)dims[2] != n3 || (int)dims[3] != nv)
    throw Error("Array dimension mismatch.");
  const float *file = reinterpret_cast<const float *>(d);
  g.prim.assign((size_t)nv * cells, 0.0f);
  auto at = [&](int v, int k, int j, int i) -> float & { return g.prim[(((size_t)v * n3 + k) * n2 + j) * n1 + i]; };
  // file: (x1, x2, x3, variable) with the variable fastest; ours: (var, k, j, i)  (:797-802)
#pragma omp parallel for schedule(static) collapse(2)
  for (int v = 0; v < nv; v++)
    for (int k = 0; k < n3; k++)
      for (int j = 0; j < n2; j++)
        for (int i = 0; i < n1; i++) {
          float value;
          std::memcpy(&value, file + (((size_t)i * n2 + j) * n3 + k) * nv + v, sizeof(float));
          at(v, k, j, i) = value;
        }
  const float gm1 = static_cast<float>(expect.plasma_gamma - 1.0);
  for (size_t c = 0; c < cells; c++) g.prim[(size_t)g.ind_pgas * cells + c] *= gm1;

  // Vector primitives (ConvertPrimitives3, simulation_geometry.cpp:95-236): the dump holds the normal-frame velocity
  // u~^i and the lab-frame field B^i on the native coordinate basis.  Per cell: native metric from the spherical
  // Kerr-Schild one through the Jacobian, u~ -> four-velocity, B -> magnetic four-vector, both pushed to the
  // (t, r, theta, phi) basis, then back to normal-frame velocity and B^i = b^i u^t - b^t u^i there.
  const double a = expect.simulation_a;
  const Metric &m = metric;
  const std::vector<double> &x2_mod = x2v_mod;
  const int vel[3] = {g.ind_uu1, g.ind_uu2, g.ind_uu3}, mag[3] = {g.ind_bb1, g.ind_bb2, g.ind_bb3};
#pragma omp parallel for schedule(static) collapse(2)
  for (int k = 0; k < n3; k++)
    for (int j = 0; j < n2; j++)
      for (int i = 0; i < n1; i++) {
        // the cell's native and spherical coordinates
        double x1, x2, r, th;
        if (m.fmks) {
          x1 = g.x1v[(size_t)i];
          x2 = g.x2v[(size_t)j];
          r = std::exp(x1);
          th = fmks_theta(m, x1, x2);
        } else {
          r = g.x1v[(size_t)i];
          th = g.x2v[(size_t)j];
          x1 = std::log(r);
          x2 = x2_mod[(size_t)j];
        }
        const Jacobian J = jacobian(m, x1, x2);
        const double s2 = std::sin(th) * std::sin(th), c2 = std::cos(th) * std::cos(th);
        // spherical Kerr-Schild metric, spatial part and time row (indices r, theta, phi), and g^{tt}, g^{tr}
        const double sigma = r * r + a * a * c2, f = 2.0 * r / sigma;
        const double time_row[3] = {f, 0.0, -a * f * s2};
        const double space[3][3] = {{1.0 + f, 0.0, -a * (1.0 + f) * s2},
                                    {0.0, sigma, 0.0},
                                    {-a * (1.0 + f) * s2, 0.0, (r * r + a * a + a * a * f * s2) * s2}};
        const double con_tt = -(1.0 + f), con_tr = f;
        // native basis vectors in spherical components: e_1 = (dr/dx1, dth/dx1, 0), e_2 = (0, dth/dx2, 0), e_3 = phi
        const double e[3][3] = {{J.dr_dx1, J.dth_dx1, 0.0}, {0.0, J.dth_dx2, 0.0}, {0.0, 0.0, 1.0}};
        double nat_time[3], nat[3][3];
        for (int p = 0; p < 3; p++) {
          nat_time[p] = 0.0;
          for (int q = 0; q < 3; q++) nat_time[p] += e[p][q] * time_row[q];
          for (int q = 0; q < 3; q++) {
            nat[p][q] = 0.0;
            for (int u = 0; u < 3; u++)
              for (int v = 0; v < 3; v++) nat[p][q] += e[p][u] * e[q][v] * space[u][v];
          }
        }
        // native g^{0i} / (-g^{00}) is the shift; g^{00} is a scalar under the spatial change of coordinates
        const double lapse = 1.0 / std::sqrt(-con_tt);
        const double con_0[3] = {con_tr / J.dr_dx1, -J.dth_dx1 * f / (J.dr_dx1 * J.dth_dx2), 0.0};
        double uu[3], bb[3];
        for (int p = 0; p < 3; p++) {
          uu[p] = at(vel[p], k, j, i);
          bb[p] = at(mag[p], k, j, i);
        }
        double norm = 1.0;
        for (int p = 0; p < 3; p++)
          for (int q = 0; q < 3; q++) norm += nat[p][q] * uu[p] * uu[q];
        const double gamma = std::sqrt(norm);
        double u[4] = {gamma / lapse, 0.0, 0.0, 0.0}, u_low[3], b[4] = {0.0, 0.0, 0.0, 0.0};
        for (int p = 0; p < 3; p++) u[p + 1] = uu[p] - lapse * con_0[p] * gamma;
        for (int p = 0; p < 3; p++) {
          u_low[p] = nat_time[p] * u[0];
          for (int q = 0; q < 3; q++) u_low[p] += nat[p][q] * u[q + 1];
          b[0] += u_low[p] * bb[p];
        }
        for (int p = 0; p < 3; p++) b[p + 1] = (bb[p] + b[0] * u[p + 1]) / u[0];
        // to the (r, theta, phi) basis: v^r = dr/dx1 v^1, v^theta = dth/dx1 v^1 + dth/dx2 v^2, v^phi = v^3
        const double us[3] = {J.dr_dx1 * u[1], J.dth_dx1 * u[1] + J.dth_dx2 * u[2], u[3]};
        const double bs[3] = {J.dr_dx1 * b[1], J.dth_dx1 * b[1] + J.dth_dx2 * b[2], b[3]};
        const double shift[3] = {lapse * lapse * con_tr, 0.0, 0.0};   // alpha^2 g^{ti}; g^{t theta} = g^{t phi} = 0
        for (int p = 0; p < 3; p++) {
          at(vel[p], k, j, i) = static_cast<float>(us[p] + shift[p] * u[0]);
          at(mag[p], k, j, i) = static_cast<float>(bs[p] * u[0] - b[0] * us[p]);
        }
      }
}

}  // namespace blh
