#include "camera.hpp"

#include <algorithm>
#include <cmath>

namespace blh {

namespace {

// Cartesian Kerr-Schild metric at (x,y,z), M = 1 (reference geodesic_geometry.cpp:38-161)
void ks_metric(double a, bool flat, double x, double y, double z, double gcov[4][4], double gcon[4][4]) {
  if (flat) {
    for (int m = 0; m < 4; m++)
      for (int n = 0; n < 4; n++) {
        double eta = m == n ? (m == 0 ? -1.0 : 1.0) : 0.0;
        if (gcov) gcov[m][n] = eta;
        if (gcon) gcon[m][n] = eta;
      }
    return;
  }
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double r2 = 0.5 * (rr2 - a2 + std::hypot(rr2 - a2, 2.0 * a * z));
  double r = std::sqrt(r2);
  double f = 2.0 * r2 * r / (r2 * r2 + a2 * z * z);
  double lo[4] = {1.0, (r * x + a * y) / (r2 + a2), (r * y - a * x) / (r2 + a2), z / r};
  double up[4] = {-1.0, lo[1], lo[2], lo[3]};
  for (int m = 0; m < 4; m++)
    for (int n = 0; n < 4; n++) {
      if (gcov) {
        double v = f * lo[m] * lo[n];
        gcov[m][n] = m == n ? (m == 0 ? v - 1.0 : v + 1.0) : v;
      }
      if (gcon) {
        double v = -f * up[m] * up[n];
        gcon[m][n] = m == n ? (m == 0 ? v - 1.0 : v + 1.0) : v;
      }
    }
}

}  // namespace

std::vector<double> image_frequencies(int num, double single, double start, double end, int spacing) {
  std::vector<double> f((size_t)num);
  if (num == 1) {
    f[0] = single;
    return f;
  }
  f[0] = start;
  f[(size_t)num - 1] = end;
  for (int l = 1; l < num - 1; l++) {
    double frac = static_cast<double>(l) / static_cast<double>(num - 1);
    if (spacing == 0)
      f[(size_t)l] = start + frac * (end - start);
    else if (spacing == 1)
      f[(size_t)l] = 1.0 / (1.0 / start + frac * (1.0 / end - 1.0 / start));
    else
      f[(size_t)l] = std::exp(std::log(start) + frac * std::log(end / start));
  }
  return f;
}

CameraFrame build_camera_frame(const CameraSetup &s) {
  CameraFrame F{};
  const double a = s.a, rc = s.r;
  double sth = std::sin(s.th), cth = std::cos(s.th);
  double sph = std::sin(s.ph), cph = std::cos(s.ph);
  double srot = std::sin(s.rotation), crot = std::cos(s.rotation);

  // position
  F.x[0] = 0.0;
  F.x[1] = sth * (rc * cph - a * sph);
  F.x[2] = sth * (rc * sph + a * cph);
  F.x[3] = rc * cth;
  if (s.flat) {
    F.x[1] = rc * sth * cph;
    F.x[2] = rc * sth * sph;
  }
  double z_sign = F.x[3] >= 0.0 ? 1.0 : -1.0;

  // spherical Kerr-Schild metric at the camera; symmetric 3x3 blocks stored as (rr, rth, rph, thth, thph, phph)
  double a2 = a * a, r2 = rc * rc;
  double delta = r2 - 2.0 * rc + a2;
  double sigma = r2 + a2 * cth * cth;
  double gl[6] = {1.0 + 2.0 * rc / sigma, 0.0, -(1.0 + 2.0 * rc / sigma) * a * sth * sth, sigma, 0.0,
                  (r2 + a2 + 2.0 * a2 * rc / sigma * sth * sth) * sth * sth};
  double gtt = -(1.0 + 2.0 * rc / sigma);
  double gt[3] = {2.0 * rc / sigma, 0.0, 0.0};  // g^{t r}, g^{t th}, g^{t ph}
  double gu[6] = {delta / sigma, 0.0, a / sigma, 1.0 / sigma, 0.0, 1.0 / (sigma * sth * sth)};
  if (s.flat && !s.pole) {
    double flat_l[6] = {1.0, 0.0, 0.0, r2, 0.0, r2 * sth * sth};
    double flat_u[6] = {1.0, 0.0, 0.0, 1.0 / r2, 0.0, 1.0 / (r2 * sth * sth)};
    std::copy(flat_l, flat_l + 6, gl);
    std::copy(flat_u, flat_u + 6, gu);
    gtt = -1.0;
    gt[0] = gt[1] = gt[2] = 0.0;
  }
  if (s.pole && !s.flat) {
    double f = 2.0 * rc / (r2 + a2);
    double pole_l[6] = {1.0 + f, 0.0, 0.0, 1.0, 0.0, 1.0};
    double pole_u[6] = {1.0 - f, 0.0, 0.0, 1.0, 0.0, 1.0};
    std::copy(pole_l, pole_l + 6, gl);
    std::copy(pole_u, pole_u + 6, gu);
    gtt = -1.0 - f;
    gt[0] = z_sign * f;
    gt[1] = gt[2] = 0.0;
  }
  if (s.flat && s.pole) {
    double id[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
    std::copy(id, id + 6, gl);
    std::copy(id, id + 6, gu);
    gtt = -1.0;
    gt[0] = gt[1] = gt[2] = 0.0;
  }

  // camera velocity in spherical coordinates from the normal-observer components
  double alpha = 1.0 / std::sqrt(-gtt);
  double beta[3] = {-gt[0] / gtt, -gt[1] / gtt, -gt[2] / gtt};
  double utn = std::sqrt(1.0 + gl[0] * s.urn * s.urn + 2.0 * gl[1] * s.urn * s.uthn + 2.0 * gl[2] * s.urn * s.uphn +
                         gl[3] * s.uthn * s.uthn + 2.0 * gl[4] * s.uthn * s.uphn + gl[5] * s.uphn * s.uphn);
  F.u_con[0] = utn / alpha;
  double ur = s.urn - beta[0] / alpha * utn;
  double uth = s.uthn - beta[1] / alpha * utn;
  double uph = s.uphn - beta[2] / alpha * utn;

  // d(x,y,z)/d(r,th,ph)
  double jx[3] = {sth * cph, cth * (rc * cph - a * sph), sth * (-rc * sph - a * cph)};
  double jy[3] = {sth * sph, cth * (rc * sph + a * cph), sth * (rc * cph - a * sph)};
  double jz[3] = {cth, -rc * sth, 0.0};
  if (s.flat && !s.pole) {
    jx[0] = sth * cph; jx[1] = rc * cth * cph; jx[2] = -rc * sth * sph;
    jy[0] = sth * sph; jy[1] = rc * cth * sph; jy[2] = rc * sth * cph;
    jz[0] = cth; jz[1] = -rc * sth; jz[2] = 0.0;
  }
  if (s.pole) {
    jx[0] = 0.0; jx[1] = 1.0; jx[2] = 0.0;
    jy[0] = 0.0; jy[1] = 0.0; jy[2] = 1.0;
    jz[0] = z_sign; jz[1] = 0.0; jz[2] = 0.0;
  }
  F.u_con[1] = jx[0] * ur + jx[1] * uth + jx[2] * uph;
  F.u_con[2] = jy[0] * ur + jy[1] * uth + jy[2] * uph;
  F.u_con[3] = jz[0] * ur + jz[1] * uth + jz[2] * uph;
  double g_cov[4][4], g_con[4][4];
  ks_metric(a, s.flat, F.x[1], F.x[2], F.x[3], g_cov, g_con);
  for (int m = 0; m < 4; m++) {
    F.u_cov[m] = 0.0;
    for (int n = 0; n < 4; n++) F.u_cov[m] += g_cov[m][n] * F.u_con[n];
  }

  // photon momentum in spherical coordinates: spatial metric of the normal observer, null condition
  double gn[6];
  gn[0] = (gtt * gu[0] - gt[0] * gt[0]) / gtt;
  gn[1] = (gtt * gu[1] - gt[0] * gt[1]) / gtt;
  gn[2] = (gtt * gu[2] - gt[0] * gt[2]) / gtt;
  gn[3] = (gtt * gu[3] - gt[1] * gt[1]) / gtt;
  gn[4] = (gtt * gu[4] - gt[1] * gt[2]) / gtt;
  gn[5] = (gtt * gu[5] - gt[2] * gt[2]) / gtt;
  double k_rn = s.k_r, k_thn = s.k_th, k_phn = s.k_ph;
  double k_tn = -std::sqrt(gn[0] * k_rn * k_rn + 2.0 * gn[1] * k_rn * k_thn + 2.0 * gn[2] * k_rn * k_phn +
                           gn[3] * k_thn * k_thn + 2.0 * gn[4] * k_thn * k_phn + gn[5] * k_phn * k_phn);
  double k_t = alpha * k_tn + (beta[0] * k_rn + beta[1] * k_thn + beta[2] * k_phn);

  // d(r,th,ph)/d(x,y,z)
  double rr2 = F.x[1] * F.x[1] + F.x[2] * F.x[2] + F.x[3] * F.x[3];
  double den = 2.0 * r2 - rr2 + a2;
  double dr[3] = {rc * F.x[1] / den, rc * F.x[2] / den, (rc * F.x[3] + a2 * F.x[3] / rc) / den};
  double dth[3] = {F.x[3] * dr[0] / (r2 * sth), F.x[3] * dr[1] / (r2 * sth), (F.x[3] * dr[2] - rc) / (r2 * sth)};
  double rho2 = F.x[1] * F.x[1] + F.x[2] * F.x[2];
  double dph[3] = {-F.x[2] / rho2 + a / (r2 + a2) * dr[0], F.x[1] / rho2 + a / (r2 + a2) * dr[1],
                   a / (r2 + a2) * dr[2]};
  if (s.flat && !s.pole) {
    dr[0] = F.x[1] / rc; dr[1] = F.x[2] / rc; dr[2] = F.x[3] / rc;
    dth[0] = cth * cph / rc; dth[1] = cth * sph / rc; dth[2] = -sth / rc;
    dph[0] = -sph / (rc * sth); dph[1] = cph / (rc * sth); dph[2] = 0.0;
  }
  if (s.pole) {
    dr[0] = 0.0; dr[1] = 0.0; dr[2] = z_sign;
    dth[0] = 1.0; dth[1] = 0.0; dth[2] = 0.0;
    dph[0] = 0.0; dph[1] = 1.0; dph[2] = 0.0;
  }
  double k_x = dr[0] * s.k_r + dth[0] * s.k_th + dph[0] * s.k_ph;
  double k_y = dr[1] * s.k_r + dth[1] * s.k_th + dph[1] * s.k_ph;
  double k_z = dr[2] * s.k_r + dth[2] * s.k_th + dph[2] * s.k_ph;
  double k_tc = F.u_con[0] * k_t + F.u_con[1] * k_x + F.u_con[2] * k_y + F.u_con[3] * k_z;

  // camera-frame spatial metric, contravariant (upper triangle xx,xy,xz,yy,yz,zz)
  double hc[6] = {g_con[1][1] + F.u_con[1] * F.u_con[1], g_con[1][2] + F.u_con[1] * F.u_con[2],
                  g_con[1][3] + F.u_con[1] * F.u_con[3], g_con[2][2] + F.u_con[2] * F.u_con[2],
                  g_con[2][3] + F.u_con[2] * F.u_con[3], g_con[3][3] + F.u_con[3] * F.u_con[3]};

  // unit normal
  double nx = k_x - F.u_cov[1] / F.u_cov[0] * k_t;
  double ny = k_y - F.u_cov[2] / F.u_cov[0] * k_t;
  double nz = k_z - F.u_cov[3] / F.u_cov[0] * k_t;
  F.norm_con_c[0] = -k_tc;
  F.norm_con_c[1] = hc[0] * nx + hc[1] * ny + hc[2] * nz;
  F.norm_con_c[2] = hc[1] * nx + hc[3] * ny + hc[4] * nz;
  F.norm_con_c[3] = hc[2] * nx + hc[4] * ny + hc[5] * nz;
  double norm_norm = std::sqrt(nx * F.norm_con_c[1] + ny * F.norm_con_c[2] + nz * F.norm_con_c[3]);
  nx /= norm_norm;
  ny /= norm_norm;
  nz /= norm_norm;
  for (int m = 0; m < 4; m++) F.norm_con_c[m] /= norm_norm;
  F.norm_con[0] = F.u_con[0] * F.norm_con_c[0] -
                  (F.u_cov[1] * F.norm_con_c[1] + F.u_cov[2] * F.norm_con_c[2] + F.u_cov[3] * F.norm_con_c[3]) / F.u_cov[0];
  F.norm_con[1] = F.norm_con_c[1] + F.u_con[1] * F.norm_con_c[0];
  F.norm_con[2] = F.norm_con_c[2] + F.u_con[2] * F.norm_con_c[0];
  F.norm_con[3] = F.norm_con_c[3] + F.u_con[3] * F.norm_con_c[0];

  // "up" before projection
  double up[3] = {0.0, 0.0, 1.0};
  if (s.pole) {
    up[1] = 1.0;
    up[2] = 0.0;
  }

  // camera-frame spatial metric, covariant (xx,xy,xz,yy,yz,zz)
  auto hcov = [&](int i, int j) {
    return g_cov[i][j] - F.u_cov[i] / F.u_cov[0] * g_cov[j][0] - F.u_cov[j] / F.u_cov[0] * g_cov[i][0] +
           F.u_cov[i] * F.u_cov[j] / (F.u_cov[0] * F.u_cov[0]) * g_cov[0][0];
  };
  double hl[6] = {hcov(1, 1), hcov(1, 2), hcov(1, 3), hcov(2, 2), hcov(2, 3), hcov(3, 3)};

  // vertical direction: project "up" off the normal, normalise
  double up_norm = up[0] * nx + up[1] * ny + up[2] * nz;
  F.vert_con_c[0] = 0.0;
  F.vert_con_c[1] = up[0] - up_norm * F.norm_con_c[1];
  F.vert_con_c[2] = up[1] - up_norm * F.norm_con_c[2];
  F.vert_con_c[3] = up[2] - up_norm * F.norm_con_c[3];
  double vx = hl[0] * F.vert_con_c[1] + hl[1] * F.vert_con_c[2] + hl[2] * F.vert_con_c[3];
  double vy = hl[1] * F.vert_con_c[1] + hl[3] * F.vert_con_c[2] + hl[4] * F.vert_con_c[3];
  double vz = hl[2] * F.vert_con_c[1] + hl[4] * F.vert_con_c[2] + hl[5] * F.vert_con_c[3];
  double vert_norm = std::sqrt(vx * F.vert_con_c[1] + vy * F.vert_con_c[2] + vz * F.vert_con_c[3]);
  vx /= vert_norm;
  vy /= vert_norm;
  vz /= vert_norm;
  F.vert_con_c[1] /= vert_norm;
  F.vert_con_c[2] /= vert_norm;
  F.vert_con_c[3] /= vert_norm;

  // horizontal direction: cross product with the metric determinant
  double det = hl[0] * (hl[3] * hl[5] - hl[4] * hl[4]) + hl[1] * (hl[4] * hl[2] - hl[1] * hl[5]) +
               hl[2] * (hl[1] * hl[4] - hl[3] * hl[2]);
  double det_sqrt = std::sqrt(det);
  F.hor_con_c[0] = 0.0;
  F.hor_con_c[1] = (vy * nz - vz * ny) / det_sqrt;
  F.hor_con_c[2] = (vz * nx - vx * nz) / det_sqrt;
  F.hor_con_c[3] = (vx * ny - vy * nx) / det_sqrt;

  // rotate about the normal
  double h0[3] = {F.hor_con_c[1], F.hor_con_c[2], F.hor_con_c[3]};
  double v0[3] = {F.vert_con_c[1], F.vert_con_c[2], F.vert_con_c[3]};
  for (int i = 0; i < 3; i++) {
    F.hor_con_c[1 + i] = h0[i] * crot - v0[i] * srot;
    F.vert_con_c[1 + i] = v0[i] * crot + h0[i] * srot;
  }
  return F;
}

void camera_pixel(const CameraSetup &s, const CameraFrame &f, double u_ind, double v_ind, double pos[4],
                  double dir[4], double *factor) {
  double u = u_ind * 1.0 * s.width;
  double v = v_ind * 1.0 * s.width;
  double p[4];
  if (s.type == 0) {
    double dc[4];
    for (int m = 0; m < 4; m++) dc[m] = u * f.hor_con_c[m] + v * f.vert_con_c[m];
    double dt = f.u_con[0] * dc[0] - (f.u_cov[1] * dc[1] + f.u_cov[2] * dc[2] + f.u_cov[3] * dc[3]) / f.u_cov[0];
    pos[0] = f.x[0] + dt;
    for (int i = 1; i < 4; i++) pos[i] = f.x[i] + (dc[i] + f.u_con[i] * dc[0]);
    p[1] = f.norm_con[1];
    p[2] = f.norm_con[2];
    p[3] = f.norm_con[3];
  } else {
    for (int m = 0; m < 4; m++) pos[m] = f.x[m];
    double normalization = std::hypot(u, v, s.r);
    double frac_norm = s.r / normalization;
    double frac_hor = -u / normalization;
    double frac_vert = -v / normalization;
    for (int i = 1; i < 4; i++) {
      double dc = frac_norm * f.norm_con_c[i] + frac_hor * f.hor_con_c[i] + frac_vert * f.vert_con_c[i];
      p[i] = dc + f.u_con[i] * f.norm_con_c[0];
    }
  }
  // p^t from the null condition g_{mu nu} p^mu p^nu = 0 (camera.cpp:553-566)
  double gcov[4][4];
  ks_metric(s.a, s.flat, pos[1], pos[2], pos[3], gcov, nullptr);
  double qa = gcov[0][0];
  double qb = 0.0;
  for (int i = 1; i < 4; i++) qb += 2.0 * gcov[0][i] * p[i];
  double qc = 0.0;
  for (int i = 1; i < 4; i++)
    for (int j = 1; j < 4; j++) qc += gcov[i][j] * p[i] * p[j];
  double qd = std::sqrt(std::max(qb * qb - 4.0 * qa * qc, 0.0));
  p[0] = qa == 0.0 ? -qc / (2.0 * qb) : (qb < 0.0 ? 2.0 * qc / (qd - qb) : -(qb + qd) / (2.0 * qa));
  for (int m = 0; m < 4; m++) {
    dir[m] = 0.0;
    for (int n = 0; n < 4; n++) dir[m] += gcov[m][n] * p[n];
  }
  double nu_local = 0.0;
  if (s.normalization == 0)
    for (int m = 0; m < 4; m++) nu_local -= dir[m] * f.u_con[m];
  else
    nu_local = -dir[0];
  *factor = 1.0 / nu_local;
}

void camera_root(const CameraSetup &s, const CameraFrame &f, std::vector<double> &pos, std::vector<double> &dir,
                 std::vector<double> &factor) {
  const int res = s.resolution;
  const size_t n = (size_t)res * res;
  pos.resize(4 * n);
  dir.resize(4 * n);
  factor.resize(n);
#pragma omp parallel for schedule(static)
  for (long m = 0; m < (long)n; m++) {
    int row = (int)(m / res), col = (int)(m % res);
    double u_ind = (col - res / 2.0 + 0.5) / res;
    double v_ind = (row - res / 2.0 + 0.5) / res;
    camera_pixel(s, f, u_ind, v_ind, &pos[4 * m], &dir[4 * m], &factor[m]);
  }
}

void camera_rows(const CameraSetup &s, const CameraFrame &f, const long long *rows, long long num_rows, double *pos, double *dir,
                 double *factor) {
  const int res = s.resolution;
#pragma omp parallel for schedule(static)
  for (long long r = 0; r < num_rows; r++) {
    const double v_ind = ((int)rows[r] - res / 2.0 + 0.5) / res;
    for (int col = 0; col < res; col++) {
      const double u_ind = (col - res / 2.0 + 0.5) / res;
      const size_t o = (size_t)r * res + col;
      camera_pixel(s, f, u_ind, v_ind, &pos[4 * o], &dir[4 * o], &factor[o]);
    }
  }
}

void child_blocks(const std::vector<int32_t> &parent_locs, const std::vector<uint8_t> &flags, std::vector<int32_t> &child_locs) {
  size_t refined = 0;
  for (uint8_t fl : flags) refined += fl ? 1 : 0;
  child_locs.resize(refined * 4 * 2);
  size_t block = 0;
  for (size_t parent = 0; parent < flags.size(); parent++) {
    if (!flags[parent]) continue;
    int pv = parent_locs[2 * parent], pu = parent_locs[2 * parent + 1];
    for (int bv = 2 * pv; bv <= 2 * pv + 1; bv++)
      for (int bu = 2 * pu; bu <= 2 * pu + 1; bu++, block++) {
        child_locs[2 * block] = bv;
        child_locs[2 * block + 1] = bu;
      }
  }
}

void camera_blocks(const CameraSetup &s, const CameraFrame &f, int level, int block_size, const int32_t *locs, long long blocks,
                   double *pos, double *dir, double *factor) {
  int eff_res = s.resolution;
  for (int n = 1; n <= level; n++) eff_res *= 2;
  const size_t bpix = (size_t)block_size * block_size;
#pragma omp parallel for schedule(static)
  for (long long b = 0; b < blocks; b++) {
    int row_off = locs[2 * (size_t)b] * block_size, col_off = locs[2 * (size_t)b + 1] * block_size;
    for (size_t m = 0; m < bpix; m++) {
      int row = (int)(m / block_size), col = (int)(m % block_size);
      double u_ind = (col + col_off - eff_res / 2.0 + 0.5) / eff_res;
      double v_ind = (row + row_off - eff_res / 2.0 + 0.5) / eff_res;
      size_t o = (size_t)b * bpix + m;
      camera_pixel(s, f, u_ind, v_ind, &pos[4 * o], &dir[4 * o], &factor[o]);
    }
  }
}

void camera_refined(const CameraSetup &s, const CameraFrame &f, int level, int block_size,
                    const std::vector<int32_t> &parent_locs, const std::vector<uint8_t> &flags,
                    std::vector<int32_t> &child_locs, std::vector<double> &pos, std::vector<double> &dir,
                    std::vector<double> &factor) {
  int eff_res = s.resolution;
  for (int n = 1; n <= level; n++) eff_res *= 2;
  // child list: parents in index order x 4 children (camera.cpp:445-459); then every pixel is independent
  child_blocks(parent_locs, flags, child_locs);
  const size_t block = child_locs.size() / 2;
  const size_t bpix = (size_t)block_size * block_size;
  pos.resize(block * bpix * 4);
  dir.resize(block * bpix * 4);
  factor.resize(block * bpix);
  camera_blocks(s, f, level, block_size, child_locs.data(), (long long)block, pos.data(), dir.data(), factor.data());
}

}  // namespace blh
