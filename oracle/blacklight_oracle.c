/* blacklight_oracle.c -- plain-C CPU restatement of the reference's per-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under blacklight_b200/ links, loads or calls this file; it exists so
 * the parity tests have an independent, dependency-free statement of the algorithm (dense 4x4 tensor form,
 * as the reference writes it) next to the unmodified reference binary (oracle/_ref/blacklight).  It is
 * pinned against the golden fixtures produced by that binary (tests/test_cpu_oracle.py): sample counts,
 * flags and every stored sample bit for bit, images to rounding.
 *
 * Covered (reference file:line):
 *   Kerr-Schild metric, inverse, inverse derivative        geodesic_geometry.cpp:19-276
 *   Hamiltonian right-hand side with proper distance       geodesics.cpp:867-893
 *   Dormand-Prince RK5(4)7M integrator with dense output   geodesics.cpp:39-324
 *   truncation, momentum renormalisation, reversal         geodesics.cpp:327-371, 808-849
 *   fixed-step RK4 / RK2 integrators                       geodesics.cpp:418-805
 *   formula-model coefficients                             formula_coefficients.cpp:25-183
 *   grid sampling (block/cell search, nearest, trilinear)  simulation_sampling.cpp:122-575, 636-1044
 *   thermal synchrotron I coefficients                     simulation_coefficients.cpp:254-524
 *   power-law synchrotron I coefficients and constants     simulation_coefficients.cpp:53-66, 559-585
 *   kappa-distribution I coefficients and constants        simulation_coefficients.cpp:82-105, 608-653, 740-773
 *   electron temperature: ti_te_beta (p or energies), code_kappa   simulation_coefficients.cpp:333-358
 *   geometric and cell-value cuts, value fallback          simulation_sampling.cpp:245-295, 695-708; simulation_coefficients.cpp:361-375
 *   inter-block trilinear anchors across refinement levels simulation_sampling.cpp:506-553, 1068-1331, 1365-1386
 *   slow light: time slice per sample, nearest / blended   simulation_sampling.cpp:297-349, 736-775, 840-905
 *   Cartesian Kerr-Schild grids (simulation_coord = cks)   radiation_geometry.cpp:37-57, 425-457
 *   fluid-frame tetrad                                     radiation_geometry.cpp:597-658
 *   unpolarized transfer, auxiliary images (both models)   unpolarized.cpp:31-221
 *   false-colour rendering (fills, threshold crossings)    rendering.cpp:25-179
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference's -O3 without -march).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.141592653589793

typedef struct {
  double a;             /* spin, M = 1 */
  int flat;
  double camera_r, r_terminate, ray_step, tol_abs, tol_rel;
  int max_steps, max_retries;
} orc_geo;

/* ---------------------------------------------------------------- geometry (geodesic_geometry.cpp) */

static double ks_radius(double a, double x, double y, double z) {
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double r2 = 0.5 * (rr2 - a2 + hypot(rr2 - a2, 2.0 * a * z));
  return sqrt(r2);
}

/* null vector pieces shared by the three metric routines */
static void ks_null(double a, double x, double y, double z, double *f, double l[4], double *r_out, double *r2_out,
                    double *rr2_out) {
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double r2 = 0.5 * (rr2 - a2 + hypot(rr2 - a2, 2.0 * a * z));
  double r = sqrt(r2);
  *f = 2.0 * 1.0 * r2 * r / (r2 * r2 + a2 * z * z);
  l[1] = (r * x + a * y) / (r2 + a2);
  l[2] = (r * y - a * x) / (r2 + a2);
  l[3] = z / r;
  *r_out = r; *r2_out = r2; *rr2_out = rr2;
}

static void metric_cov(const orc_geo *g, double x, double y, double z, double gc[4][4]) {
  int m, n;
  if (g->flat) {
    for (m = 0; m < 4; m++) for (n = 0; n < 4; n++) gc[m][n] = m == n ? (m == 0 ? -1.0 : 1.0) : 0.0;
    return;
  }
  double f, l[4], r, r2, rr2;
  ks_null(g->a, x, y, z, &f, l, &r, &r2, &rr2);
  l[0] = 1.0;
  for (m = 0; m < 4; m++)
    for (n = 0; n < 4; n++) {
      double v = f * l[m] * l[n];
      gc[m][n] = m == n ? (m == 0 ? v - 1.0 : v + 1.0) : v;
    }
}

static void metric_con(const orc_geo *g, double x, double y, double z, double gc[4][4]) {
  int m, n;
  if (g->flat) {
    for (m = 0; m < 4; m++) for (n = 0; n < 4; n++) gc[m][n] = m == n ? (m == 0 ? -1.0 : 1.0) : 0.0;
    return;
  }
  double f, l[4], r, r2, rr2;
  ks_null(g->a, x, y, z, &f, l, &r, &r2, &rr2);
  l[0] = -1.0;
  for (m = 0; m < 4; m++)
    for (n = 0; n < 4; n++) {
      double v = -f * l[m] * l[n];
      gc[m][n] = m == n ? (m == 0 ? v - 1.0 : v + 1.0) : v;
    }
}

static void metric_con_deriv(const orc_geo *g, double x, double y, double z, double dg[3][4][4]) {
  int d, m, n;
  if (g->flat) {
    memset(dg, 0, 48 * sizeof(double));
    return;
  }
  double a = g->a, a2 = a * a;
  double f, l[4], r, r2, rr2;
  ks_null(a, x, y, z, &f, l, &r, &r2, &rr2);
  l[0] = -1.0;
  double dr[3], df[3], dl[3][4];
  dr[0] = r * x / (2.0 * r2 - rr2 + a2);
  dr[1] = r * y / (2.0 * r2 - rr2 + a2);
  dr[2] = (r * z + a2 * z / r) / (2.0 * r2 - rr2 + a2);
  df[0] = -(r2 * r2 - 3.0 * a2 * z * z) * dr[0] / (r * (r2 * r2 + a2 * z * z)) * f;
  df[1] = -(r2 * r2 - 3.0 * a2 * z * z) * dr[1] / (r * (r2 * r2 + a2 * z * z)) * f;
  df[2] = -((r2 * r2 - 3.0 * a2 * z * z) * dr[2] + 2.0 * a2 * r * z) / (r * (r2 * r2 + a2 * z * z)) * f;
  for (d = 0; d < 3; d++) dl[d][0] = 0.0;
  dl[0][1] = ((x - 2.0 * r * l[1]) * dr[0] + r) / (r2 + a2);
  dl[1][1] = ((x - 2.0 * r * l[1]) * dr[1] + a) / (r2 + a2);
  dl[2][1] = (x - 2.0 * r * l[1]) * dr[2] / (r2 + a2);
  dl[0][2] = ((y - 2.0 * r * l[2]) * dr[0] - a) / (r2 + a2);
  dl[1][2] = ((y - 2.0 * r * l[2]) * dr[1] + r) / (r2 + a2);
  dl[2][2] = (y - 2.0 * r * l[2]) * dr[2] / (r2 + a2);
  dl[0][3] = -z / r2 * dr[0];
  dl[1][3] = -z / r2 * dr[1];
  dl[2][3] = -z / r2 * dr[2] + 1.0 / r;
  for (d = 0; d < 3; d++)
    for (m = 0; m < 4; m++)
      for (n = 0; n < 4; n++)
        dg[d][m][n] = -(df[d] * l[m] * l[n] + f * dl[d][m] * l[n] + f * l[m] * dl[d][n]);
}

/* dy/dlambda for y = (x^mu, p_mu, s)  (geodesics.cpp:867-893) */
static void rhs(const orc_geo *g, const double y[9], double k[9]) {
  double gcov[4][4], gcon[4][4], dg[3][4][4], t[4] = {0, 0, 0, 0};
  int a, b, m, n, p;
  metric_cov(g, y[1], y[2], y[3], gcov);
  metric_con(g, y[1], y[2], y[3], gcon);
  metric_con_deriv(g, y[1], y[2], y[3], dg);
  for (p = 0; p < 9; p++) k[p] = 0.0;
  for (m = 0; m < 4; m++)
    for (n = 0; n < 4; n++) k[m] += gcon[m][n] * y[4 + n];
  for (a = 1; a < 4; a++)
    for (m = 0; m < 4; m++)
      for (n = 0; n < 4; n++) k[4 + a] -= 0.5 * dg[a - 1][m][n] * y[4 + m] * y[4 + n];
  for (a = 1; a < 4; a++)
    for (m = 0; m < 4; m++) t[a] += (gcon[a][m] - gcon[0][a] * gcon[0][m] / gcon[0][0]) * y[4 + m];
  for (a = 1; a < 4; a++)
    for (b = 1; b < 4; b++) k[8] += gcov[a][b] * t[a] * t[b];
  k[8] = -sqrt(k[8]);
}

/* rescale p_i so that g^{mu nu} p_mu p_nu = 0 (geodesics.cpp:296-309) */
static void renormalize(const orc_geo *g, const double x[4], double p[4]) {
  double gcon[4][4];
  int a, b;
  metric_con(g, x[1], x[2], x[3], gcon);
  double qa = 0.0, qb = 0.0;
  for (a = 1; a < 4; a++)
    for (b = 1; b < 4; b++) qa += gcon[a][b] * p[a] * p[b];
  for (a = 1; a < 4; a++) qb += 2.0 * gcon[0][a] * p[0] * p[a];
  double qc = gcon[0][0] * p[0] * p[0];
  double qd = sqrt(qb * qb - 4.0 * qa * qc);
  double factor = qb < 0.0 ? (qd - qb) / (2.0 * qa) : -2.0 * qc / (qb + qd);
  for (a = 1; a < 4; a++) p[a] *= factor;
}

static double dmax(double a, double b) { return a < b ? b : a; }
static double dmin(double a, double b) { return b < a ? b : a; }

/* ---------------------------------------------------------------- Dormand-Prince (geodesics.cpp:39-396) */

/* Traces n_rays rays.  Outputs in the reference's sample_* layout (source->camera order, len > 0):
 * pos, dir: (n_rays, cap, 4); len: (n_rays, cap); returns geodesic_num_steps (max sample_num). */
int orc_trace_dp(const orc_geo *g, long n_rays, const double *cam_pos, const double *cam_dir, int cap, int *num,
                 unsigned char *flags, double *pos, double *dir, double *len) {
  static const double A[7][6] = {
      {0}, {1.0 / 5.0}, {3.0 / 40.0, 9.0 / 40.0}, {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},
      {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},
      {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},
      {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
  static const double B5[7] = {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0, 0.0};
  static const double B4[7] = {5179.0 / 57600.0, 0.0, 7571.0 / 16695.0, 393.0 / 640.0, -92097.0 / 339200.0,
                               187.0 / 2100.0, 1.0 / 40.0};
  static const double B4M[7] = {6025192743.0 / 30085553152.0, 0.0, 51252292925.0 / 65400821598.0,
                                -2691868925.0 / 45128329728.0, 187940372067.0 / 1594534317056.0,
                                -1776094331.0 / 19743644256.0, 11237099.0 / 235043384.0};
  static const double D[7] = {-12715105075.0 / 11282082432.0, 0.0, 87487479700.0 / 32700410799.0,
                              -10690763975.0 / 1880347072.0, 701980252875.0 / 199316789632.0,
                              -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0};
  int steps_max = 0;
  long m;
#pragma omp parallel for schedule(dynamic, 4) reduction(max : steps_max)
  for (m = 0; m < n_rays; m++) {
    int ms = g->max_steps, p, q, s;
    double *gp = (double *)malloc((size_t)ms * 9 * sizeof(double)); /* tracing-order scratch: pos4, dir4, len */
    double y[9], yt[9], y5[9], y4[9], ym[8], k[7][9], rv[4][8];
    for (p = 0; p < 4; p++) { y[p] = cam_pos[4 * m + p]; y[4 + p] = cam_dir[4 * m + p]; }
    y[8] = 0.0;
    for (p = 0; p < 9; p++) y5[p] = y[p];
    double r_new = ks_radius(g->a, y[1], y[2], y[3]);
    double h_new = -g->ray_step * r_new;
    int retry = 0, fail = 0, flag = 0, count = 0, n = 0;
    while (n < ms) {
      if (retry > g->max_retries) { flag = 1; break; }
      double h = h_new;
      if (!fail && n > 0) for (p = 0; p < 9; p++) { y[p] = y5[p]; k[0][p] = k[6][p]; }
      if (!fail && n == 0) rhs(g, y, k[0]);
      double r = fail ? ks_radius(g->a, y[1], y[2], y[3]) : r_new;
      for (s = 1; s < 7; s++) {
        for (p = 0; p < 9; p++) yt[p] = y[p];
        for (q = 0; q < s; q++) for (p = 0; p < 9; p++) yt[p] += A[s][q] * h * k[q][p];
        rhs(g, yt, k[s]);
      }
      for (p = 0; p < 9; p++) y5[p] = y4[p] = y[p];
      for (q = 0; q < 7; q++) for (p = 0; p < 9; p++) { y5[p] += B5[q] * h * k[q][p]; y4[p] += B4[q] * h * k[q][p]; }
      r_new = ks_radius(g->a, y5[1], y5[2], y5[3]);
      double err = 0.0;
      for (p = 0; p < 8; p++) {
        double scale = g->tol_abs + g->tol_rel * dmax(fabs(y[p]), fabs(y5[p]));
        err = dmax(err, fabs(y5[p] - y4[p]) / scale);
      }
      if (!(err <= 1.0)) {
        double fac = 0.2;
        if (isfinite(err)) fac = dmax(0.9 * pow(err, -0.2), 0.2);
        h_new = h * fac; retry++; fail = 1;
        continue;
      }
      double fac = 10.0;
      if (err > 0.0) fac = dmin(dmax(0.9 * pow(err, -0.2), 0.2), 10.0);
      if (fail) fac = dmin(fac, 1.0);
      h_new = h * fac; retry = 0; fail = 0;
      for (p = 0; p < 8; p++) ym[p] = y[p];
      for (q = 0; q < 7; q++) for (p = 0; p < 8; p++) ym[p] += B4M[q] * h * k[q][p];
      double r_mid = ks_radius(g->a, ym[1], ym[2], ym[3]);
      double ds_step = g->ray_step * r_mid, ds_full = y5[8] - y[8];
      int n_ideal = (int)ceil(ds_full / ds_step), n_sub = n_ideal;
      if (n_sub > ms - n) { n_sub = ms - n; flag = 1; }
      if (n_ideal == 1) {
        for (p = 0; p < 8; p++) gp[(size_t)n * 9 + p] = ym[p];
        gp[(size_t)n * 9 + 8] = h;
      } else if (n_ideal > 1) {
        int nn;
        for (p = 0; p < 8; p++) {
          rv[0][p] = y5[p] - y[p];
          rv[1][p] = y[p] - y5[p] + h * k[0][p];
          rv[2][p] = 2.0 * (y5[p] - y[p]) - h * (k[0][p] + k[6][p]);
          rv[3][p] = 0.0;
        }
        for (q = 0; q < 7; q++) for (p = 0; p < 8; p++) rv[3][p] += D[q] * h * k[q][p];
        for (nn = 0; nn < n_sub; nn++) {
          double fr = (nn + 0.5) / n_ideal;
          for (p = 0; p < 8; p++)
            gp[(size_t)(n + nn) * 9 + p] =
                y[p] + fr * (rv[0][p] + (1.0 - fr) * (rv[1][p] + fr * (rv[2][p] + (1.0 - fr) * rv[3][p])));
          gp[(size_t)(n + nn) * 9 + 8] = h / n_ideal;
        }
      }
      renormalize(g, y5, y5 + 4);
      count += n_sub;
      if ((r_new > g->camera_r && r_new > r) || r_new < g->r_terminate) break;
      if (n + n_sub >= ms) flag = 1;
      n += n_sub;
    }
    /* truncate at the first sample beyond the boundaries (geodesics.cpp:327-349) */
    if (count > 1) {
      double rn = ks_radius(g->a, gp[1], gp[2], gp[3]);
      for (n = 1; n < count; n++) {
        double ro = rn;
        rn = ks_radius(g->a, gp[(size_t)n * 9 + 1], gp[(size_t)n * 9 + 2], gp[(size_t)n * 9 + 3]);
        if ((rn > g->camera_r && rn > ro) || rn < g->r_terminate) { count = n; break; }
      }
    }
    /* renormalise stored momenta, reverse (geodesics.cpp:352-371, 808-849) */
    for (n = 0; n < count; n++) {
      renormalize(g, gp + (size_t)n * 9, gp + (size_t)n * 9 + 4);
      size_t o = (size_t)m * cap + (size_t)(count - 1 - n);
      for (p = 0; p < 4; p++) { pos[4 * o + p] = gp[(size_t)n * 9 + p]; dir[4 * o + p] = gp[(size_t)n * 9 + 4 + p]; }
      len[o] = -gp[(size_t)n * 9 + 8];
    }
    num[m] = count;
    flags[m] = (unsigned char)flag;
    if (count > steps_max) steps_max = count;
    free(gp);
  }
  return steps_max;
}

/* Fixed-step integrators, ray_integrator = rk4 (order 4, geodesics.cpp:418-623) and rk2 (order 2, :626-805).
 * Step h = -ray_step (r - r_horizon) from the radius at the start of the step.  rk4 stores the mean of the step's two
 * end states, rk2 the half-step state reached with the first slope; the carried momentum is renormalised after each
 * step, the last allowed step flags the ray.  Truncation, renormalisation of the stored momenta and reversal are the
 * adaptive integrator's. */
int orc_trace_rk(const orc_geo *g, int order, long n_rays, const double *cam_pos, const double *cam_dir, int cap, int *num,
                 unsigned char *flags, double *pos, double *dir, double *len) {
  int steps_max = 0;
  long m;
  double r_horizon = 1.0 + sqrt(1.0 - g->a * g->a);
#pragma omp parallel for schedule(dynamic, 4) reduction(max : steps_max)
  for (m = 0; m < n_rays; m++) {
    int ms = g->max_steps, p, n, count = 0, flag = 0;
    double *gp = (double *)malloc((size_t)ms * 9 * sizeof(double));
    double y[9], ys[9], acc[9], k[9];
    for (p = 0; p < 4; p++) { y[p] = cam_pos[4 * m + p]; y[4 + p] = cam_dir[4 * m + p]; }
    y[8] = ys[8] = 0.0;
    double r_new = ks_radius(g->a, y[1], y[2], y[3]);
    for (n = 0; n < ms; n++) {
      double r = r_new, h = -g->ray_step * (r - r_horizon);
      double *rec = gp + (size_t)n * 9;
      if (order == 4) {
        static const double node[4] = {0.0, 0.5, 0.5, 1.0}, weight[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
        int st;
        for (st = 0; st < 4; st++) {
          if (st == 0) for (p = 0; p < 8; p++) ys[p] = y[p];
          else for (p = 0; p < 8; p++) ys[p] = y[p] + node[st] * h * k[p];
          rhs(g, ys, k);
          if (st == 0) for (p = 0; p < 8; p++) acc[p] = y[p] + weight[st] * h * k[p];
          else for (p = 0; p < 8; p++) acc[p] += weight[st] * h * k[p];
        }
        for (p = 0; p < 8; p++) rec[p] = 0.5 * (y[p] + acc[p]);
        for (p = 0; p < 8; p++) y[p] = acc[p];
      } else {
        rhs(g, y, k);
        for (p = 0; p < 8; p++) ys[p] = y[p] + h * k[p];
        for (p = 0; p < 8; p++) y[p] += 1.0 / 2.0 * h * k[p];
        for (p = 0; p < 8; p++) rec[p] = y[p];
        rhs(g, ys, k);
        for (p = 0; p < 8; p++) y[p] += 1.0 / 2.0 * h * k[p];
      }
      rec[8] = h;
      renormalize(g, y, y + 4);
      count++;
      r_new = ks_radius(g->a, y[1], y[2], y[3]);
      if ((r_new > g->camera_r && r_new > r) || r_new < g->r_terminate) break;
      if (n + 1 >= ms) flag = 1;
    }
    if (count > 1) {
      double rn = ks_radius(g->a, gp[1], gp[2], gp[3]);
      for (n = 1; n < count; n++) {
        double ro = rn;
        rn = ks_radius(g->a, gp[(size_t)n * 9 + 1], gp[(size_t)n * 9 + 2], gp[(size_t)n * 9 + 3]);
        if ((rn > g->camera_r && rn > ro) || rn < g->r_terminate) { count = n; break; }
      }
    }
    for (n = 0; n < count; n++) {
      renormalize(g, gp + (size_t)n * 9, gp + (size_t)n * 9 + 4);
      size_t o = (size_t)m * cap + (size_t)(count - 1 - n);
      for (p = 0; p < 4; p++) { pos[4 * o + p] = gp[(size_t)n * 9 + p]; dir[4 * o + p] = gp[(size_t)n * 9 + 4 + p]; }
      len[o] = -gp[(size_t)n * 9 + 8];
    }
    num[m] = count;
    flags[m] = (unsigned char)flag;
    if (count > steps_max) steps_max = count;
    free(gp);
  }
  return steps_max;
}

/* ---------------------------------------------------------------- unpolarized transfer (unpolarized.cpp:98-110) */

static double transfer_step(double image, double j, double alpha, double dl_cgs) {
  double dtau = alpha * dl_cgs;
  if (alpha > 0.0) {
    if (dtau <= 100.0) return exp(-dtau) * (image + j / alpha * expm1(dtau));
    return j / alpha;
  }
  return image + j * dl_cgs;
}

/* ---------------------------------------------------------------- formula model (formula_coefficients.cpp) */

typedef struct {
  double a, camera_r, x_unit;
  double r0, h, l0, q, nup, cn0, alpha, abs_a, beta;
  int fallback_nan;
} orc_formula;

/* j / nu^2 and alpha nu of one sample (formula_coefficients.cpp:25-183); zero for samples beyond the camera radius */
static void formula_j_alpha(const orc_formula *P, double x, double y, double z, const double *k, double freq, double mom,
                            int flagged, int l, double *j_out, double *al_out) {
  *j_out = 0.0;
  *al_out = 0.0;
  {
        if (P->fallback_nan && flagged) {
          if (l == 0) *j_out = *al_out = NAN;
        } else {
          double a = P->a, r = ks_radius(a, x, y, z);
          if (!(r > P->camera_r)) {
            double rr = sqrt(r * r - z * z), cth = z / r, sth = sqrt(1.0 - cth * cth);
            double ph = atan2(y, x) - atan(a / r), sph = sin(ph), cph = cos(ph);
            double delta = r * r - 2.0 * 1.0 * r + a * a, sigma = r * r + a * a * cth * cth;
            double gtt = -(1.0 + 2.0 * 1.0 * r * (r * r + a * a) / (delta * sigma));
            double gtph = -2.0 * 1.0 * a * r / (delta * sigma);
            double grr = delta / sigma, gthth = 1.0 / sigma;
            double gphph = (sigma - 2.0 * 1.0 * r) / (delta * sigma * sth * sth);
            double ll = P->l0 / (1.0 + rr) * pow(rr, 1.0 + P->q);
            double un = 1.0 / sqrt(-gtt + 2.0 * gtph * ll - gphph * ll * ll);
            double u_t = -un, u_ph = un * ll;
            double ut_bl = gtt * u_t + gtph * u_ph, ur_bl = grr * 0.0, uth_bl = gthth * 0.0;
            double uph_bl = gtph * u_t + gphph * u_ph;
            double ut = ut_bl + 2.0 * 1.0 * r / delta * ur_bl, ur = ur_bl, uth = uth_bl, uph = uph_bl + a / delta * ur_bl;
            double u0 = ut;
            double u1 = sth * cph * ur + cth * (r * cph - a * sph) * uth + sth * (-r * sph - a * cph) * uph;
            double u2 = sth * sph * ur + cth * (r * sph + a * cph) * uth + sth * (r * cph - a * sph) * uph;
            double u3 = cth * ur - r * sth * uth;
            double nn0 = exp(-0.5 * (r * r / (P->r0 * P->r0) + P->h * P->h * cth * cth));
            double nu = -(u0 * k[0] + u1 * k[1] + u2 * k[2] + u3 * k[3]) * freq * mom;
            double jn = P->cn0 * nn0 * pow(nu / P->nup, -P->alpha);
            *j_out = jn / (nu * nu);
            double an = P->abs_a * P->cn0 * nn0 * pow(nu / P->nup, -P->beta - P->alpha);
            *al_out = an * nu;
          }
        }
  }
}

/* image: (F, n_rays).  Samples in the layout orc_trace_dp produces. */
void orc_formula_image(const orc_formula *P, long n_rays, int cap, const int *num, const unsigned char *flags,
                       const double *pos, const double *dir, const double *len, const double *mom_factor,
                       int F, const double *freqs, double *image) {
  long m;
#pragma omp parallel for schedule(dynamic, 4)
  for (m = 0; m < n_rays; m++) {
    int l, n;
    for (l = 0; l < F; l++) {
      double I = 0.0;
      for (n = 0; n < num[m]; n++) {
        size_t o = (size_t)m * cap + n;
        double j, al, dl_cgs = len[o] * P->x_unit / (freqs[l] * mom_factor[m]);
        formula_j_alpha(P, pos[4 * o + 1], pos[4 * o + 2], pos[4 * o + 3], dir + 4 * o, freqs[l], mom_factor[m], flags[m], l, &j, &al);
        I = transfer_step(I, j, al, dl_cgs);
      }
      image[(size_t)l * n_rays + m] = I * (freqs[l] * freqs[l] * freqs[l]);
    }
  }
}

/* Auxiliary images of the formula model (unpolarized.cpp:60-185): earliest coordinate time [s], proper length [cm],
 * affine length, integrated emissivity and optical depth per frequency, and the number of crossings of the plane through
 * the origin normal to the camera position.  time, length, crossings: (n_rays); lambda, emission, tau: (F, n_rays). */
void orc_formula_aux(const orc_formula *P, long n_rays, int cap, const int *num, const unsigned char *flags,
                     const double *pos, const double *dir, const double *len, const double *mom_factor, int F,
                     const double *freqs, const double *camera_x, double *time, double *length, double *lambda,
                     double *emission, double *tau, double *crossings) {
  const double t_unit = P->x_unit / 2.99792458e10;
  orc_geo geo;
  long m;
  memset(&geo, 0, sizeof geo);
  geo.a = P->a;
#pragma omp parallel for schedule(dynamic, 4)
  for (m = 0; m < n_rays; m++) {
    int l, n, a, b, mu;
    for (l = 0; l < F; l++) {
      double sum_lambda = 0.0, sum_emission = 0.0, sum_tau = 0.0, earliest = 0.0, proper = 0.0;
      int count = 0, sign = 0;
      if (num[m] > 0) {
        size_t o0 = (size_t)m * cap;
        sign = camera_x[1] * pos[4 * o0 + 1] + camera_x[2] * pos[4 * o0 + 2] + camera_x[3] * pos[4 * o0 + 3] > 0.0;
      }
      for (n = 0; n < num[m]; n++) {
        size_t o = (size_t)m * cap + n;
        double x = pos[4 * o + 1], y = pos[4 * o + 2], z = pos[4 * o + 3];
        const double *k = dir + 4 * o;
        double j, al, dl_cgs = len[o] * P->x_unit / (freqs[l] * mom_factor[m]);
        formula_j_alpha(P, x, y, z, k, freqs[l], mom_factor[m], flags[m], l, &j, &al);
        sum_lambda += dl_cgs;
        sum_emission += j * dl_cgs;
        sum_tau += al * dl_cgs;
        if (l == 0) {
          double gcov[4][4], gcon[4][4], t[4] = {0, 0, 0, 0}, sq = 0.0, t_cgs = pos[4 * o] * t_unit;
          int now;
          earliest = t_cgs < earliest ? t_cgs : earliest;   /* std::min(image, t_cgs), image starting at 0 */
          metric_cov(&geo, x, y, z, gcov);
          metric_con(&geo, x, y, z, gcon);
          for (a = 1; a < 4; a++)
            for (mu = 0; mu < 4; mu++) t[a] += (gcon[a][mu] - gcon[0][a] * gcon[0][mu] / gcon[0][0]) * k[mu];
          for (a = 1; a < 4; a++)
            for (b = 1; b < 4; b++) sq += gcov[a][b] * t[a] * t[b];
          proper += sqrt(sq) * len[o] * P->x_unit;
          now = camera_x[1] * x + camera_x[2] * y + camera_x[3] * z > 0.0;
          if (now != sign) count++;
          sign = now;
        }
      }
      lambda[(size_t)l * n_rays + m] = sum_lambda;
      emission[(size_t)l * n_rays + m] = sum_emission;
      tau[(size_t)l * n_rays + m] = sum_tau;
      if (l == 0) {
        time[m] = earliest;
        length[m] = proper;
        crossings[m] = (double)count;
      }
    }
  }
}

/* ---------------------------------------------------------------- simulation model */

typedef struct {
  double a, camera_r, x_unit;
  int n_b, n_k, n_j, n_i, interp, fallback_nan;
  double d_unit, mu, ne_ni, rat_low, rat_high;
  double cut_sigma_max;   /* < 0 disables; the other value cuts of the examples are disabled */
  int coord;              /* 0 spherical Kerr-Schild grid, 1 Cartesian Kerr-Schild grid (simulation_coord = cks) */
  /* power-law electrons (simulation_coefficients.cpp:53-66,559-585); thermal fraction = 1 - power_frac */
  double power_frac, power_p, power_gamma_min, power_gamma_max;
  /* kappa-distribution electrons (simulation_coefficients.cpp:82-105,608-664) */
  double kappa_frac, kappa, kappa_w;
  int flat;               /* ray_flat: Minkowski geodesic metric in the radiation stage (radiation_geometry.cpp:142,210) */
  double cut_omit_in, cut_omit_out;   /* sphere cuts, < 0 disables (simulation_sampling.cpp:256-261) */
  /* plasma_use_p = false: electron temperature from the internal energies (simulation_coefficients.cpp:342-346) */
  int use_energy;
  double gamma, gamma_i, gamma_e;
  /* plasma_model = code_kappa: electron entropy is variable 8 of prim (simulation_sampling.cpp:726,811-833;
     simulation_coefficients.cpp:351-358) */
  int code_kappa;
  /* remaining cuts.  geometric (simulation_sampling.cpp:245-295): camera plane (near / far, with the camera position
     cut_cam), midplane angle (radians, sign selects inside / outside) and height, arbitrary plane;
     cell values (simulation_coefficients.cpp:361-375): min / max pairs of rho, n_e, p_gas, theta_e, B, sigma, 1/beta,
     each < 0 disables (sigma max stays cut_sigma_max above) */
  int cut_omit_near, cut_omit_far, cut_plane;
  double cut_cam[3], cut_midplane_theta, cut_midplane_z, cut_plane_origin[3], cut_plane_normal[3];
  double cut_val_min[7], cut_val_max[7];
  /* fallback_nan = false: samples outside the grid take these values, zero velocity and field, stored as float
     (simulation_sampling.cpp:695-708; radiation_integrator.hpp:182-187) */
  double fallback_rho, fallback_pgas, fallback_kappa;
  /* slow light (simulation_sampling.cpp:297-349, 736-775, 840-905): n_t > 0 time slices in prim, newest first, at
     times[]; a sample at coordinate time t is looked up at t + snapshot_time; nearest slice or linear blend */
  int n_t, slow_interp;
  double snapshot_time, times[64];
  /* simulation_block_interp = true (Athena++ / AthenaK meshes): trilinear anchors looked up across MeshBlocks
     (simulation_sampling.cpp:506-553, 1068-1331, 1365-1386); levels (n_b), locations (n_b, 3) in i, j, k order,
     n_3_root = root-grid cells in the third dimension.  Not combined with slow light here. */
  int block_interp, n_3_root;
  const int *levels, *locations;
} orc_sim;

/* one feature of a false-colour render image (rendering.cpp:100-165): type 0 fill, 1 thresh, 2 rise, 3 fall */
typedef struct {
  int image, quantity, type;
  double min, max, tau_scale, thresh, opacity, xyz[3];
} orc_feature;

/* Gauss hypergeometric function by the reference's transformed, 10-term series (simulation_coefficients.cpp:740-773) */
static double hypergeometric(double alpha, double beta, double gamma, double z) {
  double a = alpha, b = gamma - beta, c_ = gamma, x = z / (z - 1.0);
  double result = 1.0, a_k = 1.0, b_k = 1.0, c_k = 1.0, xk = 1.0, k_factorial = 1.0;
  int k;
  for (k = 1; k <= 10; k++) {
    a_k *= a + k - 1.0;
    b_k *= b + k - 1.0;
    c_k *= c_ + k - 1.0;
    xk *= x;
    k_factorial *= k;
    result += a_k * b_k * xk / (c_k * k_factorial);
  }
  return result * pow(1.0 - z, -alpha);
}

static double g4(const float *prim, const orc_sim *P, int v, int b, int k, int j, int i) {
  return (double)prim[((((size_t)v * P->n_b + b) * P->n_k + k) * P->n_j + j) * P->n_i + i];
}

static double trilinear(const float *prim, const orc_sim *P, int v, int b, int k, int j, int i, double fk, double fj,
                        double fi) {
  return (1.0 - fk) * (1.0 - fj) * (1.0 - fi) * g4(prim, P, v, b, k, j, i) +
         (1.0 - fk) * (1.0 - fj) * fi * g4(prim, P, v, b, k, j, i + 1) +
         (1.0 - fk) * fj * (1.0 - fi) * g4(prim, P, v, b, k, j + 1, i) + (1.0 - fk) * fj * fi * g4(prim, P, v, b, k, j + 1, i + 1) +
         fk * (1.0 - fj) * (1.0 - fi) * g4(prim, P, v, b, k + 1, j, i) + fk * (1.0 - fj) * fi * g4(prim, P, v, b, k + 1, j, i + 1) +
         fk * fj * (1.0 - fi) * g4(prim, P, v, b, k + 1, j + 1, i) + fk * fj * fi * g4(prim, P, v, b, k + 1, j + 1, i + 1);
}

/* orthonormal tetrad (radiation_geometry.cpp:597-658) */
/* rho, pgas, uu1-3, bb1-3, entropy of one time slice at a located sample: the cell's own values, or trilinear with
   the non-positive fallback of rho, pgas and entropy (simulation_sampling.cpp:716-733, 796-835) */
static void slice_values(const float *prim, const orc_sim *P, int b, int k, int j, int i, double fk, double fj, double fi,
                         double out[9]) {
  int q;
  out[8] = 0.0;
  if (!P->interp) {
    for (q = 0; q < 8; q++) out[q] = g4(prim, P, q, b, k, j, i);
    if (P->code_kappa) out[8] = g4(prim, P, 8, b, k, j, i);
    return;
  }
  for (q = 0; q < 8; q++) out[q] = trilinear(prim, P, q, b, k, j, i, fk, fj, fi);
  if (out[0] <= 0.0) out[0] = g4(prim, P, 0, b, k, j, i);
  if (out[1] <= 0.0) out[1] = g4(prim, P, 1, b, k, j, i);
  if (P->code_kappa) {
    out[8] = trilinear(prim, P, 8, b, k, j, i, fk, fj, fi);
    if (out[8] <= 0.0) out[8] = g4(prim, P, 8, b, k, j, i);
  }
}

/* slow light: the values at time slice t_ind, or blended with slice t_ind + 1 (simulation_sampling.cpp:736-775, 840-905) */
static void slow_values(const float *prim, const orc_sim *P, int t_ind, double t_frac, int b, int k, int j, int i, double fk,
                        double fj, double fi, double out[9]) {
  size_t stride = (size_t)(P->code_kappa ? 9 : 8) * P->n_b * P->n_k * P->n_j * P->n_i;
  double next[9];
  int q;
  slice_values(prim + (size_t)t_ind * stride, P, b, k, j, i, fk, fj, fi, out);
  if (!P->slow_interp) return;
  slice_values(prim + (size_t)(t_ind + 1) * stride, P, b, k, j, i, fk, fj, fi, next);
  for (q = 0; q < 9; q++) out[q] = (1.0 - t_frac) * out[q] + t_frac * next[q];
}

/* time slice of a sample at coordinate time x0 (simulation_sampling.cpp:297-349); times[] descend */
static void slow_slice(const orc_sim *P, double x0, int *t_out, double *frac_out) {
  int t = 0, n = P->n_t;
  double frac = 0.0;
  if (x0 >= P->times[0]) {
  } else if (x0 <= P->times[n - 1]) {
    if (P->slow_interp) { t = n - 2; frac = 1.0; }
    else t = n - 1;
  } else {
    while (P->times[t] > x0) t++;
    if (P->slow_interp) { t--; frac = (x0 - P->times[t]) / (P->times[t + 1] - P->times[t]); }
    else if (P->times[t - 1] - x0 <= x0 - P->times[t]) t--;
  }
  *t_out = t;
  *frac_out = frac;
}

static int mesh_find(const orc_sim *P, int level, const int want[3]) {
  int b;
  for (b = 0; b < P->n_b; b++)
    if (P->levels[b] == level && P->locations[3 * b] == want[0] && P->locations[3 * b + 1] == want[1] &&
        P->locations[3 * b + 2] == want[2])
      return b;
  return -1;
}

/* Anchor (block, k, j, i) for the cell index idx_in = (i, j, k) of block b, where components may be -1 or n (one cell
 * beyond the block): the neighbouring block at the same level, else the coarser one, else the finer one (there shifted
 * towards the sample), with phi periodic in spherical coordinates; directions without any neighbour are clamped
 * (FindNearbyInds, simulation_sampling.cpp:1068-1331).  cell = (i, j, k) of the sample's own cell, x its coordinates,
 * xv the three cell-centre arrays.  Integer divisions truncate as in the reference. */
static void mesh_anchor(const orc_sim *P, const double *const xv[3], int b, const int idx_in[3], const int cell[3],
                        const double x[3], int out[4]) {
  int n[3] = {P->n_i, P->n_j, P->n_k}, idx[3], safe[3], upper[3], off[3], loc[3], want[3], sought[3];
  int level = P->levels[b], d, e, ba, max_level = 0, inside = 1;
  for (ba = 0; ba < P->n_b; ba++) if (P->levels[ba] > max_level) max_level = P->levels[ba];
#define N3(l) ((P->n_3_root / P->n_k) << (l))
  for (d = 0; d < 3; d++) {
    idx[d] = idx_in[d];
    loc[d] = P->locations[3 * b + d];
    upper[d] = idx[d] > n[d] / 2;
    safe[d] = idx[d] < 0 ? 0 : idx[d] > n[d] - 1 ? n[d] - 1 : idx[d];
    off[d] = idx[d] != safe[d];
    if (off[d]) inside = 0;
  }
  if (inside) { out[0] = b; out[1] = idx[2]; out[2] = idx[1]; out[3] = idx[0]; return; }
  int wrap_low = P->coord == 0 && idx[2] == -1 && loc[2] == 0;
  int wrap_high = P->coord == 0 && idx[2] == n[2] && loc[2] == N3(level) - 1;
  for (ba = 0; ba < P->n_b; ba++) {
    int la = P->levels[ba];
    const int *q = P->locations + 3 * ba;
    for (d = 0; d < 3; d++) {
      if (!off[d]) continue;
      int step = idx[d] == -1 ? -1 : 1, same = la == level, coarser = la == level - 1, finer = la == level + 1;
      for (e = 0; e < 3; e++) {
        same = same && q[e] == (e == d ? loc[e] + step : loc[e]);
        coarser = coarser && q[e] == (e == d ? (loc[e] + step) / 2 : loc[e] / 2);
        finer = finer && q[e] == (e == d ? (idx[d] == -1 ? loc[e] * 2 - 1 : loc[e] * 2 + 2) : loc[e] * 2 + upper[e]);
      }
      if (same || coarser || finer) off[d] = 0;
      if (d == 2 && off[2] && (wrap_low || wrap_high)) {
        int edge = wrap_low ? ((P->n_3_root / P->n_k) << la) - 1 : 0;
        same = la == level && q[0] == loc[0] && q[1] == loc[1] && q[2] == edge;
        coarser = la == level - 1 && q[0] == loc[0] / 2 && q[1] == loc[1] / 2 && q[2] == edge;
        finer = la == level + 1 && q[0] == loc[0] * 2 + upper[0] && q[1] == loc[1] * 2 + upper[1] && q[2] == edge;
        if (same || coarser || finer) off[2] = 0;
      }
    }
  }
  for (d = 0; d < 3; d++) if (off[d]) idx[d] = safe[d];
  /* same level */
  for (d = 0; d < 3; d++) {
    want[d] = idx[d] == safe[d] ? loc[d] : idx[d] == -1 ? loc[d] - 1 : loc[d] + 1;
    sought[d] = idx[d] == safe[d] ? idx[d] : idx[d] == -1 ? n[d] - 1 : 0;
  }
  wrap_low = P->coord == 0 && idx[2] == -1 && loc[2] == 0;
  wrap_high = P->coord == 0 && idx[2] == n[2] && loc[2] == N3(level) - 1;
  if (wrap_low) want[2] = N3(level) - 1;
  if (wrap_high) want[2] = 0;
  if ((ba = mesh_find(P, level, want)) >= 0) { out[0] = ba; out[1] = sought[2]; out[2] = sought[1]; out[3] = sought[0]; return; }
  /* coarser level */
  if (level - 1 >= 0) {
    for (d = 0; d < 3; d++) {
      want[d] = idx[d] == safe[d] ? loc[d] / 2 : idx[d] == -1 ? (loc[d] - 1) / 2 : (loc[d] + 1) / 2;
      sought[d] = idx[d] == safe[d] ? (loc[d] % 2 * n[d] + idx[d]) / 2 : idx[d] == -1 ? n[d] - 1 : 0;
    }
    if (wrap_low) want[2] = N3(level - 1) - 1;
    if (wrap_high) want[2] = 0;
    if ((ba = mesh_find(P, level - 1, want)) >= 0) { out[0] = ba; out[1] = sought[2]; out[2] = sought[1]; out[3] = sought[0]; return; }
  }
  /* finer level */
  for (d = 0; d < 3; d++) {
    want[d] = loc[d] * 2 + (idx[d] == safe[d] ? 0 : idx[d] == -1 ? -1 : 1) + upper[d];
    sought[d] = idx[d] == safe[d] ? (upper[d] ? (idx[d] - n[d] / 2) * 2 : idx[d] * 2) : idx[d] == -1 ? n[d] - 2 : 0;
  }
  if (wrap_low && level + 1 <= max_level) want[2] = N3(level + 1) - 1;
  if (wrap_high) want[2] = 0;
  ba = mesh_find(P, level + 1, want);
  if (ba < 0) { out[0] = -1; out[1] = out[2] = out[3] = 0; return; }   /* the reference throws "Grid interpolation failed." */
  for (d = 0; d < 3; d++)
    sought[d] += (idx[d] < cell[d] || (idx[d] == cell[d] && x[d] > xv[d][(size_t)b * n[d] + cell[d]])) ? 1 : 0;
  out[0] = ba; out[1] = sought[2]; out[2] = sought[1]; out[3] = sought[0];
#undef N3
}

static void tetrad(const double ucon[4], const double ucov[4], const double kcon[4], const double kcov[4],
                   const double up[4], double gcov[4][4], double gcon[4][4], double e[4][4]) {
  int m, n;
  double omega = 0.0, kup = 0.0, uup = 0.0, norm = 0.0, c[4];
  for (m = 0; m < 4; m++) omega -= kcov[m] * ucon[m];
  for (m = 0; m < 4; m++) kup += kcov[m] * up[m];
  kup /= omega;
  for (m = 0; m < 4; m++) uup += ucov[m] * up[m];
  uup /= omega;
  for (m = 0; m < 4; m++) e[0][m] = ucon[m];
  for (m = 0; m < 4; m++) e[3][m] = kcon[m] / omega - ucon[m];
  for (m = 0; m < 4; m++) e[2][m] = up[m] - kup * e[3][m] + uup * kcon[m];
  for (m = 0; m < 4; m++) for (n = 0; n < 4; n++) norm += gcov[m][n] * e[2][m] * e[2][n];
  norm = sqrt(norm);
  for (m = 0; m < 4; m++) e[2][m] /= norm;
  /* e_1 = Levi-Civita contraction of e_0, e_2, e_3 (covariant), raised */
  static const int perm[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  for (m = 0; m < 4; m++) {
    int i0 = perm[m][0], i1 = perm[m][1], i2 = perm[m][2];
    double det = e[0][i0] * (e[2][i1] * e[3][i2] - e[2][i2] * e[3][i1]) + e[0][i1] * (e[2][i2] * e[3][i0] - e[2][i0] * e[3][i2]) +
                 e[0][i2] * (e[2][i0] * e[3][i1] - e[2][i1] * e[3][i0]);
    c[m] = (m % 2 == 0) ? -det : det;
  }
  for (m = 0; m < 4; m++) {
    e[1][m] = 0.0;
    for (n = 0; n < 4; n++) e[1][m] += gcon[m][n] * c[n];
  }
}

/* Per-sample thermal coefficients + transfer on an SKS grid.  image: (n_rays); inds out: (n_rays,cap,4) or NULL
 * (entries of cut / off-grid samples are left at -1).  One frequency, ti_te_beta model with plasma_use_p.
 * aux (with camera_x), optional: (27, n_rays) auxiliary images in the order time, length, lambda, emission, tau,
 * crossings, lambda_ave[7], emission_ave[7], tau_int[7] over the cell values rho, n_e, p_gas, Theta_e, B, sigma,
 * 1/beta (unpolarized.cpp:60-200; cell values simulation_coefficients.cpp:376-387). */
void orc_simulation_image(const orc_sim *P, long n_rays, int cap, const int *num, const unsigned char *flags,
                          const double *pos, const double *dir, const double *len, const double *mom_factor,
                          double freq, const double *x1f, const double *x2f, const double *x3f, const double *x1v,
                          const double *x2v, const double *x3v, const float *prim, double *image, int *inds,
                          const double *camera_x, double *aux, int n_feat, const orc_feature *feat, int n_render,
                          double *render) {
  const double c = 2.99792458e10, hpl = 6.62607015e-27, m_p = 1.67262192369e-24, m_e = 9.1093837015e-28, qe = 4.80320425e-10;
  const double e_unit = P->d_unit * c * c, b_unit = sqrt(4.0 * PI * e_unit);
  orc_geo geo;
  memset(&geo, 0, sizeof geo);
  geo.a = P->a;
  geo.flat = P->flat;
  long m;
#pragma omp parallel for schedule(dynamic, 4)
  for (m = 0; m < n_rays; m++) {
    int n, b = 0, i, j, k, mu, nu_;
    double I = 0.0;
    double prev_cv[7] = {NAN, NAN, NAN, NAN, NAN, NAN, NAN};   /* rendering.cpp:58-61 */
    if (render) for (i = 0; i < n_render * 3; i++) render[(size_t)i * n_rays + m] = 0.0;
    int n_i = P->n_i, n_j = P->n_j, n_k = P->n_k;
    double a_time = 0.0, a_length = 0.0, a_lambda = 0.0, a_emission = 0.0, a_tau = 0.0, a_lave[7], a_eave[7], a_tint[7];
    int a_cross = 0, a_sign = 0;
    for (i = 0; i < 7; i++) a_lave[i] = a_eave[i] = a_tint[i] = 0.0;
    if (aux && num[m] > 0) {
      size_t o0 = (size_t)m * cap;
      a_sign = camera_x[1] * pos[4 * o0 + 1] + camera_x[2] * pos[4 * o0 + 2] + camera_x[3] * pos[4 * o0 + 3] > 0.0;
    }
    for (n = 0; n < num[m]; n++) {
      size_t o = (size_t)m * cap + n;
      double cv[7] = {NAN, NAN, NAN, NAN, NAN, NAN, NAN};
      double x = pos[4 * o + 1], y = pos[4 * o + 2], z = pos[4 * o + 3];
      const double *kcov = dir + 4 * o;
      double dl_cgs = len[o] * P->x_unit / (freq * mom_factor[m]);
      double jv = 0.0, av = 0.0;
      double rho, pgas, uu[3], bb[3], entropy = 0.0, sv[9], t_frac = 0.0;
      int t_ind = 0;
      int have = 0;
      if (inds) for (i = 0; i < 4; i++) inds[4 * o + i] = -1;
      if (P->fallback_nan && flags[m]) {
        rho = pgas = uu[0] = uu[1] = uu[2] = bb[0] = bb[1] = bb[2] = NAN;
        have = 1;
      } else {
        double a = P->a, r = ks_radius(a, x, y, z);
        int geom_cut = r > P->camera_r;
        if (!geom_cut && (P->cut_omit_near || P->cut_omit_far)) {
          double dot = x * P->cut_cam[0] + y * P->cut_cam[1] + z * P->cut_cam[2];
          geom_cut = (P->cut_omit_near && dot > 0.0) || (P->cut_omit_far && dot < 0.0);
        }
        if (!geom_cut) geom_cut = (P->cut_omit_in >= 0.0 && r < P->cut_omit_in) || (P->cut_omit_out >= 0.0 && r > P->cut_omit_out);
        if (!geom_cut && P->cut_midplane_theta != 0.0) {
          double dth = fabs(acos(z / r) - PI / 2.0);
          geom_cut = (P->cut_midplane_theta > 0.0 && dth > P->cut_midplane_theta) ||
                     (P->cut_midplane_theta < 0.0 && dth < -P->cut_midplane_theta);
        }
        if (!geom_cut)
          geom_cut = (P->cut_midplane_z > 0.0 && fabs(z) > P->cut_midplane_z) || (P->cut_midplane_z < 0.0 && fabs(z) < -P->cut_midplane_z);
        if (!geom_cut && P->cut_plane)
          geom_cut = (x - P->cut_plane_origin[0]) * P->cut_plane_normal[0] + (y - P->cut_plane_origin[1]) * P->cut_plane_normal[1] +
                     (z - P->cut_plane_origin[2]) * P->cut_plane_normal[2] < 0.0;
        if (!geom_cut) {
          double x1 = r, x2 = acos(z / r), x3 = atan2(y, x) - atan(a / r);
          x3 += x3 < 0.0 ? 2.0 * PI : 0.0;
          x3 -= x3 >= 2.0 * PI ? 2.0 * PI : 0.0;
          if (P->coord == 1) { x1 = x; x2 = y; x3 = z; }   /* radiation_geometry.cpp:37-57: cks keeps x, y, z */
          if (P->n_t > 0) slow_slice(P, pos[4 * o] + P->snapshot_time, &t_ind, &t_frac);
          /* block: keep while inside (inclusive), else first match (simulation_sampling.cpp:352-394) */
          if (x1 < x1f[(size_t)b * (n_i + 1)] || x1 > x1f[(size_t)b * (n_i + 1) + n_i] || x2 < x2f[(size_t)b * (n_j + 1)] ||
              x2 > x2f[(size_t)b * (n_j + 1) + n_j] || x3 < x3f[(size_t)b * (n_k + 1)] || x3 > x3f[(size_t)b * (n_k + 1) + n_k]) {
            int bn;
            for (bn = 0; bn < P->n_b; bn++)
              if (x1 >= x1f[(size_t)bn * (n_i + 1)] && x1 <= x1f[(size_t)bn * (n_i + 1) + n_i] && x2 >= x2f[(size_t)bn * (n_j + 1)] &&
                  x2 <= x2f[(size_t)bn * (n_j + 1) + n_j] && x3 >= x3f[(size_t)bn * (n_k + 1)] && x3 <= x3f[(size_t)bn * (n_k + 1) + n_k])
                break;
            if (bn == P->n_b) {
              if (P->fallback_nan) rho = pgas = uu[0] = uu[1] = uu[2] = bb[0] = bb[1] = bb[2] = NAN;
              else {
                rho = (double)(float)P->fallback_rho; pgas = (double)(float)P->fallback_pgas; entropy = (double)(float)P->fallback_kappa;
                uu[0] = uu[1] = uu[2] = bb[0] = bb[1] = bb[2] = 0.0;
              }
              have = 1;
              goto sampled;
            }
            b = bn;
          }
          for (i = 0; i < n_i; i++) if (x1f[(size_t)b * (n_i + 1) + i + 1] >= x1) break;
          for (j = 0; j < n_j; j++) if (x2f[(size_t)b * (n_j + 1) + j + 1] >= x2) break;
          for (k = 0; k < n_k; k++) if (x3f[(size_t)b * (n_k + 1) + k + 1] >= x3) break;
          if (!P->interp) {
            if (inds) { inds[4 * o] = b; inds[4 * o + 1] = k; inds[4 * o + 2] = j; inds[4 * o + 3] = i; }
            rho = g4(prim, P, 0, b, k, j, i); pgas = g4(prim, P, 1, b, k, j, i);
            for (mu = 0; mu < 3; mu++) { uu[mu] = g4(prim, P, 2 + mu, b, k, j, i); bb[mu] = g4(prim, P, 5 + mu, b, k, j, i); }
            if (P->code_kappa) entropy = g4(prim, P, 8, b, k, j, i);
            if (P->n_t > 0) slow_values(prim, P, t_ind, t_frac, b, k, j, i, 0.0, 0.0, 0.0, sv);
          } else {
            const double *v1 = x1v + (size_t)b * n_i, *v2 = x2v + (size_t)b * n_j, *v3 = x3v + (size_t)b * n_k;
            int im = (i == 0 || (i != n_i - 1 && x1 >= v1[i])) ? i : i - 1;
            int jm = (j == 0 || (j != n_j - 1 && x2 >= v2[j])) ? j : j - 1;
            int km = (k == 0 || (k != n_k - 1 && x3 >= v3[k])) ? k : k - 1;
            double fi = (x1 - v1[im]) / (v1[im + 1] - v1[im]);
            double fj = (x2 - v2[jm]) / (v2[jm + 1] - v2[jm]);
            double fk = (x3 - v3[km]) / (v3[km + 1] - v3[km]);
            if (inds) { inds[4 * o] = b; inds[4 * o + 1] = km; inds[4 * o + 2] = jm; inds[4 * o + 3] = im; }
            rho = trilinear(prim, P, 0, b, km, jm, im, fk, fj, fi);
            pgas = trilinear(prim, P, 1, b, km, jm, im, fk, fj, fi);
            if (rho <= 0.0) rho = g4(prim, P, 0, b, km, jm, im);
            if (pgas <= 0.0) pgas = g4(prim, P, 1, b, km, jm, im);
            if (P->code_kappa) {
              entropy = trilinear(prim, P, 8, b, km, jm, im, fk, fj, fi);
              if (entropy <= 0.0) entropy = g4(prim, P, 8, b, km, jm, im);
            }
            for (mu = 0; mu < 3; mu++) {
              uu[mu] = trilinear(prim, P, 2 + mu, b, km, jm, im, fk, fj, fi);
              bb[mu] = trilinear(prim, P, 5 + mu, b, km, jm, im, fk, fj, fi);
            }
            if (P->n_t > 0) slow_values(prim, P, t_ind, t_frac, b, km, jm, im, fk, fj, fi, sv);
            if (P->block_interp) {
              /* anchors one cell beyond the block are allowed; their coordinates mirror the block's own spacing
                 (simulation_sampling.cpp:506-524; the upper mirror reads x?v one past the cell, as the reference does) */
              const double *f1 = x1f + (size_t)b * (n_i + 1), *f2 = x2f + (size_t)b * (n_j + 1), *f3 = x3f + (size_t)b * (n_k + 1);
              const double *const xv[3] = {x1v, x2v, x3v};
              int lo[3] = {x1 >= v1[i] ? i : i - 1, x2 >= v2[j] ? j : j - 1, x3 >= v3[k] ? k : k - 1};
              int cell[3] = {i, j, k}, p, q, anchor[8][4];
              double xs[3] = {x1, x2, x3};
              double x1m = lo[0] == -1 ? 2.0 * f1[i] - v1[i] : v1[lo[0]], x1p = lo[0] + 1 == n_i ? 2.0 * v1[i + 1] - v1[i] : v1[lo[0] + 1];
              double x2m = lo[1] == -1 ? 2.0 * f2[j] - v2[j] : v2[lo[1]], x2p = lo[1] + 1 == n_j ? 2.0 * v2[j + 1] - v2[j] : v2[lo[1] + 1];
              double x3m = lo[2] == -1 ? 2.0 * f3[k] - v3[k] : v3[lo[2]], x3p = lo[2] + 1 == n_k ? 2.0 * v3[k + 1] - v3[k] : v3[lo[2] + 1];
              double gi = (x1 - x1m) / (x1p - x1m), gj = (x2 - x2m) / (x2p - x2m), gk = (x3 - x3m) / (x3p - x3m);
              for (p = 0; p < 8; p++) {
                int at[3] = {lo[0] + (p & 1), lo[1] + ((p >> 1) & 1), lo[2] + ((p >> 2) & 1)};
                mesh_anchor(P, xv, b, at, cell, xs, anchor[p]);
              }
              if (inds) for (q = 0; q < 4; q++) inds[4 * o + q] = anchor[0][q];
              for (q = 0; q < 9; q++) {
                double c[8];
                if (q == 8 && !P->code_kappa) { sv[8] = 0.0; break; }
                for (p = 0; p < 8; p++) c[p] = g4(prim, P, q, anchor[p][0], anchor[p][1], anchor[p][2], anchor[p][3]);
                sv[q] = (1.0 - gk) * (1.0 - gj) * (1.0 - gi) * c[0] + (1.0 - gk) * (1.0 - gj) * gi * c[1] +
                        (1.0 - gk) * gj * (1.0 - gi) * c[2] + (1.0 - gk) * gj * gi * c[3] + gk * (1.0 - gj) * (1.0 - gi) * c[4] +
                        gk * (1.0 - gj) * gi * c[5] + gk * gj * (1.0 - gi) * c[6] + gk * gj * gi * c[7];
                if ((q < 2 || q == 8) && sv[q] <= 0.0) sv[q] = c[0];
              }
            }
          }
          if (P->n_t > 0 || (P->interp && P->block_interp)) {
            rho = sv[0]; pgas = sv[1]; entropy = sv[8];
            for (mu = 0; mu < 3; mu++) { uu[mu] = sv[2 + mu]; bb[mu] = sv[5 + mu]; }
          }
          /* sampled values are stored as float (simulation_sampling.cpp:830-839) */
          rho = (double)(float)rho; pgas = (double)(float)pgas; entropy = (double)(float)entropy;
          for (mu = 0; mu < 3; mu++) { uu[mu] = (double)(float)uu[mu]; bb[mu] = (double)(float)bb[mu]; }
          have = 1;
        }
      }
    sampled:
      if (have) {
        /* plasma state (simulation_coefficients.cpp:286-348) */
        double a = P->a, a2 = a * a;
        double rr2 = x * x + y * y + z * z;
        double r2 = 0.5 * (rr2 - a2 + hypot(rr2 - a2, 2.0 * a * z)), r = sqrt(r2);
        double cth = z / r, cth2 = cth * cth, sth2 = 1.0 - cth2, sigma_ks = r2 + a2 * cth2, delta = r2 - 2.0 * r + a2;
        double gs[4][4], gc[4][4];
        memset(gs, 0, sizeof gs); memset(gc, 0, sizeof gc);
        gs[0][0] = -(1.0 - 2.0 * r / sigma_ks); gs[0][1] = gs[1][0] = 2.0 * r / sigma_ks;
        gs[0][3] = gs[3][0] = -2.0 * a * r * sth2 / sigma_ks; gs[1][1] = 1.0 + 2.0 * r / sigma_ks;
        gs[1][3] = gs[3][1] = -(1.0 + 2.0 * r / sigma_ks) * a * sth2; gs[2][2] = sigma_ks;
        gs[3][3] = (r2 + a2 + 2.0 * a2 * r * sth2 / sigma_ks) * sth2;
        gc[0][0] = -(1.0 + 2.0 * r / sigma_ks); gc[0][1] = gc[1][0] = 2.0 * r / sigma_ks; gc[1][1] = delta / sigma_ks;
        gc[1][3] = gc[3][1] = a / sigma_ks; gc[2][2] = 1.0 / sigma_ks; gc[3][3] = 1.0 / (sigma_ks * sth2);
        if (P->coord == 1) {   /* the grid's own coordinates are the Cartesian Kerr-Schild ones (radiation_geometry.cpp:425-457) */
          metric_cov(&geo, x, y, z, gs);
          metric_con(&geo, x, y, z, gc);
        }
        double rho_cgs = rho * P->d_unit, pgas_cgs = pgas * e_unit;
        double n_e = rho_cgs / (P->mu * m_p) / (1.0 + 1.0 / P->ne_ni);
        double uu0 = sqrt(1.0 + gs[1][1] * uu[0] * uu[0] + 2.0 * gs[1][2] * uu[0] * uu[1] + 2.0 * gs[1][3] * uu[0] * uu[2] +
                          gs[2][2] * uu[1] * uu[1] + 2.0 * gs[2][3] * uu[1] * uu[2] + gs[3][3] * uu[2] * uu[2]);
        double lapse = 1.0 / sqrt(-gc[0][0]);
        double us[4], ul[4] = {0, 0, 0, 0}, bs[4], bl_[4] = {0, 0, 0, 0}, b_sq = 0.0;
        us[0] = uu0 / lapse;
        for (mu = 1; mu < 4; mu++) us[mu] = uu[mu - 1] - (-gc[0][mu] / gc[0][0]) * uu0 / lapse;
        for (mu = 0; mu < 4; mu++) for (nu_ = 0; nu_ < 4; nu_++) ul[mu] += gs[mu][nu_] * us[nu_];
        bs[0] = ul[1] * bb[0] + ul[2] * bb[1] + ul[3] * bb[2];
        for (mu = 1; mu < 4; mu++) bs[mu] = (bb[mu - 1] + bs[0] * us[mu]) / us[0];
        for (mu = 0; mu < 4; mu++) for (nu_ = 0; nu_ < 4; nu_++) bl_[mu] += gs[mu][nu_] * bs[nu_];
        for (mu = 0; mu < 4; mu++) b_sq += bl_[mu] * bs[mu];
        double bb_cgs = sqrt(b_sq) * b_unit, sig = b_sq / rho, beta_inv = b_sq / (2.0 * pgas);
        double tti_tte = (P->rat_high + P->rat_low * beta_inv * beta_inv) / (1.0 + beta_inv * beta_inv);
        double kb_tot = P->mu * m_p * pgas_cgs / rho_cgs;
        double kb_te = (1.0 + P->ne_ni) / (tti_tte + P->ne_ni) * kb_tot;
        if (P->use_energy) {
          kb_te = (1.0 + P->ne_ni) * kb_tot / (P->gamma - 1.0);
          kb_te /= tti_tte / (P->gamma_i - 1.0) + P->ne_ni / (P->gamma_e - 1.0);
        }
        double theta_e = kb_te / (m_e * c * c);
        if (P->code_kappa) {
          double mu_e = P->mu * (1.0 + 1.0 / P->ne_ni);
          double rho_e = rho * m_e / (mu_e * m_p);
          double q = cbrt(rho_e * entropy);
          theta_e = 1.0 / 5.0 * (sqrt(1.0 + 25.0 * q * q) - 1.0);
          kb_te = theta_e * m_e * c * c;
        }
        int cut = P->cut_sigma_max >= 0.0 && sig > P->cut_sigma_max;
        {
          double val[7] = {rho_cgs, n_e, pgas_cgs, theta_e, bb_cgs, sig, beta_inv};
          for (i = 0; i < 7; i++)
            if ((P->cut_val_min[i] >= 0.0 && val[i] < P->cut_val_min[i]) || (P->cut_val_max[i] >= 0.0 && val[i] > P->cut_val_max[i])) cut = 1;
        }
        if (!cut) { cv[0] = rho_cgs; cv[1] = n_e; cv[2] = pgas_cgs; cv[3] = theta_e; cv[4] = bb_cgs; cv[5] = sig; cv[6] = beta_inv; }
        if (!cut && !(bb[0] == 0.0 && bb[1] == 0.0 && bb[2] == 0.0)) {
          /* to CKS, tetrad, pitch angle (simulation_coefficients.cpp:397-455) */
          double sth = sqrt(1.0 - cth * cth), ph = atan2(y, x) - atan(a / r), sph = sin(ph), cph = cos(ph);
          double jac[4][4];
          memset(jac, 0, sizeof jac);
          jac[0][0] = 1.0;
          jac[1][1] = sth * cph; jac[1][2] = cth * (r * cph - a * sph); jac[1][3] = sth * (-r * sph - a * cph);
          jac[2][1] = sth * sph; jac[2][2] = cth * (r * sph + a * cph); jac[2][3] = sth * (r * cph - a * sph);
          jac[3][1] = cth; jac[3][2] = -r * sth;
          if (P->coord == 1) {   /* no change of coordinates */
            memset(jac, 0, sizeof jac);
            jac[0][0] = jac[1][1] = jac[2][2] = jac[3][3] = 1.0;
          }
          double ucon[4] = {0, 0, 0, 0}, bcon[4] = {0, 0, 0, 0}, kcon[4] = {0, 0, 0, 0}, ucov[4] = {0, 0, 0, 0}, bcov[4] = {0, 0, 0, 0};
          double gcov[4][4], gcon[4][4], e[4][4];
          for (mu = 0; mu < 4; mu++) for (nu_ = 0; nu_ < 4; nu_++) { ucon[mu] += jac[mu][nu_] * us[nu_]; bcon[mu] += jac[mu][nu_] * bs[nu_]; }
          metric_cov(&geo, x, y, z, gcov);
          metric_con(&geo, x, y, z, gcon);
          for (mu = 0; mu < 4; mu++) for (nu_ = 0; nu_ < 4; nu_++) {
            kcon[mu] += gcon[mu][nu_] * kcov[nu_]; ucov[mu] += gcov[mu][nu_] * ucon[nu_]; bcov[mu] += gcov[mu][nu_] * bcon[nu_];
          }
          tetrad(ucon, ucov, kcon, kcov, bcon, gcov, gcon, e);
          double kt[3] = {0, 0, 0}, bt[3] = {0, 0, 0};
          for (mu = 0; mu < 4; mu++) for (i = 0; i < 3; i++) { kt[i] += e[1 + i][mu] * kcov[mu]; bt[i] += e[1 + i][mu] * bcov[mu]; }
          double ksq = kt[0] * kt[0] + kt[1] * kt[1] + kt[2] * kt[2], bsq = bt[0] * bt[0] + bt[1] * bt[1] + bt[2] * bt[2];
          double kb = kt[0] * bt[0] + kt[1] * bt[1] + kt[2] * bt[2];
          double cos2 = dmin(kb * kb / (ksq * bsq), 1.0), sin_t = sqrt(1.0 - cos2);
          /* thermal emissivity and Kirchhoff absorptivity (simulation_coefficients.cpp:458-524) */
          double nu = 0.0;
          for (mu = 0; mu < 4; mu++) nu -= kcov[mu] * ucon[mu];
          nu *= freq * mom_factor[m];
          double nu_c = qe * bb_cgs / (2.0 * PI * m_e * c), nu_s = 2.0 / 9.0 * nu_c * theta_e * theta_e * sin_t;
          double xx = nu / nu_s, x12 = sqrt(xx), x13 = cbrt(xx), x16 = sqrt(x13);
          double thermal_frac = 1.0 - (P->power_frac + P->kappa_frac);
          if (thermal_frac != 0.0) {
            double coef = thermal_frac * n_e * qe * qe * nu_c / (c * (nu * nu)) * exp(-x13);
            double va = 1.4142135623730951 * PI / 27.0 * sin_t, vb = pow(2.0, 11.0 / 12.0), vc = x12 + vb * x16;
            jv = coef * va * vc * vc;
            double bnu = 2.0 * hpl / (c * c) / expm1(hpl * nu / kb_te);
            av = jv / bnu;
            if (1.0 / (av * av) == INFINITY) av = 0.0;
          }
          if (P->power_frac != 0.0) {
            /* constants (simulation_coefficients.cpp:56-66), emissivity and absorptivity (:559-585) */
            double p = P->power_p;
            double c_a = pow(3.0, p / 2.0) * (p - 1.0), c_b = 2.0 * (p + 1.0);
            double c_c = pow(P->power_gamma_min, 1.0 - p) - pow(P->power_gamma_max, 1.0 - p);
            double c_d = tgamma((3.0 * p - 1.0) / 12.0), c_e = tgamma((3.0 * p + 19.0) / 12.0);
            double c_f = pow(3.0, (p + 1.0) / 2.0) * (p - 1.0) / 4.0;
            double c_g = tgamma((3.0 * p + 2.0) / 12.0), c_h = tgamma((3.0 * p + 22.0) / 12.0);
            double power_jj = c_a / c_b / c_c * c_d * c_e, power_aa = c_f / c_c * c_g * c_h;
            double pj = pow(nu / (nu_c * sin_t), -(p - 1.0) / 2.0);
            jv += P->power_frac * n_e * qe * qe * nu_c / (c * (nu * nu)) * power_jj * sin_t * pj;
            double pa = pow(nu / (nu_c * sin_t), -(p + 2.0) / 2.0);
            av += P->power_frac * n_e * qe * qe / (m_e * c) * power_aa * pa;
          }
          if (P->kappa_frac != 0.0) {
            /* constants (simulation_coefficients.cpp:84-105) */
            double kap = P->kappa, w = P->kappa_w;
            double k_a = 4.0 * PI * tgamma(kap - 4.0 / 3.0), k_b = pow(3.0, 7.0 / 3.0) * tgamma(kap - 2.0);
            double k_c = pow(3.0, (kap - 1.0) / 2.0), k_d = (kap - 2.0) * (kap - 1.0) / 4.0;
            double k_e = tgamma(kap / 4.0 - 1.0 / 3.0), k_f = tgamma(kap / 4.0 + 4.0 / 3.0);
            double k_g = pow(3.0, 1.0 / 6.0) * 10.0 / 41.0, k_h = w * kap;
            double k_i = 2.0 * PI * pow(k_h, kap - 10.0 / 3.0), k_j = (kap - 2.0) * (kap - 1.0) * kap;
            double k_k = 3.0 * kap - 1.0, k_l = tgamma(5.0 / 3.0);
            double k_m = hypergeometric(kap - 1.0 / 3.0, kap + 1.0, kap + 2.0 / 3.0, -k_h);
            double k_n = pow(PI, 1.5) / 3.0, k_o = k_j / (k_h * k_h * k_h);
            double k_p = 2.0 * tgamma(2.0 + kap / 2.0) / (2.0 + kap) - 1.0;
            double jj_low = k_a / k_b, jj_high = k_c * k_d * k_e * k_f, jj_x = 3.0 * pow(kap, -1.5);
            double aa_low = k_g * k_i * k_j / k_k * k_l * k_m, aa_high = k_n * k_o * k_p, aa_x = pow(-1.75 + 1.6 * kap, -0.86);
            /* kappa_aa_high_i is assigned only in polarized runs (:121); in an unpolarized run the member keeps its
             * zero, the high branch is 0, 0^(-x) is infinite and the bridged absorptivity vanishes (:649-653) */
            double aa_high_i = 0.0;
            /* emissivity (:608-622) and absorptivity (:639-653) */
            double nu_kappa = nu_c * w * w * kap * kap * sin_t, xk_ = nu / nu_kappa;
            double e_a = P->kappa_frac * n_e * qe * qe * nu_c / (c * (nu * nu));
            double e_lo = jj_low * e_a * (cbrt(xk_) * sin_t), e_hi = jj_high * e_a * (pow(xk_, -(kap - 2.0) / 2.0) * sin_t);
            jv += pow(pow(e_lo, -jj_x) + pow(e_hi, -jj_x), -1.0 / jj_x);
            double a_a = P->kappa_frac * n_e * qe * qe / (m_e * c);
            double a_lo = aa_low * a_a * pow(xk_, -2.0 / 3.0), a_hi = aa_high * a_a * pow(xk_, -(1.0 + kap) / 2.0) * aa_high_i;
            av += pow(pow(a_lo, -aa_x) + pow(a_hi, -aa_x), -1.0 / aa_x);
          }
        }
      }
      I = transfer_step(I, jv, av, dl_cgs);
      double d_length = 0.0;   /* proper length of the step (unpolarized.cpp:117-130, rendering.cpp:86-99) */
      if (aux || render) {
        double gcov[4][4], gcon[4][4], t[4] = {0, 0, 0, 0}, sq = 0.0;
        int aa, bb_;
        metric_cov(&geo, x, y, z, gcov);
        metric_con(&geo, x, y, z, gcon);
        for (aa = 1; aa < 4; aa++)
          for (mu = 0; mu < 4; mu++) t[aa] += (gcon[aa][mu] - gcon[0][aa] * gcon[0][mu] / gcon[0][0]) * kcov[mu];
        for (aa = 1; aa < 4; aa++)
          for (bb_ = 1; bb_ < 4; bb_++) sq += gcov[aa][bb_] * t[aa] * t[bb_];
        d_length = sqrt(sq) * len[o] * P->x_unit;
      }
      if (render) {   /* rendering.cpp:100-170 */
        int f, q;
        for (f = 0; f < n_feat; f++) {
          const orc_feature *ft = feat + f;
          double *px = render + ((size_t)ft->image * 3) * n_rays + m, prev = prev_cv[ft->quantity], cur = cv[ft->quantity];
          int crossed = 0;
          if (ft->type == 0 && cur >= ft->min && cur <= ft->max) {
            double dt = d_length / ft->tau_scale;
            for (q = 0; q < 3; q++)
              px[(size_t)q * n_rays] = dt <= 100.0 ? exp(-dt) * (px[(size_t)q * n_rays] + ft->xyz[q] * expm1(dt)) : ft->xyz[q];
          }
          if ((ft->type == 1 || ft->type == 2) && prev < ft->thresh && cur >= ft->thresh) crossed = 1;
          if ((ft->type == 1 || ft->type == 3) && prev > ft->thresh && cur <= ft->thresh) crossed = 1;
          if (crossed)
            for (q = 0; q < 3; q++) px[(size_t)q * n_rays] = (1.0 - ft->opacity) * px[(size_t)q * n_rays] + ft->opacity * ft->xyz[q];
        }
        for (q = 0; q < 7; q++) prev_cv[q] = cv[q];
      }
      if (aux) {
        double t_cgs = pos[4 * o] * (P->x_unit / c);
        double dtau = av * dl_cgs, e_neg = exp(-dtau), e_m1 = expm1(dtau);
        int aa, now;
        a_time = t_cgs < a_time ? t_cgs : a_time;
        a_length += d_length;
        a_lambda += dl_cgs;
        a_emission += jv * dl_cgs;
        a_tau += dtau;
        if (!isnan(cv[0]))
          for (aa = 0; aa < 7; aa++) {
            a_lave[aa] += cv[aa] * dl_cgs;
            a_eave[aa] += cv[aa] * jv * dl_cgs;
            a_tint[aa] = dtau <= 100.0 ? e_neg * (a_tint[aa] + cv[aa] * e_m1) : cv[aa];
          }
        now = camera_x[1] * x + camera_x[2] * y + camera_x[3] * z > 0.0;
        if (now != a_sign) a_cross++;
        a_sign = now;
      }
    }
    image[m] = I * (freq * freq * freq);
    if (aux) {
      size_t N = (size_t)n_rays;
      aux[0 * N + m] = a_time; aux[1 * N + m] = a_length; aux[2 * N + m] = a_lambda; aux[3 * N + m] = a_emission;
      aux[4 * N + m] = a_tau; aux[5 * N + m] = (double)a_cross;
      for (i = 0; i < 7; i++) {
        aux[(6 + i) * N + m] = a_lave[i] / a_lambda;
        aux[(13 + i) * N + m] = a_eave[i] / a_emission;
        aux[(20 + i) * N + m] = a_tint[i];
      }
    }
  }
}
