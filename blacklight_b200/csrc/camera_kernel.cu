// Camera pixels on the device: initial position, covariant momentum and frequency factor of every ray of a level,
// generated straight into the level's camera arrays in HBM (reference camera.cpp:390-413 root raster, :471-499 refined
// blocks, :528-671 SetPixelPlane / SetPixelPinhole) -- the 72 bytes per ray of the host path never cross PCIe.
//
// The arrays must be bit for bit what the reference's host code computes: the geodesic integrator's flags, counts and
// samples are discrete functions of them.  The per-pixel code uses only IEEE operations (+ - * / sqrt) and two libm
// calls; this translation unit is compiled with -fmad=false, the expressions keep the association of the host
// version (csrc/host/camera.cpp: camera_pixel, ks_metric, itself bit-identical to the reference's checkpoint), and
//   std::hypot(x, y)     -> blmath::hypot_glibc      (glibc 2.39 e_hypot.c kernel, glibc_math.cuh)
//   std::hypot(x, y, z)  -> blmath::hypot3_libstdcxx (libstdc++ <cmath> __hypot3)
// tests/test_gpu_parity.py::test_device_camera_is_bitwise_the_host_camera compares every array element.
#include "device_types.cuh"
#include "glibc_math.cuh"

namespace {

// Cartesian Kerr-Schild covariant metric (host ks_metric; reference geodesic_geometry.cpp:38-96)
__device__ __forceinline__ void ks_metric_cov(double a, bool flat, double x, double y, double z, double gcov[4][4]) {
  if (flat) {
    for (int m = 0; m < 4; m++)
      for (int n = 0; n < 4; n++) gcov[m][n] = m == n ? (m == 0 ? -1.0 : 1.0) : 0.0;
    return;
  }
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double r2 = 0.5 * (rr2 - a2 + blmath::hypot_glibc(rr2 - a2, 2.0 * a * z));
  double r = sqrt(r2);
  double f = 2.0 * r2 * r / (r2 * r2 + a2 * z * z);
  double lo[4] = {1.0, (r * x + a * y) / (r2 + a2), (r * y - a * x) / (r2 + a2), z / r};
  for (int m = 0; m < 4; m++)
    for (int n = 0; n < 4; n++) {
      double v = f * lo[m] * lo[n];
      gcov[m][n] = m == n ? (m == 0 ? v - 1.0 : v + 1.0) : v;
    }
}

// host camera_pixel (reference camera.cpp:528-671)
__device__ __forceinline__ void camera_pixel(const CameraDev &c, double u_ind, double v_ind, double pos[4], double dir[4],
                                             double *factor) {
  double u = u_ind * 1.0 * c.width;
  double v = v_ind * 1.0 * c.width;
  double p[4];
  if (c.type == 0) {
    double dc[4];
    for (int m = 0; m < 4; m++) dc[m] = u * c.hor_con_c[m] + v * c.vert_con_c[m];
    double dt = c.u_con[0] * dc[0] - (c.u_cov[1] * dc[1] + c.u_cov[2] * dc[2] + c.u_cov[3] * dc[3]) / c.u_cov[0];
    pos[0] = c.x[0] + dt;
    for (int i = 1; i < 4; i++) pos[i] = c.x[i] + (dc[i] + c.u_con[i] * dc[0]);
    p[1] = c.norm_con[1];
    p[2] = c.norm_con[2];
    p[3] = c.norm_con[3];
  } else {
    for (int m = 0; m < 4; m++) pos[m] = c.x[m];
    double normalization = blmath::hypot3_libstdcxx(u, v, c.r);
    double frac_norm = c.r / normalization;
    double frac_hor = -u / normalization;
    double frac_vert = -v / normalization;
    for (int i = 1; i < 4; i++) {
      double dc = frac_norm * c.norm_con_c[i] + frac_hor * c.hor_con_c[i] + frac_vert * c.vert_con_c[i];
      p[i] = dc + c.u_con[i] * c.norm_con_c[0];
    }
  }
  // p^t from the null condition g_{mu nu} p^mu p^nu = 0 (camera.cpp:553-566)
  double gcov[4][4];
  ks_metric_cov(c.a, c.flat != 0, pos[1], pos[2], pos[3], gcov);
  double qa = gcov[0][0];
  double qb = 0.0;
  for (int i = 1; i < 4; i++) qb += 2.0 * gcov[0][i] * p[i];
  double qc = 0.0;
  for (int i = 1; i < 4; i++)
    for (int j = 1; j < 4; j++) qc += gcov[i][j] * p[i] * p[j];
  double disc = qb * qb - 4.0 * qa * qc;
  double qd = sqrt(disc < 0.0 ? 0.0 : disc);   // std::max(disc, 0.0)
  p[0] = qa == 0.0 ? -qc / (2.0 * qb) : (qb < 0.0 ? 2.0 * qc / (qd - qb) : -(qb + qd) / (2.0 * qa));
  for (int m = 0; m < 4; m++) {
    dir[m] = 0.0;
    for (int n = 0; n < 4; n++) dir[m] += gcov[m][n] * p[n];
  }
  double nu_local = 0.0;
  if (c.normalization == 0)
    for (int m = 0; m < 4; m++) nu_local -= dir[m] * c.u_con[m];
  else
    nu_local = -dir[0];
  *factor = 1.0 / nu_local;
}

constexpr int kCamBlock = 128;

// kind 0: `units` are image rows of a raster of side eff_res (nullptr: rows 0, 1, ...); pixel o = unit * eff_res + col.
// kind 1: `units` are (v, u) block locations at effective resolution eff_res; pixel o = unit * bs^2 + row * bs + col.
__global__ void __launch_bounds__(kCamBlock) camera_pixels_kernel(CameraDev c, int kind, const int32_t *units, int eff_res,
                                                                  int block_size, int64_t num_pixels, double *cam_pos,
                                                                  double *cam_dir, double *mom_factor) {
  const int64_t o = (int64_t)blockIdx.x * kCamBlock + threadIdx.x;
  if (o >= num_pixels) return;
  int row, col;
  if (kind == 0) {
    const int64_t r = o / eff_res;
    col = (int)(o - r * eff_res);
    row = units ? units[r] : (int)r;
  } else {
    const int bpix = block_size * block_size;
    const int64_t b = o / bpix;
    const int mm = (int)(o - b * bpix);
    row = mm / block_size + units[2 * b] * block_size;
    col = mm % block_size + units[2 * b + 1] * block_size;
  }
  // (col - res / 2.0 + 0.5) / res with res an int, as the host writes it
  const double u_ind = (col - eff_res / 2.0 + 0.5) / eff_res;
  const double v_ind = (row - eff_res / 2.0 + 0.5) / eff_res;
  double pos[4], dir[4], factor;
  camera_pixel(c, u_ind, v_ind, pos, dir, &factor);
  double2 *dp = reinterpret_cast<double2 *>(cam_pos + 4 * o), *dd = reinterpret_cast<double2 *>(cam_dir + 4 * o);
  dp[0] = make_double2(pos[0], pos[1]);
  dp[1] = make_double2(pos[2], pos[3]);
  dd[0] = make_double2(dir[0], dir[1]);
  dd[1] = make_double2(dir[2], dir[3]);
  mom_factor[o] = factor;
}

}  // namespace

extern "C" cudaError_t bl_launch_camera_pixels(const CameraDev *cam, int kind, const int32_t *units, int eff_res, int block_size,
                                               int64_t num_pixels, double *cam_pos, double *cam_dir, double *mom_factor,
                                               cudaStream_t stream) {
  if (num_pixels <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((num_pixels + kCamBlock - 1) / kCamBlock);
  camera_pixels_kernel<<<grid, kCamBlock, 0, stream>>>(*cam, kind, units, eff_res, block_size, num_pixels, cam_pos, cam_dir,
                                                       mom_factor);
  return cudaGetLastError();
}
