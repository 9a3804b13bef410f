// Device-side layouts shared by the kernels and the C-ABI host glue.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

// Step buffer of one wave ("tile") of rays: one 64-byte record per stored sample,
//   rec[n * rays + m] = (t, x, y, z | p_x, p_y, p_z, len)      n = sample index in tracing order, m = ray
// i.e. two full 32-byte sectors per sample.  The geodesic kernel (lanes at unrelated n) writes a record
// with four 16-byte stores that fill both sectors completely; the radiation kernels (a warp walks 32
// adjacent rays in lock step) read 2 KB contiguous rows.  p_t is not stored: it is conserved along the ray
// (dp_t/dlambda = 0 exactly, geodesics.cpp:867-893) and equals the camera array's cam_dir[m][0].
// len is the affine step (negative while tracing backwards; consumers use -len, geodesics.cpp:840).
// Samples are stored in tracing order (n = 0 at the camera); the radiation kernels walk n downwards,
// which is the reference's source->camera order (geodesics.cpp:808-849) without the reversal copy.
struct StepBuffer {
  double *buf;
  int64_t rays;  // rays in this wave
  int32_t cap;   // sample capacity per ray (= ray_max_steps)
  static constexpr int kRecord = 8;  // doubles per sample
  __host__ __device__ size_t at(int n, int64_t m) const { return ((size_t)n * (size_t)rays + (size_t)m) * kRecord; }
};

struct GeoCounters {
  unsigned long long next_ray;
  unsigned long long attempts;
  unsigned long long accepted;
  unsigned long long bad;
  unsigned long long samples;
  int max_samples;
  unsigned int b_max_bits;   // largest impact parameter of the wave's rays (float bits), for the integrator's queue order
};

struct GeoArgs {
  const double *cam_pos;  // (rays,4) for this wave
  const double *cam_dir;  // (rays,4) covariant
  int64_t rays;
  double a, camera_r, r_terminate, r_horizon, ray_step, tol_abs, tol_rel;
  int32_t max_steps, max_retries;
  StepBuffer sb;
  int32_t *sample_num;   // (rays)
  uint8_t *sample_flags; // (rays)
  GeoCounters *counters;
  const int32_t *order;  // DP: the ray queue hands out order[0], order[1], ... (longest rays first, ray_order.cu); nullptr: 0, 1, ...
};

// What InitializeCamera leaves (reference camera.cpp:53-380; host build_camera_frame), by value to camera_pixels_kernel.
struct CameraDev {
  int32_t type, normalization, flat, pad;   // BL_CAMERA_*, BL_NORM_*, ray_flat
  double a, width, r;
  double x[4], u_con[4], u_cov[4], norm_con[4], norm_con_c[4], hor_con_c[4], vert_con_c[4];
};

#define BL_CUDA_CHECK(call)                                                        \
  do {                                                                             \
    cudaError_t err__ = (call);                                                    \
    if (err__ != cudaSuccess) return bl_fail_cuda(ctx, err__, #call, __FILE__, __LINE__); \
  } while (0)
