// Device-side layouts shared by the kernels and the C-ABI host glue.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

// Step buffer of one wave ("tile") of rays, structure-of-arrays with the ray index fastest:
//   comp c of sample n of ray m lives at buf[(c * cap + n) * rays + m],  c in [0,9):
//   0..3 = x^mu (t,x,y,z), 4..7 = covariant momentum p_mu, 8 = affine step length (negative while
//   tracing backwards; consumers use -len, see reference geodesics.cpp:840).
// Samples are stored in tracing order (n = 0 at the camera); the radiation kernels walk n downwards,
// which is the reference's source->camera order (geodesics.cpp:808-849) without the reversal copy.
struct StepBuffer {
  double *buf;
  int64_t rays;  // rays in this wave (stride between consecutive samples)
  int32_t cap;   // sample capacity per ray (= ray_max_steps)
  __host__ __device__ size_t at(int c, int n, int64_t m) const {
    return ((size_t)c * cap + n) * (size_t)rays + m;
  }
};

struct GeoCounters {
  unsigned long long next_ray;
  unsigned long long attempts;
  unsigned long long accepted;
  unsigned long long bad;
  unsigned long long samples;
  int max_samples;
  int pad;
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
};

#define BL_CUDA_CHECK(call)                                                        \
  do {                                                                             \
    cudaError_t err__ = (call);                                                    \
    if (err__ != cudaSuccess) return bl_fail_cuda(ctx, err__, #call, __FILE__, __LINE__); \
  } while (0)
