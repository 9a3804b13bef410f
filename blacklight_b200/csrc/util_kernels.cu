// Small supporting kernels: grid re-layout at upload, parity-tap unpacking, adaptive refinement
// decision (reference radiation_adaptive.cpp:19-312) and an FP64 FMA peak probe for rooflines.
#include "../../include/blacklight_b200.h"
#include "rad_types.cuh"
#include "glibc_math.cuh"

namespace {

// (var, cell) planes -> one 32-byte record per cell (see GridDev in rad_types.cuh)
__global__ void relayout_grid_kernel(const float *__restrict__ prim, const int *__restrict__ var_index,
                                     size_t cells, float4 *__restrict__ out, float *__restrict__ kappa_out) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = prim[(size_t)var_index[q] * cells + c];
  out[2 * c] = make_float4(v[0], v[1], v[2], v[3]);
  out[2 * c + 1] = make_float4(v[4], v[5], v[6], v[7]);
  if (kappa_out) kappa_out[c] = prim[(size_t)var_index[8] * cells + c];
}

// Step buffer (64-byte records, tracing order) -> the reference's sample_pos/dir/len host layout
// (N,S,4)/(N,S) in source->camera order with len > 0 (geodesics.cpp:808-849); tail zero-filled.
__global__ void unpack_samples_kernel(StepBuffer sb, const int32_t *__restrict__ num, const double *__restrict__ cam_dir,
                                      int64_t rays, int S, double *__restrict__ pos, double *__restrict__ dir,
                                      double *__restrict__ len) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_out = blockIdx.y;
  if (m >= rays) return;
  int cnt = num[m];
  size_t o = (size_t)m * S + n_out;
  if (n_out < cnt) {
    int n = cnt - 1 - n_out;
    const double *src = sb.buf + sb.at(n, m);
    if (pos) for (int c = 0; c < 4; c++) pos[4 * o + c] = src[c];
    if (dir) {
      dir[4 * o] = cam_dir[4 * m];  // p_t is conserved and not stored per sample
      for (int c = 1; c < 4; c++) dir[4 * o + c] = src[3 + c];
    }
    if (len) len[o] = -src[7];
  } else {
    if (pos) for (int c = 0; c < 4; c++) pos[4 * o + c] = 0.0;
    if (dir) for (int c = 0; c < 4; c++) dir[4 * o + c] = 0.0;
    if (len) len[o] = 0.0;
  }
}

// The reference's sample_pos/dir/len host layout (source->camera order, len > 0) -> step-buffer records
// (tracing order, len < 0): the inverse of unpack_samples_kernel, for geodesics computed elsewhere
// (checkpoint_geodesic_load).  pos/dir/len hold rays [ray0, ray0 + count) only.
__global__ void pack_samples_kernel(StepBuffer sb, const int32_t *__restrict__ num, int64_t ray0, int64_t count, int S,
                                    const double *__restrict__ pos, const double *__restrict__ dir,
                                    const double *__restrict__ len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int n_in = blockIdx.y;
  if (i >= count) return;
  int64_t m = ray0 + i;
  int cnt = num[m];
  if (n_in >= cnt) return;
  size_t o = (size_t)i * S + n_in;
  double2 *dst = reinterpret_cast<double2 *>(sb.buf + sb.at(cnt - 1 - n_in, m));
  dst[0] = make_double2(pos[4 * o + 0], pos[4 * o + 1]);
  dst[1] = make_double2(pos[4 * o + 2], pos[4 * o + 3]);
  dst[2] = make_double2(dir[4 * o + 1], dir[4 * o + 2]);
  dst[3] = make_double2(dir[4 * o + 3], -len[o]);
}

// One CTA per refinement block.  Five exceedance-fraction tests on Stokes I of the chosen frequency.
__global__ void refine_kernel(const double *__restrict__ image, int64_t stride, int level,
                              const int32_t *__restrict__ block_locs, int64_t num_blocks,
                              const bl_params *__restrict__ pp, uint8_t *__restrict__ flags) {
  const bl_params &P = *pp;
  const int bs = P.adaptive_block_size;
  const int64_t block = blockIdx.x;
  __shared__ int examined[5], exceeded[5];
  __shared__ int forced;
  if (threadIdx.x < 5) examined[threadIdx.x] = exceeded[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    forced = 0;
    if (P.adaptive_num_regions > 0) {
      int linear_blocks = P.camera_resolution / bs;
      for (int n = 1; n <= level; n++) linear_blocks *= 2;
      double y = ((block_locs[2 * block] + 0.5) / linear_blocks - 0.5) * P.camera_width;
      double x = ((block_locs[2 * block + 1] + 0.5) / linear_blocks - 0.5) * P.camera_width;
      for (int r = 0; r < P.adaptive_num_regions; r++)
        if (level < P.adaptive_region_levels[r] && x > P.adaptive_region_x_min[r] && x < P.adaptive_region_x_max[r] &&
            y > P.adaptive_region_y_min[r] && y < P.adaptive_region_y_max[r]) {
          forced = 1;
          break;
        }
    }
  }
  __syncthreads();
  if (forced) {
    if (threadIdx.x == 0) flags[block] = 1;
    return;
  }
  bool pol = P.model_type == BL_MODEL_SIMULATION && P.image_polarization;
  const double *plane = image + (size_t)(P.adaptive_frequency_num * (pol ? 4 : 1)) * stride;
  // pixel (row i, column j) of this block
  int root_blocks = P.camera_resolution / bs;
  int64_t row0 = 0, col0 = 0;
  const bool raster = level == 0 && !P.level0_block_major;
  if (raster) {
    row0 = block / root_blocks * bs;
    col0 = block % root_blocks * bs;
  }
  auto at = [&](int i, int j) -> double {
    if (raster) return plane[(row0 + i) * (int64_t)P.camera_resolution + col0 + j];
    return plane[block * (int64_t)bs * bs + (int64_t)i * bs + j];
  };
  for (int t = threadIdx.x; t < bs * bs; t += blockDim.x) {
    int i = t / bs, j = t % bs;
    double c = at(i, j);
    if (P.adaptive_val_frac >= 0.0) {
      double q = fabs(c);
      if (isfinite(q)) { atomicAdd(&examined[0], 1); if (q > P.adaptive_val_cut) atomicAdd(&exceeded[0], 1); }
    }
    double xm = j > 0 ? at(i, j - 1) : 0.0, xp = j < bs - 1 ? at(i, j + 1) : 0.0;
    double ym = i > 0 ? at(i - 1, j) : 0.0, yp = i < bs - 1 ? at(i + 1, j) : 0.0;
    if (P.adaptive_abs_grad_frac >= 0.0) {
      double qx = j == 0 ? xp - c : (j == bs - 1 ? c - xm : 0.5 * (xp - xm));
      double qy = i == 0 ? yp - c : (i == bs - 1 ? c - ym : 0.5 * (yp - ym));
      double q = hypot(qx, qy);
      if (isfinite(q)) { atomicAdd(&examined[1], 1); if (q > P.adaptive_abs_grad_cut) atomicAdd(&exceeded[1], 1); }
    }
    if (P.adaptive_rel_grad_frac >= 0.0) {
      double qx = j == 0 ? 2.0 * (xp - c) / (c + xp)
                         : (j == bs - 1 ? 2.0 * (c - xm) / (xm + c) : 2.0 * (xp - xm) / (xm + 2.0 * c + xp));
      double qy = i == 0 ? 2.0 * (yp - c) / (c + yp)
                         : (i == bs - 1 ? 2.0 * (c - ym) / (ym + c) : 2.0 * (yp - ym) / (ym + 2.0 * c + yp));
      double q = hypot(qx, qy);
      if (isfinite(q)) { atomicAdd(&examined[2], 1); if (q > P.adaptive_rel_grad_cut) atomicAdd(&exceeded[2], 1); }
    }
    if (i >= 1 && i < bs - 1 && j >= 1 && j < bs - 1) {
      if (P.adaptive_abs_lapl_frac >= 0.0) {
        double qx = xm - 2.0 * c + xp, qy = ym - 2.0 * c + yp;
        double q = fabs(qx + qy);
        if (isfinite(q)) { atomicAdd(&examined[3], 1); if (q > P.adaptive_abs_lapl_cut) atomicAdd(&exceeded[3], 1); }
      }
      if (P.adaptive_rel_lapl_frac >= 0.0) {
        double qx = 4.0 * (xm - 2.0 * c + xp) / (xm + 2.0 * c + xp);
        double qy = 4.0 * (ym - 2.0 * c + yp) / (ym + 2.0 * c + yp);
        double q = fabs(qx + qy);
        if (isfinite(q)) { atomicAdd(&examined[4], 1); if (q > P.adaptive_rel_lapl_cut) atomicAdd(&exceeded[4], 1); }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double fr[5] = {P.adaptive_val_frac, P.adaptive_abs_grad_frac, P.adaptive_rel_grad_frac,
                          P.adaptive_abs_lapl_frac, P.adaptive_rel_lapl_frac};
    bool refine = false;
    for (int t = 0; t < 5; t++)
      if (fr[t] >= 0.0) {
        double frac = (double)exceeded[t] / (double)examined[t];  // 0/0 = NaN never exceeds (reference quirk)
        if (frac > fr[t]) refine = true;
      }
    flags[block] = refine ? 1 : 0;
  }
}

// 16 independent FMA chains per thread
__global__ void fp64_peak_kernel(double *out, int iters) {
  double acc[16];
  double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9 * (blockIdx.x + 1);
#pragma unroll
  for (int q = 0; q < 16; q++) acc[q] = q * 0.125;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < 16; q++) acc[q] = fma(acc[q], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 16; q++) s += acc[q];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" cudaError_t bl_launch_relayout_grid(const float *prim, int n_var, const int *var_index, size_t cells,
                                               float4 *out, float *kappa_out, cudaStream_t stream) {
  (void)n_var;
  unsigned grid = (unsigned)((cells + 255) / 256);
  relayout_grid_kernel<<<grid, 256, 0, stream>>>(prim, var_index, cells, out, kappa_out);
  return cudaGetLastError();
}

// Self-test of blmath::div_by / sqrt_rn (branch-free sequences) against the hardware IEEE operations: every thread
// draws operand pairs from a xorshift stream -- uniformly random significands over 60 binades, plus the hard
// cases for rounding (numerators RN(q b) +- 1 ulp, whose quotients sit next to representable numbers and
// midpoints; denominators with all-ones / all-zeros significand tails) -- and counts differing bit patterns.
__global__ void division_selftest_kernel(unsigned long long seed, int iters, unsigned long long *mismatches) {
  unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
  auto next = [&]() {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return s;
  };
  auto rnd = [&](int kind) {
    unsigned long long u = next();
    unsigned long long mant = u & 0xFFFFFFFFFFFFFull;
    if (kind == 1) mant |= 0xFFFFFFFFFF000ull;          // long run of ones
    if (kind == 2) mant &= 0x0000000000FFFull;          // long run of zeros
    if (kind == 3) mant = (mant & ~0xFFFull) | 0xFFFull; // ones at the bottom
    long long e = 1023 + (long long)((u >> 52) % 61) - 30;
    unsigned long long sign = (u >> 63) << 63;
    return __longlong_as_double((long long)(sign | ((unsigned long long)e << 52) | mant));
  };
  unsigned long long bad = 0;
  for (int it = 0; it < iters; it++) {
    int kb = (int)(next() & 3), ka = (int)(next() & 3);
    double b = rnd(kb);
    double a = rnd(ka);
    if ((it & 3) == 1) {
      // quotient next to a representable number or a midpoint
      double q = rnd(0);
      a = __dmul_rn(q, b);
      long long bits = __double_as_longlong(a) + (long long)(next() % 5) - 2;
      a = __longlong_as_double(bits);
    }
    blmath::Recip d = blmath::recip_of(b);
    double got = blmath::div_by(a, d);
    double want = __ddiv_rn(a, b);
    if (__double_as_longlong(got) != __double_as_longlong(want)) bad++;
    double x = fabs(a);
    if (__double_as_longlong(blmath::sqrt_rn(x)) != __double_as_longlong(__dsqrt_rn(x))) bad++;
  }
  if (bad) atomicAdd(mismatches, bad);
}

extern "C" cudaError_t bl_launch_division_selftest(unsigned long long seed, int blocks, int iters,
                                                   unsigned long long *mismatches, cudaStream_t stream) {
  division_selftest_kernel<<<blocks, 256, 0, stream>>>(seed, iters, mismatches);
  return cudaGetLastError();
}

extern "C" cudaError_t bl_launch_unpack_samples(const StepBuffer *sb, const int32_t *num, const double *cam_dir,
                                                int64_t rays, int S, double *pos, double *dir, double *len,
                                                cudaStream_t stream) {
  if (rays <= 0 || S <= 0) return cudaSuccess;
  dim3 grid((unsigned)((rays + 127) / 128), (unsigned)S);
  unpack_samples_kernel<<<grid, 128, 0, stream>>>(*sb, num, cam_dir, rays, S, pos, dir, len);
  return cudaGetLastError();
}

extern "C" cudaError_t bl_launch_pack_samples(const StepBuffer *sb, const int32_t *num, int64_t ray0, int64_t count, int S,
                                              const double *pos, const double *dir, const double *len,
                                              cudaStream_t stream) {
  if (count <= 0 || S <= 0) return cudaSuccess;
  dim3 grid((unsigned)((count + 127) / 128), (unsigned)S);
  pack_samples_kernel<<<grid, 128, 0, stream>>>(*sb, num, ray0, count, S, pos, dir, len);
  return cudaGetLastError();
}

extern "C" cudaError_t bl_launch_refine(const double *image, int64_t stride, int level, const int32_t *block_locs,
                                        int64_t num_blocks, const bl_params *params_dev, uint8_t *flags,
                                        cudaStream_t stream) {
  refine_kernel<<<(unsigned)num_blocks, 128, 0, stream>>>(image, stride, level, block_locs, num_blocks, params_dev, flags);
  return cudaGetLastError();
}

extern "C" cudaError_t bl_launch_fp64_peak(double *out, int blocks, int iters, cudaStream_t stream) {
  fp64_peak_kernel<<<blocks, 256, 0, stream>>>(out, iters);
  return cudaGetLastError();
}
