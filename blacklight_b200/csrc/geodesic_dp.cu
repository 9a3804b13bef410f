// Backward null-geodesic integration, Dormand-Prince RK5(4)7M with dense output -- one FP64
// sm_100a kernel: one ray per thread, persistent warps that refill finished lanes from a global
// ray queue (warp ballot + one aggregated atomic), stage derivatives in shared memory, step buffer
// written as 64-byte records (device_types.cuh).  Truncation (reference geodesics.cpp:327-349), momentum
// renormalisation of stored samples (:352-371), the max/bad-ray reductions (:374-382) are fused
// into the same kernel; the reversal copy (:808-849) is eliminated (consumers walk backwards).
//
// State machine follows reference src/geodesic_integrator/geodesics.cpp:109-324 and reproduces its
// floating-point dataflow exactly (compiled with -fmad=false; hypot/pow from glibc_math.cuh), so that
// sample_flags / sample_num and the stored samples are bit-identical to the reference's.
#include "device_types.cuh"
#include "ks_exact.cuh"

namespace {

#ifndef BL_GEO_BLOCK
#define BL_GEO_BLOCK 128
#endif
constexpr int kBlock = BL_GEO_BLOCK;
constexpr int kComp = 7;  // t, x, y, z, p_x, p_y, p_z in shared memory (dp_t/dlambda is identically zero);
                          // the proper-distance rate ds/dlambda of a stage is consumed at once (registers)

// Butcher tableau of RK5(4)7M (Dormand & Prince 1980) and the dense-output weights of Shampine (1986)
__constant__ double c_a[7][6] = {
    {0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {1.0 / 5.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {3.0 / 40.0, 9.0 / 40.0, 0.0, 0.0, 0.0, 0.0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0.0, 0.0, 0.0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0.0, 0.0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0.0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
__constant__ double c_b5[7] = {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0,
                               -2187.0 / 6784.0, 11.0 / 84.0, 0.0};
__constant__ double c_b4[7] = {5179.0 / 57600.0, 0.0, 7571.0 / 16695.0, 393.0 / 640.0,
                               -92097.0 / 339200.0, 187.0 / 2100.0, 1.0 / 40.0};
__constant__ double c_b4m[7] = {6025192743.0 / 30085553152.0, 0.0, 51252292925.0 / 65400821598.0,
                                -2691868925.0 / 45128329728.0, 187940372067.0 / 1594534317056.0,
                                -1776094331.0 / 19743644256.0, 11237099.0 / 235043384.0};
__constant__ double c_d[7] = {-12715105075.0 / 11282082432.0, 0.0, 87487479700.0 / 32700410799.0,
                              -10690763975.0 / 1880347072.0, 701980252875.0 / 199316789632.0,
                              -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0};

struct Ray {
  double y[9];    // x^mu, p_mu, s at the start of the current step
  double y5[9];   // 5th-order solution at the end of the last attempted step
  double h_new, r_new;
  double r_prev_sample;
  int64_t m;      // ray slot in the wave
  int n;          // samples stored so far
  int num_retry;
  int trunc;      // first truncated sample index, -1 if none
  double ds0;     // ds/dlambda of stage 0 (first-same-as-last)
  bool prev_fail, flag, need_k0;
};

template <bool flat>
__device__ __forceinline__ double eval_rhs(const GeoArgs &g, const double pos[3], const double p[4],
                                           double *ks, int q) {
  double dx[4], dp[3], ds;
  ksx::rhs<flat>(g.a, pos[0], pos[1], pos[2], p, dx, dp, ds);
  double *kq = ks + (size_t)q * kComp * kBlock;
  kq[0 * kBlock] = dx[0];
  kq[1 * kBlock] = dx[1];
  kq[2 * kBlock] = dx[2];
  kq[3 * kBlock] = dx[3];
  kq[4 * kBlock] = dp[0];
  kq[5 * kBlock] = dp[1];
  kq[6 * kBlock] = dp[2];
  return ds;
}

// Store one sample: truncation test on its radius, renormalise its spatial momentum, write SoA.
template <bool flat>
__device__ __forceinline__ void store_sample(const GeoArgs &g, Ray &ray, int idx, const double v[8],
                                             double len) {
  if (ray.trunc >= 0) return;  // beyond the truncation point nothing is ever read
  double rs = ksx::radius(g.a, v[1], v[2], v[3]);
  if (idx >= 1) {
    bool outer = rs > g.camera_r && rs > ray.r_prev_sample;
    bool inner = rs < g.r_terminate;
    if (outer || inner) {
      ray.trunc = idx;
      return;
    }
  }
  ray.r_prev_sample = rs;
  double p[4] = {v[4], v[5], v[6], v[7]};
  ksx::renormalize_momentum<flat>(g.a, v[1], v[2], v[3], p);
  // one 64-byte record = two full sectors, written with four 16-byte streaming stores
  double2 *dst = reinterpret_cast<double2 *>(g.sb.buf + g.sb.at(idx, ray.m));
  __stcs(dst + 0, make_double2(v[0], v[1]));
  __stcs(dst + 1, make_double2(v[2], v[3]));
  __stcs(dst + 2, make_double2(p[1], p[2]));
  __stcs(dst + 3, make_double2(p[3], len));
}

template <bool flat, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) geodesic_dp_kernel(GeoArgs g) {
  extern __shared__ double smem[];
  double *ks = smem + threadIdx.x;  // k[q][p] at ks[(q*kComp + p) * kBlock]
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;

  Ray ray;
  bool active = false;
  bool exhausted = false;
  double ds_last = 0.0;  // ds/dlambda of stage 6 of the last attempt
  unsigned long long n_attempts = 0, n_accepted = 0;

  for (;;) {
    // ---- refill idle lanes from the ray queue (one atomic per warp) ----
    unsigned idle = __ballot_sync(full, !active);
    if (idle && !exhausted) {
      int leader = __ffs(idle) - 1;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(&g.counters->next_ray, (unsigned long long)__popc(idle));
      base = __shfl_sync(full, base, leader);
      if (!active) {
        unsigned long long idx = base + __popc(idle & ((1u << lane) - 1u));
        if (idx < (unsigned long long)g.rays) {
          ray.m = g.order ? (int64_t)g.order[idx] : (int64_t)idx;
          const double *cp = g.cam_pos + 4 * ray.m;
          const double *cd = g.cam_dir + 4 * ray.m;
          for (int c = 0; c < 4; c++) {
            ray.y[c] = cp[c];
            ray.y[4 + c] = cd[c];
          }
          ray.y[8] = 0.0;
          for (int c = 0; c < 9; c++) ray.y5[c] = ray.y[c];
          ray.r_new = ksx::radius(g.a, ray.y[1], ray.y[2], ray.y[3]);
          ray.h_new = -g.ray_step * ray.r_new;
          ray.n = 0;
          ray.num_retry = 0;
          ray.trunc = -1;
          ray.r_prev_sample = 0.0;
          ray.prev_fail = false;
          ray.flag = false;
          ray.need_k0 = true;
          active = true;
        }
      }
      if (base + __popc(idle) >= (unsigned long long)g.rays) exhausted = true;
    }
#if !defined(BL_GEO_NOSYNC)
    // keep the CTA's warps in the same region of the (instruction-cache-sized) loop body
    if (!__syncthreads_or(active ? 1 : 0)) break;
#else
    if (!__any_sync(full, active)) break;
#endif
    if (!active) continue;

    // ---- one step attempt ----
    bool finished = false;
    if (ray.num_retry > g.max_retries) {
      ray.flag = true;
      finished = true;
    } else {
      n_attempts++;
      double h = ray.h_new;
      if (!ray.prev_fail && ray.n > 0) {
        for (int c = 0; c < 9; c++) ray.y[c] = ray.y5[c];
        for (int p = 0; p < kComp; p++) ks[p * kBlock] = ks[(6 * kComp + p) * kBlock];  // FSAL
        ray.ds0 = ds_last;
      }
      double r = ray.prev_fail ? ksx::radius(g.a, ray.y[1], ray.y[2], ray.y[3]) : ray.r_new;

      // stages (stage 0 only for a fresh ray); the distance component of the 5th-order solution is
      // accumulated stage by stage, in the same order as the other components below
      double s5 = ray.y[8];
      if (!ray.need_k0) s5 += c_b5[0] * h * ray.ds0;
      for (int s = ray.need_k0 ? 0 : 1; s < 7; s++) {
        double pos[3] = {ray.y[1], ray.y[2], ray.y[3]};
        double mom[4] = {ray.y[4], ray.y[5], ray.y[6], ray.y[7]};
        for (int q = 0; q < s; q++) {
          double ah = c_a[s][q] * h;
          const double *kq = ks + (size_t)q * kComp * kBlock;
          pos[0] += ah * kq[1 * kBlock];
          pos[1] += ah * kq[2 * kBlock];
          pos[2] += ah * kq[3 * kBlock];
          mom[1] += ah * kq[4 * kBlock];
          mom[2] += ah * kq[5 * kBlock];
          mom[3] += ah * kq[6 * kBlock];
        }
        double ds = eval_rhs<flat>(g, pos, mom, ks, s);
        if (s == 0) ray.ds0 = ds;
        s5 += c_b5[s] * h * ds;
        ds_last = ds;
      }
      ray.need_k0 = false;

      // 5th- and 4th-order solutions, error estimate over t,x,y,z,p_i (s excluded; p_t is constant)
      double y4[8];
      for (int c = 0; c < 9; c++) ray.y5[c] = ray.y[c];
      for (int c = 0; c < 8; c++) y4[c] = ray.y[c];
      for (int q = 0; q < 7; q++) {
        double b5h = c_b5[q] * h, b4h = c_b4[q] * h;
        const double *kq = ks + (size_t)q * kComp * kBlock;
        double kv;
        kv = kq[0 * kBlock]; ray.y5[0] += b5h * kv; y4[0] += b4h * kv;
        kv = kq[1 * kBlock]; ray.y5[1] += b5h * kv; y4[1] += b4h * kv;
        kv = kq[2 * kBlock]; ray.y5[2] += b5h * kv; y4[2] += b4h * kv;
        kv = kq[3 * kBlock]; ray.y5[3] += b5h * kv; y4[3] += b4h * kv;
        kv = kq[4 * kBlock]; ray.y5[5] += b5h * kv; y4[5] += b4h * kv;
        kv = kq[5 * kBlock]; ray.y5[6] += b5h * kv; y4[6] += b4h * kv;
        kv = kq[6 * kBlock]; ray.y5[7] += b5h * kv; y4[7] += b4h * kv;
      }
      ray.y5[8] = s5;
      ray.r_new = ksx::radius(g.a, ray.y5[1], ray.y5[2], ray.y5[3]);
      double error = 0.0;
      for (int c = 0; c < 8; c++) {
        double ya = fabs(ray.y[c]), yb = fabs(ray.y5[c]);
        double y_abs = ya < yb ? yb : ya;
        double scale = g.tol_abs + g.tol_rel * y_abs;
        double ratio = blmath::div_rn(fabs(ray.y5[c] - y4[c]), scale);
        error = error < ratio ? ratio : error;
      }

      if (!(error <= 1.0)) {
        // reject: shrink and retry from the same state
        double factor = 0.2;
        if (isfinite(error)) {
          double ideal = 0.9 * blmath::pow_glibc(error, -0.2);
          factor = ideal < 0.2 ? 0.2 : ideal;
        }
        ray.h_new = h * factor;
        ray.num_retry++;
        ray.prev_fail = true;
      } else {
        n_accepted++;
        double factor = 10.0;
        if (error > 0.0) {
          factor = 0.9 * blmath::pow_glibc(error, -0.2);
          factor = factor < 0.2 ? 0.2 : factor;
          factor = 10.0 < factor ? 10.0 : factor;
        }
        if (ray.prev_fail) factor = 1.0 < factor ? 1.0 : factor;
        ray.h_new = h * factor;
        ray.num_retry = 0;
        ray.prev_fail = false;

        // 4th-order midpoint
        double ym[8];
        for (int c = 0; c < 8; c++) ym[c] = ray.y[c];
        for (int q = 0; q < 7; q++) {
          double bh = c_b4m[q] * h;
          const double *kq = ks + (size_t)q * kComp * kBlock;
          ym[0] += bh * kq[0 * kBlock];
          ym[1] += bh * kq[1 * kBlock];
          ym[2] += bh * kq[2 * kBlock];
          ym[3] += bh * kq[3 * kBlock];
          ym[5] += bh * kq[4 * kBlock];
          ym[6] += bh * kq[5 * kBlock];
          ym[7] += bh * kq[6 * kBlock];
        }

        // subdivide so each stored piece is at most ray_step * r long
        double r_mid = ksx::radius(g.a, ym[1], ym[2], ym[3]);
        double ds_step = g.ray_step * r_mid;
        double ds_full = ray.y5[8] - ray.y[8];
        int n_ideal = (int)ceil(blmath::div_rn(ds_full, ds_step));
        int n_room = g.max_steps - ray.n;
        int n_sub = n_ideal;
        if (n_sub > n_room) {
          n_sub = n_room;
          ray.flag = true;
        }

        if (n_ideal == 1) {
          store_sample<flat>(g, ray, ray.n, ym, h);
        } else if (n_ideal > 1) {
          // quartic through both endpoints, both end slopes and the 4th-order midpoint
          double q0[7], q1[7], q2[7], q3[7];
          const int yc[7] = {0, 1, 2, 3, 5, 6, 7};
#pragma unroll
          for (int j = 0; j < 7; j++) {
            int c = yc[j];
            double k0 = ks[j * kBlock], k6 = ks[(6 * kComp + j) * kBlock];
            q0[j] = ray.y5[c] - ray.y[c];
            q1[j] = ray.y[c] - ray.y5[c] + h * k0;
            q2[j] = 2.0 * (ray.y5[c] - ray.y[c]) - h * (k0 + k6);
            q3[j] = 0.0;
          }
          for (int q = 0; q < 7; q++) {
            double dh = c_d[q] * h;
            const double *kq = ks + (size_t)q * kComp * kBlock;
#pragma unroll
            for (int j = 0; j < 7; j++) q3[j] += dh * kq[j * kBlock];
          }
          const blmath::Recip inv_n = blmath::recip_of((double)n_ideal);
          double len = blmath::div_by(h, inv_n);
          for (int nn = 0; nn < n_sub; nn++) {
            double frac = blmath::div_by(nn + 0.5, inv_n);
            double v[8];
            v[4] = ray.y[4];
#pragma unroll
            for (int j = 0; j < 7; j++) {
              int c = yc[j];
              v[c] = ray.y[c] +
                     frac * (q0[j] + (1.0 - frac) * (q1[j] + frac * (q2[j] + (1.0 - frac) * q3[j])));
            }
            store_sample<flat>(g, ray, ray.n + nn, v, len);
          }
        }

        // renormalise the end-of-step momentum to the null cone
        ksx::renormalize_momentum<flat>(g.a, ray.y5[1], ray.y5[2], ray.y5[3], &ray.y5[4]);

        bool outer = ray.r_new > g.camera_r && ray.r_new > r;
        bool inner = ray.r_new < g.r_terminate;
        int n_end = ray.n + n_sub;
        if (outer || inner) {
          ray.n = n_end;
          finished = true;
        } else {
          if (n_end >= g.max_steps) ray.flag = true;
          ray.n = n_end;
          if (ray.n >= g.max_steps || n_sub <= 0) finished = true;
        }
      }
    }

    if (finished) {
      int count = ray.n;
      if (ray.trunc >= 0 && count > 1) count = ray.trunc;
      g.sample_num[ray.m] = count;
      g.sample_flags[ray.m] = ray.flag ? 1 : 0;
      atomicAdd(&g.counters->samples, (unsigned long long)count);
      if (ray.flag) atomicAdd(&g.counters->bad, 1ull);
      atomicMax(&g.counters->max_samples, count);
      active = false;
    }
  }

  // per-warp totals of attempts / accepted steps (roofline accounting)
  for (int off = 16; off > 0; off >>= 1) {
    n_attempts += __shfl_down_sync(full, n_attempts, off);
    n_accepted += __shfl_down_sync(full, n_accepted, off);
  }
  if (lane == 0) {
    atomicAdd(&g.counters->attempts, n_attempts);
    atomicAdd(&g.counters->accepted, n_accepted);
  }
}


// Fixed-fraction-step integrators (reference geodesics.cpp:418-606 RK4, :626-795 RK2): one sample per
// step, h = -ray_step (r - r_horizon).  Same store path (truncation + renormalisation) as DP.
template <bool flat, int order>
__global__ void __launch_bounds__(kBlock) geodesic_rk_kernel(GeoArgs g) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.rays) return;
  Ray ray;
  ray.m = m;
  double y[8];
  for (int c = 0; c < 4; c++) {
    y[c] = g.cam_pos[4 * m + c];
    y[4 + c] = g.cam_dir[4 * m + c];
  }
  ray.trunc = -1;
  ray.r_prev_sample = 0.0;
  ray.flag = false;
  double r_new = ksx::radius(g.a, y[1], y[2], y[3]);
  int count = 0;
  auto deriv = [&](const double *v, double *k) {
    double dx[4], dp[3], ds;
    ksx::rhs<flat>(g.a, v[1], v[2], v[3], v + 4, dx, dp, ds);
    k[0] = dx[0]; k[1] = dx[1]; k[2] = dx[2]; k[3] = dx[3];
    k[4] = 0.0; k[5] = dp[0]; k[6] = dp[1]; k[7] = dp[2];
  };
  for (int n = 0; n < g.max_steps; n++) {
    double r = r_new;
    double h = -g.ray_step * (r - g.r_horizon);
    double k[8], sub[8], mid[8];
    if (order == 4) {
      double acc[8];
      deriv(y, k);
      for (int p = 0; p < 8; p++) acc[p] = y[p] + 1.0 / 6.0 * h * k[p];
      for (int p = 0; p < 8; p++) sub[p] = y[p] + 0.5 * h * k[p];
      deriv(sub, k);
      for (int p = 0; p < 8; p++) acc[p] += 1.0 / 3.0 * h * k[p];
      for (int p = 0; p < 8; p++) sub[p] = y[p] + 0.5 * h * k[p];
      deriv(sub, k);
      for (int p = 0; p < 8; p++) acc[p] += 1.0 / 3.0 * h * k[p];
      for (int p = 0; p < 8; p++) sub[p] = y[p] + h * k[p];
      deriv(sub, k);
      for (int p = 0; p < 8; p++) acc[p] += 1.0 / 6.0 * h * k[p];
      for (int p = 0; p < 8; p++) mid[p] = 0.5 * (y[p] + acc[p]);
      for (int p = 0; p < 8; p++) y[p] = acc[p];
    } else {
      deriv(y, k);
      for (int p = 0; p < 8; p++) sub[p] = y[p] + h * k[p];
      for (int p = 0; p < 8; p++) y[p] += 1.0 / 2.0 * h * k[p];
      for (int p = 0; p < 8; p++) mid[p] = y[p];
      deriv(sub, k);
      for (int p = 0; p < 8; p++) y[p] += 1.0 / 2.0 * h * k[p];
    }
    store_sample<flat>(g, ray, n, mid, h);
    ksx::renormalize_momentum<flat>(g.a, y[1], y[2], y[3], &y[4]);
    count++;
    r_new = ksx::radius(g.a, y[1], y[2], y[3]);
    bool outer = r_new > g.camera_r && r_new > r;
    bool inner = r_new < g.r_terminate;
    if (outer || inner) break;
    if (n + 1 >= g.max_steps) ray.flag = true;
  }
  if (ray.trunc >= 0 && count > 1) count = ray.trunc;
  g.sample_num[m] = count;
  g.sample_flags[m] = ray.flag ? 1 : 0;
  atomicAdd(&g.counters->samples, (unsigned long long)count);
  if (ray.flag) atomicAdd(&g.counters->bad, 1ull);
  atomicMax(&g.counters->max_samples, count);
}

}  // namespace

namespace {
template <bool flat, int MINB>
cudaError_t launch_dp(const GeoArgs *args, int sm_count, cudaStream_t stream) {
  size_t smem = (size_t)7 * kComp * kBlock * sizeof(double);
  cudaError_t err = cudaFuncSetAttribute(geodesic_dp_kernel<flat, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, geodesic_dp_kernel<flat, MINB>, kBlock, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  long long want = (args->rays + kBlock - 1) / kBlock;
  long long grid = (long long)sm_count * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  geodesic_dp_kernel<flat, MINB><<<(unsigned)grid, kBlock, smem, stream>>>(*args);
  return cudaGetLastError();
}
}  // namespace

// Launch the DP integrator for one wave of rays on `stream`: a persistent grid of (SM count x resident
// blocks).  min_blocks selects the occupancy variant (registers capped for 2, 3 or 4 blocks per SM).
extern "C" cudaError_t bl_launch_geodesic_dp(const GeoArgs *args, int flat, int sm_count, int min_blocks,
                                             cudaStream_t stream) {
  if (flat) return launch_dp<true, 2>(args, sm_count, stream);
  if (min_blocks >= 4) return launch_dp<false, 4>(args, sm_count, stream);
  if (min_blocks == 3) return launch_dp<false, 3>(args, sm_count, stream);
  return launch_dp<false, 2>(args, sm_count, stream);
}

extern "C" cudaError_t bl_launch_geodesic_rk(const GeoArgs *args, int flat, int order, int sm_count,
                                             cudaStream_t stream) {
  (void)sm_count;
  unsigned grid = (unsigned)((args->rays + kBlock - 1) / kBlock);
  if (order == 4) {
    if (flat) geodesic_rk_kernel<true, 4><<<grid, kBlock, 0, stream>>>(*args);
    else geodesic_rk_kernel<false, 4><<<grid, kBlock, 0, stream>>>(*args);
  } else {
    if (flat) geodesic_rk_kernel<true, 2><<<grid, kBlock, 0, stream>>>(*args);
    else geodesic_rk_kernel<false, 2><<<grid, kBlock, 0, stream>>>(*args);
  }
  return cudaGetLastError();
}
