// Per-sample stages shared by the unpolarized and polarized transfer kernels:
//   geometric cuts -> CKS->SKS -> block search -> cell search -> nearest / trilinear gather
//   (reference src/radiation_integrator/simulation_sampling.cpp:205-575, :667-1033)
//   -> plasma state in CGS, value cuts, cell values (simulation_coefficients.cpp:286-387)
//   -> fluid four-velocity and magnetic field in Cartesian Kerr-Schild (:397-408)
// Nothing here is written to HBM: the fused kernels consume the results in registers, so the
// reference's N x S x (inds, fracs, 9 primitives, 8 coefficients) arrays never exist.
#pragma once
#include "rad_types.cuh"

namespace rad {

enum SampleStatus : int { kSampleOk = 0, kSampleCut = 1, kSampleNan = 2, kSampleFallback = 3 };

struct Prims {
  float rho, pgas, kappa, uu1, uu2, uu3, bb1, bb2, bb3;
};

struct SampleIndex {  // what the reference stores in sample_inds / sample_fracs
  int b, k, j, i;
  double fk, fj, fi;
};

// Kerr-Schild radius and its reciprocal with ordinary (fused) arithmetic; agrees with the reference
// (radiation_geometry.cpp:18-25) to rounding.  One rsqrt yields both r and 1/r.
__device__ __forceinline__ double ks_radius(double a, double x, double y, double z, double &inv_r) {
  double a2 = a * a;
  double rr2 = x * x + y * y + z * z;
  double d = rr2 - a2, e = 2.0 * a * z;
  double r2 = 0.5 * (d + sqrt(d * d + e * e));
  inv_r = rsqrt(r2);
  return r2 * inv_r;
}
__device__ __forceinline__ double ks_radius(double a, double x, double y, double z) {
  double inv_r;
  return ks_radius(a, x, y, z, inv_r);
}

// first index i in [0, n) with faces[i+1] >= x, else n (reference linear scan :458-466)
__device__ __forceinline__ int find_cell(const double *__restrict__ faces, int n, double x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(faces + mid + 1) >= x)
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

// Same result as find_cell, starting from the cell the previous sample of this ray lay in: consecutive
// samples are ray_step * r apart, so the answer is almost always the hint or a neighbour and the two
// (independent) face loads replace a chain of log2(n) dependent ones.  Falls back to bisection of the
// remaining range, so the returned index is the reference's for any hint.
__device__ __forceinline__ int find_cell_hint(const double *__restrict__ faces, int n, double x, int hint) {
  int h = hint < 0 ? 0 : (hint > n - 1 ? n - 1 : hint);
  double fa = __ldg(faces + h), fb = __ldg(faces + h + 1);
  int lo, hi;
  if (fb >= x) {
    if (h == 0 || fa < x) return h;
    if (h == 1 || __ldg(faces + h - 1) < x) return h - 1;
    lo = 0;
    hi = h - 2;
  } else {
    if (h + 1 >= n) return n;
    if (__ldg(faces + h + 2) >= x) return h + 1;
    lo = h + 2;
    hi = n;
  }
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(faces + mid + 1) >= x)
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ bool in_block(const double *__restrict__ bd, double x1, double x2, double x3) {
  return x1 >= bd[0] && x1 <= bd[1] && x2 >= bd[2] && x2 <= bd[3] && x3 >= bd[4] && x3 <= bd[5];
}

__device__ __forceinline__ void load_cell(const GridDev &g, size_t c, float v[8]) {
  float4 lo = __ldg(g.cells + 2 * c), hi = __ldg(g.cells + 2 * c + 1);
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
  v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// FMKS only: cell index c >= the number of cells, i.e. past the end of every variable's plane of the reader's
// (variable, cell) array -- each variable reads the first cells of the variable stored after it (GridDev::next_slot).
__device__ __forceinline__ void load_cell_past_end(const GridDev &g, size_t slot, size_t over, float v[8], float &kappa) {
  float t[8];
  load_cell(g, slot + over, t);
  float tk = g.kappa ? __ldg(g.kappa + slot + over) : 0.0f;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    int ns = g.next_slot[q];
    float val = 0.0f;
#pragma unroll
    for (int u = 0; u < 8; u++) val = ns == u ? t[u] : val;
    v[q] = ns == 8 ? tk : val;
  }
  int nk = g.next_slot[8];
  float val = 0.0f;
#pragma unroll
  for (int u = 0; u < 8; u++) val = nk == u ? t[u] : val;
  kappa = val;
}

// The radiation kernels walk a ray's records one after the other and the first use of a record sits at the head of the
// sample's dependency chain (ncu: ~10 % of the unpolarized kernel's stall samples wait on that load).  One instruction, no
// register: ask L2 for the record `ahead` samples further down the walk.
// ahead = 1 ... 9: into L2; 10 + k: k samples ahead into L1 (BL_RAD_PREFETCH, tuning).
__device__ __forceinline__ void prefetch_record(const StepBuffer &sb, int n, int64_t m, int ahead) {
  if (ahead >= 10) {
    if (n - (ahead - 10) >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(sb.buf + sb.at(n - (ahead - 10), m)));
  } else if (ahead > 0 && n - ahead >= 0) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(sb.buf + sb.at(n - ahead, m)));
  }
}

// Geometric cuts of one sample (simulation_sampling.cpp:237-292, formula_coefficients.cpp:75-119).
// Returns true if the sample is cut.  r is the Kerr-Schild radius of (x,y,z).
__device__ __forceinline__ bool geometric_cut(const RadParams &P, double x, double y, double z, double r) {
  if (r > P.camera_r) return true;
  if (P.cut_omit_near || P.cut_omit_far) {
    double dot = x * P.camera_x[1] + y * P.camera_x[2] + z * P.camera_x[3];
    if ((P.cut_omit_near && dot > 0.0) || (P.cut_omit_far && dot < 0.0)) return true;
  }
  if ((P.cut_omit_in >= 0.0 && r < P.cut_omit_in) || (P.cut_omit_out >= 0.0 && r > P.cut_omit_out)) return true;
  if (P.cut_midplane_theta > 0.0 || P.cut_midplane_theta < 0.0) {
    double th = acos(z / r);
    double off = fabs(th - phys::pi / 2.0);
    if ((P.cut_midplane_theta > 0.0 && off > P.cut_midplane_theta) ||
        (P.cut_midplane_theta < 0.0 && off < -P.cut_midplane_theta))
      return true;
  }
  if ((P.cut_midplane_z > 0.0 && fabs(z) > P.cut_midplane_z) ||
      (P.cut_midplane_z < 0.0 && fabs(z) < -P.cut_midplane_z))
    return true;
  if (P.cut_plane) {
    double dot = (x - P.cut_plane_origin[0]) * P.cut_plane_normal[0] +
                 (y - P.cut_plane_origin[1]) * P.cut_plane_normal[1] +
                 (z - P.cut_plane_origin[2]) * P.cut_plane_normal[2];
    if (dot < 0.0) return true;
  }
  return false;
}

// ---- inter-block interpolation (simulation_block_interp; reference simulation_sampling.cpp:504-549) ----

// Block with the given refinement level and logical location, or -1.
__device__ __forceinline__ int find_block(const GridDev &g, int level, int li, int lj, int lk) {
  if (level < 0 || level > g.max_level || li < 0 || lj < 0 || lk < 0) return -1;
  unsigned long long key = block_key(level, li, lj, lk);
  uint32_t h = block_hash(key) & g.hash_mask;
  for (;;) {
    unsigned long long kk = __ldg(g.hash_keys + h);
    if (kk == key) return __ldg(g.hash_vals + h);
    if (kk == ~0ull) return -1;
    h = (h + 1) & g.hash_mask;
  }
}

// Cell (block, k, j, i) standing in for cell (k, j, i) of block b when an index is one beyond the block
// (reference FindNearbyInds, simulation_sampling.cpp:1068-1321): the ghost cell's owner at the same level,
// else the coarser cell containing it, else the nearest finer cell; constant extrapolation off the grid;
// periodic in x3 for spherical coordinates.  (k_c, j_c, i_c) is the cell containing the sample point.
// The reference scans all blocks for each of its existence tests; here each test is one hash lookup.
static __device__ __noinline__ void find_nearby_inds(const RadParams &P, const GridDev &g, int b, int k, int j, int i, int k_c,
                                              int j_c, int i_c, double x3, double x2, double x1, int inds[4]) {
  const int n_i = g.n_i, n_j = g.n_j, n_k = g.n_k;
  const int level = __ldg(g.levels + b);
  const int li = __ldg(g.locs + 3 * b), lj = __ldg(g.locs + 3 * b + 1), lk = __ldg(g.locs + 3 * b + 2);
  const bool upper_i = i > n_i / 2, upper_j = j > n_j / 2, upper_k = k > n_k / 2;
  const int i_safe = max(min(i, n_i - 1), 0), j_safe = max(min(j, n_j - 1), 0), k_safe = max(min(k, n_k - 1), 0);
  inds[0] = b; inds[1] = k; inds[2] = j; inds[3] = i;
  if (i == i_safe && j == j_safe && k == k_safe) return;
  const bool sks = P.coord != 0;
  auto n3 = [&](int lev) { return g.n3_root << lev; };   // blocks around x3 at a level
  const int ui = upper_i ? 1 : 0, uj = upper_j ? 1 : 0, uk = upper_k ? 1 : 0;

  // does the grid continue in each direction in which the cell lies outside the block?
  bool x1_off = false, x2_off = false, x3_off = false;
  if (i != i_safe) {
    int s = i == -1 ? -1 : 1;
    x1_off = find_block(g, level, li + s, lj, lk) < 0 && find_block(g, level - 1, (li + s) / 2, lj / 2, lk / 2) < 0 &&
             find_block(g, level + 1, i == -1 ? li * 2 - 1 : li * 2 + 2, lj * 2 + uj, lk * 2 + uk) < 0;
  }
  if (j != j_safe) {
    int s = j == -1 ? -1 : 1;
    x2_off = find_block(g, level, li, lj + s, lk) < 0 && find_block(g, level - 1, li / 2, (lj + s) / 2, lk / 2) < 0 &&
             find_block(g, level + 1, li * 2 + ui, j == -1 ? lj * 2 - 1 : lj * 2 + 2, lk * 2 + uk) < 0;
  }
  if (k != k_safe) {
    int s = k == -1 ? -1 : 1;
    x3_off = find_block(g, level, li, lj, lk + s) < 0 && find_block(g, level - 1, li / 2, lj / 2, (lk + s) / 2) < 0 &&
             find_block(g, level + 1, li * 2 + ui, lj * 2 + uj, k == -1 ? lk * 2 - 1 : lk * 2 + 2) < 0;
    // across the periodic boundary
    if (x3_off && sks && ((k == -1 && lk == 0) || (k == n_k && lk == n3(level) - 1))) {
      bool low = k == -1;
      x3_off = find_block(g, level, li, lj, low ? n3(level) - 1 : 0) < 0 &&
               (level < 1 || find_block(g, level - 1, li / 2, lj / 2, low ? n3(level - 1) - 1 : 0) < 0) &&
               (level + 1 > g.max_level ||
                find_block(g, level + 1, li * 2 + ui, lj * 2 + uj, low ? n3(level + 1) - 1 : 0) < 0);
    }
  }
  if (x1_off) i = i_safe;
  if (x2_off) j = j_safe;
  if (x3_off) k = k_safe;
  const bool wrap_low = sks && k == -1 && lk == 0;
  const bool wrap_high = sks && k == n_k && lk == n3(level) - 1;

  // same level
  {
    int si = i == i_safe ? li : (i == -1 ? li - 1 : li + 1);
    int sj = j == j_safe ? lj : (j == -1 ? lj - 1 : lj + 1);
    int sk = k == k_safe ? lk : (k == -1 ? lk - 1 : lk + 1);
    if (wrap_low) sk = n3(level) - 1;
    if (wrap_high) sk = 0;
    int bb = find_block(g, level, si, sj, sk);
    if (bb >= 0) {
      inds[0] = bb;
      inds[1] = k == k_safe ? k : (k == -1 ? n_k - 1 : 0);
      inds[2] = j == j_safe ? j : (j == -1 ? n_j - 1 : 0);
      inds[3] = i == i_safe ? i : (i == -1 ? n_i - 1 : 0);
      return;
    }
  }
  // coarser level
  if (level - 1 >= 0) {
    int si = i == i_safe ? li / 2 : (i == -1 ? (li - 1) / 2 : (li + 1) / 2);
    int sj = j == j_safe ? lj / 2 : (j == -1 ? (lj - 1) / 2 : (lj + 1) / 2);
    int sk = k == k_safe ? lk / 2 : (k == -1 ? (lk - 1) / 2 : (lk + 1) / 2);
    if (wrap_low) sk = n3(level - 1) - 1;
    if (wrap_high) sk = 0;
    int bb = find_block(g, level - 1, si, sj, sk);
    if (bb >= 0) {
      inds[0] = bb;
      inds[1] = k == k_safe ? (lk % 2 * n_k + k) / 2 : (k == -1 ? n_k - 1 : 0);
      inds[2] = j == j_safe ? (lj % 2 * n_j + j) / 2 : (j == -1 ? n_j - 1 : 0);
      inds[3] = i == i_safe ? (li % 2 * n_i + i) / 2 : (i == -1 ? n_i - 1 : 0);
      return;
    }
  }
  // finer level
  {
    int si = li * 2 + (i == i_safe ? 0 : (i == -1 ? -1 : 1)) + ui;
    int sj = lj * 2 + (j == j_safe ? 0 : (j == -1 ? -1 : 1)) + uj;
    int sk = lk * 2 + (k == k_safe ? 0 : (k == -1 ? -1 : 1)) + uk;
    if (wrap_low && level + 1 <= g.max_level) sk = n3(level + 1) - 1;
    if (wrap_high) sk = 0;
    int bb = find_block(g, level + 1, si, sj, sk);
    if (bb >= 0) {
      inds[0] = bb;
      inds[1] = k == k_safe ? (upper_k ? (k - n_k / 2) * 2 : k * 2) : (k == -1 ? n_k - 2 : 0);
      inds[2] = j == j_safe ? (upper_j ? (j - n_j / 2) * 2 : j * 2) : (j == -1 ? n_j - 2 : 0);
      inds[3] = i == i_safe ? (upper_i ? (i - n_i / 2) * 2 : i * 2) : (i == -1 ? n_i - 2 : 0);
      inds[1] += (k < k_c || (k == k_c && x3 > __ldg(g.x3v + (size_t)b * n_k + k_c))) ? 1 : 0;
      inds[2] += (j < j_c || (j == j_c && x2 > __ldg(g.x2v + (size_t)b * n_j + j_c))) ? 1 : 0;
      inds[3] += (i < i_c || (i == i_c && x1 > __ldg(g.x1v + (size_t)b * n_i + i_c))) ? 1 : 0;
      return;
    }
  }
  // inconsistent mesh (the reference throws "Grid interpolation failed."): fall back to the nearest own cell
  inds[0] = b; inds[1] = k_safe; inds[2] = j_safe; inds[3] = i_safe;
}

// Where the previous sample of this ray was found: the block (the reference keeps the same cache per
// OpenMP thread, simulation_sampling.cpp:180-189,352-394) and, as search hints only, its cell.
struct CellCache {
  int b, i, j, k;
};

// Locate and gather.  smem_bounds: block bounds staged in shared memory (n_b*6 doubles) or nullptr to
// read them from HBM.  inv_r = 1/r.
// Which time slice(s) of the resident window a sample at coordinate time x0 reads (slow light; reference
// simulation_sampling.cpp:298-349): entry t_ind, blended with entry t_ind + 1 by t_frac when interpolating.
struct SlowLight {   // per-ray accumulation of the extrapolation accounting (simulation_sampling.cpp:556-575)
  int extrap;        // bit 0: camera side small, 1: camera side large, 2: source side small, 3: source side large
  double ext[4];     // largest extrapolation seen in each category
};

__device__ __forceinline__ void time_slice(const RadParams &P, double x0, int &t_ind, double &t_frac, SlowLight &sl) {
  t_ind = 0;
  t_frac = 0.0;
  const int last = P.slow_count - 1;
  if (x0 >= P.slow_time[0]) {
    if (x0 > P.slow_time[0] + P.extrap_tol) { sl.extrap |= 2; sl.ext[1] = fmax(sl.ext[1], x0 - P.slow_time[0]); }
    else if (x0 > P.slow_time[0]) { sl.extrap |= 1; sl.ext[0] = fmax(sl.ext[0], x0 - P.slow_time[0]); }
  } else if (x0 <= P.slow_time[last]) {
    if (x0 < P.slow_time[last] - P.extrap_tol) { sl.extrap |= 8; sl.ext[3] = fmax(sl.ext[3], P.slow_time[last] - x0); }
    else if (x0 < P.slow_time[last]) { sl.extrap |= 4; sl.ext[2] = fmax(sl.ext[2], P.slow_time[last] - x0); }
    if (P.slow_interp) { t_ind = last - 1; t_frac = 1.0; }
    else t_ind = last;
  } else {
    while (P.slow_time[t_ind++] > x0) {}
    t_ind--;
    if (P.slow_interp) {
      t_ind--;
      t_frac = (x0 - P.slow_time[t_ind]) / (P.slow_time[t_ind + 1] - P.slow_time[t_ind]);
    } else if (P.slow_time[t_ind - 1] - x0 <= x0 - P.slow_time[t_ind]) {
      t_ind--;
    }
  }
}

// End of a ray: add its extrapolation flags to the per-launch counters (pixels per category, largest values;
// positive doubles order like their bit patterns, so atomicMax on the bits is a max on the values).
__device__ __forceinline__ void flush_slow_light(unsigned long long *counters, const SlowLight &sl) {
  if (!counters || !sl.extrap) return;
  for (int c = 0; c < 4; c++)
    if (sl.extrap & (1 << c)) {
      atomicAdd(counters + c, 1ull);
      atomicMax(counters + 4 + c, (unsigned long long)__double_as_longlong(sl.ext[c]));
    }
}

// FMKS grids (simulation_coord = fmks): zone and fractional position by scaling in the native coordinates, found
// through the reader's (r, theta) -> (x1, x2) table (simulation_sampling.cpp:397-452).  The arithmetic is spelled out
// operation by operation (no contraction) because truncations of its results are cell indices.  Indices one past a
// row (the reference forms them for the last zone) address the following cells of the flat array, as they do there;
// past the last cell see load_cell_past_end.  Kept apart from sample_grid so that the other coordinate systems'
// code is not perturbed.
static __device__ __noinline__ SampleStatus sample_grid_fmks(const RadParams &P, const GridDev &g, int b, double x1, double x2,
                                                      double x3, size_t slot_a, size_t slot_b, bool two_slices,
                                                      double t_frac, CellCache &cache, Prims &out, SampleIndex &si) {
  const int n_i = g.n_i, n_j = g.n_j, n_k = g.n_k;
  const size_t last_cell = (size_t)g.n_b * n_k * n_j * n_i - 1;
  int k = find_cell_hint(g.x3f + (size_t)b * (n_k + 1), n_k, x3, cache.k);
  cache.k = k;
  si.b = b;
  const int m1 = g.map_n1, m2 = g.map_n2;
  double i_ind, j_ind;
  double t_i = modf(__ddiv_rn(__dsub_rn(x1, g.map_r_in), g.map_dr), &i_ind);
  double t_j = modf(__ddiv_rn(x2, g.map_dtheta), &j_ind);
  int mi = min(max((int)i_ind, 0), m1 - 1), mj = min(max((int)j_ind, 0), m2 - 1);
  const double *map_x1 = g.sks_map + (size_t)mj * m1, *map_x2 = g.sks_map + ((size_t)m2 + min(mj + 1, m2 - 1)) * m1;
  double a_lo = __ldg(map_x1 + mi), a_hi = __ldg(map_x1 + min(mi + 1, m1 - 1)), b_hi = __ldg(map_x2 + mi);
  double nat_x1 = __dadd_rn(__dmul_rn(__dsub_rn(1.0, t_i), a_lo), __dmul_rn(t_i, a_hi));
  double nat_x2 = __dadd_rn(__dmul_rn(__dsub_rn(1.0, t_j), b_hi), __dmul_rn(t_j, b_hi));   // both terms at j+1, as the reference
  double x1_0 = __ldg(g.x1f), dx1 = __dsub_rn(__ldg(g.x1f + 1), x1_0), dx2 = __dsub_rn(__ldg(g.x2f + 1), __ldg(g.x2f));
  double f_i = modf(__ddiv_rn(__dsub_rn(nat_x1, x1_0), dx1), &i_ind);
  double f_j = modf(__ddiv_rn(nat_x2, dx2), &j_ind);
  int i_m = (int)i_ind, j_m = (int)j_ind;

  auto load = [&](size_t slot, size_t cell, float f[8], float &kv) {
    if (cell > last_cell) {
      load_cell_past_end(g, slot, cell - last_cell - 1, f, kv);
    } else {
      load_cell(g, slot + cell, f);
      kv = g.kappa ? __ldg(g.kappa + slot + cell) : 0.0f;
    }
  };
  auto finish = [&](auto gather) {
    double v[9];
    gather(slot_a, v);
    if (two_slices) {
      double w[9];
      gather(slot_b, w);
      for (int q = 0; q < 9; q++) v[q] = (1.0 - t_frac) * v[q] + t_frac * w[q];
    }
    out.rho = (float)v[0]; out.pgas = (float)v[1]; out.uu1 = (float)v[2]; out.uu2 = (float)v[3];
    out.uu3 = (float)v[4]; out.bb1 = (float)v[5]; out.bb2 = (float)v[6]; out.bb3 = (float)v[7];
    out.kappa = (float)v[8];
  };

  if (!P.interp) {
    int i = f_i >= 0.5 ? i_m + 1 : i_m, j = f_j >= 0.5 ? j_m + 1 : j_m;
    si.k = k; si.j = j; si.i = i;
    si.fk = si.fj = si.fi = 0.0;
    size_t c = (((size_t)b * n_k + k) * n_j + j) * n_i + i;
    finish([&](size_t slot, double v[9]) {
      float f[8], kv;
      load(slot, c, f, kv);
      for (int q = 0; q < 8; q++) v[q] = (double)f[q];
      v[8] = (double)kv;
    });
    return kSampleOk;
  }
  const double *x3v = g.x3v + (size_t)b * n_k;
  int k_m = (k == 0 || (k != n_k - 1 && x3 >= __ldg(x3v + k))) ? k : k - 1;
  double f_k = (x3 - __ldg(x3v + k_m)) * __ldg(g.x3d + (size_t)b * n_k + k_m);
  si.k = k_m; si.j = j_m; si.i = i_m;
  si.fk = f_k; si.fj = f_j; si.fi = f_i;
  double gk = 1.0 - f_k, gj = 1.0 - f_j, gi = 1.0 - f_i;
  double w[8] = {gk * gj * gi, gk * gj * f_i, gk * f_j * gi, gk * f_j * f_i,
                 f_k * gj * gi, f_k * gj * f_i, f_k * f_j * gi, f_k * f_j * f_i};
  size_t c0 = (((size_t)b * n_k + k_m) * n_j + j_m) * n_i + i_m;
  size_t sj = (size_t)n_i, sk = (size_t)n_j * n_i;
  size_t off[8] = {0, 1, sj, sj + 1, sk, sk + 1, sk + sj, sk + sj + 1};
  finish([&](size_t slot, double v[9]) {
    float corner[8] = {0, 0, 0, 0, 0, 0, 0, 0}, corner_kappa = 0.0f;
    for (int q = 0; q < 9; q++) v[q] = 0.0;
    for (int p = 0; p < 8; p++) {
      float f[8], kv;
      load(slot, c0 + off[p], f, kv);
      if (p == 0) {
        for (int q = 0; q < 8; q++) corner[q] = f[q];
        corner_kappa = kv;
      }
      for (int q = 0; q < 8; q++) v[q] += w[p] * (double)f[q];
      v[8] += w[p] * (double)kv;
    }
    // non-positive interpolated rho / pgas / kappa fall back to the anchor cell (:822-827)
    if (v[0] <= 0.0) v[0] = (double)corner[0];
    if (v[1] <= 0.0) v[1] = (double)corner[1];
    if (g.kappa && v[8] <= 0.0) v[8] = (double)corner_kappa;
  });
  return kSampleOk;
}

// EXT: inter-block interpolation and slow light compiled in (selected at run time by P.block_interp /
// P.slow_light); the light-only unpolarized kernel is instantiated without them so that the common path keeps
// its register budget.  x0 = coordinate time of the sample + camera time of the image (slow light only).
template <bool EXT>
__device__ __forceinline__ SampleStatus sample_grid(const RadParams &P, const GridDev &g,
                                                    const double *smem_bounds, double x, double y,
                                                    double z, double r, double inv_r, double x0, CellCache &cache,
                                                    Prims &out, SampleIndex &si, SlowLight &slow) {
  // simulation coordinates of the point (radiation_geometry.cpp:37-57)
  double x1 = x, x2 = y, x3 = z;
  if (P.coord != 0) {
    double th = acos(z * inv_r);
    // atan2(y, x) - atan(a / r) as one angle: arg((x + i y)(r - i a))
    double ph = atan2(y * r - P.a * x, x * r + P.a * y);
    ph += ph < 0.0 ? 2.0 * phys::pi : 0.0;
    ph -= ph >= 2.0 * phys::pi ? 2.0 * phys::pi : 0.0;
    x1 = r; x2 = th; x3 = ph;
  }
  // time slices (slow light): slot_a always, slot_b blended in with weight t_frac when interpolating in time
  size_t slot_a = 0, slot_b = 0;
  double t_frac = 0.0;
  bool two_slices = false;
  if (EXT && P.slow_light) {
    int t_ind;
    time_slice(P, x0, t_ind, t_frac, slow);
    slot_a = (size_t)P.slow_slot[t_ind] * g.slice_cells;
    two_slices = P.slow_interp != 0;
    if (two_slices) slot_b = (size_t)P.slow_slot[t_ind + 1] * g.slice_cells;
  }
  // block: keep the cached one while it still contains the point, else first match in index order
  const double *bounds = smem_bounds ? smem_bounds : g.bounds;
  int b = cache.b;
  const double *bd = bounds + 6 * b;
  if (x1 < bd[0] || x1 > bd[1] || x2 < bd[2] || x2 > bd[3] || x3 < bd[4] || x3 > bd[5]) {
    int bn = 0;
    for (; bn < g.n_b; bn++)
      if (in_block(bounds + 6 * bn, x1, x2, x3)) break;
    if (bn == g.n_b) return P.fallback_nan ? kSampleNan : kSampleFallback;
    b = bn;
    cache.b = bn;
  }
  const int n_i = g.n_i, n_j = g.n_j, n_k = g.n_k;
  if (EXT && P.coord == 2) return sample_grid_fmks(P, g, b, x1, x2, x3, slot_a, slot_b, two_slices, t_frac, cache, out, si);
  int i = find_cell_hint(g.x1f + (size_t)b * (n_i + 1), n_i, x1, cache.i);
  int j = find_cell_hint(g.x2f + (size_t)b * (n_j + 1), n_j, x2, cache.j);
  int k = find_cell_hint(g.x3f + (size_t)b * (n_k + 1), n_k, x3, cache.k);
  cache.i = i; cache.j = j; cache.k = k;
  si.b = b;
  const double *x1v = g.x1v + (size_t)b * n_i, *x2v = g.x2v + (size_t)b * n_j, *x3v = g.x3v + (size_t)b * n_k;

  // The three sampling modes below only differ in which cells they read and with what weights; `finish`
  // evaluates a mode on the time slice(s) and stores the float primitives the way the reference does
  // (per-slice fallback of non-positive rho / pgas / kappa to the anchor cell, blend in double, cast).
  auto finish = [&](auto gather) {
    double v[9];
    gather(slot_a, v);
    if (EXT && two_slices) {
      double w[9];
      gather(slot_b, w);
#pragma unroll
      for (int q = 0; q < 9; q++) v[q] = (1.0 - t_frac) * v[q] + t_frac * w[q];
    }
    out.rho = (float)v[0]; out.pgas = (float)v[1]; out.uu1 = (float)v[2]; out.uu2 = (float)v[3];
    out.uu3 = (float)v[4]; out.bb1 = (float)v[5]; out.bb2 = (float)v[6]; out.bb3 = (float)v[7];
    out.kappa = (float)v[8];
  };

  if (!P.interp) {
    si.k = k; si.j = j; si.i = i;
    si.fk = si.fj = si.fi = 0.0;
    size_t c = (((size_t)b * n_k + k) * n_j + j) * n_i + i;
    finish([&](size_t slot, double v[9]) {
      float f[8];
      load_cell(g, slot + c, f);
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = (double)f[q];
      v[8] = g.kappa ? (double)__ldg(g.kappa + slot + c) : 0.0;
    });
    return kSampleOk;
  }

  if (EXT && P.block_interp) {
    // inter-block trilinear: anchors may be ghost cells, resolved on neighbouring blocks of any level
    // (simulation_sampling.cpp:504-549).  Upper ghost coordinates are formed exactly as the reference
    // does, from x?v(b, i+1) -- for i = n-1 the next block's first centre (the arrays carry one element of
    // padding for the last block).
    const double *x1fb = g.x1f + (size_t)b * (n_i + 1), *x2fb = g.x2f + (size_t)b * (n_j + 1), *x3fb = g.x3f + (size_t)b * (n_k + 1);
    int i_m = x1 >= __ldg(x1v + i) ? i : i - 1, i_p = i_m + 1;
    int j_m = x2 >= __ldg(x2v + j) ? j : j - 1, j_p = j_m + 1;
    int k_m = x3 >= __ldg(x3v + k) ? k : k - 1, k_p = k_m + 1;
    double x1_m = i_m == -1 ? 2.0 * __ldg(x1fb + i) - __ldg(x1v + i) : __ldg(x1v + i_m);
    double x2_m = j_m == -1 ? 2.0 * __ldg(x2fb + j) - __ldg(x2v + j) : __ldg(x2v + j_m);
    double x3_m = k_m == -1 ? 2.0 * __ldg(x3fb + k) - __ldg(x3v + k) : __ldg(x3v + k_m);
    double x1_p = i_p == n_i ? 2.0 * __ldg(x1v + i + 1) - __ldg(x1v + i) : __ldg(x1v + i_p);
    double x2_p = j_p == n_j ? 2.0 * __ldg(x2v + j + 1) - __ldg(x2v + j) : __ldg(x2v + j_p);
    double x3_p = k_p == n_k ? 2.0 * __ldg(x3v + k + 1) - __ldg(x3v + k) : __ldg(x3v + k_p);
    double f_i = (x1 - x1_m) / (x1_p - x1_m);
    double f_j = (x2 - x2_m) / (x2_p - x2_m);
    double f_k = (x3 - x3_m) / (x3_p - x3_m);
    double gk = 1.0 - f_k, gj = 1.0 - f_j, gi = 1.0 - f_i;
    double w[8] = {gk * gj * gi, gk * gj * f_i, gk * f_j * gi, gk * f_j * f_i,
                   f_k * gj * gi, f_k * gj * f_i, f_k * f_j * gi, f_k * f_j * f_i};
    size_t cell[8];
#pragma unroll 1
    for (int p = 0; p < 8; p++) {
      int inds[4];
      find_nearby_inds(P, g, b, (p & 4) ? k_p : k_m, (p & 2) ? j_p : j_m, (p & 1) ? i_p : i_m, k, j, i, x3, x2, x1, inds);
      if (p == 0) {
        si.b = inds[0]; si.k = inds[1]; si.j = inds[2]; si.i = inds[3];
      }
      cell[p] = (((size_t)inds[0] * n_k + inds[1]) * n_j + inds[2]) * n_i + inds[3];
    }
    si.fk = f_k; si.fj = f_j; si.fi = f_i;
    finish([&](size_t slot, double v[9]) {
      float corner[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      float corner_kappa = 0.0f;
#pragma unroll
      for (int q = 0; q < 9; q++) v[q] = 0.0;
      for (int p = 0; p < 8; p++) {
        float f[8];
        load_cell(g, slot + cell[p], f);
        if (p == 0)
          for (int q = 0; q < 8; q++) corner[q] = f[q];
        for (int q = 0; q < 8; q++) v[q] += w[p] * (double)f[q];
        if (g.kappa) {
          float kv = __ldg(g.kappa + slot + cell[p]);
          if (p == 0) corner_kappa = kv;
          v[8] += w[p] * (double)kv;
        }
      }
      if (v[0] <= 0.0) v[0] = (double)corner[0];
      if (v[1] <= 0.0) v[1] = (double)corner[1];
      if (g.kappa && v[8] <= 0.0) v[8] = (double)corner_kappa;
    });
    return kSampleOk;
  }

  // intra-block trilinear with extrapolation at block edges (simulation_sampling.cpp:485-502)
  int i_m = (i == 0 || (i != n_i - 1 && x1 >= __ldg(x1v + i))) ? i : i - 1;
  int j_m = (j == 0 || (j != n_j - 1 && x2 >= __ldg(x2v + j))) ? j : j - 1;
  int k_m = (k == 0 || (k != n_k - 1 && x3 >= __ldg(x3v + k))) ? k : k - 1;
  double f_i = (x1 - __ldg(x1v + i_m)) * __ldg(g.x1d + (size_t)b * n_i + i_m);
  double f_j = (x2 - __ldg(x2v + j_m)) * __ldg(g.x2d + (size_t)b * n_j + j_m);
  double f_k = (x3 - __ldg(x3v + k_m)) * __ldg(g.x3d + (size_t)b * n_k + k_m);
  si.k = k_m; si.j = j_m; si.i = i_m;
  si.fk = f_k; si.fj = f_j; si.fi = f_i;
  // weights in the reference's term order (InterpolateSimple, simulation_sampling.cpp:1334-1351)
  double gk = 1.0 - f_k, gj = 1.0 - f_j, gi = 1.0 - f_i;
  double w[8] = {gk * gj * gi, gk * gj * f_i, gk * f_j * gi, gk * f_j * f_i,
                 f_k * gj * gi, f_k * gj * f_i, f_k * f_j * gi, f_k * f_j * f_i};
  size_t c0 = (((size_t)b * n_k + k_m) * n_j + j_m) * n_i + i_m;
  size_t sj = (size_t)n_i, sk = (size_t)n_j * n_i;
  size_t off[8] = {0, 1, sj, sj + 1, sk, sk + 1, sk + sj, sk + sj + 1};
  finish([&](size_t slot, double v[9]) {
    float corner[8];
    float corner_kappa = 0.0f;
#pragma unroll
    for (int q = 0; q < 9; q++) v[q] = 0.0;
#pragma unroll
    for (int p = 0; p < 8; p++) {
      float f[8];
      load_cell(g, slot + c0 + off[p], f);
      if (p == 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) corner[q] = f[q];
      }
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] += w[p] * (double)f[q];
      if (g.kappa) {
        float kv = __ldg(g.kappa + slot + c0 + off[p]);
        if (p == 0) corner_kappa = kv;
        v[8] += w[p] * (double)kv;
      }
    }
    // non-positive interpolated rho / pgas / kappa fall back to the anchor cell (:822-827)
    if (v[0] <= 0.0) v[0] = (double)corner[0];
    if (v[1] <= 0.0) v[1] = (double)corner[1];
    if (g.kappa && v[8] <= 0.0) v[8] = (double)corner_kappa;
  });
  return kSampleOk;
}

struct Plasma {
  double rho_cgs, n_e_cgs, pgas_cgs, theta_e, inv_theta_e, kb_tt_e_cgs, bb_cgs, sigma, beta_inv, b_sq;
  double ucon[4], bcon[4];  // Cartesian Kerr-Schild components
  bool value_cut;           // true: skip coupling (simulation_coefficients.cpp:361-375)
  bool b_zero;              // all three simulation field components vanish (:394)
};

// Plasma state of one sample (simulation_coefficients.cpp:286-408).  (x,y,z) CKS position, r its radius,
// inv_r = 1/r.  vectors: 0 = stop after the value cuts (cell values only); 1 = CKS u^mu, b^mu only for
// samples that couple to the radiation; 2 = always (the polarized transport needs the fluid frame at every
// sample).  Algebraically the reference's formulas; reciprocals are shared and divisions by parameters are
// folded into host-side constants (the results agree to a few ulp, far inside the 1e-6 image tolerance).
__device__ __forceinline__ void plasma_state(const RadParams &P, double x, double y, double z, double r,
                                             double inv_r, const Prims &pr, int vectors, Plasma &s) {
  const double a = P.a;
  double rho = pr.rho, pgas = pr.pgas, kappa = pr.kappa;
  double uu1 = pr.uu1, uu2 = pr.uu2, uu3 = pr.uu3, bb1 = pr.bb1, bb2 = pr.bb2, bb3 = pr.bb3;
  s.rho_cgs = rho * P.d_unit;
  s.pgas_cgs = pgas * P.e_unit;
  s.n_e_cgs = s.rho_cgs * P.n_e_factor;

  // simulation-coordinate metric (radiation_geometry.cpp:421-573); only the entries that are used
  double ucon_sim[4], bcon_sim[4];
  double r2 = r * r, a2 = a * a;
  double cth = z * inv_r;
  double cth2 = cth * cth;
  double sth2 = 1.0 - cth2;
  if (P.coord != 0) {
    // spherical Kerr-Schild: nonzero g_{tt} g_{tr} g_{tph} g_{rr} g_{rph} g_{thth} g_{phph}
    double sigma = r2 + a2 * cth2;
    double tr = 2.0 * r / sigma;
    double g00 = -(1.0 - tr), g01 = tr, g03 = -tr * a * sth2;
    double g11 = 1.0 + tr, g13 = -(1.0 + tr) * a * sth2, g22 = sigma;
    double g33 = (r2 + a2 + tr * a2 * sth2) * sth2;
    double uu0 = sqrt(1.0 + g11 * uu1 * uu1 + 2.0 * g13 * uu1 * uu3 + g22 * uu2 * uu2 + g33 * uu3 * uu3);
    // lapse = 1/sqrt(1 + tr), shift^r = tr/(1 + tr)
    double rs = rsqrt(g11);
    ucon_sim[0] = uu0 * (g11 * rs);
    ucon_sim[1] = uu1 - tr * (rs * rs) * ucon_sim[0];
    ucon_sim[2] = uu2;
    ucon_sim[3] = uu3;
    double ucov1 = g01 * ucon_sim[0] + g11 * ucon_sim[1] + g13 * ucon_sim[3];
    double ucov2 = g22 * ucon_sim[2];
    double ucov3 = g03 * ucon_sim[0] + g13 * ucon_sim[1] + g33 * ucon_sim[3];
    bcon_sim[0] = ucov1 * bb1 + ucov2 * bb2 + ucov3 * bb3;
    double inv_u0 = 1.0 / ucon_sim[0];
    bcon_sim[1] = (bb1 + bcon_sim[0] * ucon_sim[1]) * inv_u0;
    bcon_sim[2] = (bb2 + bcon_sim[0] * ucon_sim[2]) * inv_u0;
    bcon_sim[3] = (bb3 + bcon_sim[0] * ucon_sim[3]) * inv_u0;
    double bcov0 = g00 * bcon_sim[0] + g01 * bcon_sim[1] + g03 * bcon_sim[3];
    double bcov1 = g01 * bcon_sim[0] + g11 * bcon_sim[1] + g13 * bcon_sim[3];
    double bcov2 = g22 * bcon_sim[2];
    double bcov3 = g03 * bcon_sim[0] + g13 * bcon_sim[1] + g33 * bcon_sim[3];
    s.b_sq = bcov0 * bcon_sim[0] + bcov1 * bcon_sim[1] + bcov2 * bcon_sim[2] + bcov3 * bcon_sim[3];
  } else {
    // Cartesian Kerr-Schild simulation: g = eta + f l l
    double f = 2.0 * r2 * r / (r2 * r2 + a2 * z * z);
    double inv_ra2 = 1.0 / (r2 + a2);
    double l[4] = {1.0, (r * x + a * y) * inv_ra2, (r * y - a * x) * inv_ra2, cth};
    double uv[4] = {0.0, uu1, uu2, uu3};
    double lu = l[1] * uu1 + l[2] * uu2 + l[3] * uu3;
    double uu0 = sqrt(1.0 + uu1 * uu1 + uu2 * uu2 + uu3 * uu3 + f * lu * lu);
    double g11 = 1.0 + f;  // -g^{00}
    double rs = rsqrt(g11);
    ucon_sim[0] = uu0 * (g11 * rs);
    for (int q = 1; q < 4; q++) ucon_sim[q] = uv[q] - f * l[q] * (rs * rs) * ucon_sim[0];
    double lucon = l[0] * ucon_sim[0] + l[1] * ucon_sim[1] + l[2] * ucon_sim[2] + l[3] * ucon_sim[3];
    double ucov[4];
    ucov[0] = -ucon_sim[0] + f * l[0] * lucon;
    for (int q = 1; q < 4; q++) ucov[q] = ucon_sim[q] + f * l[q] * lucon;
    bcon_sim[0] = ucov[1] * bb1 + ucov[2] * bb2 + ucov[3] * bb3;
    double inv_u0 = 1.0 / ucon_sim[0];
    bcon_sim[1] = (bb1 + bcon_sim[0] * ucon_sim[1]) * inv_u0;
    bcon_sim[2] = (bb2 + bcon_sim[0] * ucon_sim[2]) * inv_u0;
    bcon_sim[3] = (bb3 + bcon_sim[0] * ucon_sim[3]) * inv_u0;
    double lb = l[0] * bcon_sim[0] + l[1] * bcon_sim[1] + l[2] * bcon_sim[2] + l[3] * bcon_sim[3];
    s.b_sq = -bcon_sim[0] * bcon_sim[0] + bcon_sim[1] * bcon_sim[1] + bcon_sim[2] * bcon_sim[2] +
             bcon_sim[3] * bcon_sim[3] + f * lb * lb;
  }
  s.bb_cgs = sqrt(s.b_sq) * P.b_unit;
  s.sigma = s.beta_inv = nan("");
  if (P.need_sigma_beta) {
    s.sigma = s.b_sq / rho;
    s.beta_inv = s.b_sq / (2.0 * pgas);
  }

  // electron temperature
  s.kb_tt_e_cgs = nan("");
  s.theta_e = s.inv_theta_e = nan("");
  const double me_c2 = phys::m_e * phys::c * phys::c;
  if (P.thermal_frac != 0.0 && P.plasma_model == 0) {
    // T_i/T_e = (R_high beta^-2 ... ) with beta^-1 = b^2 / (2 p): written over the common denominator so
    // that kT_e and its reciprocal come from one division
    double p2 = 4.0 * pgas * pgas, b4 = s.b_sq * s.b_sq;
    double rat_num = P.plasma_rat_high * p2 + P.plasma_rat_low * b4;  // (T_i/T_e) * (p2 + b4)
    double one_num = p2 + b4;
    double num, den;
    if (P.plasma_use_p) {
      num = (1.0 + P.plasma_ne_ni) * P.plasma_mu * phys::m_p * s.pgas_cgs * one_num;
      den = (rat_num + P.plasma_ne_ni * one_num) * s.rho_cgs;
    } else {
      num = (1.0 + P.plasma_ne_ni) * P.plasma_mu * phys::m_p * s.pgas_cgs * one_num;
      den = (P.plasma_gamma - 1.0) * s.rho_cgs *
            (rat_num / (P.plasma_gamma_i - 1.0) + P.plasma_ne_ni * one_num / (P.plasma_gamma_e - 1.0));
    }
    double q = 1.0 / (num * den);
    s.kb_tt_e_cgs = num * num * q;
    s.theta_e = s.kb_tt_e_cgs * (1.0 / me_c2);
    s.inv_theta_e = den * den * q * me_c2;
  }
  if (P.thermal_frac != 0.0 && P.plasma_model == 1) {
    double mu_e = P.plasma_mu * (1.0 + 1.0 / P.plasma_ne_ni);
    double rho_e = rho * phys::m_e / (mu_e * phys::m_p);
    double cb = cbrt(rho_e * kappa);
    s.theta_e = 1.0 / 5.0 * (sqrt(1.0 + 25.0 * cb * cb) - 1.0);
    s.inv_theta_e = 1.0 / s.theta_e;
    s.kb_tt_e_cgs = s.theta_e * me_c2;
  }

  s.value_cut = false;
  if (P.any_value_cut) {
    // sigma = b^2/rho and beta^-1 = b^2/(2 p) are compared in product form (rho, p > 0)
    s.value_cut =
        (P.cut_rho_min >= 0.0 && s.rho_cgs < P.cut_rho_min) || (P.cut_rho_max >= 0.0 && s.rho_cgs > P.cut_rho_max) ||
        (P.cut_n_e_min >= 0.0 && s.n_e_cgs < P.cut_n_e_min) || (P.cut_n_e_max >= 0.0 && s.n_e_cgs > P.cut_n_e_max) ||
        (P.cut_p_gas_min >= 0.0 && s.pgas_cgs < P.cut_p_gas_min) || (P.cut_p_gas_max >= 0.0 && s.pgas_cgs > P.cut_p_gas_max) ||
        (P.cut_theta_e_min >= 0.0 && s.theta_e < P.cut_theta_e_min) || (P.cut_theta_e_max >= 0.0 && s.theta_e > P.cut_theta_e_max) ||
        (P.cut_b_min >= 0.0 && s.bb_cgs < P.cut_b_min) || (P.cut_b_max >= 0.0 && s.bb_cgs > P.cut_b_max) ||
        (P.cut_sigma_min >= 0.0 && s.b_sq < P.cut_sigma_min * rho) || (P.cut_sigma_max >= 0.0 && s.b_sq > P.cut_sigma_max * rho) ||
        (P.cut_beta_inverse_min >= 0.0 && s.b_sq < P.cut_beta_inverse_min * (2.0 * pgas)) ||
        (P.cut_beta_inverse_max >= 0.0 && s.b_sq > P.cut_beta_inverse_max * (2.0 * pgas));
  }
  s.b_zero = bb1 == 0.0 && bb2 == 0.0 && bb3 == 0.0;
  if (vectors == 0 || (vectors == 1 && (s.value_cut || s.b_zero))) return;

  // to Cartesian Kerr-Schild (CoordinateJacobian, radiation_geometry.cpp:69-126)
  if (P.coord != 0) {
    // x^2 + y^2 = (r^2 + a^2) sin^2(theta):  sin(theta) and the azimuth come from two rsqrt
    double rc2 = x * x + y * y;
    double inv_ra = rsqrt(r2 + a2);
    double inv_rc = rc2 > 0.0 ? rsqrt(rc2) : 0.0;
    double sth = rc2 * inv_rc * inv_ra;
    double cx = rc2 > 0.0 ? x * inv_rc : 1.0, cy = rc2 > 0.0 ? y * inv_rc : 0.0;
    double cph = (cx * r + cy * a) * inv_ra;  // cos(atan2(y,x) - atan(a/r))
    double sph = (cy * r - cx * a) * inv_ra;
    double j11 = sth * cph, j12 = cth * (r * cph - a * sph), j13 = sth * (-r * sph - a * cph);
    double j21 = sth * sph, j22 = cth * (r * sph + a * cph), j23 = sth * (r * cph - a * sph);
    double j31 = cth, j32 = -r * sth;
    s.ucon[0] = ucon_sim[0];
    s.ucon[1] = j11 * ucon_sim[1] + j12 * ucon_sim[2] + j13 * ucon_sim[3];
    s.ucon[2] = j21 * ucon_sim[1] + j22 * ucon_sim[2] + j23 * ucon_sim[3];
    s.ucon[3] = j31 * ucon_sim[1] + j32 * ucon_sim[2];
    s.bcon[0] = bcon_sim[0];
    s.bcon[1] = j11 * bcon_sim[1] + j12 * bcon_sim[2] + j13 * bcon_sim[3];
    s.bcon[2] = j21 * bcon_sim[1] + j22 * bcon_sim[2] + j23 * bcon_sim[3];
    s.bcon[3] = j31 * bcon_sim[1] + j32 * bcon_sim[2];
  } else {
    for (int q = 0; q < 4; q++) {
      s.ucon[q] = ucon_sim[q];
      s.bcon[q] = bcon_sim[q];
    }
  }
}

__device__ __forceinline__ void cell_values_of(const Plasma &s, double cv[RAD_NUM_CELL_VALUES]) {
  cv[0] = s.rho_cgs; cv[1] = s.n_e_cgs; cv[2] = s.pgas_cgs; cv[3] = s.theta_e;
  cv[4] = s.bb_cgs; cv[5] = s.sigma; cv[6] = s.beta_inv;
}

// Proper length per unit affine parameter, sqrt(g_ij t^i t^j) with t^i the spatial projection of the
// momentum (unpolarized.cpp:118-130, rendering.cpp:86-99).  kc = covariant momentum.
__device__ __forceinline__ double proper_length_rate(const RadParams &P, double x, double y, double z,
                                                     const double kc[4]) {
  if (P.ray_flat) return sqrt(kc[1] * kc[1] + kc[2] * kc[2] + kc[3] * kc[3]);
  double a = P.a, a2 = a * a;
  double r = ks_radius(a, x, y, z), r2 = r * r;
  double f = 2.0 * r2 * r / (r2 * r2 + a2 * z * z);
  double l[3] = {(r * x + a * y) / (r2 + a2), (r * y - a * x) / (r2 + a2), z / r};
  // t_a = (g^{a mu} - g^{0a} g^{0 mu}/g^{00}) k_mu with g^{00} = -(1+f), g^{0a} = f l_a, g^{ab} = delta - f l_a l_b
  double g00 = -(1.0 + f);
  double lk = l[0] * kc[1] + l[1] * kc[2] + l[2] * kc[3];
  double t[3];
  for (int q = 0; q < 3; q++) {
    double g0a = f * l[q];
    double spatial = kc[1 + q] - g0a * lk;                     // g^{ab} k_b
    double corr = g0a * (f * lk) / g00;                        // g^{0a} g^{0b} k_b / g^{00}
    t[q] = spatial - corr;                                     // the mu = 0 terms cancel identically
  }
  double lt = l[0] * t[0] + l[1] * t[1] + l[2] * t[2];
  return sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + f * lt * lt);
}

// False-colour compositing of one sample into the (R,3) XYZ accumulators of a ray (rendering.cpp:101-166)
__device__ __forceinline__ void render_update(const RadParams &P, double *render, int64_t stride,
                                              const double prev[RAD_NUM_CELL_VALUES],
                                              const double cur[RAD_NUM_CELL_VALUES], double delta_length) {
  for (int im = 0; im < P.render_num_images; im++) {
    double *px = render + (size_t)(3 * im) * stride;
    double cx = px[0], cy = px[stride], cz = px[2 * stride];
    bool touched = false;
    for (int f = P.render_feature_start[im]; f < P.render_feature_start[im + 1]; f++) {
      int q = P.render_quantities[f];
      int type = P.render_types[f];
      double pv = prev[q], cv = cur[q];
      if (type == 0 && cv >= P.render_min_vals[f] && cv <= P.render_max_vals[f]) {
        double delta_tau = delta_length / P.render_tau_scales[f];
        if (delta_tau <= 100.0) {
          double en = exp(-delta_tau), em = expm1(delta_tau);
          cx = en * (cx + P.render_x_vals[f] * em);
          cy = en * (cy + P.render_y_vals[f] * em);
          cz = en * (cz + P.render_z_vals[f] * em);
        } else {
          cx = P.render_x_vals[f];
          cy = P.render_y_vals[f];
          cz = P.render_z_vals[f];
        }
        touched = true;
      }
      bool crossed = false;
      double th = P.render_thresh_vals[f];
      if ((type == 1 || type == 2) && pv < th && cv >= th) crossed = true;
      if ((type == 1 || type == 3) && pv > th && cv <= th) crossed = true;
      if (crossed) {
        double op = P.render_opacities[f];
        cx = (1.0 - op) * cx + op * P.render_x_vals[f];
        cy = (1.0 - op) * cy + op * P.render_y_vals[f];
        cz = (1.0 - op) * cz + op * P.render_z_vals[f];
        touched = true;
      }
    }
    if (touched) {
      px[0] = cx;
      px[stride] = cy;
      px[2 * stride] = cz;
    }
  }
}

}  // namespace rad
