// Rays of a wave ordered by length for the radiation kernels.
//
// The radiation kernels give a thread a ray; a warp runs as long as its longest ray, and the polarized
// pipeline (radiate_pol_split.cu) launches every slab of 64 samples over all rays although beyond the median length
// most rays have already ended (mock snapshot, 1024^2: 32 % of all samples sit in slabs in which fewer than half of the
// rays are alive, and such a slab ran at about half the efficiency of a full one).  A stable counting sort of the rays
// by their slab count ceil(num / unit), longest first, turns "the rays alive in slab s" into a prefix of one list: the
// pipeline launches each slab over that prefix only and addresses its scratch by list position (dense, coalesced), and
// the fused kernels' warps hold rays of nearly equal length.  Stability keeps image neighbours together inside a
// bucket, so gathers still share cells.  Nothing about a ray's own arithmetic changes: images are bit for bit the same.
//
// Three launches: per-chunk bucket histograms (one warp per chunk of kChunk rays), one block scanning them in
// (bucket descending, chunk ascending) order, one scatter with warp-level stable ranks (__match_any_sync).
#include "device_types.cuh"

namespace {

constexpr int kChunk = 1024;       // rays per warp
constexpr int kMaxBuckets = 2048;  // shared-memory histogram of one warp

__device__ __forceinline__ int bucket_of(int num, int unit, int buckets) {
  int key = (num + unit - 1) / unit;   // slabs this ray lives in
  key = key < 0 ? 0 : (key >= buckets ? buckets - 1 : key);
  return buckets - 1 - key;            // longest rays first
}

__global__ void __launch_bounds__(32) order_hist_kernel(const int32_t *__restrict__ num, int64_t rays, int unit, int buckets,
                                                        int chunks, int32_t *__restrict__ hist) {
  __shared__ int h[kMaxBuckets];
  const int lane = threadIdx.x, chunk = blockIdx.x;
  for (int b = lane; b < buckets; b += 32) h[b] = 0;
  __syncwarp();
  const int64_t base = (int64_t)chunk * kChunk;
  for (int t = 0; t < kChunk; t += 32) {
    const int64_t m = base + t + lane;
    if (m < rays) atomicAdd(&h[bucket_of(num[m], unit, buckets)], 1);
  }
  __syncwarp();
  for (int b = lane; b < buckets; b += 32) hist[(size_t)b * chunks + chunk] = h[b];
}

// exclusive scan of hist in (bucket, chunk) order, in place; totals[b] = rays in bucket b
__global__ void __launch_bounds__(1024) order_scan_kernel(int32_t *__restrict__ hist, int buckets, int chunks,
                                                          int32_t *__restrict__ totals) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  const size_t n = (size_t)buckets * chunks;
  for (size_t start = 0; start < n; start += 1024) {
    const size_t idx = start + tid;
    const int v = idx < n ? hist[idx] : 0;
    int x = v;
    for (int off = 1; off < 32; off <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
      for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, off);
        if (lane >= off) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int before = carry + (warp > 0 ? warp_sum[warp - 1] : 0) + x - v;
    if (idx < n) hist[idx] = before;
    __syncthreads();
    if (tid == 1023) carry = before + v;
    __syncthreads();
  }
  // bucket totals from the scanned offsets: start of bucket b+1 minus start of bucket b
  for (int b = tid; b < buckets; b += 1024) {
    const int begin = hist[(size_t)b * chunks];
    const int end = b + 1 < buckets ? hist[(size_t)(b + 1) * chunks] : carry;
    totals[b] = end - begin;
  }
}

__global__ void __launch_bounds__(32) order_scatter_kernel(const int32_t *__restrict__ num, int64_t rays, int unit, int buckets,
                                                           int chunks, const int32_t *__restrict__ offsets,
                                                           int32_t *__restrict__ order) {
  __shared__ int cursor[kMaxBuckets];
  const int lane = threadIdx.x, chunk = blockIdx.x;
  for (int b = lane; b < buckets; b += 32) cursor[b] = offsets[(size_t)b * chunks + chunk];
  __syncwarp();
  const int64_t base = (int64_t)chunk * kChunk;
  for (int t = 0; t < kChunk; t += 32) {
    const int64_t m = base + t + lane;
    const bool valid = m < rays;
    const int b = valid ? bucket_of(num[m], unit, buckets) : -1 - lane;   // invalid lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int pos = 0;
    if (valid) pos = cursor[b] + rank;
    __syncwarp();
    if (valid && rank == __popc(peers) - 1) cursor[b] = pos + 1;   // the group's last lane advances the cursor
    __syncwarp();
    if (valid) order[pos] = (int32_t)m;
  }
}

}  // namespace

extern "C" int bl_ray_order_max_buckets(void) { return kMaxBuckets; }
extern "C" size_t bl_ray_order_workspace(int64_t rays, int buckets) {
  const int64_t chunks = (rays + kChunk - 1) / kChunk;
  return (size_t)buckets * (size_t)chunks + (size_t)buckets;   // int32 entries: histogram / offsets, then totals
}

// order[i], i < rays: ray indices sorted by ceil(num / unit) descending, stable; workspace + buckets*chunks holds the
// bucket totals (bucket b = rays with ceil(num / unit) == buckets - 1 - b).
extern "C" cudaError_t bl_launch_ray_order(const int32_t *num, int64_t rays, int unit, int buckets, int32_t *workspace,
                                           int32_t *order, cudaStream_t stream) {
  if (rays <= 0) return cudaSuccess;
  const int chunks = (int)((rays + kChunk - 1) / kChunk);
  int32_t *totals = workspace + (size_t)buckets * chunks;
  order_hist_kernel<<<chunks, 32, 0, stream>>>(num, rays, unit, buckets, chunks, workspace);
  order_scan_kernel<<<1, 1024, 0, stream>>>(workspace, buckets, chunks, totals);
  order_scatter_kernel<<<chunks, 32, 0, stream>>>(num, rays, unit, buckets, chunks, workspace, order);
  return cudaGetLastError();
}

// ---- longest rays first for the geodesic integrator -------------------------------------------------------------------
// The integrator's persistent warps take rays from a queue; what limits a far camera is the serial chain of its
// longest rays (example_formula: median 115 Dormand-Prince attempts per ray, but the 4 % that plunge towards the horizon
// take 3000 - 4400), so those should start first and the short ones fill the gaps (longest-processing-time-first).
// The cost of a ray is not known before it is traced, but it falls monotonically with the impact parameter
// b = |x cross p| / |p| of its initial condition: measured on both benchmark cameras, attempts per ray drop from
// > 1000 (formula) / 130 (simulation) inside b = 5 to ~100 / ~45 at the image corners.  Rays are therefore queued in
// order of increasing b (1024 buckets, stable): a scheduling hint only, results do not depend on it.
namespace {

__global__ void impact_parameter_kernel(const double *__restrict__ cam_pos, const double *__restrict__ cam_dir, int64_t rays,
                                        float *__restrict__ b_out, unsigned int *__restrict__ b_max_bits) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float b = 0.0f;
  if (m < rays) {
    const double x = cam_pos[4 * m + 1], y = cam_pos[4 * m + 2], z = cam_pos[4 * m + 3];
    const double px = cam_dir[4 * m + 1], py = cam_dir[4 * m + 2], pz = cam_dir[4 * m + 3];
    const double cx = y * pz - z * py, cy = z * px - x * pz, cz = x * py - y * px;
    const double p2 = px * px + py * py + pz * pz;
    b = p2 > 0.0 ? (float)sqrt((cx * cx + cy * cy + cz * cz) / p2) : 0.0f;
    if (!(b >= 0.0f) || isinf(b)) b = 0.0f;
    b_out[m] = b;
  }
  // non-negative floats order like their bit patterns
  unsigned int bits = __float_as_uint(b);
  for (int off = 16; off > 0; off >>= 1) {
    unsigned int o = __shfl_xor_sync(0xffffffffu, bits, off);
    bits = o > bits ? o : bits;
  }
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(b_max_bits, bits);
}

__global__ void impact_key_kernel(int32_t *__restrict__ keys_inout, int64_t rays, const unsigned int *__restrict__ b_max_bits,
                                  int buckets) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rays) return;
  const float b = __int_as_float(keys_inout[m]), b_max = __uint_as_float(*b_max_bits);
  int q = b_max > 0.0f ? (int)(b / b_max * (float)(buckets - 2)) : 0;
  q = q < 0 ? 0 : (q > buckets - 2 ? buckets - 2 : q);
  keys_inout[m] = buckets - 2 - q;   // smallest impact parameter = largest key = first in the list
}

}  // namespace

// order[i]: rays by increasing impact parameter.  keys: int32 scratch of `rays` entries (the level's sample_num array,
// which the integrator overwrites afterwards); b_max_bits: one zeroed unsigned int on the device.
extern "C" cudaError_t bl_launch_impact_order(const double *cam_pos, const double *cam_dir, int64_t rays, int32_t *keys,
                                              unsigned int *b_max_bits, int buckets, int32_t *workspace, int32_t *order,
                                              cudaStream_t stream) {
  if (rays <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((rays + 255) / 256);
  impact_parameter_kernel<<<grid, 256, 0, stream>>>(cam_pos, cam_dir, rays, reinterpret_cast<float *>(keys), b_max_bits);
  impact_key_kernel<<<grid, 256, 0, stream>>>(keys, rays, b_max_bits, buckets);
  return bl_launch_ray_order(keys, rays, 1, buckets, workspace, order, stream);
}
