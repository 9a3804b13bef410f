// Rays of a wave ordered by length for the radiation kernels.
//
// The radiation kernels give a thread a ray; a warp runs as long as its longest ray, and the three-stage polarized
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
