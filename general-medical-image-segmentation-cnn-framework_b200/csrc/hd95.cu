// 95th-percentile Hausdorff distance of two binary masks on the GPU (utils/metric.py:29-32 calls
// monai.metrics.compute_hausdorff_distance(pred, gt, percentile=95, spacing=...)).
//
// MONAI 1.3.1 (requirements.txt; not vendored, not installed) computes, per direction, the Euclidean distance transform of
// the complement of one mask's EDGE set and samples it at the other mask's edge voxels; the edge set is
// seg ^ binary_erosion(seg) with the 6-neighbourhood cross and a zero border.  The distance-transform value at p is
// min_q |(p - q) * spacing| over the edge voxels q, which is what these kernels evaluate directly:
//   mask_edge_count / mask_edge_points : the edge voxels of a mask as physical coordinates (two passes: count, fill);
//   min_distances                      : for every point of A the distance to the nearest point of B, B tiled through
//                                        shared memory (|A| x |B| distance evaluations; edge sets are surfaces, 1e4-1e6 points).
// The percentile itself (linear interpolation, numpy semantics) is taken by the caller on the |A| distances.
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ bool is_edge(const uint8_t* __restrict__ m, int x, int y, int z, int W, int H, int D) {
  const int64_t i = (static_cast<int64_t>(x) * H + y) * D + z;
  if (!m[i]) return false;
  // eroded away <=> some 6-neighbour is background, the outside of the volume counting as background
  if (x == 0 || x == W - 1 || y == 0 || y == H - 1 || z == 0 || z == D - 1) return true;
  return !(m[i - static_cast<int64_t>(H) * D] && m[i + static_cast<int64_t>(H) * D] && m[i - D] && m[i + D] && m[i - 1] && m[i + 1]);
}

__global__ void __launch_bounds__(256) mask_edge_points_kernel(const uint8_t* __restrict__ m, int W, int H, int D, float sx, float sy,
                                                               float sz, float* __restrict__ pts, unsigned long long* __restrict__ count,
                                                               unsigned long long cap) {
  const int64_t total = static_cast<int64_t>(W) * H * D;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int z = static_cast<int>(i % D), y = static_cast<int>((i / D) % H), x = static_cast<int>(i / (static_cast<int64_t>(D) * H));
    if (!is_edge(m, x, y, z, W, H, D)) continue;
    const unsigned long long k = atomicAdd(count, 1ULL);
    if (pts != nullptr && k < cap) {
      pts[3 * k + 0] = x * sx;
      pts[3 * k + 1] = y * sy;
      pts[3 * k + 2] = z * sz;
    }
  }
}

constexpr int kTileB = 1024;

__global__ void __launch_bounds__(256) min_distances_kernel(const float* __restrict__ a, long long na, const float* __restrict__ b,
                                                            long long nb, float* __restrict__ out) {
  __shared__ float sb[kTileB * 3];
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (i < na) {
    ax = a[3 * i];
    ay = a[3 * i + 1];
    az = a[3 * i + 2];
  }
  float best = INFINITY;
  for (long long j0 = 0; j0 < nb; j0 += kTileB) {
    const int cnt = static_cast<int>(min(static_cast<long long>(kTileB), nb - j0));
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += blockDim.x) sb[t] = b[3 * j0 + t];
    __syncthreads();
    if (i < na) {
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const float dx = ax - sb[3 * j], dy = ay - sb[3 * j + 1], dz = az - sb[3 * j + 2];
        best = fminf(best, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      }
    }
  }
  if (i < na) out[i] = sqrtf(best);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200seg_mask_edge_points(const uint8_t* mask, int w, int h, int d, float sx, float sy, float sz, float* points,
                             unsigned long long capacity, unsigned long long* count, void* stream) {
  B200_CHECK_ARG(mask && count && w > 0 && h > 0 && d > 0, "mask_edge_points: bad arguments");
  const int64_t total = static_cast<int64_t>(w) * h * d;
  mask_edge_points_kernel<<<grid_for(total, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mask, w, h, d, sx, sy, sz, points, count, capacity);
  B200_CHECK_LAUNCH("mask_edge_points");
  return 0;
}

int b200seg_min_distances(const float* a, int64_t na, const float* b, int64_t nb, float* out, void* stream) {
  B200_CHECK_ARG(a && b && out && na > 0 && nb > 0, "min_distances: empty point sets are the caller's case (inf / nan)");
  min_distances_kernel<<<static_cast<unsigned>((na + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, na, b, nb, out);
  B200_CHECK_LAUNCH("min_distances");
  return 0;
}

}  // extern "C"
