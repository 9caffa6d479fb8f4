// GPU side of the training data path (dataloader.py:52-67): per-volume z-normalisation statistics and uniform patch
// cropping from volumes that stay resident in HBM.  Replaces torchio's ZNormalization transform + UniformSampler + Queue
// (host-side, num_workers = 0 in the reference) for data that fits device memory -- 180 GB holds ~600 volumes of 512x512x256
// fp32.  Bandwidth-bound: every patch voxel is read once and written once.
#include "common.cuh"

namespace b200 {

// sums[0] += sum x, sums[1] += sum x^2 (double), over n floats
__global__ void __launch_bounds__(256) volume_stats_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ sums) {
  double s = 0.0, q = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n4 = (reinterpret_cast<uintptr_t>(x) & 15) == 0 ? n / 4 : 0;
  for (int64_t j = i; j < n4; j += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[j];
    // fp32 partial per 4 values, double across iterations: 1e8-voxel volumes keep ~12 significant digits
    s += static_cast<double>((v.x + v.y) + (v.z + v.w));
    q += static_cast<double>(fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w));
  }
  for (int64_t j = n4 * 4 + i; j < n; j += stride) {
    const float v = x[j];
    s += v;
    q += static_cast<double>(v) * v;
  }
  __shared__ double red[2];
  if (threadIdx.x < 2) red[threadIdx.x] = 0.0;
  __syncthreads();
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&red[0], s);
    atomicAdd(&red[1], q);
  }
  __syncthreads();
  if (threadIdx.x < 2) atomicAdd(&sums[threadIdx.x], red[threadIdx.x]);
}

// torchio ZNormalization (masking_method=None): mean over all voxels, torch.std = UNBIASED standard deviation
__global__ void znorm_finalize_kernel(const double* __restrict__ sums, double n, float* __restrict__ out) {
  const double mean = sums[0] / n;
  double var = (sums[1] - sums[0] * mean) / (n - 1.0);
  if (var < 0) var = 0;
  out[0] = static_cast<float>(mean);
  out[1] = static_cast<float>(1.0 / sqrt(var));
}

// out[c][i][j][k] = (vol[c][x0+i][y0+j][z0+k] - mean) * inv_std   (T = float), or a plain copy (T = uint8_t labels)
template <typename T>
__global__ void __launch_bounds__(256) crop_patch_kernel(const T* __restrict__ vol, int C, int W, int H, int D, int x0, int y0,
                                                         int z0, int pw, int ph, int pd, const float* __restrict__ norm,
                                                         T* __restrict__ out) {
  const float mean = norm ? norm[0] : 0.f, inv_std = norm ? norm[1] : 1.f;
  const int64_t total = static_cast<int64_t>(C) * pw * ph * pd;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % pd);
    int64_t r = i / pd;
    const int j = static_cast<int>(r % ph);
    r /= ph;
    const int ii = static_cast<int>(r % pw);
    const int c = static_cast<int>(r / pw);
    const T v = vol[((static_cast<int64_t>(c) * W + x0 + ii) * H + y0 + j) * D + z0 + k];
    if constexpr (sizeof(T) == 4) out[i] = (v - mean) * inv_std;
    else out[i] = v;
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200seg_volume_stats(const float* x, int64_t n, double* sums, void* stream) {
  B200_CHECK_ARG(x && sums && n > 0, "volume_stats: bad arguments");
  volume_stats_kernel<<<grid_for(n / 4 + 1, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, sums);
  B200_CHECK_LAUNCH("volume_stats");
  return 0;
}

int b200seg_znorm_finalize(const double* sums, int64_t n, float* mean_inv_std, void* stream) {
  B200_CHECK_ARG(sums && mean_inv_std && n > 1, "znorm_finalize: needs at least two voxels");
  znorm_finalize_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(sums, static_cast<double>(n), mean_inv_std);
  B200_CHECK_LAUNCH("znorm_finalize");
  return 0;
}

int b200seg_crop_patch(const void* vol, int is_label, int c, int w, int h, int d, int x0, int y0, int z0, int pw, int ph, int pd,
                       const float* mean_inv_std, void* out, void* stream) {
  B200_CHECK_ARG(vol && out && c > 0 && pw > 0 && ph > 0 && pd > 0, "crop_patch: bad arguments");
  B200_CHECK_ARG(x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + pw <= w && y0 + ph <= h && z0 + pd <= d,
                 "crop_patch: patch [%d:%d, %d:%d, %d:%d] leaves the %dx%dx%d volume", x0, x0 + pw, y0, y0 + ph, z0, z0 + pd, w, h, d);
  const int64_t total = static_cast<int64_t>(c) * pw * ph * pd;
  auto st = static_cast<cudaStream_t>(stream);
  if (is_label)
    crop_patch_kernel<uint8_t><<<grid_for(total, 256, kNumSMs * 8), 256, 0, st>>>(static_cast<const uint8_t*>(vol), c, w, h, d, x0, y0,
                                                                              z0, pw, ph, pd, nullptr, static_cast<uint8_t*>(out));
  else
    crop_patch_kernel<float><<<grid_for(total, 256, kNumSMs * 8), 256, 0, st>>>(static_cast<const float*>(vol), c, w, h, d, x0, y0, z0,
                                                                            pw, ph, pd, mean_inv_std, static_cast<float*>(out));
  B200_CHECK_LAUNCH("crop_patch");
  return 0;
}

}  // extern "C"
