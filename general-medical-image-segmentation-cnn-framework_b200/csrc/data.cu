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

// ---- Pad3d 'reflect' / 'replicate' (utils/convolution.py:78-86 = F.pad(x, 6*[pad], mode)) on NDHWC bf16 rows ----------
// source index of padded position o along an axis of n voxels (pad p on both sides)
__device__ __forceinline__ int pad_src(int o, int n, int p, int mode) {
  int i = o - p;
  if (mode == 1) {            // reflect (no edge repeat): -1 -> 1, n -> n-2
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
  } else {                    // replicate
    i = i < 0 ? 0 : (i >= n ? n - 1 : i);
  }
  return i;
}

__global__ void __launch_bounds__(256) pad3d_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t xp, __nv_bfloat16* __restrict__ y,
                                                        int64_t yp, int n, int d, int h, int w, int c, int p, int mode) {
  const int od = d + 2 * p, oh = h + 2 * p, ow = w + 2 * p;
  const int64_t total = static_cast<int64_t>(n) * od * oh * ow * c;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    int64_t v = i / c;
    const int xo = static_cast<int>(v % ow);
    v /= ow;
    const int yo = static_cast<int>(v % oh);
    v /= oh;
    const int zo = static_cast<int>(v % od);
    const int nn = static_cast<int>(v / od);
    const int64_t src = ((static_cast<int64_t>(nn) * d + pad_src(zo, d, p, mode)) * h + pad_src(yo, h, p, mode)) * w +
                        pad_src(xo, w, p, mode);
    y[(i / c) * yp + ch] = x[src * xp + ch];
  }
}

// Adjoint: dx[i] = sum of dy over every padded position that reads i.  Along one axis those are: the centre i + p, and
// reflect: p - i (1 <= i <= p) and 2(n-1) + p - i (n-1-p <= i <= n-2); replicate: 0..p-1 for i = 0, n+p..n+2p-1 for i = n-1.
__device__ __forceinline__ int pad_sources(int i, int n, int p, int mode, int* out) {
  int k = 0;
  out[k++] = i + p;
  if (mode == 1) {
    if (i >= 1 && i <= p) out[k++] = p - i;
    if (i >= n - 1 - p && i <= n - 2) out[k++] = 2 * (n - 1) + p - i;
  } else {
    if (i == 0)
      for (int o = 0; o < p; ++o) out[k++] = o;
    if (i == n - 1)
      for (int o = n + p; o < n + 2 * p; ++o) out[k++] = o;
  }
  return k;
}

constexpr int kMaxPad = 8;

__global__ void __launch_bounds__(256) pad3d_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dyp, __nv_bfloat16* __restrict__ dx,
                                                        int64_t dxp, int n, int d, int h, int w, int c, int p, int mode) {
  const int oh = h + 2 * p, ow = w + 2 * p, od = d + 2 * p;
  const int64_t total = static_cast<int64_t>(n) * d * h * w * c;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    int64_t v = i / c;
    const int xi = static_cast<int>(v % w);
    v /= w;
    const int yi = static_cast<int>(v % h);
    v /= h;
    const int zi = static_cast<int>(v % d);
    const int nn = static_cast<int>(v / d);
    int zs[2 * kMaxPad + 1], ys[2 * kMaxPad + 1], xs[2 * kMaxPad + 1];
    const int nz = pad_sources(zi, d, p, mode, zs), ny = pad_sources(yi, h, p, mode, ys), nx = pad_sources(xi, w, p, mode, xs);
    float acc = 0.f;
    for (int a = 0; a < nz; ++a)
      for (int b = 0; b < ny; ++b)
        for (int e = 0; e < nx; ++e)
          acc += __bfloat162float(dy[(((static_cast<int64_t>(nn) * od + zs[a]) * oh + ys[b]) * ow + xs[e]) * dyp + ch]);
    dx[(i / c) * dxp + ch] = __float2bfloat16(acc);
  }
}

// ------------------------------------------------------------------------------------------------ space <-> depth
// Non-overlapping strided convolutions (kernel k <= stride s, no padding: csrnet.py:115-154 uses Conv3d(k3, s4) and
// ConvTranspose3d(k4, s4)) are plain GEMMs once the k^3 taps of every s^3 cell sit next to each other in the channel
// dimension: y[n, oz, oy, ox, ((a*k + b)*k + e)*C + c] = x[n, oz*s + a, oy*s + b, ox*s + e, c].  V = channels per thread.
template <int V>
__global__ void space_to_depth_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch, __nv_bfloat16* __restrict__ y,
                                      int n, int d, int h, int w, int C, int k, int s, int od, int oh, int ow) {
  const int cv = C / V;
  const int k3 = k * k * k;
  const int64_t total = static_cast<int64_t>(n) * od * oh * ow * k3 * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    int64_t r = i / cv;
    const int t = static_cast<int>(r % k3);
    r /= k3;
    const int ox = static_cast<int>(r % ow);
    r /= ow;
    const int oy = static_cast<int>(r % oh);
    r /= oh;
    const int oz = static_cast<int>(r % od);
    const int nn = static_cast<int>(r / od);
    const int e = t % k, b = (t / k) % k, a = t / (k * k);
    const int64_t src = ((static_cast<int64_t>(nn) * d + oz * s + a) * h + oy * s + b) * w + ox * s + e;
    if constexpr (V == 8) st8(y + i * 8, ld8(x + src * x_pitch + ch * 8));
    else y[i] = x[src * x_pitch + ch];
  }
}

// The adjoint / inverse: x[n, z, yy, xx, c] = y[n, z/s, yy/s, xx/s, tap(z%s, yy%s, xx%s), c], zero where the position is not
// covered (index >= k inside its cell, or a cell beyond the last output).
template <int V>
__global__ void depth_to_space_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                                      int n, int d, int h, int w, int C, int k, int s, int od, int oh, int ow) {
  const int cv = C / V;
  const int k3 = k * k * k;
  const int64_t total = static_cast<int64_t>(n) * d * h * w * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    int64_t r = i / cv;
    const int64_t row = r;
    const int xx = static_cast<int>(r % w);
    r /= w;
    const int yy = static_cast<int>(r % h);
    r /= h;
    const int z = static_cast<int>(r % d);
    const int nn = static_cast<int>(r / d);
    const int oz = z / s, a = z % s, oy = yy / s, b = yy % s, ox = xx / s, e = xx % s;
    const bool ok = a < k && b < k && e < k && oz < od && oy < oh && ox < ow;
    const int64_t src = ((((static_cast<int64_t>(nn) * od + oz) * oh + oy) * ow + ox) * k3 + (a * k + b) * k + e) * cv + ch;
    if constexpr (V == 8) {
      bf16x8 v;
      if (ok) v = ld8(y + src * 8);
      else {
        const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        v = pack8(z8);
      }
      st8(x + row * x_pitch + ch * 8, v);
    } else {
      x[row * x_pitch + ch] = ok ? y[src] : __float2bfloat16(0.f);
    }
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

static int check_pad3d(int n, int d, int h, int w, int c, int pad, int mode, const char* who) {
  B200_CHECK_ARG(n > 0 && d > 0 && h > 0 && w > 0 && c > 0 && pad >= 0 && pad <= kMaxPad, "%s: bad extents (pad <= %d)", who, kMaxPad);
  B200_CHECK_ARG(mode == 1 || mode == 2, "%s: mode must be 1 (reflect) or 2 (replicate); 'constant' is folded into the convolutions", who);
  B200_CHECK_ARG(mode != 1 || (pad < d && pad < h && pad < w), "%s: reflect padding must be smaller than every extent", who);
  return 0;
}

int b200seg_pad3d_fwd(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int n, int d, int h, int w, int c, int pad, int mode,
                      void* stream) {
  B200_CHECK_ARG(x && y && x_pitch >= c && y_pitch >= c, "pad3d_fwd: bad buffers");
  if (int rc = check_pad3d(n, d, h, w, c, pad, mode, "pad3d_fwd")) return rc;
  const int64_t total = static_cast<int64_t>(n) * (d + 2 * pad) * (h + 2 * pad) * (w + 2 * pad) * c;
  pad3d_fwd_kernel<<<grid_for(total, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), y_pitch, n, d, h, w, c, pad, mode);
  B200_CHECK_LAUNCH("pad3d_fwd");
  return 0;
}

int b200seg_pad3d_bwd(const void* dy, int64_t dy_pitch, void* dx, int64_t dx_pitch, int n, int d, int h, int w, int c, int pad,
                      int mode, void* stream) {
  B200_CHECK_ARG(dy && dx && dy_pitch >= c && dx_pitch >= c, "pad3d_bwd: bad buffers");
  if (int rc = check_pad3d(n, d, h, w, c, pad, mode, "pad3d_bwd")) return rc;
  const int64_t total = static_cast<int64_t>(n) * d * h * w * c;
  pad3d_bwd_kernel<<<grid_for(total, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), dy_pitch, static_cast<__nv_bfloat16*>(dx), dx_pitch, n, d, h, w, c, pad, mode);
  B200_CHECK_LAUNCH("pad3d_bwd");
  return 0;
}

static int check_s2d(int n, int d, int h, int w, int c, int k, int s, int od, int oh, int ow, const char* who) {
  B200_CHECK_ARG(n > 0 && c > 0 && k >= 1 && k <= s && s >= 1 && d >= k && h >= k && w >= k, "%s: bad geometry (needs k <= s)", who);
  B200_CHECK_ARG(od >= 1 && oh >= 1 && ow >= 1 && (od - 1) * s + k <= d && (oh - 1) * s + k <= h && (ow - 1) * s + k <= w,
                 "%s: output extent does not fit the input", who);
  return 0;
}

int b200seg_space_to_depth(const void* x, int64_t x_pitch, void* y, int n, int d, int h, int w, int c, int k, int s, int od,
                           int oh, int ow, void* stream) {
  B200_CHECK_ARG(x && y && x_pitch >= c, "space_to_depth: bad buffers");
  if (int rc = check_s2d(n, d, h, w, c, k, s, od, oh, ow, "space_to_depth")) return rc;
  const bool vec = c % 8 == 0 && x_pitch % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const int64_t total = static_cast<int64_t>(n) * od * oh * ow * k * k * k * (vec ? c / 8 : c);
  const int grid = grid_for(total, 256, kNumSMs * 8);
  if (vec)
    space_to_depth_kernel<8><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), n, d, h, w, c, k, s, od, oh, ow);
  else
    space_to_depth_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), n, d, h, w, c, k, s, od, oh, ow);
  B200_CHECK_LAUNCH("space_to_depth");
  return 0;
}

int b200seg_depth_to_space(const void* y, void* x, int64_t x_pitch, int n, int d, int h, int w, int c, int k, int s, int od,
                           int oh, int ow, void* stream) {
  B200_CHECK_ARG(x && y && x_pitch >= c, "depth_to_space: bad buffers");
  if (int rc = check_s2d(n, d, h, w, c, k, s, od, oh, ow, "depth_to_space")) return rc;
  const bool vec = c % 8 == 0 && x_pitch % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const int64_t total = static_cast<int64_t>(n) * d * h * w * (vec ? c / 8 : c);
  const int grid = grid_for(total, 256, kNumSMs * 8);
  if (vec)
    depth_to_space_kernel<8><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(x), x_pitch, n, d, h, w, c, k, s, od, oh, ow);
  else
    depth_to_space_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(x), x_pitch, n, d, h, w, c, k, s, od, oh, ow);
  B200_CHECK_LAUNCH("depth_to_space");
  return 0;
}

}  // extern "C"
