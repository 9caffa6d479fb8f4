// Attention gates of the ER-Net / RE-Net / Double-UNet model files (models/three_d/ER_net.py, RE_net.py, Double_Unet.py,
// SE.py): all of them are bandwidth-bound element-wise passes or per-(sample, channel) reductions.
//
//   * reverse attention (ER_net.py:184-187, RE_net.py:118-121):  g = ConvTranspose3d(1, 1, k2, s2)(Conv3d(C, 1, k1)(coarse));
//     out = fine * (1 - sigmoid(g)) + fine = fine * (2 - sigmoid(g)).  The C -> 1 projection is the head kernel
//     (head.cu); here: the single-channel k2s2 transposed convolution on fp32 maps and the gating pass.
//   * channel blend (SE.py:41-49 `x + x * y`, ER_net.py:86-105 selective fusion `x1 * a1 + x2 * a2`):
//     out[v][c] = x1[v][c] * w1[n][c] (+ x2[v][c] * w2[n][c]) with per-(sample, channel) fp32 weights, and its two
//     backward passes (per-(sample, channel) dot products, then the data gradients).
//   * sigmoid on fp32 class maps (RE_net.py:158).
#include <algorithm>

#include "common.cuh"

namespace b200 {

template <int V>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float (&f)[V]) {
  static_assert(V == 8, "8 bf16 = one 128-bit access");
  unpack8(ld8(p), f);
}
template <int V>
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&f)[V]) {
  static_assert(V == 8, "8 bf16 = one 128-bit access");
  st8(p, pack8(f));
}

// ------------------------------------------------------------------------------------- ConvTranspose3d(1, 1, k2, s2)
// in [planes][d][h][w] fp32 -> out [planes][2d][2h][2w]: out[2z+a][2y+b][2x+c] = in[z][y][x] * w[a][b][c] + bias
__global__ void convt1_k2s2_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                       const float* __restrict__ bias, float* __restrict__ out, int64_t planes, int d, int h,
                                       int wd) {
  const int64_t total = planes * d * h * wd;
  float wr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wr[i] = w[i];
  const float b = bias ? bias[0] : 0.f;
  const int W2 = 2 * wd, H2 = 2 * h;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % wd);
    const int y = static_cast<int>((i / wd) % h);
    const int64_t pz = i / (static_cast<int64_t>(wd) * h);      // plane * d + z
    const float v = in[i];
    float* o = out + ((pz * 2) * H2 + 2 * y) * W2 + 2 * x;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        float2 q;
        q.x = fmaf(v, wr[a * 4 + bb * 2 + 0], b);
        q.y = fmaf(v, wr[a * 4 + bb * 2 + 1], b);
        *reinterpret_cast<float2*>(o + (static_cast<int64_t>(a) * H2 + bb) * W2) = q;
      }
  }
}

// din[z][y][x] = sum_abc dout[2z+a][2y+b][2x+c] * w[abc];  sums[abc] += sum dout * in;  sums[8] += sum dout
__global__ void convt1_k2s2_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ in,
                                       const float* __restrict__ w, float* __restrict__ din, float* __restrict__ sums,
                                       int64_t planes, int d, int h, int wd) {
  const int64_t total = planes * d * h * wd;
  float wr[8], acc[9];
#pragma unroll
  for (int i = 0; i < 8; ++i) wr[i] = w[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  const int W2 = 2 * wd, H2 = 2 * h;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % wd);
    const int y = static_cast<int>((i / wd) % h);
    const int64_t pz = i / (static_cast<int64_t>(wd) * h);
    const float v = in[i];
    const float* o = dout + ((pz * 2) * H2 + 2 * y) * W2 + 2 * x;
    float g = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const float2 q = *reinterpret_cast<const float2*>(o + (static_cast<int64_t>(a) * H2 + bb) * W2);
        g = fmaf(q.x, wr[a * 4 + bb * 2 + 0], g);
        g = fmaf(q.y, wr[a * 4 + bb * 2 + 1], g);
        acc[a * 4 + bb * 2 + 0] = fmaf(q.x, v, acc[a * 4 + bb * 2 + 0]);
        acc[a * 4 + bb * 2 + 1] = fmaf(q.y, v, acc[a * 4 + bb * 2 + 1]);
        acc[8] += q.x + q.y;
      }
    if (din) din[i] = g;
  }
  __shared__ float red[9];
  if (threadIdx.x < 9) red[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const float s = warp_sum(acc[j]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[j], s);
  }
  __syncthreads();
  if (threadIdx.x < 9) atomicAdd(&sums[threadIdx.x], red[threadIdx.x]);
}

// ------------------------------------------------------------------------------------- reverse-attention gate
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// out[v][c] = fine[v][c] * (2 - sigmoid(g[v]));  one thread per (voxel, 8 channels)
__global__ void reverse_gate_fwd_kernel(const __nv_bfloat16* __restrict__ fine, int64_t f_pitch, const float* __restrict__ g,
                                        __nv_bfloat16* __restrict__ out, int64_t o_pitch, int64_t rows, int C) {
  const int cv = C >> 3;
  const int64_t total = rows * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    const int64_t row = i / cv;
    const float s = 2.f - sigmoidf_(g[row]);
    float f[8];
    load_vec<8>(fine + row * f_pitch + ch * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= s;
    store_vec<8>(out + row * o_pitch + ch * 8, f);
  }
}

// dfine[v][c] = dout[v][c] * (2 - s);  dg[v] = -s (1 - s) * sum_c dout[v][c] * fine[v][c];  a group of C/8 threads per voxel
// (C / 8 is a power of two <= 32), the dot product closes with shuffles inside the group
__global__ void reverse_gate_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int64_t d_pitch,
                                        const __nv_bfloat16* __restrict__ fine, int64_t f_pitch, const float* __restrict__ g,
                                        __nv_bfloat16* __restrict__ dfine, int64_t df_pitch, float* __restrict__ dg,
                                        int64_t rows, int C) {
  const int cv = C >> 3;
  const int64_t total = ((rows * cv + 31) / 32) * 32;    // whole warps stay in the loop: the shuffles need every lane
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    const int64_t row = i / cv;
    const bool ok = row < rows;
    float dot = 0.f;
    if (ok) {
      const float sg = sigmoidf_(g[row]);
      const float s = 2.f - sg;
      float a[8], f[8];
      load_vec<8>(dout + row * d_pitch + ch * 8, a);
      load_vec<8>(fine + row * f_pitch + ch * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dot = fmaf(a[j], f[j], dot);
        a[j] *= s;
      }
      store_vec<8>(dfine + row * df_pitch + ch * 8, a);
      dot *= -sg * (1.f - sg);
    }
    for (int o = cv >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (ok && ch == 0) dg[row] = dot;
  }
}

// ------------------------------------------------------------------------------------- per-(sample, channel) blends
// out = x1 * w1[n][c] (+ x2 * w2[n][c])
__global__ void channel_blend_fwd_kernel(const __nv_bfloat16* __restrict__ x1, int64_t p1, const float* __restrict__ w1,
                                         const __nv_bfloat16* __restrict__ x2, int64_t p2, const float* __restrict__ w2,
                                         __nv_bfloat16* __restrict__ out, int64_t po, int64_t rows_per_sample, int n, int C) {
  const int cv = C >> 3;
  const int64_t total = rows_per_sample * n * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    const int64_t row = i / cv;
    const int s = static_cast<int>(row / rows_per_sample);
    float a[8], wa[8];
    load_vec<8>(x1 + row * p1 + ch * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) wa[j] = w1[static_cast<int64_t>(s) * C + ch * 8 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= wa[j];
    if (x2) {
      float b[8];
      load_vec<8>(x2 + row * p2 + ch * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fmaf(b[j], w2[static_cast<int64_t>(s) * C + ch * 8 + j], a[j]);
    }
    store_vec<8>(out + row * po + ch * 8, a);
  }
}

// dots1[n][c] = sum_v dout * x1, dots2[n][c] = sum_v dout * x2.  grid = (row blocks, n); a thread keeps one 8-channel chunk.
__global__ void __launch_bounds__(256) channel_blend_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ dout, int64_t pd, const __nv_bfloat16* __restrict__ x1, int64_t p1,
    const __nv_bfloat16* __restrict__ x2, int64_t p2, float* __restrict__ dots1, float* __restrict__ dots2,
    int64_t rows_per_sample, int C) {
  extern __shared__ float sred[];      // [2][C]
  const int cv = C >> 3;               // host guarantees blockDim.x % cv == 0
  const int s = blockIdx.y;
  const int ch = threadIdx.x % cv;
  const int rpb = blockDim.x / cv;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  const int64_t base = static_cast<int64_t>(s) * rows_per_sample;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * rpb + threadIdx.x / cv; r < rows_per_sample;
       r += static_cast<int64_t>(gridDim.x) * rpb) {
    float g[8], a[8];
    load_vec<8>(dout + (base + r) * pd + ch * 8, g);
    load_vec<8>(x1 + (base + r) * p1 + ch * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) a1[j] = fmaf(g[j], a[j], a1[j]);
    if (x2) {
      load_vec<8>(x2 + (base + r) * p2 + ch * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) a2[j] = fmaf(g[j], a[j], a2[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sred[ch * 8 + j], a1[j]);
    if (x2) atomicAdd(&sred[C + ch * 8 + j], a2[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dots1[static_cast<int64_t>(s) * C + i], sred[i]);
    if (x2) atomicAdd(&dots2[static_cast<int64_t>(s) * C + i], sred[C + i]);
  }
}

// dx1 = dout * w1[n][c] + add[n][c], dx2 = dout * w2[n][c] + add[n][c]  (`add`: the gradient that reaches the inputs through
// the global average pool feeding the weights, already divided by the voxel count; may be null)
__global__ void channel_blend_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, int64_t pd,
                                               const float* __restrict__ w1, const float* __restrict__ w2,
                                               const float* __restrict__ add, __nv_bfloat16* __restrict__ dx1, int64_t p1,
                                               __nv_bfloat16* __restrict__ dx2, int64_t p2, int64_t rows_per_sample, int n,
                                               int C) {
  const int cv = C >> 3;
  const int64_t total = rows_per_sample * n * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    const int64_t row = i / cv;
    const int64_t sc = static_cast<int64_t>(row / rows_per_sample) * C + ch * 8;
    float g[8], o[8];
    load_vec<8>(dout + row * pd + ch * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(g[j], w1[sc + j], add ? add[sc + j] : 0.f);
    store_vec<8>(dx1 + row * p1 + ch * 8, o);
    if (dx2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(g[j], w2[sc + j], add ? add[sc + j] : 0.f);
      store_vec<8>(dx2 + row * p2 + ch * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------- sigmoid on fp32 maps
__global__ void f32_sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t numel) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = 1.f / (1.f + expf(-x[i]));
}
// dx = dy * y * (1 - y)
__global__ void f32_sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                       int64_t numel) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float s = y[i];
    dx[i] = dy[i] * s * (1.f - s);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200seg_convt1_k2s2_fwd(const float* in, const float* w, const float* bias, float* out, int64_t planes, int d, int h,
                            int wd, void* stream) {
  B200_CHECK_ARG(in && w && out && planes > 0 && d > 0 && h > 0 && wd > 0, "convt1_k2s2_fwd: bad arguments");
  convt1_k2s2_fwd_kernel<<<grid_for(planes * d * h * wd, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, w, bias, out, planes, d, h, wd);
  B200_CHECK_LAUNCH("convt1_k2s2_fwd");
  return 0;
}

int b200seg_convt1_k2s2_bwd(const float* dout, const float* in, const float* w, float* din, float* sums, int64_t planes,
                            int d, int h, int wd, void* stream) {
  B200_CHECK_ARG(dout && in && w && sums && planes > 0 && d > 0 && h > 0 && wd > 0, "convt1_k2s2_bwd: bad arguments");
  convt1_k2s2_bwd_kernel<<<grid_for(planes * d * h * wd, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dout, in, w, din, sums, planes, d, h, wd);
  B200_CHECK_LAUNCH("convt1_k2s2_bwd");
  return 0;
}

static bool gate_channels_ok(int c) { return c >= 8 && c <= 256 && (c & (c - 1)) == 0; }

int b200seg_reverse_gate_fwd(const void* fine, int64_t fine_pitch, const float* g, void* out, int64_t out_pitch,
                             int64_t rows, int c, void* stream) {
  B200_CHECK_ARG(fine && g && out && rows > 0 && c % 8 == 0 && c > 0 && fine_pitch % 8 == 0 && out_pitch % 8 == 0,
                 "reverse_gate_fwd: bad arguments (C must be a multiple of 8)");
  reverse_gate_fwd_kernel<<<grid_for(rows * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(fine), fine_pitch, g, static_cast<__nv_bfloat16*>(out), out_pitch, rows, c);
  B200_CHECK_LAUNCH("reverse_gate_fwd");
  return 0;
}

int b200seg_reverse_gate_bwd(const void* dout, int64_t dout_pitch, const void* fine, int64_t fine_pitch, const float* g,
                             void* dfine, int64_t dfine_pitch, float* dg, int64_t rows, int c, void* stream) {
  B200_CHECK_ARG(dout && fine && g && dfine && dg && rows > 0 && gate_channels_ok(c) && dout_pitch % 8 == 0 &&
                     fine_pitch % 8 == 0 && dfine_pitch % 8 == 0,
                 "reverse_gate_bwd: bad arguments (C must be a power of two in [8, 256])");
  reverse_gate_bwd_kernel<<<grid_for(rows * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), dout_pitch, static_cast<const __nv_bfloat16*>(fine), fine_pitch, g,
      static_cast<__nv_bfloat16*>(dfine), dfine_pitch, dg, rows, c);
  B200_CHECK_LAUNCH("reverse_gate_bwd");
  return 0;
}

int b200seg_channel_blend_fwd(const void* x1, int64_t x1_pitch, const float* w1, const void* x2, int64_t x2_pitch,
                              const float* w2, void* out, int64_t out_pitch, int64_t rows_per_sample, int n, int c,
                              void* stream) {
  B200_CHECK_ARG(x1 && w1 && out && rows_per_sample > 0 && n > 0 && c > 0 && c % 8 == 0 && (!x2 || w2) &&
                     x1_pitch % 8 == 0 && out_pitch % 8 == 0 && (!x2 || x2_pitch % 8 == 0),
                 "channel_blend_fwd: bad arguments (C must be a multiple of 8)");
  channel_blend_fwd_kernel<<<grid_for(rows_per_sample * n * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x1), x1_pitch, w1, static_cast<const __nv_bfloat16*>(x2), x2_pitch, w2,
      static_cast<__nv_bfloat16*>(out), out_pitch, rows_per_sample, n, c);
  B200_CHECK_LAUNCH("channel_blend_fwd");
  return 0;
}

int b200seg_channel_blend_bwd_reduce(const void* dout, int64_t dout_pitch, const void* x1, int64_t x1_pitch, const void* x2,
                                     int64_t x2_pitch, float* dots1, float* dots2, int64_t rows_per_sample, int n, int c,
                                     void* stream) {
  B200_CHECK_ARG(dout && x1 && dots1 && (!x2 || dots2) && rows_per_sample > 0 && n > 0 && c % 8 == 0 && c > 0 &&
                     256 % (c / 8) == 0 && dout_pitch % 8 == 0 && x1_pitch % 8 == 0 && (!x2 || x2_pitch % 8 == 0),
                 "channel_blend_bwd_reduce: bad arguments (C / 8 must divide 256)");
  const int rpb = 256 / (c / 8);
  const int64_t want = (rows_per_sample + rpb - 1) / rpb;
  const int gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, (kNumSMs * 8 + n - 1) / n)));
  channel_blend_bwd_reduce_kernel<<<dim3(gx, n), 256, 2 * c * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), dout_pitch, static_cast<const __nv_bfloat16*>(x1), x1_pitch,
      static_cast<const __nv_bfloat16*>(x2), x2_pitch, dots1, dots2, rows_per_sample, c);
  B200_CHECK_LAUNCH("channel_blend_bwd_reduce");
  return 0;
}

int b200seg_channel_blend_bwd_apply(const void* dout, int64_t dout_pitch, const float* w1, const float* w2, const float* add,
                                    void* dx1, int64_t dx1_pitch, void* dx2, int64_t dx2_pitch, int64_t rows_per_sample,
                                    int n, int c, void* stream) {
  B200_CHECK_ARG(dout && w1 && dx1 && (!dx2 || w2) && rows_per_sample > 0 && n > 0 && c % 8 == 0 && c > 0 &&
                     dout_pitch % 8 == 0 && dx1_pitch % 8 == 0 && (!dx2 || dx2_pitch % 8 == 0),
                 "channel_blend_bwd_apply: bad arguments");
  channel_blend_bwd_apply_kernel<<<grid_for(rows_per_sample * n * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), dout_pitch, w1, w2, add, static_cast<__nv_bfloat16*>(dx1), dx1_pitch,
      static_cast<__nv_bfloat16*>(dx2), dx2_pitch, rows_per_sample, n, c);
  B200_CHECK_LAUNCH("channel_blend_bwd_apply");
  return 0;
}

int b200seg_f32_sigmoid_fwd(const float* x, float* y, int64_t numel, void* stream) {
  B200_CHECK_ARG(x && y && numel > 0, "f32_sigmoid_fwd: bad arguments");
  f32_sigmoid_fwd_kernel<<<grid_for(numel, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, numel);
  B200_CHECK_LAUNCH("f32_sigmoid_fwd");
  return 0;
}

int b200seg_f32_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t numel, void* stream) {
  B200_CHECK_ARG(dy && y && dx && numel > 0, "f32_sigmoid_bwd: bad arguments");
  f32_sigmoid_bwd_kernel<<<grid_for(numel, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, dx, numel);
  B200_CHECK_LAUNCH("f32_sigmoid_bwd");
  return 0;
}

}  // extern "C"
