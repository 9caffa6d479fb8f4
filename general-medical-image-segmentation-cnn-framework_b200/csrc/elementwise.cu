// Bandwidth-bound kernels: layout conversion, normalisation (+activation) forward/backward, MaxPool3d(2,2),
// nearest x2 upsample, add, fused Adam.  All operate on NDHWC bf16 rows with 128-bit accesses where C % 8 == 0.
// Roofline for every kernel here is HBM: algorithmic bytes are stated per kernel in DESIGN.md.
#include <stdarg.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ------------------------------------------------------------------------------------------------ layout
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int c,
                                      int64_t spatial, int64_t pitch) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t s0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i;
    const int64_t s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (cc < c && s < spatial) ? src[(static_cast<int64_t>(n) * c + cc) * spatial + s] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t s = s0 + i;
    const int cc = c0 + threadIdx.x;
    if (cc < c && s < spatial) dst[(static_cast<int64_t>(n) * spatial + s) * pitch + cc] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

// C == 1: the two layouts coincide, only the element type changes (the model input, train.py:195)
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t numel) {
  const int64_t nvec = numel / 8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    st8(dst + 8 * i, pack8(t));
  }
  if (blockIdx.x == 0 && threadIdx.x < numel - nvec * 8) {
    const int64_t i = nvec * 8 + threadIdx.x;
    dst[i] = __float2bfloat16(src[i]);
  }
}

__global__ void ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int c,
                                      int64_t spatial, int64_t pitch) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t s0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t s = s0 + i;
    const int cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (cc < c && s < spatial) ? __bfloat162float(src[(static_cast<int64_t>(n) * spatial + s) * pitch + cc]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i;
    const int64_t s = s0 + threadIdx.x;
    if (cc < c && s < spatial) dst[(static_cast<int64_t>(n) * c + cc) * spatial + s] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------ norm
// Row-major [groups][rows][C] with pitch; V = channels per thread access (8 = 128-bit, 1 = scalar tail shapes).
template <int V>
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float (&f)[V]) {
  if constexpr (V == 8) {
    float t[8];
    unpack8(ld8(p), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = t[i];
  } else {
    f[0] = __bfloat162float(p[0]);
  }
}
template <int V>
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&f)[V]) {
  if constexpr (V == 8) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = f[i];
    st8(p, pack8(t));
  } else {
    p[0] = __float2bfloat16(f[0]);
  }
}

// Raw (unconverted) vector: lets a thread keep several loads in flight at 4 registers each.
template <int V> struct RawVec { bf16x8 v; };
template <> struct RawVec<1> { __nv_bfloat16 v; };
template <int V>
__device__ __forceinline__ RawVec<V> ld_raw(const __nv_bfloat16* p) {
  RawVec<V> r;
  if constexpr (V == 8) r.v = ld8(p); else r.v = p[0];
  return r;
}
template <int V>
__device__ __forceinline__ void cvt_raw(const RawVec<V>& r, float (&f)[V]) {
  if constexpr (V == 8) {
    float t[8];
    unpack8(r.v, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = t[i];
  } else {
    f[0] = __bfloat162float(r.v);
  }
}

// dst[0..7] += v[0..7] with two 4-wide vector reductions (REDG.E.ADD.F32x4) when dst is 16-byte aligned.  Every block of a
// column reduction ends with 2-4 such updates per channel chunk onto the SAME few cache lines; the L2 retires ~3.5 atomic
// operations per clock in total, so 296 blocks x 4 x C scalar atomics cost 25 us on a 128-channel layer (more than the
// 8 us its 34 MB take to stream).  Vector reductions cut the operation count by four.
__device__ __forceinline__ void red_add8(float* dst, const float* v) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dst + i, v[i]);
  }
}

// Block = (TX chunk lanes) x (TY row lanes), TX*TY = 256.  NS sums of V channels each are reduced over rows.
template <int V, int NS, int NT, typename FL, typename FA>
__device__ __forceinline__ void column_reduce(int64_t rows, int cv, float* __restrict__ out_group, int C, FL&& load,
                                              FA&& accum, float* __restrict__ affine = nullptr) {
  // load(r, ch, raw[NT]) fetches the NT tensors of row r unconverted; accum(ch, raw[NT], acc) folds one row in.
  extern __shared__ float red[];  // [TY][TX][NS*V]
  const int TX = blockDim.x, TY = blockDim.y;
  for (int ch = threadIdx.x; ch < cv; ch += TX) {
    float acc[NS][V];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[s][i] = 0.f;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * TY;
    int64_t r = static_cast<int64_t>(blockIdx.x) * TY + threadIdx.y;
    constexpr int U = NT == 1 ? 8 : 2;   // rows in flight; 4 registers per pending vector
    for (; r + (U - 1) * stride < rows; r += U * stride) {
      RawVec<V> raw[U][NT];
#pragma unroll
      for (int u = 0; u < U; ++u) load(r + u * stride, ch, raw[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) accum(ch, raw[u], acc);
    }
    for (; r < rows; r += stride) {
      RawVec<V> raw[NT];
      load(r, ch, raw);
      accum(ch, raw, acc);
    }
    float* mine = red + (static_cast<size_t>(threadIdx.y) * TX + threadIdx.x) * (NS * V);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int i = 0; i < V; ++i) mine[s * V + i] = acc[s][i];
    __syncthreads();
    for (int st = TY >> 1; st > 0; st >>= 1) {
      if (threadIdx.y < st) {
        const float* other = red + (static_cast<size_t>(threadIdx.y + st) * TX + threadIdx.x) * (NS * V);
#pragma unroll
        for (int j = 0; j < NS * V; ++j) mine[j] += other[j];
      }
      __syncthreads();
    }
    if (threadIdx.y == 0) {
      if constexpr (V == 8) {
#pragma unroll
        for (int s = 0; s < NS; ++s) red_add8(out_group + static_cast<size_t>(s) * C + ch * V, mine + s * V);
        if (affine != nullptr) {   // {dgamma[C], dbeta[C]} = {row 1, row 0}, straight into the parameter gradients
          red_add8(affine + ch * V, mine + 1 * V);
          red_add8(affine + C + ch * V, mine + 0 * V);
        }
      } else {
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int i = 0; i < V; ++i) atomicAdd(out_group + static_cast<size_t>(s) * C + ch * V + i, mine[s * V + i]);
        if (affine != nullptr) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            atomicAdd(affine + ch * V + i, mine[1 * V + i]);
            atomicAdd(affine + C + ch * V + i, mine[0 * V + i]);
          }
        }
      }
    }
    __syncthreads();
  }
}

template <int V>
__global__ void __launch_bounds__(256, 4) channel_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t pitch, int64_t rows, int C,
                                     float* __restrict__ stats) {
  const int g = blockIdx.y;
  const __nv_bfloat16* xg = x + static_cast<int64_t>(g) * rows * pitch;
  column_reduce<V, 2, 1>(
      rows, C / V, stats + static_cast<size_t>(g) * 2 * C, C,
      [&](int64_t r, int ch, RawVec<V>(&raw)[1]) { raw[0] = ld_raw<V>(xg + r * pitch + ch * V); },
      [&](int, const RawVec<V>(&raw)[1], float(&acc)[2][V]) {
        float f[V];
        cvt_raw<V>(raw[0], f);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          acc[0][i] += f[i];
          acc[1][i] += f[i] * f[i];
        }
      });
}

// {sum, sum of squares} of one channel -> {mean, inv_std, scale, shift} (and the running statistics when `update`).
// One definition for the stand-alone finalisation kernel and for the prologue of the fused normalise kernel: bit-identical.
struct NormCoef {
  float mean, inv_std, scale, shift;
};
__device__ __forceinline__ NormCoef norm_finalize_channel(double s, double q, double count, int c,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var,
                                                          float momentum, float eps, int clamp_eps, bool update) {
  const double mean = s / count;
  double var = (q - s * mean) / count;  // batchnorm.py:116-120: sumvar = ssum - sum*mean
  if (var < 0) var = 0;
  const double inv_std = clamp_eps ? 1.0 / sqrt(var < eps ? static_cast<double>(eps) : var) : 1.0 / sqrt(var + eps);
  if (update && running_mean != nullptr) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
  const double ga = gamma ? gamma[c] : 1.0;
  const double be = beta ? beta[c] : 0.0;
  NormCoef r;
  r.mean = static_cast<float>(mean);
  r.inv_std = static_cast<float>(inv_std);
  r.scale = static_cast<float>(ga * inv_std);
  r.shift = static_cast<float>(be - mean * ga * inv_std);
  return r;
}

__global__ void norm_finalize_kernel(const float* __restrict__ stats, double count, int groups, int C,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                     float eps, int clamp_eps, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * C) return;
  const int g = i / C, c = i % C;
  const double s = stats[(static_cast<size_t>(g) * 2 + 0) * C + c];
  const double q = stats[(static_cast<size_t>(g) * 2 + 1) * C + c];
  const NormCoef r = norm_finalize_channel(s, q, count, c, gamma, beta, running_mean, running_var, momentum, eps, clamp_eps,
                                           groups == 1);
  float* o = out + static_cast<size_t>(g) * 4 * C;
  o[0 * C + c] = r.mean;
  o[1 * C + c] = r.inv_std;
  o[2 * C + c] = r.scale;
  o[3 * C + c] = r.shift;
}

// Training-mode BatchNorm without a separate finalisation launch: every block of the normalise kernel derives the
// constants of all C <= kFusedNormMaxC channels from the {sum, sumsq} vector in its prologue (shared memory); block 0 also
// stores them for the backward pass and updates the running statistics.
constexpr int kFusedNormMaxC = 512;
struct NormFinalizeArgs {
  const float* stats;     // {sum[C], sumsq[C]}; nullptr: `coef` holds finished constants (the unfused form)
  double count;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float momentum, eps;
  int clamp_eps;
  float* coef_out;        // [4][C]
};

// Inference-mode BatchNorm constants from the running statistics (F.batch_norm(training=False)): rows {mean, inv_std,
// scale = gamma * inv_std, shift = beta - mean * scale [+ scale * conv_bias]} -- the optional conv bias is folded into the
// shift so that the convolution's epilogue can apply scale / shift / activation to the raw accumulator.
__global__ void norm_eval_coef_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      const float* __restrict__ conv_bias, float eps, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = running_mean[c];
  const float inv_std = 1.f / sqrtf(running_var[c] + eps);
  const float scale = (gamma ? gamma[c] : 1.f) * inv_std;
  float shift = (beta ? beta[c] : 0.f) - mean * scale;
  if (conv_bias) shift = fmaf(scale, conv_bias[c], shift);
  out[0 * C + c] = mean;
  out[1 * C + c] = inv_std;
  out[2 * C + c] = scale;
  out[3 * C + c] = shift;
}

template <int V, int ACT>
__global__ void __launch_bounds__(256, 3) norm_act_fwd_kernel(const __nv_bfloat16* __restrict__ y, int64_t y_pitch,
                                    const float* __restrict__ coef, int64_t rows_per_group, int groups, int C, int act_rt,
                                    float act_param, const float* __restrict__ prelu_w,
                                    const __nv_bfloat16* __restrict__ res, int64_t res_pitch,
                                    __nv_bfloat16* __restrict__ z, int64_t z_pitch, const NormFinalizeArgs fin) {
  constexpr int act = ACT;
  (void)act_rt;
  __shared__ float s_fin[2][kFusedNormMaxC];     // scale, shift (fused finalisation only)
  if (fin.stats != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const NormCoef r = norm_finalize_channel(fin.stats[c], fin.stats[C + c], fin.count, c, fin.gamma, fin.beta,
                                               fin.running_mean, fin.running_var, fin.momentum, fin.eps, fin.clamp_eps,
                                               blockIdx.x == 0);
      s_fin[0][c] = r.scale;
      s_fin[1][c] = r.shift;
      if (blockIdx.x == 0) {
        fin.coef_out[0 * C + c] = r.mean;
        fin.coef_out[1 * C + c] = r.inv_std;
        fin.coef_out[2 * C + c] = r.scale;
        fin.coef_out[3 * C + c] = r.shift;
      }
    }
    __syncthreads();
  }
  // Host guarantees (gridDim.x * blockDim.x) % (C / V) == 0: a thread keeps one channel chunk for its whole life,
  // so scale / shift / slope live in registers; rows are walked 4 at a time to keep 4 loads in flight.
  const int cv = C / V;
  const int64_t gtid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int ch = static_cast<int>(gtid % cv);
  const int64_t row0 = gtid / cv, rstep = nthr / cv;
  const int64_t rows = rows_per_group * groups;
  float sc[V], sh[V], sl[V];
  int cur_g = -1;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    sc[j] = 1.f;
    sh[j] = 0.f;
    sl[j] = (act == B200SEG_ACT_PRELU) ? prelu_w[ch * V + j] : act_param;
  }
  auto load_coef = [&](int g) {
    if (fin.stats != nullptr) {      // groups == 1: the prologue's constants
      if (cur_g == 0) return;
      cur_g = 0;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        sc[j] = s_fin[0][ch * V + j];
        sh[j] = s_fin[1][ch * V + j];
      }
      return;
    }
    if (coef == nullptr || g == cur_g) return;
    cur_g = g;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      sc[j] = coef[(static_cast<size_t>(g) * 4 + 2) * C + ch * V + j];
      sh[j] = coef[(static_cast<size_t>(g) * 4 + 3) * C + ch * V + j];
    }
  };
  auto one = [&](int64_t row, const float (&f)[V], const float (&r)[V]) {
    float o[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float pre = f[j] * sc[j] + sh[j];
      if (res) pre += r[j];
      o[j] = act_fwd(pre, act, sl[j]);
    }
    store_vec<V>(z + row * z_pitch + ch * V, o);
  };
  load_coef(0);
  int64_t row = row0;
  // Without a residual: 6 rows (96 bytes per thread) in flight -- B200 needs ~90 KB outstanding per SM to saturate HBM.
  // With one: 3 rows of both tensors (the two loops keep the register count of either path below the 85 of 3 CTAs/SM).
  if (res == nullptr) {
    for (; row + 5 * rstep < rows; row += 6 * rstep) {
      RawVec<V> ry[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) ry[u] = ld_raw<V>(y + (row + u * rstep) * y_pitch + ch * V);
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        float f[V], r[V];
        cvt_raw<V>(ry[u], f);
        if (groups > 1) load_coef(static_cast<int>((row + u * rstep) / rows_per_group));
        one(row + u * rstep, f, r);
      }
    }
  } else {
    for (; row + 2 * rstep < rows; row += 3 * rstep) {
      RawVec<V> ry[3], rr[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        ry[u] = ld_raw<V>(y + (row + u * rstep) * y_pitch + ch * V);
        rr[u] = ld_raw<V>(res + (row + u * rstep) * res_pitch + ch * V);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        float f[V], r[V];
        cvt_raw<V>(ry[u], f);
        cvt_raw<V>(rr[u], r);
        if (groups > 1) load_coef(static_cast<int>((row + u * rstep) / rows_per_group));
        one(row + u * rstep, f, r);
      }
    }
  }
  for (; row < rows; row += rstep) {
    float f[V], r[V];
    load_vec<V>(y + row * y_pitch + ch * V, f);
    if (res) load_vec<V>(res + row * res_pitch + ch * V, r);
    if (groups > 1) load_coef(static_cast<int>(row / rows_per_group));
    one(row, f, r);
  }
}

template <int V, int ACT>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dz, int64_t dz_pitch,
                                           const __nv_bfloat16* __restrict__ y, int64_t y_pitch,
                                           const float* __restrict__ coef, int64_t rows, int C, int act_rt,
                                           float act_param, const float* __restrict__ prelu_w,
                                           const __nv_bfloat16* __restrict__ res, int64_t res_pitch,
                                           float* __restrict__ sums, float* __restrict__ dprelu,
                                           float* __restrict__ affine) {
  constexpr int act = ACT;
  (void)act_rt;
  const int g = blockIdx.y;
  const int64_t base = static_cast<int64_t>(g) * rows;
  const float* cg = coef ? coef + static_cast<size_t>(g) * 4 * C : nullptr;
  // each thread normally owns exactly one channel chunk (cv <= blockDim.x): keep its coefficients in registers
  const int ch0 = threadIdx.x;
  const bool hoisted = (C / V) <= static_cast<int>(blockDim.x);
  float h_mean[V], h_istd[V], h_sc[V], h_sh[V], h_sl[V];   // h_mean holds mean*inv_std: xhat = y*inv_std - h_mean
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int c = ch0 * V + j;
    const bool ok = hoisted && c < C;
    h_istd[j] = (ok && cg) ? cg[C + c] : 1.f;
    h_mean[j] = (ok && cg) ? cg[c] * h_istd[j] : 0.f;
    h_sc[j] = (ok && cg) ? cg[2 * C + c] : 1.f;
    h_sh[j] = (ok && cg) ? cg[3 * C + c] : 0.f;
    h_sl[j] = (ok && act == B200SEG_ACT_PRELU) ? prelu_w[c] : act_param;
  }
  auto load = [&](int64_t r, int ch, RawVec<V>(&raw)[3]) {
    raw[0] = ld_raw<V>(y + (base + r) * y_pitch + ch * V);
    raw[1] = ld_raw<V>(dz + (base + r) * dz_pitch + ch * V);
    if (res) raw[2] = ld_raw<V>(res + (base + r) * res_pitch + ch * V);
  };
  auto accum = [&](int ch, const RawVec<V>(&raw)[3], auto& acc) {
    float fy[V], fd[V], fr[V];
    cvt_raw<V>(raw[0], fy);
    cvt_raw<V>(raw[1], fd);
    if (res) cvt_raw<V>(raw[2], fr);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float mean = h_mean[j], istd = h_istd[j], sc = h_sc[j], sh = h_sh[j], slope = h_sl[j];
      if (!hoisted) {
        const int c = ch * V + j;
        if (cg) {
          istd = cg[C + c];
          mean = cg[c] * istd;
          sc = cg[2 * C + c];
          sh = cg[3 * C + c];
        }
        slope = (act == B200SEG_ACT_PRELU) ? prelu_w[c] : act_param;
      }
      float pre = fy[j] * sc + sh;
      const float xhat = fy[j] * istd - mean;
      if (res) pre += fr[j];
      const float dpre = fd[j] * act_bwd(pre, act, slope);
      acc[0][j] += dpre;
      acc[1][j] += dpre * xhat;
      if constexpr (sizeof(acc) / sizeof(acc[0]) == 3) acc[2][j] += (pre > 0.f) ? 0.f : fd[j] * pre;
    }
  };
  if (dprelu) {
    // third running sum = PReLU slope gradient; lands in a scratch row after the two sums of this group
    column_reduce<V, 3, 3>(rows, C / V, sums + static_cast<size_t>(g) * 3 * C, C, load,
                           [&](int ch, const RawVec<V>(&raw)[3], float(&acc)[3][V]) { accum(ch, raw, acc); }, affine);
  } else {
    column_reduce<V, 2, 3>(rows, C / V, sums + static_cast<size_t>(g) * 2 * C, C, load,
                           [&](int ch, const RawVec<V>(&raw)[3], float(&acc)[2][V]) { accum(ch, raw, acc); }, affine);
  }
}

// Fast path of the backward reduction (128-bit rows, no residual, no PReLU slope gradient, one channel chunk per thread):
// xhat only enters the second sum linearly, sum(dpre * xhat) = inv_std * sum(dpre * (y - mean)), so the loop keeps just
// {scale, shift, mean} per channel and folds inv_std in once at the end.  That leaves registers for 4 rows x 2 tensors of
// 128-bit loads in flight per thread at 3 CTAs/SM (~98 KB outstanding per SM, what B200 needs to stream at HBM speed).
template <int ACT>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_reduce_fast_kernel(
    const __nv_bfloat16* __restrict__ dz, int64_t dz_pitch, const __nv_bfloat16* __restrict__ y, int64_t y_pitch,
    const float* __restrict__ coef, int64_t rows, int C, float act_param, const float* __restrict__ prelu_w,
    float* __restrict__ sums, float* __restrict__ affine) {
  constexpr int V = 8;
  extern __shared__ float red[];  // [TY][TX][2*V]
  const int TX = blockDim.x, TY = blockDim.y;
  const int g = blockIdx.y;
  const int64_t base = static_cast<int64_t>(g) * rows;
  const float* cg = coef ? coef + static_cast<size_t>(g) * 4 * C : nullptr;
  const int ch = threadIdx.x;      // host guarantees C / 8 == blockDim.x
  float sc[V], sh[V], mu[V], sl[V], s1[V], s2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int c = ch * V + j;
    mu[j] = cg ? cg[c] : 0.f;
    sc[j] = cg ? cg[2 * C + c] : 1.f;
    sh[j] = cg ? cg[3 * C + c] : 0.f;
    sl[j] = (ACT == B200SEG_ACT_PRELU) ? prelu_w[c] : act_param;
    s1[j] = s2[j] = 0.f;
  }
  auto fold = [&](const bf16x8& ry, const bf16x8& rd) {
    float fy[V], fd[V];
    unpack8(ry, fy);
    unpack8(rd, fd);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float pre = fy[j] * sc[j] + sh[j];
      const float dpre = fd[j] * act_bwd(pre, ACT, sl[j]);
      s1[j] += dpre;
      s2[j] = fmaf(dpre, fy[j] - mu[j], s2[j]);
    }
  };
  const int64_t stride = static_cast<int64_t>(gridDim.x) * TY;
  int64_t r = static_cast<int64_t>(blockIdx.x) * TY + threadIdx.y;
#ifndef B200_REDUCE_ROWS
#define B200_REDUCE_ROWS 4
#endif
  // rows in flight per thread (2 * UR 128-bit loads).  6 and 8 were measured (probes/norm_bw.py, -DB200_REDUCE_ROWS): no
  // gain -- at 5.1 TB/s of reads the kernel is bound by its ~10 instructions per element, not by bytes in flight
  constexpr int UR = B200_REDUCE_ROWS;
  for (; r + (UR - 1) * stride < rows; r += UR * stride) {
    bf16x8 ry[UR], rd[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      ry[u] = ld8(y + (base + r + u * stride) * y_pitch + ch * V);
      rd[u] = ld8(dz + (base + r + u * stride) * dz_pitch + ch * V);
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) fold(ry[u], rd[u]);
  }
  for (; r < rows; r += stride) fold(ld8(y + (base + r) * y_pitch + ch * V), ld8(dz + (base + r) * dz_pitch + ch * V));
  float* mine = red + (static_cast<size_t>(threadIdx.y) * TX + threadIdx.x) * (2 * V);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    mine[j] = s1[j];
    mine[V + j] = s2[j];
  }
  __syncthreads();
  for (int st = TY >> 1; st > 0; st >>= 1) {
    if (threadIdx.y < st) {
      const float* other = red + (static_cast<size_t>(threadIdx.y + st) * TX + threadIdx.x) * (2 * V);
#pragma unroll
      for (int j = 0; j < 2 * V; ++j) mine[j] += other[j];
    }
    __syncthreads();
  }
  if (threadIdx.y == 0) {
    float* out = sums + static_cast<size_t>(g) * 2 * C;
    float r0[V], r1[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = ch * V + j;
      const float istd = cg ? cg[C + c] : 1.f;
      r0[j] = mine[j];
      r1[j] = mine[V + j] * istd;
    }
    red_add8(out + ch * V, r0);
    red_add8(out + C + ch * V, r1);
    if (affine != nullptr) {   // {dgamma[C], dbeta[C]} straight into the parameter gradients
      red_add8(affine + ch * V, r1);
      red_add8(affine + C + ch * V, r0);
    }
  }
}

template <int V, int ACT>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dz, int64_t dz_pitch,
                                          const __nv_bfloat16* __restrict__ y, int64_t y_pitch,
                                          const float* __restrict__ coef, const float* __restrict__ sums,
                                          int sums_stride, float inv_count, int64_t rows_per_group, int groups, int C,
                                          int act_rt, float act_param, const float* __restrict__ prelu_w,
                                          const __nv_bfloat16* __restrict__ res, int64_t res_pitch,
                                          __nv_bfloat16* __restrict__ dy, int64_t dy_pitch,
                                          __nv_bfloat16* __restrict__ dres, int64_t dres_pitch,
                                          const __nv_bfloat16* acc, int64_t acc_pitch) {
  constexpr int act = ACT;
  (void)act_rt;
  // acc != nullptr: dy = acc + (the gradient below); acc may alias dy (every thread reads its own elements before it
  // writes them) -- a tensor that also feeds a concatenation accumulates both gradients in one buffer (dense blocks).
  // dy = scale * (dpre - m1 - xhat * m2), xhat = (y - mean) * inv_std, m1 = sum(dpre)/n, m2 = sum(dpre*xhat)/n.
  // Same thread -> channel-chunk binding as the forward kernel (host guarantees divisibility).
  const int cv = C / V;
  const int64_t gtid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int ch = static_cast<int>(gtid % cv);
  const int64_t row0 = gtid / cv, rstep = nthr / cv;
  const int64_t rows = rows_per_group * groups;
  // dy = sc*dpre - (A + y*B) with B = inv_std*m2*sc, A = m1*sc - mean*B  (4 fused coefficients per channel)
  float sc[V], sh[V], ca[V], cb[V], sl[V];
  int cur_g = -1;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    sc[j] = 1.f;
    sh[j] = 0.f;
    ca[j] = cb[j] = 0.f;
    sl[j] = (act == B200SEG_ACT_PRELU) ? prelu_w[ch * V + j] : act_param;
  }
  auto load_coef = [&](int g) {
    if (g == cur_g) return;
    cur_g = g;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = ch * V + j;
      float mean = 0.f, istd = 1.f, m1 = 0.f, m2 = 0.f;
      if (coef) {
        const float* cg = coef + static_cast<size_t>(g) * 4 * C;
        mean = cg[c];
        istd = cg[C + c];
        sc[j] = cg[2 * C + c];
        sh[j] = cg[3 * C + c];
      }
      if (sums) {
        const float* sg = sums + static_cast<size_t>(g) * sums_stride * C;
        m1 = sg[c] * inv_count;
        m2 = sg[C + c] * inv_count;
      }
      // without normalisation (coef == nullptr) xhat is never subtracted: sums are null in that case
      cb[j] = coef ? istd * m2 * sc[j] : 0.f;
      ca[j] = m1 * sc[j] - mean * cb[j];
    }
  };
  auto one = [&](int64_t row, const float (&fy)[V], const float (&fd)[V], const float (&fr)[V]) {
    float o[V], dr[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float pre = fy[j] * sc[j] + sh[j];
      if (res) pre += fr[j];
      const float dpre = fd[j] * act_bwd(pre, act, sl[j]);
      dr[j] = dpre;
      o[j] = dpre * sc[j] - (ca[j] + fy[j] * cb[j]);
    }
    if (acc != nullptr) {
      float fa[V];
      load_vec<V>(acc + row * acc_pitch + ch * V, fa);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] += fa[j];
    }
    store_vec<V>(dy + row * dy_pitch + ch * V, o);
    if (dres) store_vec<V>(dres + row * dres_pitch + ch * V, dr);
  };
  load_coef(0);
  int64_t row = row0;
  if (res == nullptr) {
    for (; row + 3 * rstep < rows; row += 4 * rstep) {
      RawVec<V> ry[4], rd[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ry[u] = ld_raw<V>(y + (row + u * rstep) * y_pitch + ch * V);
        rd[u] = ld_raw<V>(dz + (row + u * rstep) * dz_pitch + ch * V);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float fy[V], fd[V], fr[V];
        cvt_raw<V>(ry[u], fy);
        cvt_raw<V>(rd[u], fd);
        if (groups > 1) load_coef(static_cast<int>((row + u * rstep) / rows_per_group));
        one(row + u * rstep, fy, fd, fr);
      }
    }
  } else {
    for (; row + rstep < rows; row += 2 * rstep) {
      RawVec<V> ry[2], rd[2], rr[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        ry[u] = ld_raw<V>(y + (row + u * rstep) * y_pitch + ch * V);
        rd[u] = ld_raw<V>(dz + (row + u * rstep) * dz_pitch + ch * V);
        rr[u] = ld_raw<V>(res + (row + u * rstep) * res_pitch + ch * V);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float fy[V], fd[V], fr[V];
        cvt_raw<V>(ry[u], fy);
        cvt_raw<V>(rd[u], fd);
        cvt_raw<V>(rr[u], fr);
        if (groups > 1) load_coef(static_cast<int>((row + u * rstep) / rows_per_group));
        one(row + u * rstep, fy, fd, fr);
      }
    }
  }
  for (; row < rows; row += rstep) {
    float fy[V], fd[V], fr[V];
    load_vec<V>(y + row * y_pitch + ch * V, fy);
    load_vec<V>(dz + row * dz_pitch + ch * V, fd);
    if (res) load_vec<V>(res + row * res_pitch + ch * V, fr);
    if (groups > 1) load_coef(static_cast<int>(row / rows_per_group));
    one(row, fy, fd, fr);
  }
}

// ------------------------------------------------------------------------------------------------ pooling
template <int V>
__global__ void maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                                    __nv_bfloat16* __restrict__ y, int64_t y_pitch, uint8_t* __restrict__ idx, int n,
                                    int d, int h, int w, int C) {
  const int od = d / 2, oh = h / 2, ow = w / 2, cv = C / V;
  const int64_t total = static_cast<int64_t>(n) * od * oh * ow * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    int64_t v = i / cv;
    const int xo = static_cast<int>(v % ow);
    v /= ow;
    const int yo = static_cast<int>(v % oh);
    v /= oh;
    const int zo = static_cast<int>(v % od);
    const int nn = static_cast<int>(v / od);
    float best[V];
    uint8_t bi[V];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int a = k >> 2, b = (k >> 1) & 1, e = k & 1;
      const int64_t row = ((static_cast<int64_t>(nn) * d + 2 * zo + a) * h + 2 * yo + b) * w + 2 * xo + e;
      float f[V];
      load_vec<V>(x + row * x_pitch + ch * V, f);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        // torch: strictly-greater replaces; a NaN candidate replaces a non-NaN best and then sticks
        const bool take = (k == 0) || (f[j] > best[j]) || (f[j] != f[j] && best[j] == best[j]);
        if (take) {
          best[j] = f[j];
          bi[j] = static_cast<uint8_t>(k);
        }
      }
    }
    const int64_t orow = i / cv;
    store_vec<V>(y + orow * y_pitch + ch * V, best);
#pragma unroll
    for (int j = 0; j < V; ++j) idx[orow * C + ch * V + j] = bi[j];
  }
}

template <int V>
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_pitch,
                                    const uint8_t* __restrict__ idx, __nv_bfloat16* __restrict__ dx, int64_t dx_pitch,
                                    const __nv_bfloat16* __restrict__ addend, int64_t add_pitch,
                                    int n, int d, int h, int w, int C) {
  const int od = d / 2, oh = h / 2, ow = w / 2, cv = C / V;
  const int64_t total = static_cast<int64_t>(n) * od * oh * ow * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    const int64_t orow = i / cv;
    int64_t v = orow;
    const int xo = static_cast<int>(v % ow);
    v /= ow;
    const int yo = static_cast<int>(v % oh);
    v /= oh;
    const int zo = static_cast<int>(v % od);
    const int nn = static_cast<int>(v / od);
    float g[V];
    load_vec<V>(dy + orow * dy_pitch + ch * V, g);
    uint8_t bi[V];
#pragma unroll
    for (int j = 0; j < V; ++j) bi[j] = idx[orow * C + ch * V + j];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int a = k >> 2, b = (k >> 1) & 1, e = k & 1;
      const int64_t row = ((static_cast<int64_t>(nn) * d + 2 * zo + a) * h + 2 * yo + b) * w + 2 * xo + e;
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = (bi[j] == k) ? g[j] : 0.f;
      if (addend != nullptr) {   // gradient arriving through the skip connection of the same tensor
        float s[V];
        load_vec<V>(addend + row * add_pitch + ch * V, s);
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] += s[j];
      }
      store_vec<V>(dx + row * dx_pitch + ch * V, o);
    }
  }
}

__global__ void maxpool2_idx_to_torch_kernel(const uint8_t* __restrict__ idx, int64_t* __restrict__ out, int n, int d,
                                             int h, int w, int C) {
  const int od = d / 2, oh = h / 2, ow = w / 2;
  const int64_t total = static_cast<int64_t>(n) * C * od * oh * ow;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t v = i;  // NCDHW order of the output
    const int xo = static_cast<int>(v % ow);
    v /= ow;
    const int yo = static_cast<int>(v % oh);
    v /= oh;
    const int zo = static_cast<int>(v % od);
    v /= od;
    const int c = static_cast<int>(v % C);
    const int nn = static_cast<int>(v / C);
    const int64_t orow = ((static_cast<int64_t>(nn) * od + zo) * oh + yo) * ow + xo;
    const int k = idx[orow * C + c];
    const int a = k >> 2, b = (k >> 1) & 1, e = k & 1;
    out[i] = (static_cast<int64_t>(2 * zo + a) * h + 2 * yo + b) * w + 2 * xo + e;
  }
}

template <int V>
__global__ void upsample2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                                     __nv_bfloat16* __restrict__ y, int64_t y_pitch, int n, int d, int h, int w, int C) {
  const int cv = C / V;
  const int64_t total = static_cast<int64_t>(n) * d * h * w * 8 * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    int64_t v = i / cv;  // output voxel index in [n][2d][2h][2w]
    const int64_t orow = v;
    const int xo = static_cast<int>(v % (2 * w));
    v /= 2 * w;
    const int yo = static_cast<int>(v % (2 * h));
    v /= 2 * h;
    const int zo = static_cast<int>(v % (2 * d));
    const int nn = static_cast<int>(v / (2 * d));
    const int64_t irow = ((static_cast<int64_t>(nn) * d + zo / 2) * h + yo / 2) * w + xo / 2;
    float f[V];
    load_vec<V>(x + irow * x_pitch + ch * V, f);
    store_vec<V>(y + orow * y_pitch + ch * V, f);
  }
}

template <int V>
__global__ void upsample2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_pitch,
                                     __nv_bfloat16* __restrict__ dx, int64_t dx_pitch, int n, int d, int h, int w,
                                     int C) {
  const int cv = C / V;
  const int64_t total = static_cast<int64_t>(n) * d * h * w * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    const int64_t irow = i / cv;
    int64_t v = irow;
    const int xi = static_cast<int>(v % w);
    v /= w;
    const int yi = static_cast<int>(v % h);
    v /= h;
    const int zi = static_cast<int>(v % d);
    const int nn = static_cast<int>(v / d);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int a = k >> 2, b = (k >> 1) & 1, e = k & 1;
      const int64_t orow = ((static_cast<int64_t>(nn) * 2 * d + 2 * zi + a) * 2 * h + 2 * yi + b) * 2 * w + 2 * xi + e;
      float f[V];
      load_vec<V>(dy + orow * dy_pitch + ch * V, f);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += f[j];
    }
    store_vec<V>(dx + irow * dx_pitch + ch * V, acc);
  }
}

template <int V>
__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, int64_t a_pitch, const __nv_bfloat16* __restrict__ b,
                           int64_t b_pitch, __nv_bfloat16* __restrict__ out, int64_t out_pitch, int64_t rows, int C) {
  const int cv = C / V;
  const int64_t total = rows * cv;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv);
    const int64_t row = i / cv;
    float fa[V], fb[V];
    load_vec<V>(a + row * a_pitch + ch * V, fa);
    load_vec<V>(b + row * b_pitch + ch * V, fb);
#pragma unroll
    for (int j = 0; j < V; ++j) fa[j] += fb[j];
    store_vec<V>(out + row * out_pitch + ch * V, fa);
  }
}


// ------------------------------------------------------------------------------------------------ dropout
// Counter-based mask: keep = hash(seed, index) >= p * 2^32.  Forward and backward call the same kernel with the same
// (seed, salt), so no mask is stored.  `seed` lives in device memory (a captured CUDA graph sees a fresh value on every
// replay once the host bumps it); `salt` distinguishes the call sites of one step.
__device__ __forceinline__ uint32_t mix_hash(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return static_cast<uint32_t>(x >> 16);
}

// channel_mode = 0: nn.Dropout (one draw per element); 1: nn.Dropout3d (one draw per (sample, channel)).
// salt2 != 0 applies a second, independent mask in the same pass (densevoxelnet3d.py:25-32 runs its dropout twice);
// channels [C, c_out) of y are written as zeros (a gradient widened for the 16-channel tensor-core tiles).
__global__ void dropout_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch, __nv_bfloat16* __restrict__ y,
                               int64_t y_pitch, int64_t rows, int64_t rows_per_sample, int C, int c_out, float p,
                               const unsigned long long* __restrict__ seed, unsigned long long salt,
                               unsigned long long salt2, int channel_mode) {
  const unsigned long long sd = (seed ? *seed : 0ULL) * 0x9E3779B97F4A7C15ULL;
  const unsigned long long s = sd + salt * 0xD1B54A32D192ED03ULL, s2 = sd + salt2 * 0xD1B54A32D192ED03ULL;
  const uint32_t thresh = p >= 1.f ? 0xFFFFFFFFu : static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  const float scale = p >= 1.f ? 0.f : 1.f / (1.f - p);
  const int64_t total = rows * c_out;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c_out);
    const int64_t row = i / c_out;
    if (c >= C) {
      y[row * y_pitch + c] = __float2bfloat16(0.f);
      continue;
    }
    const unsigned long long idx = channel_mode ? static_cast<unsigned long long>((row / rows_per_sample) * C + c)
                                                : static_cast<unsigned long long>(row * C + c);
    float v = __bfloat162float(x[row * x_pitch + c]);
    v = mix_hash(s + idx) >= thresh ? v * scale : 0.f;
    if (salt2) {   // the reference rounds to the storage type between the two modules; so does this
      v = __bfloat162float(__float2bfloat16(v));
      v = mix_hash(s2 + idx) >= thresh ? v * scale : 0.f;
    }
    y[row * y_pitch + c] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------ class maps
// fp32 NCDHW class-score maps (deep supervision, residual_unet3d.py:196-202): out = nearest_x2(coarse) [+ fine].
__global__ void classmap_up2_add_kernel(const float* __restrict__ coarse, const float* __restrict__ fine,
                                        float* __restrict__ out, int64_t planes, int d, int h, int w) {
  const int64_t total = planes * 8 * d * h * w;
  const int W2 = 2 * w, H2 = 2 * h, D2 = 2 * d;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W2);
    const int y = static_cast<int>((i / W2) % H2);
    const int z = static_cast<int>((i / (static_cast<int64_t>(W2) * H2)) % D2);
    const int64_t pl = i / (static_cast<int64_t>(W2) * H2 * D2);
    const float v = coarse[((pl * d + (z >> 1)) * h + (y >> 1)) * w + (x >> 1)];
    out[i] = fine ? v + fine[i] : v;
  }
}
// backward of the up-sampling: dcoarse = sum of the 8 fine gradients.
__global__ void classmap_down2_sum_kernel(const float* __restrict__ dfine, float* __restrict__ dcoarse, int64_t planes,
                                          int d, int h, int w) {
  const int64_t total = planes * d * h * w;
  const int W2 = 2 * w, H2 = 2 * h;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const int z = static_cast<int>((i / (static_cast<int64_t>(w) * h)) % d);
    const int64_t pl = i / (static_cast<int64_t>(w) * h * d);
    const float* src = dfine + ((pl * 2 * d + 2 * z) * H2 + 2 * y) * W2 + 2 * x;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) acc += src[(static_cast<int64_t>(a) * H2 + b) * W2] + src[(static_cast<int64_t>(a) * H2 + b) * W2 + 1];
    dcoarse[i] = acc;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t numel, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float grad = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) grad += wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * grad;
    const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// Graph-capturable variant: step count and hyper-parameters live in device memory, so a captured launch stays correct
// when it is replayed (state[0] = step, incremented here by the last block to finish; hyper = {lr, b1, b2, eps, wd, gscale}).
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t numel, const float* __restrict__ hyper,
                                int* __restrict__ state) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gscale = hyper[5];
  const float step = static_cast<float>(state[0] + 1);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float grad = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) grad += wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * grad;
    const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int done = atomicAdd(&state[1], 1);
    if (done == static_cast<int>(gridDim.x) - 1) {   // every block has read state[0]: safe to advance it
      state[1] = 0;
      state[0] = state[0] + 1;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

const char* b200seg_version(void) { return "b200seg 0.1 (sm_100a)"; }
const char* b200seg_last_error(void) { return b200::get_error(); }

int b200seg_ncdhw_f32_to_ndhwc_bf16_pitched(const float* src, void* dst, int64_t dst_pitch, int n, int c, int64_t spatial,
                                            void* stream) {
  B200_CHECK_ARG(src && dst && n > 0 && c > 0 && spatial > 0 && dst_pitch >= c, "ncdhw_to_ndhwc: bad arguments");
  if (c == 1 && dst_pitch == 1 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int64_t numel = static_cast<int64_t>(n) * spatial;
    f32_to_bf16_kernel<<<grid_for(numel / 8 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<__nv_bfloat16*>(dst), numel);
    B200_CHECK_LAUNCH("ncdhw_to_ndhwc");
    return 0;
  }
  dim3 grid(static_cast<unsigned>((spatial + 31) / 32), (c + 31) / 32, n), block(32, 8);
  ncdhw_to_ndhwc_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), c, spatial, dst_pitch);
  B200_CHECK_LAUNCH("ncdhw_to_ndhwc");
  return 0;
}
int b200seg_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int n, int c, int64_t spatial, void* stream) {
  return b200seg_ncdhw_f32_to_ndhwc_bf16_pitched(src, dst, c, n, c, spatial, stream);
}
int b200seg_ndhwc_bf16_to_ncdhw_f32_pitched(const void* src, int64_t src_pitch, float* dst, int n, int c, int64_t spatial,
                                            void* stream) {
  B200_CHECK_ARG(src && dst && n > 0 && c > 0 && spatial > 0 && src_pitch >= c, "ndhwc_to_ncdhw: bad arguments");
  dim3 grid(static_cast<unsigned>((spatial + 31) / 32), (c + 31) / 32, n), block(32, 8);
  ndhwc_to_ncdhw_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, c, spatial, src_pitch);
  B200_CHECK_LAUNCH("ndhwc_to_ncdhw");
  return 0;
}
int b200seg_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int n, int c, int64_t spatial, void* stream) {
  return b200seg_ndhwc_bf16_to_ncdhw_f32_pitched(src, c, dst, n, c, spatial, stream);
}

// ACT is a template parameter so every instantiation carries only its own activation code (the runtime switch made the
// kernels ~50 KB of SASS and instruction-cache bound).
#define B200_ACT_DISPATCH(ACTV, ...)                                              \
  switch (ACTV) {                                                                 \
    case B200SEG_ACT_NONE: { constexpr int A_ = B200SEG_ACT_NONE; __VA_ARGS__; } break;   \
    case B200SEG_ACT_RELU: { constexpr int A_ = B200SEG_ACT_RELU; __VA_ARGS__; } break;   \
    case B200SEG_ACT_LEAKY: { constexpr int A_ = B200SEG_ACT_LEAKY; __VA_ARGS__; } break; \
    case B200SEG_ACT_ELU: { constexpr int A_ = B200SEG_ACT_ELU; __VA_ARGS__; } break;     \
    default: { constexpr int A_ = B200SEG_ACT_PRELU; __VA_ARGS__; } break;                \
  }

static inline bool vec_ok(int c, int64_t p0, int64_t p1 = 0, int64_t p2 = 0, int64_t p3 = 0) {
  return c % 8 == 0 && p0 % 8 == 0 && p1 % 8 == 0 && p2 % 8 == 0 && p3 % 8 == 0;
}
// Grid for the norm/act elementwise kernels: gridDim.x * 256 must be a multiple of the channel-chunk count cv.
static inline int ew_grid(int64_t total_items, int cv) {
  int blocks = grid_for(total_items, 256, kNumSMs * 8);
  int64_t m = cv;           // smallest block multiple: lcm(cv, 256) / 256
  int64_t a = cv, b = 256;
  while (b) { const int64_t t = a % b; a = b; b = t; }
  m = cv / a;
  blocks = static_cast<int>((blocks + m - 1) / m * m);
  return blocks;
}
// Blocks of a column reduction: every block ends with a shared-memory tree and 2-3 atomics per channel, so a block must
// own a few rows per thread row (>= 8: two 4-row batches) to amortise that -- but not more: the 16^3 / 8^3 layers are
// latency-bound (8 blocks walking 32 dependent rows each took 27 us for 2 MB), so small tensors still spread over the SMs.
static inline int reduce_grid(int64_t rows, int ty, int cap) {
  int64_t b = (rows + static_cast<int64_t>(ty) * 8 - 1) / (static_cast<int64_t>(ty) * 8);
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return static_cast<int>(b);
}
static inline dim3 reduce_block(int cv) {
  int tx = 1;
  while (tx < cv && tx < 64) tx <<= 1;
  return dim3(tx, 256 / tx);
}

int b200seg_channel_stats(const void* x, int64_t pitch, int64_t rows_per_group, int groups, int c, float* stats,
                          void* stream) {
  B200_CHECK_ARG(x && stats && rows_per_group > 0 && groups > 0 && c > 0 && pitch >= c, "channel_stats: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  const auto* xp = static_cast<const __nv_bfloat16*>(x);
  if (vec_ok(c, pitch)) {
    dim3 block = reduce_block(c / 8);
    dim3 grid(reduce_grid(rows_per_group, block.y, kNumSMs * 4 / (groups > 4 ? 4 : 1)), groups);
    channel_stats_kernel<8><<<grid, block, 256 * 16 * sizeof(float), st>>>(xp, pitch, rows_per_group, c, stats);
  } else {
    dim3 block = reduce_block(c);
    dim3 grid(reduce_grid(rows_per_group, block.y, kNumSMs * 4 / (groups > 4 ? 4 : 1)), groups);
    channel_stats_kernel<1><<<grid, block, 256 * 2 * sizeof(float), st>>>(xp, pitch, rows_per_group, c, stats);
  }
  B200_CHECK_LAUNCH("channel_stats");
  return 0;
}

int b200seg_norm_finalize(const float* stats, double count, int groups, int c, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float momentum, float eps, int clamp_eps,
                          float* out, void* stream) {
  B200_CHECK_ARG(stats && out && count > 0 && groups > 0 && c > 0, "norm_finalize: bad arguments");
  const int total = groups * c;
  norm_finalize_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, count, groups, c, gamma, beta, running_mean, running_var, momentum, eps, clamp_eps, out);
  B200_CHECK_LAUNCH("norm_finalize");
  return 0;
}

int b200seg_norm_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                           const float* conv_bias, float eps, int c, float* out, void* stream) {
  B200_CHECK_ARG(running_mean && running_var && out && c > 0, "norm_eval_coef: bad arguments");
  norm_eval_coef_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(gamma, beta, running_mean, running_var,
                                                                                       conv_bias, eps, c, out);
  B200_CHECK_LAUNCH("norm_eval_coef");
  return 0;
}

static int norm_act_fwd_launch(const void* y, int64_t y_pitch, const float* coef, int64_t rows_per_group, int groups, int c,
                               int act, float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                               void* z, int64_t z_pitch, const b200::NormFinalizeArgs& fin, void* stream);

int b200seg_norm_act_fwd(const void* y, int64_t y_pitch, const float* coef, int64_t rows_per_group, int groups, int c,
                         int act, float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                         void* z, int64_t z_pitch, void* stream) {
  return norm_act_fwd_launch(y, y_pitch, coef, rows_per_group, groups, c, act, act_param, prelu_w, residual, res_pitch, z,
                             z_pitch, b200::NormFinalizeArgs{}, stream);
}

int b200seg_norm_act_fwd_stats(const void* y, int64_t y_pitch, const float* stats, double count, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                               int clamp_eps, float* coef_out, int64_t rows, int c, int act, float act_param,
                               const float* prelu_w, const void* residual, int64_t res_pitch, void* z, int64_t z_pitch,
                               void* stream) {
  B200_CHECK_ARG(stats && coef_out && count > 0 && c > 0 && c <= b200::kFusedNormMaxC,
                 "norm_act_fwd_stats: needs statistics, a coefficient buffer and C <= %d", b200::kFusedNormMaxC);
  b200::NormFinalizeArgs fin{stats, count, gamma, beta, running_mean, running_var, momentum, eps, clamp_eps, coef_out};
  return norm_act_fwd_launch(y, y_pitch, nullptr, rows, 1, c, act, act_param, prelu_w, residual, res_pitch, z, z_pitch, fin,
                             stream);
}

static int norm_act_fwd_launch(const void* y, int64_t y_pitch, const float* coef, int64_t rows_per_group, int groups, int c,
                               int act, float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                               void* z, int64_t z_pitch, const b200::NormFinalizeArgs& fin, void* stream) {
  B200_CHECK_ARG(y && z && rows_per_group > 0 && groups > 0 && c > 0, "norm_act_fwd: bad arguments");
  B200_CHECK_ARG(act != B200SEG_ACT_PRELU || prelu_w, "norm_act_fwd: PReLU needs a slope vector");
  auto st = static_cast<cudaStream_t>(stream);
  const auto* yp = static_cast<const __nv_bfloat16*>(y);
  const auto* rp = static_cast<const __nv_bfloat16*>(residual);
  auto* zp = static_cast<__nv_bfloat16*>(z);
  if (vec_ok(c, y_pitch, z_pitch, residual ? res_pitch : 0)) {
    const int64_t total = rows_per_group * groups * (c / 8);
    B200_ACT_DISPATCH(act, norm_act_fwd_kernel<8, A_><<<ew_grid(total, c / 8), 256, 0, st>>>(yp, y_pitch, coef, rows_per_group, groups, c, act,
                                                                 act_param, prelu_w, rp, res_pitch, zp, z_pitch, fin));
  } else {
    const int64_t total = rows_per_group * groups * c;
    B200_ACT_DISPATCH(act, norm_act_fwd_kernel<1, A_><<<ew_grid(total, c), 256, 0, st>>>(yp, y_pitch, coef, rows_per_group, groups, c, act,
                                                                 act_param, prelu_w, rp, res_pitch, zp, z_pitch, fin));
  }
  B200_CHECK_LAUNCH("norm_act_fwd");
  return 0;
}

int b200seg_norm_act_bwd_reduce(const void* dz, int64_t dz_pitch, const void* y, int64_t y_pitch, const float* coef,
                                int64_t rows_per_group, int groups, int c, int act, float act_param,
                                const float* prelu_w, const void* residual, int64_t res_pitch, float* sums,
                                float* dprelu, float* grad_affine, void* stream) {
  B200_CHECK_ARG(dz && y && sums && rows_per_group > 0 && groups > 0 && c > 0, "norm_act_bwd_reduce: bad arguments");
  B200_CHECK_ARG(!dprelu || groups == 1, "norm_act_bwd_reduce: PReLU gradient needs groups == 1");
  auto st = static_cast<cudaStream_t>(stream);
  const auto* dzp = static_cast<const __nv_bfloat16*>(dz);
  const auto* yp = static_cast<const __nv_bfloat16*>(y);
  const auto* rp = static_cast<const __nv_bfloat16*>(residual);
  // with dprelu the caller passes sums sized [3][c]; the third row is the slope gradient (== dprelu target)
  if (vec_ok(c, dz_pitch, y_pitch) && residual == nullptr && dprelu == nullptr && c / 8 <= 64 && ((c / 8) & (c / 8 - 1)) == 0 &&
      ((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    dim3 block(c / 8, 256 / (c / 8));
    static const int waves = getenv("B200SEG_REDUCE_WAVES") ? std::max(1, atoi(getenv("B200SEG_REDUCE_WAVES"))) : 2;
    dim3 grid(reduce_grid(rows_per_group, block.y, kNumSMs * waves / (groups > 4 ? 4 : 1)), groups);
    B200_ACT_DISPATCH(act, norm_act_bwd_reduce_fast_kernel<A_><<<grid, block, 256 * 16 * sizeof(float), st>>>(
        dzp, dz_pitch, yp, y_pitch, coef, rows_per_group, c, act_param, prelu_w, sums, grad_affine));
  } else if (vec_ok(c, dz_pitch, y_pitch, residual ? res_pitch : 0)) {
    dim3 block = reduce_block(c / 8);
    dim3 grid(reduce_grid(rows_per_group, block.y, kNumSMs * 4 / (groups > 4 ? 4 : 1)), groups);
    B200_ACT_DISPATCH(act, norm_act_bwd_reduce_kernel<8, A_><<<grid, block, 256 * 24 * sizeof(float), st>>>(
        dzp, dz_pitch, yp, y_pitch, coef, rows_per_group, c, act, act_param, prelu_w, rp, res_pitch, sums, dprelu, grad_affine));
  } else {
    dim3 block = reduce_block(c);
    dim3 grid(reduce_grid(rows_per_group, block.y, kNumSMs * 4 / (groups > 4 ? 4 : 1)), groups);
    B200_ACT_DISPATCH(act, norm_act_bwd_reduce_kernel<1, A_><<<grid, block, 256 * 3 * sizeof(float), st>>>(
        dzp, dz_pitch, yp, y_pitch, coef, rows_per_group, c, act, act_param, prelu_w, rp, res_pitch, sums, dprelu, grad_affine));
  }
  B200_CHECK_LAUNCH("norm_act_bwd_reduce");
  return 0;
}

int b200seg_norm_act_bwd_apply(const void* dz, int64_t dz_pitch, const void* y, int64_t y_pitch, const float* coef,
                               const float* sums, double count, int64_t rows_per_group, int groups, int c, int act,
                               float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                               void* dy, int64_t dy_pitch, void* dres, int64_t dres_pitch, void* stream) {
  return b200seg_norm_act_bwd_apply_acc(dz, dz_pitch, y, y_pitch, coef, sums, count, rows_per_group, groups, c, act, act_param,
                                        prelu_w, residual, res_pitch, dy, dy_pitch, dres, dres_pitch, nullptr, 0, stream);
}

int b200seg_norm_act_bwd_apply_acc(const void* dz, int64_t dz_pitch, const void* y, int64_t y_pitch, const float* coef,
                                   const float* sums, double count, int64_t rows_per_group, int groups, int c, int act,
                                   float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                                   void* dy, int64_t dy_pitch, void* dres, int64_t dres_pitch, const void* acc,
                                   int64_t acc_pitch, void* stream) {
  B200_CHECK_ARG(dz && y && dy && rows_per_group > 0 && groups > 0 && c > 0, "norm_act_bwd_apply: bad arguments");
  const auto* ap = static_cast<const __nv_bfloat16*>(acc);
  auto st = static_cast<cudaStream_t>(stream);
  const auto* dzp = static_cast<const __nv_bfloat16*>(dz);
  const auto* yp = static_cast<const __nv_bfloat16*>(y);
  const auto* rp = static_cast<const __nv_bfloat16*>(residual);
  auto* dyp = static_cast<__nv_bfloat16*>(dy);
  auto* drp = static_cast<__nv_bfloat16*>(dres);
  const float inv_count = sums ? static_cast<float>(1.0 / count) : 0.f;
  const int sums_stride = (act == B200SEG_ACT_PRELU) ? 3 : 2;
  if (vec_ok(c, dz_pitch, y_pitch, dy_pitch, (residual ? res_pitch : 0) | (dres ? dres_pitch : 0) | (acc ? acc_pitch : 0))) {
    const int64_t total = rows_per_group * groups * (c / 8);
    B200_ACT_DISPATCH(act, norm_act_bwd_apply_kernel<8, A_><<<ew_grid(total, c / 8), 256, 0, st>>>(
        dzp, dz_pitch, yp, y_pitch, coef, sums, sums_stride, inv_count, rows_per_group, groups, c, act, act_param,
        prelu_w, rp, res_pitch, dyp, dy_pitch, drp, dres_pitch, ap, acc_pitch));
  } else {
    const int64_t total = rows_per_group * groups * c;
    B200_ACT_DISPATCH(act, norm_act_bwd_apply_kernel<1, A_><<<ew_grid(total, c), 256, 0, st>>>(
        dzp, dz_pitch, yp, y_pitch, coef, sums, sums_stride, inv_count, rows_per_group, groups, c, act, act_param,
        prelu_w, rp, res_pitch, dyp, dy_pitch, drp, dres_pitch, ap, acc_pitch));
  }
  B200_CHECK_LAUNCH("norm_act_bwd_apply");
  return 0;
}

int b200seg_maxpool2_fwd(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, uint8_t* idx, int n, int d, int h,
                         int w, int c, void* stream) {
  B200_CHECK_ARG(x && y && idx && n > 0 && d >= 2 && h >= 2 && w >= 2 && c > 0, "maxpool2_fwd: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t ov = static_cast<int64_t>(n) * (d / 2) * (h / 2) * (w / 2);
  if (vec_ok(c, x_pitch, y_pitch))
    maxpool2_fwd_kernel<8><<<grid_for(ov * (c / 8), 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), y_pitch, idx, n, d, h, w, c);
  else
    maxpool2_fwd_kernel<1><<<grid_for(ov * c, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), y_pitch, idx, n, d, h, w, c);
  B200_CHECK_LAUNCH("maxpool2_fwd");
  return 0;
}
int b200seg_maxpool2_bwd(const void* dy, int64_t dy_pitch, const uint8_t* idx, void* dx, int64_t dx_pitch,
                         const void* addend, int64_t addend_pitch, int n, int d, int h, int w, int c, void* stream) {
  B200_CHECK_ARG(dy && dx && idx && n > 0 && d >= 2 && h >= 2 && w >= 2 && c > 0, "maxpool2_bwd: bad arguments");
  B200_CHECK_ARG(d % 2 == 0 && h % 2 == 0 && w % 2 == 0, "maxpool2_bwd: odd extents leave uncovered voxels; zero dx first");
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t ov = static_cast<int64_t>(n) * (d / 2) * (h / 2) * (w / 2);
  const auto* ap = static_cast<const __nv_bfloat16*>(addend);
  if (vec_ok(c, dy_pitch, dx_pitch, addend ? addend_pitch : 0) && (reinterpret_cast<uintptr_t>(addend) & 15) == 0)
    maxpool2_bwd_kernel<8><<<grid_for(ov * (c / 8), 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pitch, idx, static_cast<__nv_bfloat16*>(dx), dx_pitch, ap, addend_pitch,
        n, d, h, w, c);
  else
    maxpool2_bwd_kernel<1><<<grid_for(ov * c, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pitch, idx, static_cast<__nv_bfloat16*>(dx), dx_pitch, ap, addend_pitch,
        n, d, h, w, c);
  B200_CHECK_LAUNCH("maxpool2_bwd");
  return 0;
}
int b200seg_maxpool2_idx_to_torch(const uint8_t* idx, int64_t* out, int n, int d, int h, int w, int c, void* stream) {
  B200_CHECK_ARG(idx && out && n > 0 && c > 0, "maxpool2_idx_to_torch: bad arguments");
  const int64_t total = static_cast<int64_t>(n) * c * (d / 2) * (h / 2) * (w / 2);
  maxpool2_idx_to_torch_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, out, n, d, h,
                                                                                                   w, c);
  B200_CHECK_LAUNCH("maxpool2_idx_to_torch");
  return 0;
}

int b200seg_upsample2_fwd(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int n, int d, int h, int w, int c,
                          void* stream) {
  B200_CHECK_ARG(x && y && n > 0 && c > 0, "upsample2_fwd: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t ov = static_cast<int64_t>(n) * d * h * w * 8;
  if (vec_ok(c, x_pitch, y_pitch))
    upsample2_fwd_kernel<8><<<grid_for(ov * (c / 8), 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), y_pitch, n, d, h, w, c);
  else
    upsample2_fwd_kernel<1><<<grid_for(ov * c, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), y_pitch, n, d, h, w, c);
  B200_CHECK_LAUNCH("upsample2_fwd");
  return 0;
}
int b200seg_upsample2_bwd(const void* dy, int64_t dy_pitch, void* dx, int64_t dx_pitch, int n, int d, int h, int w,
                          int c, void* stream) {
  B200_CHECK_ARG(dy && dx && n > 0 && c > 0, "upsample2_bwd: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t iv = static_cast<int64_t>(n) * d * h * w;
  if (vec_ok(c, dy_pitch, dx_pitch))
    upsample2_bwd_kernel<8><<<grid_for(iv * (c / 8), 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pitch, static_cast<__nv_bfloat16*>(dx), dx_pitch, n, d, h, w, c);
  else
    upsample2_bwd_kernel<1><<<grid_for(iv * c, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pitch, static_cast<__nv_bfloat16*>(dx), dx_pitch, n, d, h, w, c);
  B200_CHECK_LAUNCH("upsample2_bwd");
  return 0;
}
int b200seg_add(const void* a, int64_t a_pitch, const void* b, int64_t b_pitch, void* out, int64_t out_pitch,
                int64_t rows, int c, void* stream) {
  B200_CHECK_ARG(a && b && out && rows > 0 && c > 0, "add: bad arguments");
  auto st = static_cast<cudaStream_t>(stream);
  if (vec_ok(c, a_pitch, b_pitch, out_pitch))
    add_kernel<8><<<grid_for(rows * (c / 8), 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(a), a_pitch, static_cast<const __nv_bfloat16*>(b), b_pitch,
        static_cast<__nv_bfloat16*>(out), out_pitch, rows, c);
  else
    add_kernel<1><<<grid_for(rows * c, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(a), a_pitch, static_cast<const __nv_bfloat16*>(b), b_pitch,
        static_cast<__nv_bfloat16*>(out), out_pitch, rows, c);
  B200_CHECK_LAUNCH("add");
  return 0;
}


int b200seg_dropout2(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int64_t rows, int64_t rows_per_sample,
                     int c, int c_out, float p, const unsigned long long* seed, unsigned long long salt,
                     unsigned long long salt2, int channel_mode, void* stream) {
  B200_CHECK_ARG(x && y && rows > 0 && rows_per_sample > 0 && c > 0 && c_out >= c && y_pitch >= c_out && p >= 0.f && p <= 1.f,
                 "dropout: bad arguments");
  dropout_kernel<<<grid_for(rows * c_out, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), y_pitch, rows, rows_per_sample, c,
      c_out, p, seed, salt, salt2, channel_mode);
  B200_CHECK_LAUNCH("dropout");
  return 0;
}

int b200seg_dropout(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int64_t rows, int64_t rows_per_sample,
                    int c, float p, const unsigned long long* seed, unsigned long long salt, int channel_mode,
                    void* stream) {
  return b200seg_dropout2(x, x_pitch, y, y_pitch, rows, rows_per_sample, c, c, p, seed, salt, 0ULL, channel_mode, stream);
}

int b200seg_classmap_up2_add(const float* coarse, const float* fine, float* out, int64_t planes, int d, int h, int w,
                             void* stream) {
  B200_CHECK_ARG(coarse && out && planes > 0 && d > 0 && h > 0 && w > 0, "classmap_up2_add: bad arguments");
  classmap_up2_add_kernel<<<grid_for(planes * 8 * d * h * w, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      coarse, fine, out, planes, d, h, w);
  B200_CHECK_LAUNCH("classmap_up2_add");
  return 0;
}
int b200seg_classmap_down2_sum(const float* dfine, float* dcoarse, int64_t planes, int d, int h, int w, void* stream) {
  B200_CHECK_ARG(dfine && dcoarse && planes > 0 && d > 0 && h > 0 && w > 0, "classmap_down2_sum: bad arguments");
  classmap_down2_sum_kernel<<<grid_for(planes * d * h * w, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dfine, dcoarse, planes, d, h, w);
  B200_CHECK_LAUNCH("classmap_down2_sum");
  return 0;
}

int b200seg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                      void* stream) {
  B200_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && numel > 0 && step >= 1, "adam_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  adam_kernel<<<grid_for(numel, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
  B200_CHECK_LAUNCH("adam_step");
  return 0;
}

int b200seg_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                          const float* hyper, int32_t* state, void* stream) {
  B200_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && numel > 0 && hyper && state, "adam_step_dev: bad arguments");
  adam_dev_kernel<<<grid_for(numel, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, numel, hyper, state);
  B200_CHECK_LAUNCH("adam_step_dev");
  return 0;
}

}  // extern "C"
