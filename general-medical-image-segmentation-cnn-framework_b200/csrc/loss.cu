// Head (1x1x1 conv to class logits), fused softmax + Dice + CE (+ sigmoid-Dice, BCE) reduction and gradient,
// arg-max label maps, segmentation counts (Dice/IoU metric) and the sliding-window aggregator.
// All HBM-bound: one pass over their inputs, 128-bit loads where the layout allows.
#include <algorithm>

#include "common.cuh"

namespace b200 {

constexpr int kMaxClasses = 8;

// ------------------------------------------------------------------------------------------------ head conv
// x: [n*spatial][cin] bf16 (pitch), logits: [n][classes][spatial] fp32.
__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch, const float* __restrict__ w,
                                const float* __restrict__ b, float* __restrict__ logits, int n, int64_t spatial,
                                int cin, int classes) {
  extern __shared__ float sw[];  // [classes][cin] + [classes]
  for (int i = threadIdx.x; i < classes * cin; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < classes; i += blockDim.x) sw[classes * cin + i] = b ? b[i] : 0.f;
  __syncthreads();
  const int64_t total = static_cast<int64_t>(n) * spatial;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc[kMaxClasses];
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) acc[k] = (k < classes) ? sw[classes * cin + k] : 0.f;
    const __nv_bfloat16* xr = x + v * x_pitch;
    if ((cin & 7) == 0 && (x_pitch & 7) == 0) {
      for (int c0 = 0; c0 < cin; c0 += 8) {
        float f[8];
        unpack8(ld8(xr + c0), f);
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k)
          if (k < classes) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k] += f[j] * sw[k * cin + c0 + j];
          }
      }
    } else {
      for (int c = 0; c < cin; ++c) {
        const float f = __bfloat162float(xr[c]);
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k)
          if (k < classes) acc[k] += f * sw[k * cin + c];
      }
    }
    const int64_t nn = v / spatial, s = v % spatial;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) logits[(nn * classes + k) * spatial + s] = acc[k];
  }
}

// dx[v][c] = sum_k dl[k][v] w[k][c];  grad_w[k][c] += sum_v dl[k][v] x[v][c];  grad_b[k] += sum_v dl[k][v]
// lane = channel (c, c+32, ... up to 4 per lane): x rows and dx rows are read / written coalesced, the class
// gradients of a voxel are warp-uniform broadcast loads, and every lane keeps its own grad_w accumulators in
// registers -- no shuffles in the loop.
template <int KC_, int CPL>   // KC_ = class slots kept in registers (>= classes), CPL = channels per lane
__global__ void __launch_bounds__(256)
    head_bwd_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                    const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int64_t dx_pitch,
                    float* __restrict__ grad_w, float* __restrict__ grad_b, int n, int64_t spatial, int cin,
                    int classes) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float wr[KC_][CPL], gw[KC_][CPL], gb[KC_];
#pragma unroll
  for (int k = 0; k < KC_; ++k) {
    gb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
      wr[k][j] = (k < classes && c < cin) ? w[k * cin + c] : 0.f;
      gw[k][j] = 0.f;
    }
  }
  for (int nn = 0; nn < n; ++nn) {
    const float* dl_n = dlogits + static_cast<int64_t>(nn) * classes * spatial;
    const __nv_bfloat16* x_n = x + static_cast<int64_t>(nn) * spatial * x_pitch;
    __nv_bfloat16* dx_n = dx + static_cast<int64_t>(nn) * spatial * dx_pitch;
    for (int64_t s = warp_id; s < spatial; s += nwarps) {
      float dl[KC_];
#pragma unroll
      for (int k = 0; k < KC_; ++k) dl[k] = (k < classes) ? dl_n[k * spatial + s] : 0.f;
#pragma unroll
      for (int k = 0; k < KC_; ++k) gb[k] += dl[k];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        if (c < cin) {
          const float xv = __bfloat162float(x_n[s * x_pitch + c]);
          float d = 0.f;
#pragma unroll
          for (int k = 0; k < KC_; ++k) {
            d += dl[k] * wr[k][j];
            gw[k][j] += dl[k] * xv;
          }
          dx_n[s * dx_pitch + c] = __float2bfloat16(d);
        }
      }
    }
  }
  __shared__ float sgw[kMaxClasses * 512], sgb[kMaxClasses];
  for (int i = threadIdx.x; i < classes * cin; i += blockDim.x) sgw[i] = 0.f;
  if (threadIdx.x < kMaxClasses) sgb[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KC_; ++k) {
    if (k < classes) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        if (c < cin) atomicAdd(&sgw[k * cin + c], gw[k][j]);
      }
      if (lane == 0) atomicAdd(&sgb[k], gb[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < classes * cin; i += blockDim.x) atomicAdd(&grad_w[i], sgw[i]);
  if (threadIdx.x < classes) atomicAdd(&grad_b[threadIdx.x], sgb[threadIdx.x]);
}

// Fast path of the head backward pass (<= 2 classes, C_in in {16, 32, 64}, vectorisable rows): LPV = C_in / 8 lanes
// share a voxel, each owning 8 channels (one 128-bit load of x, one 128-bit store of dx), so a warp streams
// 32 / LPV whole feature rows per step; the 2 x 8 weight-gradient partial sums per lane are folded across the lanes that
// own the same channels at the very end.
template <int LPV>
__global__ void __launch_bounds__(256)
    head_bwd_vec_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                        const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int64_t dx_pitch,
                        float* __restrict__ grad_w, float* __restrict__ grad_b, int n, int64_t spatial, int classes) {
  constexpr int CIN = LPV * 8;
  constexpr int U = 4;                              // voxels in flight per thread
  __shared__ float sgw[2 * CIN + 2];
  for (int i = threadIdx.x; i < 2 * CIN + 2; i += blockDim.x) sgw[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPV;                       // which 8-channel group of the voxel
  const int c0 = sub * 8;
  float w0[8], w1[8], gw0[8], gw1[8], gb0 = 0.f, gb1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w0[j] = w[c0 + j];
    w1[j] = classes > 1 ? w[CIN + c0 + j] : 0.f;
    gw0[j] = gw1[j] = 0.f;
  }
  (void)n;
  const int64_t nn = blockIdx.y;                    // one sample per grid row: no per-voxel 64-bit division
  const float* dl0 = dlogits + nn * classes * spatial;
  const float* dl1 = classes > 1 ? dl0 + spatial : dl0;
  const __nv_bfloat16* xs = x + nn * spatial * x_pitch + c0;
  __nv_bfloat16* dxs = dx + nn * spatial * dx_pitch + c0;
  const int64_t gthread = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x / LPV;
  auto one = [&](int64_t v, float d0, float d1, const bf16x8& raw) {
    float f[8], o[8];
    unpack8(raw, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = d0 * w0[j] + d1 * w1[j];
      gw0[j] += d0 * f[j];
      gw1[j] += d1 * f[j];
    }
    st8(dxs + v * dx_pitch, pack8(o));
    if (sub == 0) {
      gb0 += d0;
      gb1 += d1;
    }
  };
  int64_t v = gthread / LPV;
  for (; v + (U - 1) * step < spatial; v += U * step) {
    float d0[U], d1[U];
    bf16x8 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      d0[u] = dl0[v + u * step];
      d1[u] = classes > 1 ? dl1[v + u * step] : 0.f;
      raw[u] = ld8(xs + (v + u * step) * x_pitch);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(v + u * step, d0[u], d1[u], raw[u]);
  }
  for (; v < spatial; v += step) one(v, dl0[v], classes > 1 ? dl1[v] : 0.f, ld8(xs + v * x_pitch));
  // fold the lanes that own the same channel group (lane, lane + LPV, lane + 2 LPV, ...)
#pragma unroll
  for (int off = LPV; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gw0[j] += __shfl_xor_sync(0xffffffffu, gw0[j], off);
      gw1[j] += __shfl_xor_sync(0xffffffffu, gw1[j], off);
    }
  }
  gb0 = warp_sum(gb0);
  gb1 = warp_sum(gb1);
  if (lane < LPV) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sgw[c0 + j], gw0[j]);
      atomicAdd(&sgw[CIN + c0 + j], gw1[j]);
    }
  }
  if (lane == 0) {
    atomicAdd(&sgw[2 * CIN], gb0);
    atomicAdd(&sgw[2 * CIN + 1], gb1);
  }
  __syncthreads();
  // ~1200 blocks end on the same three cache lines: 4-wide vector reductions keep the L2 atomic unit out of the profile
  if ((reinterpret_cast<uintptr_t>(grad_w) & 15) == 0) {
    for (int i = threadIdx.x * 4; i < classes * CIN; i += blockDim.x * 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(grad_w + i), "f"(sgw[i]), "f"(sgw[i + 1]),
                   "f"(sgw[i + 2]), "f"(sgw[i + 3])
                   : "memory");
  } else {
    for (int i = threadIdx.x; i < classes * CIN; i += blockDim.x) atomicAdd(&grad_w[i], sgw[i]);
  }
  if (threadIdx.x < classes) atomicAdd(&grad_b[threadIdx.x], sgw[2 * CIN + threadIdx.x]);
}

__global__ void argmax_kernel(const float* __restrict__ logits, uint8_t* __restrict__ labels, int n, int64_t spatial,
                              int classes) {
  const int64_t total = static_cast<int64_t>(n) * spatial;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t nn = v / spatial, s = v % spatial;
    float best = logits[(nn * classes) * spatial + s];
    int bi = 0;
    for (int k = 1; k < classes; ++k) {
      const float f = logits[(nn * classes + k) * spatial + s];
      if (f > best || (f != f && best == best)) {  // ties keep the lowest index; NaN wins like torch.argmax
        best = f;
        bi = k;
      }
    }
    labels[v] = static_cast<uint8_t>(bi);
  }
}

// ------------------------------------------------------------------------------------------------ loss
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void loss_reduce_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, int n,
                                   int64_t spatial, int classes, const float* __restrict__ ce_weight,
                                   double* __restrict__ partial) {
  const int nsum = 1 + 3 * classes + 4;
  float acc[1 + 3 * kMaxClasses + 4];
#pragma unroll
  for (int i = 0; i < 1 + 3 * kMaxClasses + 4; ++i) acc[i] = 0.f;
  const int64_t total = static_cast<int64_t>(n) * spatial;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t nn = v / spatial, s = v % spatial;
    const int t = labels[v];
    float l[kMaxClasses];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) {
        l[k] = logits[(nn * classes + k) * spatial + s];
        m = fmaxf(m, l[k]);
      }
    float z = 0.f, e[kMaxClasses];
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) {
        e[k] = expf(l[k] - m);
        z += e[k];
      }
    const float inv_z = 1.f / z, logz = logf(z);
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) {
        const float p = e[k] * inv_z;
        const float tk = (k == t) ? 1.f : 0.f;
        if (k == t) acc[0] += -(l[k] - m - logz) * (ce_weight ? ce_weight[k] : 1.f);   // nll_loss(weight=...), :13
        acc[1 + 3 * k + 0] += p * tk;
        acc[1 + 3 * k + 1] += p * p;
        acc[1 + 3 * k + 2] += tk;
        const float sg = sigmoidf_(l[k]);
        acc[1 + 3 * classes + 0] += sg * tk;
        acc[1 + 3 * classes + 1] += sg;
        acc[1 + 3 * classes + 2] += tk;
        acc[1 + 3 * classes + 3] += fmaxf(l[k], 0.f) - l[k] * tk + log1pf(expf(-fabsf(l[k])));
      }
  }
  __shared__ double red[1 + 3 * kMaxClasses + 4];
  for (int i = threadIdx.x; i < nsum; i += blockDim.x) red[i] = 0.0;
  __syncthreads();
  for (int i = 0; i < nsum; ++i) {
    const double t = warp_sum_d(static_cast<double>(acc[i]));
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], t);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nsum; i += blockDim.x) atomicAdd(&partial[i], red[i]);
}

__global__ void loss_grad_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, int n,
                                 int64_t spatial, int classes, const double* __restrict__ partial, float w_ce,
                                 float w_dice, float w_sdice, float w_bce, const float* __restrict__ gscale_p,
                                 const float* __restrict__ class_weights, float* __restrict__ dlogits) {
  const float gscale = gscale_p ? *gscale_p : 1.f;
  __shared__ float ck_t[kMaxClasses], ck_p[kMaxClasses];  // dDice/dp_k = ck_t*t_k + ck_p*p_k
  __shared__ float ce_w[kMaxClasses];                      // class_weights = {ce[classes], dice[classes]} or null
  __shared__ float sd_t, sd_c;
  const int64_t total = static_cast<int64_t>(n) * spatial;
  if (threadIdx.x < classes) {
    const int k = threadIdx.x;
    const double smooth = 1e-5;
    const double I = partial[1 + 3 * k], Z = partial[2 + 3 * k], Y = partial[3 + 3 * k];
    const double D = Z + Y + smooth;
    // L_k = 1 - (2I+s)/D ; dL_k/dp = -2 t / D + (2I+s) * 2p / D^2 ; averaged over classes
    const double wd = class_weights ? class_weights[classes + k] : 1.0;   // loss += dice_k * weight[k], :183
    ck_t[k] = static_cast<float>(-2.0 / D / classes * wd);
    ck_p[k] = static_cast<float>(2.0 * (2.0 * I + smooth) / (D * D) / classes * wd);
    ce_w[k] = class_weights ? class_weights[k] : 1.f;
  }
  if (threadIdx.x == 0) {
    const double eps = 1e-5;
    const double I = partial[1 + 3 * classes], U = partial[2 + 3 * classes] + partial[3 + 3 * classes];
    // L = 1 - 2(I+e)/(U+e): dL/dsig = -2 [ t (U+e) - (I+e) ] / (U+e)^2
    sd_t = static_cast<float>(-2.0 / (U + eps));
    sd_c = static_cast<float>(2.0 * (I + eps) / ((U + eps) * (U + eps)));
  }
  __syncthreads();
  const float inv_vox = 1.f / static_cast<float>(total);
  const float inv_elems = inv_vox / classes;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t nn = v / spatial, s = v % spatial;
    const int t = labels[v];
    float l[kMaxClasses], p[kMaxClasses], a[kMaxClasses];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) {
        l[k] = logits[(nn * classes + k) * spatial + s];
        m = fmaxf(m, l[k]);
      }
    float z = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) {
        p[k] = expf(l[k] - m);
        z += p[k];
      }
    const float inv_z = 1.f / z;
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) {
        p[k] *= inv_z;
        const float tk = (k == t) ? 1.f : 0.f;
        a[k] = ck_t[k] * tk + ck_p[k] * p[k];
        dot += a[k] * p[k];
      }
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k)
      if (k < classes) {
        const float tk = (k == t) ? 1.f : 0.f;
        float g = w_ce * ce_w[t < classes ? t : 0] * (p[k] - tk) * inv_vox + w_dice * p[k] * (a[k] - dot);
        if (w_sdice != 0.f || w_bce != 0.f) {
          const float sg = sigmoidf_(l[k]);
          g += w_sdice * (sd_t * tk + sd_c) * sg * (1.f - sg) + w_bce * (sg - tk) * inv_elems;
        }
        dlogits[(nn * classes + k) * spatial + s] = g * gscale;
      }
  }
}

// ---- Dice on PROBABILITIES supplied by the caller: BinaryDiceLoss (loss_function.py:61-99, one loss per sample) and
// DiceLossss without its soft-max (:172-184, one loss per class).  pred is fp32 [N][K][S]; the target is either a float
// tensor of the same layout or uint8 labels [N][S] (target of class k = (label == k), the comparison of :154-160).
// Row r of the reduction is the sample (by_class = 0) or the class (by_class = 1): partial[r] = {sum x*t, sum x^p, sum t^p}.
__device__ __forceinline__ float pow_p(float v, float p) { return p == 2.f ? v * v : (p == 1.f ? v : powf(v, p)); }

__global__ void __launch_bounds__(256) dice_sums_kernel(const float* __restrict__ pred, const float* __restrict__ tf,
                                                        const uint8_t* __restrict__ tl, int classes, int64_t spatial,
                                                        int by_class, float p_exp, double* __restrict__ partial) {
  const int plane = blockIdx.y, nn = plane / classes, k = plane % classes;
  const float* x = pred + static_cast<int64_t>(plane) * spatial;
  const float* t_f = tf ? tf + static_cast<int64_t>(plane) * spatial : nullptr;
  const uint8_t* t_l = tl ? tl + static_cast<int64_t>(nn) * spatial : nullptr;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < spatial;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    const float t = t_f ? t_f[i] : (t_l[i] == k ? 1.f : 0.f);
    a0 = fmaf(v, t, a0);
    a1 += pow_p(v, p_exp);
    a2 += t_f ? pow_p(t, p_exp) : t;
  }
  __shared__ double red[3];
  if (threadIdx.x < 3) red[threadIdx.x] = 0.0;
  __syncthreads();
  const double s0 = warp_sum_d(a0), s1 = warp_sum_d(a1), s2 = warp_sum_d(a2);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&red[0], s0);
    atomicAdd(&red[1], s1);
    atomicAdd(&red[2], s2);
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicAdd(&partial[3 * (by_class ? k : nn) + threadIdx.x], red[threadIdx.x]);
}

// dpred = gscale * (coef_t[r] * t + coef_x[r] * p * x^(p-1))
__global__ void __launch_bounds__(256) dice_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tf,
                                                        const uint8_t* __restrict__ tl, int classes, int64_t spatial,
                                                        int by_class, float p_exp, const float* __restrict__ coef_t,
                                                        const float* __restrict__ coef_x, const float* __restrict__ gscale_p,
                                                        float* __restrict__ dpred) {
  const int plane = blockIdx.y, nn = plane / classes, k = plane % classes;
  const int r = by_class ? k : nn;
  const float gs = gscale_p ? *gscale_p : 1.f;
  const float ct = coef_t[r] * gs, cx = coef_x[r] * gs * p_exp;
  const float* x = pred + static_cast<int64_t>(plane) * spatial;
  float* dx = dpred + static_cast<int64_t>(plane) * spatial;
  const float* t_f = tf ? tf + static_cast<int64_t>(plane) * spatial : nullptr;
  const uint8_t* t_l = tl ? tl + static_cast<int64_t>(nn) * spatial : nullptr;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < spatial;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    const float t = t_f ? t_f[i] : (t_l[i] == k ? 1.f : 0.f);
    const float vp = p_exp == 2.f ? v : (p_exp == 1.f ? 1.f : powf(v, p_exp - 1.f));
    dx[i] = ct * t + cx * vp;
  }
}

// ---- two-class fast paths (the reference's segmentation setting, train.py:331): four consecutive voxels per thread and
// iteration (one 128-bit load per class plane, one 32-bit load of labels), one sample per grid row (no per-voxel 64-bit
// division), two iterations in flight, and only the sums the configured loss needs.  With two classes the soft-max needs
// ONE exponential: the larger logit has e = 1 and the other e = exp(-|l1 - l0|).
struct Soft2 {
  float p0, p1, nll_if0, nll_if1;   // probabilities and -log p_k
};
__device__ __forceinline__ Soft2 softmax2(float l0, float l1) {
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  const float z = e0 + e1, inv_z = 1.f / z, logz = logf(z);
  Soft2 r;
  r.p0 = e0 * inv_z;
  r.p1 = e1 * inv_z;
  r.nll_if0 = -(l0 - m - logz);
  r.nll_if1 = -(l1 - m - logz);
  return r;
}

template <bool SOFT, bool SIG>
__global__ void __launch_bounds__(256)
    loss_reduce2_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, int64_t spatial,
                        double* __restrict__ partial) {
  constexpr int NS = 11;   // 1 + 3*2 + 4
  float acc[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) acc[i] = 0.f;
  const int64_t nn = blockIdx.y;
  const float4* l0 = reinterpret_cast<const float4*>(logits + nn * 2 * spatial);
  const float4* l1 = reinterpret_cast<const float4*>(logits + (nn * 2 + 1) * spatial);
  const uchar4* lab = reinterpret_cast<const uchar4*>(labels + nn * spatial);
  const int64_t quads = spatial >> 2;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto one = [&](float a, float b, int t) {
    const float t0 = (t == 0) ? 1.f : 0.f, t1 = (t == 1) ? 1.f : 0.f;
    if (SOFT) {
      const Soft2 r = softmax2(a, b);
      acc[0] += (t == 0) ? r.nll_if0 : ((t == 1) ? r.nll_if1 : 0.f);
      acc[1] += r.p0 * t0;
      acc[2] += r.p0 * r.p0;
      acc[3] += t0;
      acc[4] += r.p1 * t1;
      acc[5] += r.p1 * r.p1;
      acc[6] += t1;
    }
    if (SIG) {
      const float s0 = sigmoidf_(a), s1 = sigmoidf_(b);
      acc[7] += s0 * t0 + s1 * t1;
      acc[8] += s0 + s1;
      acc[9] += t0 + t1;
      acc[10] += (fmaxf(a, 0.f) - a * t0 + log1pf(expf(-fabsf(a)))) + (fmaxf(b, 0.f) - b * t1 + log1pf(expf(-fabsf(b))));
    }
  };
  auto quad = [&](const float4& a, const float4& b, const uchar4& t) {
    one(a.x, b.x, t.x);
    one(a.y, b.y, t.y);
    one(a.z, b.z, t.z);
    one(a.w, b.w, t.w);
  };
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; q + step < quads; q += 2 * step) {
    const float4 a0 = l0[q], b0 = l1[q], a1 = l0[q + step], b1 = l1[q + step];
    const uchar4 t0 = lab[q], t1 = lab[q + step];
    quad(a0, b0, t0);
    quad(a1, b1, t1);
  }
  for (; q < quads; q += step) quad(l0[q], l1[q], lab[q]);
  __shared__ double red[NS];
  if (threadIdx.x < NS) red[threadIdx.x] = 0.0;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    if ((i < 7 && !SOFT) || (i >= 7 && !SIG)) continue;
    const float t = warp_sum(acc[i]);   // <= 32 x (a few hundred) fp32 terms per warp; doubles from here on
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], static_cast<double>(t));
  }
  __syncthreads();
  if (threadIdx.x < NS && ((threadIdx.x < 7) ? SOFT : SIG)) atomicAdd(&partial[threadIdx.x], red[threadIdx.x]);
}

template <bool SIG>
__global__ void __launch_bounds__(256)
    loss_grad2_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, int n, int64_t spatial,
                      const double* __restrict__ partial, float w_ce, float w_dice, float w_sdice, float w_bce,
                      const float* __restrict__ gscale_p, float* __restrict__ dlogits) {
  const float gscale = gscale_p ? *gscale_p : 1.f;
  float ck_t[2], ck_p[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {   // same coefficients as loss_grad_kernel, recomputed per thread (a handful of flops)
    const double smooth = 1e-5;
    const double I = partial[1 + 3 * k], Z = partial[2 + 3 * k], Y = partial[3 + 3 * k];
    const double D = Z + Y + smooth;
    ck_t[k] = static_cast<float>(-2.0 / D / 2);
    ck_p[k] = static_cast<float>(2.0 * (2.0 * I + smooth) / (D * D) / 2);
  }
  float sd_t = 0.f, sd_c = 0.f;
  if (SIG) {
    const double eps = 1e-5;
    const double I = partial[7], U = partial[8] + partial[9];
    sd_t = static_cast<float>(-2.0 / (U + eps));
    sd_c = static_cast<float>(2.0 * (I + eps) / ((U + eps) * (U + eps)));
  }
  const float inv_vox = 1.f / static_cast<float>(static_cast<int64_t>(n) * spatial);
  const float inv_elems = inv_vox / 2;
  const int64_t nn = blockIdx.y;
  const float4* l0 = reinterpret_cast<const float4*>(logits + nn * 2 * spatial);
  const float4* l1 = reinterpret_cast<const float4*>(logits + (nn * 2 + 1) * spatial);
  float4* g0 = reinterpret_cast<float4*>(dlogits + nn * 2 * spatial);
  float4* g1 = reinterpret_cast<float4*>(dlogits + (nn * 2 + 1) * spatial);
  const uchar4* lab = reinterpret_cast<const uchar4*>(labels + nn * spatial);
  const int64_t quads = spatial >> 2;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto one = [&](float a, float b, int t, float& ga, float& gb) {
    const float t0 = (t == 0) ? 1.f : 0.f, t1 = (t == 1) ? 1.f : 0.f;
    const Soft2 r = softmax2(a, b);
    const float a0 = ck_t[0] * t0 + ck_p[0] * r.p0, a1 = ck_t[1] * t1 + ck_p[1] * r.p1;
    const float dot = a0 * r.p0 + a1 * r.p1;
    ga = w_ce * (r.p0 - t0) * inv_vox + w_dice * r.p0 * (a0 - dot);
    gb = w_ce * (r.p1 - t1) * inv_vox + w_dice * r.p1 * (a1 - dot);
    if (SIG) {
      const float s0 = sigmoidf_(a), s1 = sigmoidf_(b);
      ga += w_sdice * (sd_t * t0 + sd_c) * s0 * (1.f - s0) + w_bce * (s0 - t0) * inv_elems;
      gb += w_sdice * (sd_t * t1 + sd_c) * s1 * (1.f - s1) + w_bce * (s1 - t1) * inv_elems;
    }
    ga *= gscale;
    gb *= gscale;
  };
  auto quad = [&](int64_t q, const float4& a, const float4& b, const uchar4& t) {
    float4 oa, ob;
    one(a.x, b.x, t.x, oa.x, ob.x);
    one(a.y, b.y, t.y, oa.y, ob.y);
    one(a.z, b.z, t.z, oa.z, ob.z);
    one(a.w, b.w, t.w, oa.w, ob.w);
    g0[q] = oa;
    g1[q] = ob;
  };
  int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; q + step < quads; q += 2 * step) {
    const float4 a0 = l0[q], b0 = l1[q], a1 = l0[q + step], b1 = l1[q + step];
    const uchar4 t0 = lab[q], t1 = lab[q + step];
    quad(q, a0, b0, t0);
    quad(q + step, a1, b1, t1);
  }
  for (; q < quads; q += step) quad(q, l0[q], l1[q], lab[q]);
}

// ------------------------------------------------------------------------------------------------ metric
__global__ void seg_counts_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred, int64_t numel,
                                  unsigned long long* __restrict__ counts) {
  unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  const int64_t nvec = numel / 16;
  const uint4* g4 = reinterpret_cast<const uint4*>(gt);
  const uint4* p4 = reinterpret_cast<const uint4*>(pred);
  const bool aligned = ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
  auto one = [&](unsigned g, unsigned p) {
    c0 += g;
    c1 += p;
    c2 += (g & p) != 0;
    c3 += (g | p) != 0;
  };
  int64_t done = 0;
  if (aligned) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const uint4 a = g4[i], b = p4[i];
      const unsigned aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) one((aw[q] >> (8 * j)) & 0xFF, (bw[q] >> (8 * j)) & 0xFF);
    }
    done = nvec * 16;
  }
  for (int64_t i = done + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    one(gt[i], pred[i]);
  unsigned long long vals[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    unsigned long long t = vals[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(&counts[q], t);
  }
}

// ------------------------------------------------------------------------------------------------ sliding window
// Crop mode keeps torchio's semantics where cropped interiors overlap (the extra window at the far border): the patch
// that comes LATER in sampler order wins.  Every voxel holds a key ((patch id + 1) << 8 | label) and patches are merged
// with atomicMax, which is order-independent: correct within a batch, across batches and across ranks (all-reduce MAX).
__global__ void window_crop_kernel(const uint8_t* __restrict__ patches, const int64_t* __restrict__ loc,
                                   const int64_t* __restrict__ patch_ids, int64_t first_id, int pw, int ph, int pd,
                                   int ow, int oh, int od, int* __restrict__ keys, int vw, int vh, int vd) {
  const int b = blockIdx.y;
  const int64_t* L = loc + static_cast<int64_t>(b) * 6;
  const int i0 = static_cast<int>(L[0]), j0 = static_cast<int>(L[1]), k0 = static_cast<int>(L[2]);
  const int i1 = static_cast<int>(L[3]), j1 = static_cast<int>(L[4]), k1 = static_cast<int>(L[5]);
  const int64_t id = patch_ids ? patch_ids[b] : first_id + b;
  const int keyhi = static_cast<int>((id + 1) << 8);
  // trim overlap/2 from every face that is not on the volume border
  const int li = i0 > 0 ? ow / 2 : 0, lj = j0 > 0 ? oh / 2 : 0, lk = k0 > 0 ? od / 2 : 0;
  const int hi = i1 < vw ? ow / 2 : 0, hj = j1 < vh ? oh / 2 : 0, hk = k1 < vd ? od / 2 : 0;
  const int64_t pvox = static_cast<int64_t>(pw) * ph * pd;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < pvox;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(t % pd);
    const int j = static_cast<int>((t / pd) % ph);
    const int i = static_cast<int>(t / (static_cast<int64_t>(pd) * ph));
    if (i < li || i >= pw - hi || j < lj || j >= ph - hj || k < lk || k >= pd - hk) continue;
    atomicMax(&keys[(static_cast<int64_t>(i0 + i) * vh + (j0 + j)) * vd + (k0 + k)], keyhi | patches[b * pvox + t]);
  }
}

__global__ void window_keys_to_labels_kernel(const int* __restrict__ keys, uint8_t* __restrict__ labels, int64_t voxels) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < voxels;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x)
    labels[v] = static_cast<uint8_t>(keys[v] & 255);
}

__global__ void window_avg_kernel(const float* __restrict__ patches, const int64_t* __restrict__ loc, int c, int pw,
                                  int ph, int pd, float* __restrict__ acc, float* __restrict__ count, int vw, int vh,
                                  int vd) {
  const int b = blockIdx.y;
  const int64_t* L = loc + static_cast<int64_t>(b) * 6;
  const int i0 = static_cast<int>(L[0]), j0 = static_cast<int>(L[1]), k0 = static_cast<int>(L[2]);
  const int64_t pvox = static_cast<int64_t>(pw) * ph * pd;
  const int64_t vvox = static_cast<int64_t>(vw) * vh * vd;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < pvox;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(t % pd);
    const int j = static_cast<int>((t / pd) % ph);
    const int i = static_cast<int>(t / (static_cast<int64_t>(pd) * ph));
    const int64_t o = (static_cast<int64_t>(i0 + i) * vh + (j0 + j)) * vd + (k0 + k);
    for (int ch = 0; ch < c; ++ch) atomicAdd(&acc[ch * vvox + o], patches[(static_cast<int64_t>(b) * c + ch) * pvox + t]);
    atomicAdd(&count[o], 1.f);
  }
}

__global__ void window_finalize_kernel(float* __restrict__ acc, const float* __restrict__ count, int c, int64_t voxels,
                                       uint8_t* __restrict__ labels) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < voxels;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float inv = 1.f / fmaxf(count[v], 1.f);
    float best = 0.f;
    int bi = 0;
    for (int ch = 0; ch < c; ++ch) {
      const float f = acc[ch * voxels + v] * inv;
      acc[ch * voxels + v] = f;
      if (ch == 0 || f > best) {
        best = f;
        bi = ch;
      }
    }
    if (labels) labels[v] = static_cast<uint8_t>(bi);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200seg_head_conv1x1_fwd(const void* x, int64_t x_pitch, const float* w, const float* b, float* logits, int n,
                             int64_t spatial, int cin, int classes, void* stream) {
  B200_CHECK_ARG(x && w && logits && n > 0 && spatial > 0 && cin > 0, "head_conv1x1_fwd: bad arguments");
  B200_CHECK_ARG(classes >= 1 && classes <= kMaxClasses, "head_conv1x1_fwd: classes must be in [1,%d]", kMaxClasses);
  const size_t smem = (static_cast<size_t>(classes) * cin + classes) * sizeof(float);
  head_fwd_kernel<<<grid_for(static_cast<int64_t>(n) * spatial, 256), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_pitch, w, b, logits, n, spatial, cin, classes);
  B200_CHECK_LAUNCH("head_conv1x1_fwd");
  return 0;
}

int b200seg_head_conv1x1_bwd(const float* dlogits, const void* x, int64_t x_pitch, const float* w, void* dx,
                             int64_t dx_pitch, float* grad_w, float* grad_b, int n, int64_t spatial, int cin,
                             int classes, void* stream) {
  B200_CHECK_ARG(dlogits && x && w && dx && grad_w && grad_b && n > 0 && spatial > 0 && cin > 0,
                 "head_conv1x1_bwd: bad arguments");
  B200_CHECK_ARG(classes >= 1 && classes <= kMaxClasses, "head_conv1x1_bwd: classes must be in [1,%d]", kMaxClasses);
  B200_CHECK_ARG(cin <= 128 || (cin <= 512 && classes <= 4), "head_conv1x1_bwd: at most 128 input channels (512 for <= 4 classes)");
  auto st = static_cast<cudaStream_t>(stream);
  const auto* xp = static_cast<const __nv_bfloat16*>(x);
  auto* dxp = static_cast<__nv_bfloat16*>(dx);
  if (classes <= 2 && (cin == 16 || cin == 32 || cin == 64) && x_pitch % 8 == 0 && dx_pitch % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
    const int per_sample = std::max(1, kNumSMs * 8 / n);
    const dim3 g(static_cast<unsigned>(std::min<int64_t>(per_sample, (spatial * (cin / 8) + 255) / 256)), n);
    if (cin == 16)
      head_bwd_vec_kernel<2><<<g, 256, 0, st>>>(dlogits, xp, x_pitch, w, dxp, dx_pitch, grad_w, grad_b, n, spatial, classes);
    else if (cin == 32)
      head_bwd_vec_kernel<4><<<g, 256, 0, st>>>(dlogits, xp, x_pitch, w, dxp, dx_pitch, grad_w, grad_b, n, spatial, classes);
    else
      head_bwd_vec_kernel<8><<<g, 256, 0, st>>>(dlogits, xp, x_pitch, w, dxp, dx_pitch, grad_w, grad_b, n, spatial, classes);
    B200_CHECK_LAUNCH("head_conv1x1_bwd");
    return 0;
  }
  const int grid = kNumSMs * 8;
#define B200_HEAD_BWD(KC_, CPL_) \
  head_bwd_kernel<KC_, CPL_><<<grid, 256, 0, st>>>(dlogits, xp, x_pitch, w, dxp, dx_pitch, grad_w, grad_b, n, spatial, cin, classes)
  const int cpl = (cin + 31) / 32;
  if (classes <= 2 && cpl == 1) B200_HEAD_BWD(2, 1);
  else if (classes <= 4 && cpl == 1) B200_HEAD_BWD(4, 1);
  else if (classes <= 4 && cpl <= 2) B200_HEAD_BWD(4, 2);
  else if (cpl <= 4) B200_HEAD_BWD(8, 4);
  else if (cpl <= 8) B200_HEAD_BWD(4, 8);      // deep-supervision heads of the residual U-Net (256 -> classes)
  else B200_HEAD_BWD(4, 16);
#undef B200_HEAD_BWD
  B200_CHECK_LAUNCH("head_conv1x1_bwd");
  return 0;
}

int b200seg_argmax_labels(const float* logits, uint8_t* labels, int n, int64_t spatial, int classes, void* stream) {
  B200_CHECK_ARG(logits && labels && n > 0 && spatial > 0 && classes >= 1 && classes <= 255, "argmax_labels: bad arguments");
  argmax_kernel<<<grid_for(static_cast<int64_t>(n) * spatial, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, labels, n, spatial, classes);
  B200_CHECK_LAUNCH("argmax_labels");
  return 0;
}

int b200seg_loss_reduce(const float* logits, const uint8_t* labels, int n, int64_t spatial, int classes, int terms,
                        const float* class_weights, double* partial, void* stream) {
  B200_CHECK_ARG(logits && labels && partial && n > 0 && spatial > 0, "loss_reduce: bad arguments");
  B200_CHECK_ARG(classes >= 1 && classes <= kMaxClasses, "loss_reduce: classes must be in [1,%d]", kMaxClasses);
  B200_CHECK_ARG(terms >= 1 && terms <= 3, "loss_reduce: terms must be a mask of 1 (soft-max sums) | 2 (sigmoid sums)");
  if (classes == 2 && !class_weights && spatial % 4 == 0 && n <= 65535 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(labels) & 3) == 0) {
    const int per_sample = std::max(1, kNumSMs * 4 / n);
    const dim3 grid(static_cast<unsigned>(std::min<int64_t>(per_sample, (spatial / 4 + 511) / 512)), n);
    auto st = static_cast<cudaStream_t>(stream);
    if (terms == 1) loss_reduce2_kernel<true, false><<<grid, 256, 0, st>>>(logits, labels, spatial, partial);
    else if (terms == 2) loss_reduce2_kernel<false, true><<<grid, 256, 0, st>>>(logits, labels, spatial, partial);
    else loss_reduce2_kernel<true, true><<<grid, 256, 0, st>>>(logits, labels, spatial, partial);
    B200_CHECK_LAUNCH("loss_reduce");
    return 0;
  }
  loss_reduce_kernel<<<grid_for(static_cast<int64_t>(n) * spatial, 256, kNumSMs * 8), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(logits, labels, n, spatial, classes, class_weights, partial);
  B200_CHECK_LAUNCH("loss_reduce");
  return 0;
}

int b200seg_loss_grad(const float* logits, const uint8_t* labels, int n, int64_t spatial, int classes,
                      const double* partial, float w_ce, float w_dice, float w_sdice, float w_bce,
                      const float* gscale, const float* class_weights, float* dlogits, void* stream) {
  B200_CHECK_ARG(logits && labels && partial && dlogits && n > 0 && spatial > 0, "loss_grad: bad arguments");
  B200_CHECK_ARG(classes >= 1 && classes <= kMaxClasses, "loss_grad: classes must be in [1,%d]", kMaxClasses);
  if (classes == 2 && !class_weights && spatial % 4 == 0 && n <= 65535 &&
      ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(labels) & 3) == 0) {
    const int per_sample = std::max(1, kNumSMs * 8 / n);
    const dim3 grid(static_cast<unsigned>(std::min<int64_t>(per_sample, (spatial / 4 + 511) / 512)), n);
    auto st = static_cast<cudaStream_t>(stream);
    if (w_sdice != 0.f || w_bce != 0.f)
      loss_grad2_kernel<true><<<grid, 256, 0, st>>>(logits, labels, n, spatial, partial, w_ce, w_dice, w_sdice, w_bce, gscale, dlogits);
    else
      loss_grad2_kernel<false><<<grid, 256, 0, st>>>(logits, labels, n, spatial, partial, w_ce, w_dice, w_sdice, w_bce, gscale, dlogits);
    B200_CHECK_LAUNCH("loss_grad");
    return 0;
  }
  loss_grad_kernel<<<grid_for(static_cast<int64_t>(n) * spatial, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, labels, n, spatial, classes, partial, w_ce, w_dice, w_sdice, w_bce, gscale, class_weights, dlogits);
  B200_CHECK_LAUNCH("loss_grad");
  return 0;
}

int b200seg_dice_sums(const float* pred, const float* target_f, const uint8_t* target_l, int n, int classes, int64_t spatial,
                      int by_class, float p_exp, double* partial, void* stream) {
  B200_CHECK_ARG(pred && partial && (target_f != nullptr) != (target_l != nullptr) && n > 0 && classes > 0 && spatial > 0 &&
                     p_exp > 0.f, "dice_sums: bad arguments (exactly one of the two target forms)");
  const int64_t planes = static_cast<int64_t>(n) * classes;
  B200_CHECK_ARG(planes <= 65535, "dice_sums: at most 65535 (sample, class) planes");
  const int per_plane = static_cast<int>(std::min<int64_t>(std::max<int64_t>(1, kNumSMs * 8 / planes), (spatial + 2047) / 2048));
  dice_sums_kernel<<<dim3(per_plane, static_cast<unsigned>(planes)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, target_f, target_l, classes, spatial, by_class, p_exp, partial);
  B200_CHECK_LAUNCH("dice_sums");
  return 0;
}

int b200seg_dice_grad(const float* pred, const float* target_f, const uint8_t* target_l, int n, int classes, int64_t spatial,
                      int by_class, float p_exp, const float* coef_t, const float* coef_x, const float* gscale, float* dpred,
                      void* stream) {
  B200_CHECK_ARG(pred && dpred && coef_t && coef_x && (target_f != nullptr) != (target_l != nullptr) && n > 0 && classes > 0 &&
                     spatial > 0 && p_exp > 0.f, "dice_grad: bad arguments");
  const int64_t planes = static_cast<int64_t>(n) * classes;
  B200_CHECK_ARG(planes <= 65535, "dice_grad: at most 65535 (sample, class) planes");
  const int per_plane = static_cast<int>(std::min<int64_t>(std::max<int64_t>(1, kNumSMs * 8 / planes), (spatial + 1023) / 1024));
  dice_grad_kernel<<<dim3(per_plane, static_cast<unsigned>(planes)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, target_f, target_l, classes, spatial, by_class, p_exp, coef_t, coef_x, gscale, dpred);
  B200_CHECK_LAUNCH("dice_grad");
  return 0;
}

int b200seg_seg_counts(const uint8_t* gt, const uint8_t* pred, int64_t numel, unsigned long long* counts,
                       void* stream) {
  B200_CHECK_ARG(gt && pred && counts && numel >= 0, "seg_counts: bad arguments");
  if (numel == 0) return 0;
  seg_counts_kernel<<<grid_for((numel + 15) / 16, 256, kNumSMs * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gt, pred, numel, counts);
  B200_CHECK_LAUNCH("seg_counts");
  return 0;
}

int b200seg_window_accumulate_crop(const uint8_t* patches, const int64_t* locations, const int64_t* patch_ids,
                                   int64_t first_id, int batch, int pw, int ph, int pd, int ow, int oh, int od,
                                   int32_t* keys, int vw, int vh, int vd, void* stream) {
  B200_CHECK_ARG(patches && locations && keys && batch > 0 && pw > 0 && ph > 0 && pd > 0, "window_crop: bad arguments");
  B200_CHECK_ARG(ow % 2 == 0 && oh % 2 == 0 && od % 2 == 0, "window_crop: overlap must be even");
  B200_CHECK_ARG(first_id >= 0 && first_id + batch < (1 << 23), "window_crop: patch ids must fit 23 bits");
  dim3 grid(grid_for(static_cast<int64_t>(pw) * ph * pd, 256, kNumSMs * 4), batch);
  window_crop_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(patches, locations, patch_ids, first_id, pw, ph,
                                                                          pd, ow, oh, od, keys, vw, vh, vd);
  B200_CHECK_LAUNCH("window_crop");
  return 0;
}

int b200seg_window_keys_to_labels(const int32_t* keys, uint8_t* labels, int64_t voxels, void* stream) {
  B200_CHECK_ARG(keys && labels && voxels > 0, "window_keys_to_labels: bad arguments");
  window_keys_to_labels_kernel<<<grid_for(voxels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, labels, voxels);
  B200_CHECK_LAUNCH("window_keys_to_labels");
  return 0;
}

int b200seg_window_accumulate_average(const float* patches, const int64_t* locations, int batch, int c, int pw, int ph,
                                      int pd, float* acc, float* count, int vw, int vh, int vd, void* stream) {
  B200_CHECK_ARG(patches && locations && acc && count && batch > 0 && c > 0, "window_average: bad arguments");
  dim3 grid(grid_for(static_cast<int64_t>(pw) * ph * pd, 256, kNumSMs * 4), batch);
  window_avg_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(patches, locations, c, pw, ph, pd, acc, count,
                                                                         vw, vh, vd);
  B200_CHECK_LAUNCH("window_average");
  return 0;
}

int b200seg_window_finalize(float* acc, const float* count, int c, int64_t voxels, uint8_t* labels, void* stream) {
  B200_CHECK_ARG(acc && count && c > 0 && c <= 255 && voxels > 0, "window_finalize: bad arguments");
  window_finalize_kernel<<<grid_for(voxels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, count, c, voxels,
                                                                                              labels);
  B200_CHECK_LAUNCH("window_finalize");
  return 0;
}

}  // extern "C"
