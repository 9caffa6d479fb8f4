// Direct (CUDA-core) convolution path: every geometry the tensor-core path does not take -- C_in = 1 stems, class
// heads, channel counts that are not multiples of 16, strided (k2s2 / k3s2) convolutions and their transposes.
// bf16 operands, fp32 accumulation.  Also weight packing / gradient unpacking shared with the tcgen05 path.
#include <algorithm>

#include "common.cuh"
#include "conv_impl.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------ packing
// w [cout][cin][k3] fp32 -> fprop: p[t][co][ci_local] ; dgrad: p[k3-1-t][ci_local][co]   (bf16)
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ p, int cout, int cin,
                                        int k3, int cin_off, int cin_cnt, int dgrad) {
  const int64_t total = static_cast<int64_t>(k3) * cout * cin_cnt;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int t, co, ci;
    if (!dgrad) {
      ci = static_cast<int>(i % cin_cnt);
      co = static_cast<int>((i / cin_cnt) % cout);
      t = static_cast<int>(i / (static_cast<int64_t>(cin_cnt) * cout));
      p[i] = cin_off + ci < cin ? __float2bfloat16(w[(static_cast<int64_t>(co) * cin + cin_off + ci) * k3 + t])
                                : __float2bfloat16(0.f);   // channels past the tensor: zero padding of the K dimension
    } else {
      co = static_cast<int>(i % cout);
      ci = static_cast<int>((i / cout) % cin_cnt);
      t = static_cast<int>(i / (static_cast<int64_t>(cin_cnt) * cout));
      p[i] = cin_off + ci < cin ? __float2bfloat16(w[(static_cast<int64_t>(co) * cin + cin_off + ci) * k3 + (k3 - 1 - t)])
                                : __float2bfloat16(0.f);
    }
  }
}

// The same packs with both channel counts padded (zero rows / columns): p[t][cout_p][cin_p] or p[k3-1-t][cin_p][cout_p].
__global__ void pack_conv_weight_padded_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ p, int cout,
                                               int cin, int k3, int cout_p, int cin_p, int dgrad) {
  const int64_t total = static_cast<int64_t>(k3) * cout_p * cin_p;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int t, co, ci;
    if (!dgrad) {
      ci = static_cast<int>(i % cin_p);
      co = static_cast<int>((i / cin_p) % cout_p);
    } else {
      co = static_cast<int>(i % cout_p);
      ci = static_cast<int>((i / cout_p) % cin_p);
    }
    t = static_cast<int>(i / (static_cast<int64_t>(cin_p) * cout_p));
    if (dgrad) t = k3 - 1 - t;
    p[i] = (co < cout && ci < cin) ? __float2bfloat16(w[(static_cast<int64_t>(co) * cin + ci) * k3 + t])
                                   : __float2bfloat16(0.f);
  }
}

// y[row][0:c] = x[row][0:c], y[row][c:cpad] = 0 (cpad a multiple of 8): widens a 1-4 channel tensor so that it can feed
// the tensor-core kernels, whose K dimension comes in chunks of 16 channels.
// C = 1 -> 16 (the U-Net stem input): one thread per voxel, one 256-bit store
__global__ void pad_1_to_16_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch, __nv_bfloat16* __restrict__ y,
                                   int64_t rows) {
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t v = xs[i * x_pitch], z = 0u;
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %2, %2, %2, %2, %2, %2};" ::"l"(y + i * 16), "r"(v), "r"(z) : "memory");
  }
}

__global__ void pad_channels_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch, int c,
                                    __nv_bfloat16* __restrict__ y, int cpad, int64_t rows) {
  const int groups = cpad / 8;
  const int64_t total = rows * groups;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / groups;
    const int c0 = static_cast<int>(i % groups) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < c) ? __bfloat162float(x[row * x_pitch + c0 + j]) : 0.f;
    st8(y + row * cpad + c0, pack8(f));
  }
}

// Both packs of MANY weight tensors in one launch (blockIdx.y = descriptor): the optimiser calls it once per step over its
// parameter arena instead of two small launches per convolution.  desc = {src offset (floats), dst offset (bf16 elements),
// cout, cin, k3, dgrad}.
struct PackDesc {
  long long src, dst;
  int cout, cin, k3, dgrad;
};
__global__ void pack_weights_batched_kernel(const float* __restrict__ arena, __nv_bfloat16* __restrict__ packs,
                                            const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  const float* w = arena + d.src;
  __nv_bfloat16* p = packs + d.dst;
  const int64_t total = static_cast<int64_t>(d.k3) * d.cout * d.cin;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int t, co, ci;
    if (!d.dgrad) {
      ci = static_cast<int>(i % d.cin);
      co = static_cast<int>((i / d.cin) % d.cout);
      t = static_cast<int>(i / (static_cast<int64_t>(d.cin) * d.cout));
      p[i] = __float2bfloat16(w[(static_cast<int64_t>(co) * d.cin + ci) * d.k3 + t]);
    } else {
      co = static_cast<int>(i % d.cout);
      ci = static_cast<int>((i / d.cout) % d.cin);
      t = static_cast<int>(i / (static_cast<int64_t>(d.cin) * d.cout));
      p[i] = __float2bfloat16(w[(static_cast<int64_t>(co) * d.cin + ci) * d.k3 + (d.k3 - 1 - t)]);
    }
  }
}

// Tiled versions of the two layout changes for MANY tensors in one launch (blockIdx.y = descriptor).  A block moves a
// 32 (co) x 8 (ci) x k3 brick through shared memory so that both the torch side ([co][ci][k3], k3 fastest) and the packed
// side ([k3][co][ci] or [k3][ci][co]) are accessed in runs of consecutive addresses.
constexpr int kTileCo = 32, kTileCi = 8;

__global__ void __launch_bounds__(256)
    pack_weights_tiled_kernel(const float* __restrict__ arena, __nv_bfloat16* __restrict__ packs,
                              const PackDesc* __restrict__ descs, int ndesc, int total_tiles) {
  extern __shared__ float tile[];   // [kTileCo][kTileCi * k3 + 1]
  // bricks of all tensors form one list dealt round-robin to the blocks (tensor sizes differ by four orders of magnitude)
  for (int gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
    int di = 0, tix = gt;
    for (; di < ndesc; ++di) {
      const int nt = ((descs[di].cin + kTileCi - 1) / kTileCi) * ((descs[di].cout + kTileCo - 1) / kTileCo);
      if (tix < nt) break;
      tix -= nt;
    }
    const PackDesc d = descs[di];
    const float* w = arena + d.src;
    __nv_bfloat16* p = packs + d.dst;
    const int k3 = d.k3, cout = d.cout, cin = d.cin;
    const int tiles_ci = (cin + kTileCi - 1) / kTileCi;
    const int row = kTileCi * k3 + 1;
    const int co0 = (tix / tiles_ci) * kTileCo, ci0 = (tix % tiles_ci) * kTileCi;
    const int nci = min(kTileCi, cin - ci0), nco = min(kTileCo, cout - co0);
    __syncthreads();
    // torch layout: for each co a run of nci * k3 consecutive floats
    for (int i = threadIdx.x; i < nco * nci * k3; i += blockDim.x) {
      const int co = i / (nci * k3), r = i % (nci * k3);
      tile[co * row + r] = w[(static_cast<int64_t>(co0 + co) * cin + ci0) * k3 + r];
    }
    __syncthreads();
    if (!d.dgrad) {   // p[t][co][ci]: runs of nci
      for (int i = threadIdx.x; i < k3 * nco * nci; i += blockDim.x) {
        const int ci = i % nci, co = (i / nci) % nco, t = i / (nci * nco);
        p[(static_cast<int64_t>(t) * cout + co0 + co) * cin + ci0 + ci] = __float2bfloat16(tile[co * row + ci * k3 + t]);
      }
    } else {          // p[t][ci][co] with the taps flipped: runs of nco
      for (int i = threadIdx.x; i < k3 * nci * nco; i += blockDim.x) {
        const int co = i % nco, ci = (i / nco) % nci, t = i / (nco * nci);
        p[(static_cast<int64_t>(t) * cin + ci0 + ci) * cout + co0 + co] =
            __float2bfloat16(tile[co * row + ci * k3 + (k3 - 1 - t)]);
      }
    }
  }
}

// grad[co][ci][t] += dwp[t][ci][co] for many tensors: desc.src = float offset of dwp in `packed`, desc.dst = float offset of
// the gradient in `grads` (cout, cin, k3 as above; dgrad unused).
__global__ void __launch_bounds__(256)
    unpack_wgrads_tiled_kernel(const float* __restrict__ packed, float* __restrict__ grads,
                               const PackDesc* __restrict__ descs, int ndesc, int total_tiles) {
  extern __shared__ float tile[];   // [kTileCo][kTileCi * k3 + 1]
  for (int gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
    int di = 0, tix = gt;
    for (; di < ndesc; ++di) {
      const int nt = ((descs[di].cin + kTileCi - 1) / kTileCi) * ((descs[di].cout + kTileCo - 1) / kTileCo);
      if (tix < nt) break;
      tix -= nt;
    }
    const PackDesc d = descs[di];
    const float* src = packed + d.src;
    float* dst = grads + d.dst;
    const int k3 = d.k3, cout = d.cout, cin = d.cin;
    const int tiles_ci = (cin + kTileCi - 1) / kTileCi;
    const int row = kTileCi * k3 + 1;
    const int co0 = (tix / tiles_ci) * kTileCo, ci0 = (tix % tiles_ci) * kTileCi;
    const int nci = min(kTileCi, cin - ci0), nco = min(kTileCo, cout - co0);
    __syncthreads();
    for (int i = threadIdx.x; i < k3 * nci * nco; i += blockDim.x) {   // packed side: runs of nco
      const int co = i % nco, ci = (i / nco) % nci, t = i / (nco * nci);
      tile[co * row + ci * k3 + t] = src[(static_cast<int64_t>(t) * cin + ci0 + ci) * cout + co0 + co];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nco * nci * k3; i += blockDim.x) {   // torch side: runs of nci * k3
      const int co = i / (nci * k3), r = i % (nci * k3);
      dst[(static_cast<int64_t>(co0 + co) * cin + ci0) * k3 + r] += tile[co * row + r];
    }
  }
}

// dw_packed [t][ci_local][co] fp32 -> grad [co][cin][k3] (+=)
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ grad, int cout, int cin,
                                         int k3, int cin_off, int cin_cnt, int accumulate) {
  const int64_t total = static_cast<int64_t>(k3) * cout * cin_cnt;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // iterate in destination order for coalesced writes
    const int t = static_cast<int>(i % k3);
    const int ci = static_cast<int>((i / k3) % cin_cnt);
    const int co = static_cast<int>(i / (static_cast<int64_t>(k3) * cin_cnt));
    float* srcp = const_cast<float*>(dwp) + (static_cast<int64_t>(t) * cin_cnt + ci) * cout + co;
    const float v = *srcp;
    if (accumulate & 2) *srcp = 0.f;   // read-and-clear: the packed accumulator is ready for its next use
    float* dst = grad + (static_cast<int64_t>(co) * cin + cin_off + ci) * k3 + t;
    *dst = (accumulate & 1) ? *dst + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ fprop
// thread = (output voxel, 8 output channels).  wp: [t][co][ci].
template <bool VEC>
__global__ void conv_direct_fprop_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                                         const __nv_bfloat16* __restrict__ wp, const float* __restrict__ bias,
                                         __nv_bfloat16* __restrict__ y, int64_t y_pitch, float* __restrict__ stats) {
  const int cog = (g.cout + 7) / 8;
  const int64_t total = static_cast<int64_t>(g.n) * g.od * g.oh * g.ow * cog;
  const int k = g.k;
  float ssum[8], ssq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ssum[j] = ssq[j] = 0.f;
  int my_cg = -1;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % cog);
    my_cg = cg;  // grid stride is a multiple of cog (host guarantees), so cg is loop-invariant per thread
    int64_t v = i / cog;
    const int64_t orow = v;
    const int xo = static_cast<int>(v % g.ow);
    v /= g.ow;
    const int yo = static_cast<int>(v % g.oh);
    v /= g.oh;
    const int zo = static_cast<int>(v % g.od);
    const int nn = static_cast<int>(v / g.od);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int co0 = cg * 8;
    for (int a = 0; a < k; ++a) {
      const int zi = zo * g.stride - g.pad + a * g.dil;
      if (zi < 0 || zi >= g.d) continue;
      for (int b = 0; b < k; ++b) {
        const int yi = yo * g.stride - g.pad + b * g.dil;
        if (yi < 0 || yi >= g.h) continue;
        for (int e = 0; e < k; ++e) {
          const int xi = xo * g.stride - g.pad + e * g.dil;
          if (xi < 0 || xi >= g.w) continue;
          const int t = (a * k + b) * k + e;
          const __nv_bfloat16* xr = x + (((static_cast<int64_t>(nn) * g.d + zi) * g.h + yi) * g.w + xi) * x_pitch;
          const __nv_bfloat16* wt = wp + static_cast<int64_t>(t) * g.cout * g.cin;
          if constexpr (VEC) {
            for (int c0 = 0; c0 < g.cin; c0 += 8) {
              float xf[8];
              unpack8(ld8(xr + c0), xf);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (co0 + j < g.cout) {
                  float wf[8];
                  unpack8(ld8(wt + static_cast<int64_t>(co0 + j) * g.cin + c0), wf);
#pragma unroll
                  for (int q = 0; q < 8; ++q) acc[j] += xf[q] * wf[q];
                }
              }
            }
          } else {
            for (int c = 0; c < g.cin; ++c) {
              const float xf = __bfloat162float(xr[c]);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (co0 + j < g.cout) acc[j] += xf * __bfloat162float(wt[static_cast<int64_t>(co0 + j) * g.cin + c]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (co0 + j < g.cout) {
        if (bias) acc[j] += bias[co0 + j];
        ssum[j] += acc[j];
        ssq[j] += acc[j] * acc[j];
        y[orow * y_pitch + co0 + j] = __float2bfloat16(acc[j]);
      }
    }
  }
  if (stats) {
    // block-level reduction per channel through shared memory, then one atomic per channel per block
    extern __shared__ float sred[];  // [2][cout]
    for (int i = threadIdx.x; i < 2 * g.cout; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
    if (my_cg >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (my_cg * 8 + j < g.cout) {
          atomicAdd(&sred[my_cg * 8 + j], ssum[j]);
          atomicAdd(&sred[g.cout + my_cg * 8 + j], ssq[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * g.cout; i += blockDim.x) atomicAdd(&stats[i], sred[i]);
  }
}

// ------------------------------------------------------------------------------------------------ dgrad (gather)
// thread = (input voxel, 8 input channels).  wd: flipped pack [k3-1-t][ci][co].
template <bool VEC>
__global__ void conv_direct_dgrad_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ dy, int64_t dy_pitch,
                                         const __nv_bfloat16* __restrict__ wd, const float* __restrict__ bias,
                                         __nv_bfloat16* __restrict__ dx, int64_t dx_pitch) {
  const int cig = (g.cin + 7) / 8;
  const int64_t total = static_cast<int64_t>(g.n) * g.d * g.h * g.w * cig;
  const int k = g.k, k3 = k * k * k;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % cig);
    int64_t v = i / cig;
    const int64_t irow = v;
    const int xi = static_cast<int>(v % g.w);
    v /= g.w;
    const int yi = static_cast<int>(v % g.h);
    v /= g.h;
    const int zi = static_cast<int>(v % g.d);
    const int nn = static_cast<int>(v / g.d);
    const int ci0 = cg * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < k; ++a) {
      const int zn = zi + g.pad - a * g.dil;
      if (zn < 0 || zn % g.stride) continue;
      const int zo = zn / g.stride;
      if (zo >= g.od) continue;
      for (int b = 0; b < k; ++b) {
        const int yn = yi + g.pad - b * g.dil;
        if (yn < 0 || yn % g.stride) continue;
        const int yo = yn / g.stride;
        if (yo >= g.oh) continue;
        for (int e = 0; e < k; ++e) {
          const int xn = xi + g.pad - e * g.dil;
          if (xn < 0 || xn % g.stride) continue;
          const int xo = xn / g.stride;
          if (xo >= g.ow) continue;
          const int t = (a * k + b) * k + e;
          const __nv_bfloat16* dr = dy + (((static_cast<int64_t>(nn) * g.od + zo) * g.oh + yo) * g.ow + xo) * dy_pitch;
          const __nv_bfloat16* wt = wd + static_cast<int64_t>(k3 - 1 - t) * g.cin * g.cout;
          if constexpr (VEC) {
            for (int c0 = 0; c0 < g.cout; c0 += 8) {
              float df[8];
              unpack8(ld8(dr + c0), df);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (ci0 + j < g.cin) {
                  float wf[8];
                  unpack8(ld8(wt + static_cast<int64_t>(ci0 + j) * g.cout + c0), wf);
#pragma unroll
                  for (int q = 0; q < 8; ++q) acc[j] += df[q] * wf[q];
                }
              }
            }
          } else {
            for (int c = 0; c < g.cout; ++c) {
              const float df = __bfloat162float(dr[c]);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (ci0 + j < g.cin) acc[j] += df * __bfloat162float(wt[static_cast<int64_t>(ci0 + j) * g.cout + c]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (ci0 + j < g.cin) dx[irow * dx_pitch + ci0 + j] = __float2bfloat16(acc[j] + (bias ? bias[ci0 + j] : 0.f));
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// block = one tap, one 32x32 (ci, co) tile, one chunk of output voxels.  256 threads, 4 accumulators each.
__global__ void __launch_bounds__(256)
    conv_direct_wgrad_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                             const __nv_bfloat16* __restrict__ dy, int64_t dy_pitch, float* __restrict__ dwp,
                             int64_t vox_per_block) {
  __shared__ float xs[32][33];   // [voxel][ci]
  __shared__ float ds[32][33];   // [voxel][co]
  const int co_tiles = (g.cout + 31) / 32;
  const int ci0 = (blockIdx.x / co_tiles) * 32, co0 = (blockIdx.x % co_tiles) * 32;
  const int t = blockIdx.y, k = g.k;
  const int a = t / (k * k), b = (t / k) % k, e = t % k;
  const int64_t total = static_cast<int64_t>(g.n) * g.od * g.oh * g.ow;
  const int64_t v_begin = static_cast<int64_t>(blockIdx.z) * vox_per_block;
  const int64_t v_end = v_begin + vox_per_block < total ? v_begin + vox_per_block : total;
  const int tid = threadIdx.x;
  const int ci_l = tid >> 3;          // 0..31
  const int co_l = (tid & 7) * 4;     // 0,4,..28
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t vb = v_begin; vb < v_end; vb += 32) {
    // stage 32 voxels: thread (vv = tid/8, 4 channels at (tid%8)*4)
    {
      const int vv = tid >> 3, c4 = (tid & 7) * 4;
      const int64_t v = vb + vv;
      bool ok = v < v_end;
      int64_t irow = 0;
      if (ok) {
        int64_t r = v;
        const int xo = static_cast<int>(r % g.ow);
        r /= g.ow;
        const int yo = static_cast<int>(r % g.oh);
        r /= g.oh;
        const int zo = static_cast<int>(r % g.od);
        const int nn = static_cast<int>(r / g.od);
        const int zi = zo * g.stride - g.pad + a * g.dil, yi = yo * g.stride - g.pad + b * g.dil,
                  xi = xo * g.stride - g.pad + e * g.dil;
        const bool inb = zi >= 0 && zi < g.d && yi >= 0 && yi < g.h && xi >= 0 && xi < g.w;
        irow = ((static_cast<int64_t>(nn) * g.d + zi) * g.h + yi) * g.w + xi;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ci = ci0 + c4 + q, co = co0 + c4 + q;
          xs[vv][c4 + q] = (inb && ci < g.cin) ? __bfloat162float(x[irow * x_pitch + ci]) : 0.f;
          ds[vv][c4 + q] = (co < g.cout) ? __bfloat162float(dy[v * dy_pitch + co]) : 0.f;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) xs[vv][c4 + q] = ds[vv][c4 + q] = 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int vv = 0; vv < 32; ++vv) {
      const float xv = xs[vv][ci_l];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += xv * ds[vv][co_l + q];
    }
    __syncthreads();
  }
  const int ci = ci0 + ci_l;
  if (ci < g.cin) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = co0 + co_l + q;
      if (co < g.cout && acc[q] != 0.f) atomicAdd(&dwp[(static_cast<int64_t>(t) * g.cin + ci) * g.cout + co], acc[q]);
    }
  }
}


// ------------------------------------------------------------------------------------------------ stem (C_in = 1)
// First layer of every model on the path: 1 input channel, k^3 taps, COUT <= 32 channels ('same' padding, stride 1).
// 26 FLOP/B: HBM-bound on the bf16 output.  fprop: thread = voxel, all COUT accumulators in registers, weights
// broadcast from shared memory.  wgrad: lane = output channel, 27 accumulators per lane, the 27 input taps of a voxel
// are fetched by 27 lanes and broadcast with shuffles.
template <int COUT>
__global__ void __launch_bounds__(256)
    stem_fprop_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                      const __nv_bfloat16* __restrict__ wp, const float* __restrict__ bias,
                      __nv_bfloat16* __restrict__ y, int64_t y_pitch, float* __restrict__ stats) {
  __shared__ __align__(16) float sw[125 * COUT];
  __shared__ float sred[2 * COUT];
  const int k = g.k, k3 = k * k * k;
  for (int i = threadIdx.x; i < k3 * COUT; i += blockDim.x) sw[i] = __bfloat162float(wp[i]);  // [t][co][ci=1]
  for (int i = threadIdx.x; i < 2 * COUT; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float s1[COUT], s2[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) s1[j] = s2[j] = 0.f;
  const int64_t total = static_cast<int64_t>(g.n) * g.d * g.h * g.w;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = v;
    const int xo = static_cast<int>(r % g.w);
    r /= g.w;
    const int yo = static_cast<int>(r % g.h);
    r /= g.h;
    const int zo = static_cast<int>(r % g.d);
    const int nn = static_cast<int>(r / g.d);
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = bias ? bias[j] : 0.f;
    for (int a = 0; a < k; ++a) {
      const int zi = zo - g.pad + a;
      if (zi < 0 || zi >= g.d) continue;
      for (int b = 0; b < k; ++b) {
        const int yi = yo - g.pad + b;
        if (yi < 0 || yi >= g.h) continue;
        const __nv_bfloat16* row = x + ((static_cast<int64_t>(nn) * g.d + zi) * g.h + yi) * g.w * x_pitch;
        for (int e = 0; e < k; ++e) {
          const int xi = xo - g.pad + e;
          if (xi < 0 || xi >= g.w) continue;
          const float xv = __bfloat162float(row[xi * x_pitch]);
          const float4* w4 = reinterpret_cast<const float4*>(sw + ((a * k + b) * k + e) * COUT);
#pragma unroll
          for (int j = 0; j < COUT / 4; ++j) {
            const float4 t = w4[j];
            acc[4 * j + 0] += xv * t.x;
            acc[4 * j + 1] += xv * t.y;
            acc[4 * j + 2] += xv * t.z;
            acc[4 * j + 3] += xv * t.w;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < COUT; j += 8) {
      float t8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t8[i] = acc[j + i];
        s1[j + i] += acc[j + i];
        s2[j + i] += acc[j + i] * acc[j + i];
      }
      st8(y + v * y_pitch + j, pack8(t8));
    }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
      const float a1 = warp_sum(s1[j]), a2 = warp_sum(s2[j]);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sred[j], a1);
        atomicAdd(&sred[COUT + j], a2);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * COUT; i += blockDim.x) atomicAdd(&stats[i], sred[i]);
  }
}

// dwp[t][0][co] += sum_v x[v + off_t] * dy[v][co];  k = 3, COUT = 32 (lane = co).
__global__ void __launch_bounds__(256)
    stem_wgrad_k3_c32_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                             const __nv_bfloat16* __restrict__ dy, int64_t dy_pitch, float* __restrict__ dwp) {
  __shared__ float sacc[27 * 32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t warp_id = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * warps_per_block;
  // lane t < 27 fetches tap t = (a, b, e)
  const int ta = lane / 9 - 1, tb = (lane / 3) % 3 - 1, te = lane % 3 - 1;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  const int64_t rows = static_cast<int64_t>(g.n) * g.d * g.h;   // one (n, z, y) line of w voxels per iteration
  for (int64_t r = warp_id; r < rows; r += nwarps) {
    const int yo = static_cast<int>(r % g.h);
    const int zo = static_cast<int>((r / g.h) % g.d);
    const int nn = static_cast<int>(r / (static_cast<int64_t>(g.h) * g.d));
    const int zi = zo + ta, yi = yo + tb;
    const bool row_ok = lane < 27 && zi >= 0 && zi < g.d && yi >= 0 && yi < g.h;
    const __nv_bfloat16* xrow = x + ((static_cast<int64_t>(nn) * g.d + zi) * g.h + yi) * g.w * x_pitch;
    const __nv_bfloat16* drow = dy + r * g.w * dy_pitch + lane;
    for (int xo = 0; xo < g.w; ++xo) {
      const int xi = xo + te;
      const float xv = (row_ok && xi >= 0 && xi < g.w) ? __bfloat162float(xrow[xi * x_pitch]) : 0.f;
      const float dv = __bfloat162float(drow[xo * dy_pitch]);
#pragma unroll
      for (int t = 0; t < 27; ++t) acc[t] += __shfl_sync(0xffffffffu, xv, t) * dv;
    }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t) atomicAdd(&sacc[t * 32 + lane], acc[t]);
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) atomicAdd(&dwp[i], sacc[i]);
}

// Tiled C_in = 1 weight gradient: a block stages a 4 x 4 x 32 voxel tile of dy (as fp32) and the matching x halo in shared
// memory; each thread owns 2 taps x 8 output channels = 16 accumulators (14 tap pairs x COUT/8 channel groups active
// threads per 64-thread stream, 4 streams = 4 z-slices of the tile), so a voxel costs it 2 + 2 shared loads for 16 FMAs
// instead of the 27 shuffles + 27 FMAs per lane of the warp-broadcast version above.
template <int COUT>
__global__ void __launch_bounds__(256)
    stem_wgrad_k3_tiled_kernel(ConvGeom g, const __nv_bfloat16* __restrict__ x, int64_t x_pitch,
                               const __nv_bfloat16* __restrict__ dy, int64_t dy_pitch, float* __restrict__ dwp) {
  constexpr int TZ = 4, TY = 4, TX = 32, CG = COUT / 8;
  extern __shared__ float smem_f[];
  float* sdy = smem_f;                                   // [TZ*TY*TX][COUT]
  float* sx = sdy + TZ * TY * TX * COUT;                 // [TZ+2][TY+2][TX+2]
  float* sacc = sx + (TZ + 2) * (TY + 2) * (TX + 2);     // [27][COUT]
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sacc[i] = 0.f;
  const int stream = threadIdx.x >> 6, lt = threadIdx.x & 63;
  const bool active = lt < 14 * CG;
  const int tp = lt / CG, cg = lt % CG;
  const int t0 = 2 * tp, t1 = 2 * tp + 1;
  const bool has1 = t1 < 27;
  const int off0 = ((t0 / 9) * (TY + 2) + (t0 / 3) % 3) * (TX + 2) + t0 % 3;
  const int off1 = has1 ? ((t1 / 9) * (TY + 2) + (t1 / 3) % 3) * (TX + 2) + t1 % 3 : off0;
  float acc0[8], acc1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.f;
  const int tz = (g.d + TZ - 1) / TZ, ty = (g.h + TY - 1) / TY, tx = (g.w + TX - 1) / TX;
  const long long tiles = static_cast<long long>(g.n) * tz * ty * tx;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    long long r = tile;
    const int x0 = static_cast<int>(r % tx) * TX;
    r /= tx;
    const int y0 = static_cast<int>(r % ty) * TY;
    r /= ty;
    const int z0 = static_cast<int>(r % tz) * TZ;
    const int nn = static_cast<int>(r / tz);
    __syncthreads();   // previous tile fully consumed
    // dy tile: TZ*TY rows of TX voxels x COUT channels, 8 channels (one 128-bit load) per thread step
    for (int i = threadIdx.x; i < TZ * TY * TX * CG; i += blockDim.x) {
      const int c8 = i % CG, v = i / CG;
      const int xx = v % TX, yy = (v / TX) % TY, zz = v / (TX * TY);
      const int gz = z0 + zz, gy = y0 + yy, gx = x0 + xx;
      float f[8];
      if (gz < g.d && gy < g.h && gx < g.w) {
        const long long vox = ((static_cast<long long>(nn) * g.d + gz) * g.h + gy) * g.w + gx;
        unpack8(ld8(dy + vox * dy_pitch + c8 * 8), f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(sdy + v * COUT + c8 * 8);
      dst[0] = make_float4(f[0], f[1], f[2], f[3]);
      dst[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int i = threadIdx.x; i < (TZ + 2) * (TY + 2) * (TX + 2); i += blockDim.x) {
      const int xx = i % (TX + 2), yy = (i / (TX + 2)) % (TY + 2), zz = i / ((TX + 2) * (TY + 2));
      const int gz = z0 - 1 + zz, gy = y0 - 1 + yy, gx = x0 - 1 + xx;
      float v = 0.f;
      if (gz >= 0 && gz < g.d && gy >= 0 && gy < g.h && gx >= 0 && gx < g.w)
        v = __bfloat162float(x[(((static_cast<long long>(nn) * g.d + gz) * g.h + gy) * g.w + gx) * x_pitch]);
      sx[i] = v;
    }
    __syncthreads();
    if (active) {
      const float* sxs = sx + stream * (TY + 2) * (TX + 2);
      const float* sds = sdy + stream * TY * TX * COUT + cg * 8;
#pragma unroll 1
      for (int yy = 0; yy < TY; ++yy) {
#pragma unroll 4
        for (int xx = 0; xx < TX; ++xx) {
          const int hb = yy * (TX + 2) + xx;
          const float xv0 = sxs[hb + off0];
          const float xv1 = has1 ? sxs[hb + off1] : 0.f;
          const float4 da = *reinterpret_cast<const float4*>(sds + (yy * TX + xx) * COUT);
          const float4 db = *reinterpret_cast<const float4*>(sds + (yy * TX + xx) * COUT + 4);
          const float d8[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc0[j] = fmaf(xv0, d8[j], acc0[j]);
            acc1[j] = fmaf(xv1, d8[j], acc1[j]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sacc[t0 * COUT + cg * 8 + j], acc0[j]);
      if (has1) atomicAdd(&sacc[t1 * COUT + cg * 8 + j], acc1[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) atomicAdd(&dwp[i], sacc[i]);
}

static bool is_stem(const ConvGeom& g) {
  return g.cin == 1 && g.stride == 1 && g.dil == 1 && g.k <= 5 && (g.k & 1) && g.pad == (g.k - 1) / 2 &&
         (g.cout == 32 || g.cout == 16);
}

// ------------------------------------------------------------------------------------------------ host side
int conv_direct_fprop(const ConvGeom& g, const void* x, int64_t x_pitch, const void* wp, const float* bias, void* y,
                      int64_t y_pitch, float* stats, cudaStream_t st) {
  if (is_stem(g) && y_pitch % 8 == 0) {
    const int64_t total = static_cast<int64_t>(g.n) * g.d * g.h * g.w;
    const int blocks = grid_for(total, 256, kNumSMs * 4);
    if (g.cout == 32)
      stem_fprop_kernel<32><<<blocks, 256, 0, st>>>(g, static_cast<const __nv_bfloat16*>(x), x_pitch,
                                                     static_cast<const __nv_bfloat16*>(wp), bias,
                                                     static_cast<__nv_bfloat16*>(y), y_pitch, stats);
    else
      stem_fprop_kernel<16><<<blocks, 256, 0, st>>>(g, static_cast<const __nv_bfloat16*>(x), x_pitch,
                                                     static_cast<const __nv_bfloat16*>(wp), bias,
                                                     static_cast<__nv_bfloat16*>(y), y_pitch, stats);
    B200_CHECK_LAUNCH("stem_fprop");
    return 0;
  }
  const int cog = (g.cout + 7) / 8;
  const int64_t total = static_cast<int64_t>(g.n) * g.od * g.oh * g.ow * cog;
  // blockDim * gridDim must be a multiple of cog so each thread keeps one channel group (see kernel)
  int threads = 256;
  int blocks = grid_for(total, threads, kNumSMs * 8);
  if ((static_cast<int64_t>(threads) * blocks) % cog != 0) {
    threads = cog * (256 / cog > 0 ? 256 / cog : 1);
    if (threads > 1024 || threads < 1) {
      set_error("conv_direct_fprop: cout=%d not supported by the direct path", g.cout);
      return B200SEG_ERR_INVALID;
    }
    blocks = grid_for(total, threads, kNumSMs * 8);
  }
  const bool vec = (g.cin % 8 == 0) && (x_pitch % 8 == 0);
  const size_t smem = stats ? 2 * g.cout * sizeof(float) : 0;
  if (vec)
    conv_direct_fprop_kernel<true><<<blocks, threads, smem, st>>>(
        g, static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<const __nv_bfloat16*>(wp), bias,
        static_cast<__nv_bfloat16*>(y), y_pitch, stats);
  else
    conv_direct_fprop_kernel<false><<<blocks, threads, smem, st>>>(
        g, static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<const __nv_bfloat16*>(wp), bias,
        static_cast<__nv_bfloat16*>(y), y_pitch, stats);
  B200_CHECK_LAUNCH("conv_direct_fprop");
  return 0;
}

int conv_direct_dgrad(const ConvGeom& g, const void* dy, int64_t dy_pitch, const void* wd, const float* bias, void* dx,
                      int64_t dx_pitch, cudaStream_t st) {
  const int cig = (g.cin + 7) / 8;
  const int64_t total = static_cast<int64_t>(g.n) * g.d * g.h * g.w * cig;
  const bool vec = (g.cout % 8 == 0) && (dy_pitch % 8 == 0);
  if (vec)
    conv_direct_dgrad_kernel<true><<<grid_for(total, 256, kNumSMs * 8), 256, 0, st>>>(
        g, static_cast<const __nv_bfloat16*>(dy), dy_pitch, static_cast<const __nv_bfloat16*>(wd), bias,
        static_cast<__nv_bfloat16*>(dx), dx_pitch);
  else
    conv_direct_dgrad_kernel<false><<<grid_for(total, 256, kNumSMs * 8), 256, 0, st>>>(
        g, static_cast<const __nv_bfloat16*>(dy), dy_pitch, static_cast<const __nv_bfloat16*>(wd), bias,
        static_cast<__nv_bfloat16*>(dx), dx_pitch);
  B200_CHECK_LAUNCH("conv_direct_dgrad");
  return 0;
}

// ---- C_in = 1 stem on the tensor cores: im2col in the CHANNEL dimension ----------------------------------------------
// xcol[v][j] = x[v + tap_j - 1] for the 27 taps j = (a*3 + b)*3 + e of a 3x3x3 / pad 1 kernel (zero outside the volume),
// channels 27..31 zero: the stem's weight gradient dW[tap][0][co] = sum_v x[v + tap - 1] dy[v][co] then IS the weight
// gradient of a 1x1x1 convolution with 32 input channels, which the tcgen05 wgrad kernel runs at HBM speed.
// One thread per voxel: 27 two-byte loads (neighbours hit L1), four 128-bit stores; a warp writes 2 KB contiguously.
__global__ void __launch_bounds__(256)
    stem_im2col_k3_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_pitch, __nv_bfloat16* __restrict__ xcol, int n,
                          int d, int h, int w) {
  // rows of the (n, d, h) index space are dealt to warps; lanes walk along w, so every tap is one coalesced load per warp
  const int64_t lines = static_cast<int64_t>(n) * d * h;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (int64_t line = warp0; line < lines; line += nwarps) {
    const int yh = static_cast<int>(line % h);
    const int zd = static_cast<int>((line / h) % d);
    const int64_t v0 = line * w;
    for (int xw = lane; xw < w; xw += 32) {
      const int64_t v = v0 + xw;
      uint32_t pk[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) pk[q] = 0u;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            const int t = (a * 3 + b) * 3 + e;
            const int zz = zd + a - 1, yy = yh + b - 1, xx = xw + e - 1;
            const bool ok = zz >= 0 && zz < d && yy >= 0 && yy < h && xx >= 0 && xx < w;
            const int64_t off = v + (static_cast<int64_t>(a - 1) * h + (b - 1)) * w + (e - 1);
            const uint32_t bits = ok ? static_cast<uint32_t>(xs[off * x_pitch]) : 0u;
            pk[t >> 1] |= (t & 1) ? (bits << 16) : bits;
          }
      __nv_bfloat16* dst = xcol + v * 32;     // 64-byte rows of a 256-byte aligned workspace: two 256-bit stores
#pragma unroll
      for (int q = 0; q < 2; ++q)
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * q), "r"(pk[8 * q]),
                     "r"(pk[8 * q + 1]), "r"(pk[8 * q + 2]), "r"(pk[8 * q + 3]), "r"(pk[8 * q + 4]), "r"(pk[8 * q + 5]),
                     "r"(pk[8 * q + 6]), "r"(pk[8 * q + 7])
                     : "memory");
    }
  }
}

__global__ void add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] += src[i];
}

int stem_im2col_k3(const void* x, int64_t x_pitch, void* xcol, int n, int d, int h, int w, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(n) * d * h * w;
  stem_im2col_k3_kernel<<<grid_for(total, 256, kNumSMs * 16), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(xcol), n, d, h, w);
  B200_CHECK_LAUNCH("stem_im2col");
  return 0;
}

int add_f32(float* dst, const float* src, int n, cudaStream_t st) {
  add_f32_kernel<<<grid_for(n, 256, 64), 256, 0, st>>>(dst, src, n);
  B200_CHECK_LAUNCH("add_f32");
  return 0;
}

int conv_direct_wgrad(const ConvGeom& g, const void* x, int64_t x_pitch, const void* dy, int64_t dy_pitch, float* dwp,
                      cudaStream_t st) {
  if (is_stem(g) && g.k == 3 && g.pad == 1 && (g.cout == 32 || g.cout == 16) && dy_pitch % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && static_cast<long long>(g.n) * g.d * g.h * g.w >= (1 << 15)) {
    const size_t smem = (static_cast<size_t>(4 * 4 * 32) * g.cout + 6 * 6 * 34 + 27 * g.cout) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(stem_wgrad_k3_tiled_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(stem_wgrad_k3_tiled_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      attr_set = true;
    }
    if (g.cout == 32)
      stem_wgrad_k3_tiled_kernel<32><<<kNumSMs * 3, 256, smem, st>>>(g, static_cast<const __nv_bfloat16*>(x), x_pitch,
                                                                      static_cast<const __nv_bfloat16*>(dy), dy_pitch, dwp);
    else
      stem_wgrad_k3_tiled_kernel<16><<<kNumSMs * 3, 256, smem, st>>>(g, static_cast<const __nv_bfloat16*>(x), x_pitch,
                                                                      static_cast<const __nv_bfloat16*>(dy), dy_pitch, dwp);
    B200_CHECK_LAUNCH("stem_wgrad_tiled");
    return 0;
  }
  if (is_stem(g) && g.k == 3 && g.cout == 32) {
    stem_wgrad_k3_c32_kernel<<<kNumSMs * 4, 256, 0, st>>>(g, static_cast<const __nv_bfloat16*>(x), x_pitch,
                                                           static_cast<const __nv_bfloat16*>(dy), dy_pitch, dwp);
    B200_CHECK_LAUNCH("stem_wgrad");
    return 0;
  }
  const int k3 = g.k * g.k * g.k;
  const int tiles = ((g.cin + 31) / 32) * ((g.cout + 31) / 32);
  const int64_t total = static_cast<int64_t>(g.n) * g.od * g.oh * g.ow;
  // aim for ~8 blocks per SM overall, at least 256 voxels per block
  int64_t want_z = (static_cast<int64_t>(kNumSMs) * 8 + static_cast<int64_t>(tiles) * k3 - 1) / (static_cast<int64_t>(tiles) * k3);
  if (want_z < 1) want_z = 1;
  int64_t vpb = (total + want_z - 1) / want_z;
  if (vpb < 256) vpb = 256;
  vpb = (vpb + 31) / 32 * 32;
  const int64_t gz = (total + vpb - 1) / vpb;
  if (tiles > 2147483647 || gz > 65535 || k3 > 65535) {
    set_error("conv_direct_wgrad: grid too large");
    return B200SEG_ERR_INVALID;
  }
  dim3 grid(tiles, k3, static_cast<unsigned>(gz));
  conv_direct_wgrad_kernel<<<grid, 256, 0, st>>>(g, static_cast<const __nv_bfloat16*>(x), x_pitch,
                                                 static_cast<const __nv_bfloat16*>(dy), dy_pitch, dwp, vpb);
  B200_CHECK_LAUNCH("conv_direct_wgrad");
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200seg_pad_channels(const void* x, int64_t x_pitch, int c, void* y, int cpad, int64_t rows, void* stream) {
  B200_CHECK_ARG(x && y && c > 0 && cpad >= c && cpad % 8 == 0 && rows > 0 && x_pitch >= c, "pad_channels: bad arguments");
  if (c == 1 && cpad == 16 && (reinterpret_cast<uintptr_t>(y) & 31) == 0)
    b200::pad_1_to_16_kernel<<<grid_for(rows, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, static_cast<__nv_bfloat16*>(y), rows);
  else
    b200::pad_channels_kernel<<<grid_for(rows * (cpad / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), x_pitch, c, static_cast<__nv_bfloat16*>(y), cpad, rows);
  B200_CHECK_LAUNCH("pad_channels");
  return 0;
}

int b200seg_pack_weights_batched(const float* arena, void* packs, const void* descs, int ndesc, int max_k3,
                                 int total_tiles, void* stream) {
  B200_CHECK_ARG(arena && packs && descs && ndesc > 0 && ndesc <= 65535 && max_k3 >= 1 && max_k3 <= 125,
                 "pack_weights_batched: bad arguments");
  static_assert(sizeof(b200::PackDesc) == 32, "descriptor layout is part of the ABI");
  // brick of 32 x 8 x k3 floats in shared memory, sized for the largest kernel in the table (k3 <= 125 = 5x5x5)
  const size_t smem = static_cast<size_t>(b200::kTileCo) * (b200::kTileCi * max_k3 + 1) * sizeof(float);
  const size_t smem_max = static_cast<size_t>(b200::kTileCo) * (b200::kTileCi * 125 + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(b200::pack_weights_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
    attr = true;
  }
  b200::pack_weights_tiled_kernel<<<std::min(total_tiles, b200::kNumSMs * 6), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      arena, static_cast<__nv_bfloat16*>(packs), static_cast<const b200::PackDesc*>(descs), ndesc, total_tiles);
  B200_CHECK_LAUNCH("pack_weights_batched");
  return 0;
}

int b200seg_unpack_wgrads_batched(const float* packed, float* grads, const void* descs, int ndesc, int max_k3,
                                  int total_tiles, void* stream) {
  B200_CHECK_ARG(packed && grads && descs && ndesc > 0 && ndesc <= 65535 && max_k3 >= 1 && max_k3 <= 125,
                 "unpack_wgrads_batched: bad arguments");
  const size_t smem = static_cast<size_t>(b200::kTileCo) * (b200::kTileCi * max_k3 + 1) * sizeof(float);
  const size_t smem_max = static_cast<size_t>(b200::kTileCo) * (b200::kTileCi * 125 + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(b200::unpack_wgrads_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
    attr = true;
  }
  b200::unpack_wgrads_tiled_kernel<<<std::min(total_tiles, b200::kNumSMs * 6), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      packed, grads, static_cast<const b200::PackDesc*>(descs), ndesc, total_tiles);
  B200_CHECK_LAUNCH("unpack_wgrads_batched");
  return 0;
}

int b200seg_pack_conv_weight(const float* w, void* packed, int cout, int cin, int k, int cin_off, int cin_cnt,
                             int dgrad, void* stream) {
  B200_CHECK_ARG(w && packed && cout > 0 && cin > 0 && k > 0 && cin_off >= 0 && cin_cnt > 0 && cin_off < cin,
                 "pack_conv_weight: bad arguments");   // cin_off + cin_cnt > cin: the excess channels are packed as zeros
  const int k3 = k * k * k;
  const int64_t total = static_cast<int64_t>(k3) * cout * cin_cnt;
  pack_conv_weight_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(packed), cout, cin, k3, cin_off, cin_cnt, dgrad);
  B200_CHECK_LAUNCH("pack_conv_weight");
  return 0;
}

int b200seg_pack_conv_weight_padded(const float* w, void* packed, int cout, int cin, int k, int cout_pad, int cin_pad,
                                    int dgrad, void* stream) {
  B200_CHECK_ARG(w && packed && cout > 0 && cin > 0 && k > 0 && cout_pad >= cout && cin_pad >= cin,
                 "pack_conv_weight_padded: bad arguments");
  const int k3 = k * k * k;
  const int64_t total = static_cast<int64_t>(k3) * cout_pad * cin_pad;
  pack_conv_weight_padded_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(packed), cout, cin, k3, cout_pad, cin_pad, dgrad);
  B200_CHECK_LAUNCH("pack_conv_weight_padded");
  return 0;
}

int b200seg_unpack_conv_wgrad(const float* dw_packed, float* grad_w, int cout, int cin, int k, int cin_off,
                              int cin_cnt, int accumulate, void* stream) {
  B200_CHECK_ARG(dw_packed && grad_w && cout > 0 && cin > 0 && k > 0 && cin_off >= 0 && cin_cnt > 0 &&
                     cin_off + cin_cnt <= cin,
                 "unpack_conv_wgrad: bad arguments");
  const int k3 = k * k * k;
  const int64_t total = static_cast<int64_t>(k3) * cout * cin_cnt;
  unpack_conv_wgrad_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dw_packed, grad_w, cout, cin, k3, cin_off, cin_cnt, accumulate);
  B200_CHECK_LAUNCH("unpack_conv_wgrad");
  return 0;
}

}  // extern "C"
