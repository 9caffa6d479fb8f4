// Fused optimiser step for the flat parameter arena (torch.optim.Adam, train.py:109,214) that also does the two layout
// changes the convolution kernels need, so that a training step has ONE pass over the 22.6 M U-Net parameters instead of
// three (gradient transpose -> Adam -> bf16 weight packs):
//
//   * the weight-gradient kernels leave dW in the packed layout [tap][C_in][C_out] (fp32, `dw`); the parameters, moments
//     and torch-visible gradients are [C_out][C_in][tap].  A thread block owns a brick of 32 (C_out) x TCI (C_in) x k^3
//     elements: it gathers the brick of `dw` through shared memory (global reads in runs of 32 C_out floats), runs Adam over
//     the brick in torch order (p, g, m, v accessed in runs of TCI*k^3 consecutive floats), leaves the new parameters in
//     shared memory and writes both bf16 packs from there: fprop [tap][C_out][C_in] and dgrad [k^3-1-tap][C_in][C_out].
//   * everything that is not a cubic conv weight (biases, norm affine parameters, PReLU slopes) is a "plain range"
//     record handled by the same launch in chunks of 4096 elements.
//
// Step counter and hyper-parameters live in device memory (CUDA-graph safe), exactly as in adam_dev_kernel.
#include <algorithm>

#include "common.cuh"

namespace b200 {

struct AdamDesc {   // 40 bytes, part of the ABI (include/b200seg.h)
  long long src;    // element offset of the tensor in the parameter / gradient / moment arenas (and in dw)
  long long dst;    // bf16 element offset of the fprop pack in `packs`; the dgrad pack follows at dst + cout*cin*k3
  int cout, cin, k3;
  int kind;         // 0 = cubic conv weight [cout][cin][k3], 1 = plain range of `cout` elements
  int tile0;        // index of this record's first brick / chunk in the launch-wide list
  int pad;
};
static_assert(sizeof(AdamDesc) == 40, "descriptor layout is part of the ABI");

constexpr int kBrickCo = 32;
constexpr int kPlainChunk = 4096;

struct AdamCoef {
  float lr_bc1, b1, b2, eps, wd, gscale, bc2_sqrt;
};

__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
               : "l"(p));
}
__device__ __forceinline__ void st8f(float* p, const float (&f)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]),
               "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7])
               : "memory");
}

__device__ __forceinline__ float adam_one(float pi, float grad, float& mi, float& vi, const AdamCoef& c) {
  grad *= c.gscale;
  if (c.wd != 0.f) grad += c.wd * pi;
  mi = c.b1 * mi + (1.f - c.b1) * grad;
  vi = c.b2 * vi + (1.f - c.b2) * grad * grad;
  const float denom = sqrtf(vi) / c.bc2_sqrt + c.eps;   // same operation order as adam_dev_kernel: bit-identical updates
  return pi - c.lr_bc1 * (mi / denom);
}

// K3 > 0: compile-time tap count (27, 8, 125); K3 == 0: run-time.  TCI: C_in extent of a brick (32, or 8 when 5x5x5 weights
// are present).  FULL: the brick is complete, 32 x TCI (all index arithmetic is shifts / constant divisions).
template <int K3, int TCI, bool FULL>
__device__ __forceinline__ void adam_brick(const AdamDesc& d, int tix, float* __restrict__ tile,
                                           float* __restrict__ p, const float* __restrict__ g,
                                           const float* __restrict__ dw, float* __restrict__ m, float* __restrict__ v,
                                           __nv_bfloat16* __restrict__ packs, const AdamCoef& c, bool use_dw) {
  const int k3 = K3 ? K3 : d.k3;
  const int cout = d.cout, cin = d.cin;
  constexpr int tile_ci = TCI;
  const int tiles_ci = (cin + tile_ci - 1) / tile_ci;
  const int co0 = (tix / tiles_ci) * kBrickCo, ci0 = (tix % tiles_ci) * tile_ci;
  const int nco = FULL ? 32 : min(kBrickCo, cout - co0), nci = FULL ? TCI : min(tile_ci, cin - ci0);
  const int row = tile_ci * k3 + 1;     // odd: column walks through the brick are bank-conflict free
  const int run = nci * k3;             // consecutive floats per output channel in torch layout
  const int nthr = blockDim.x, tid = threadIdx.x;
  const int total = nco * run;
  __syncthreads();                      // the previous brick's pack writers are done with the tile
  if (use_dw) {
    // packed side: for every (tap, ci) a run of nco consecutive floats
    const float* src = dw + d.src;
    if (FULL && (cout & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      // 128-bit loads along C_out (8 per (tap, ci) row of the brick); the four values go to four tile rows
#pragma unroll 4
      for (int i = tid; i < total / 4; i += nthr) {
        const int co4 = (i & 7) << 2, r = i >> 3;     // r = t * TCI + ci
        const int ci = r % TCI, t = r / TCI;
        const float4 q = *reinterpret_cast<const float4*>(src + (static_cast<long long>(t) * cin + ci0 + ci) * cout + co0 + co4);
        float* tl = tile + co4 * row + ci * k3 + t;
        tl[0] = q.x;
        tl[row] = q.y;
        tl[2 * row] = q.z;
        tl[3 * row] = q.w;
      }
    } else
#pragma unroll 8
    for (int i = tid; i < total; i += nthr) {
      const int co = i % nco, r = i / nco;          // r = t * nci + ci
      const int ci = r % nci, t = r / nci;
      tile[co * row + ci * k3 + t] = src[(static_cast<long long>(t) * cin + ci0 + ci) * cout + co0 + co];
    }
    __syncthreads();
  }
  // torch side: Adam over the brick, new parameters stay in the tile
  {
    const long long base = d.src + (static_cast<long long>(co0) * cin + ci0) * k3;
    const long long co_stride = static_cast<long long>(cin) * k3;
    const bool vec = (run % 4 == 0) && (co_stride % 4 == 0) && (base % 4 == 0);
    const bool vec8 = (run % 8 == 0) && (co_stride % 8 == 0) && (base % 8 == 0);
    if (vec8) {
      // 256-bit accesses (sm_100 LDG/STG.256): half the memory requests of the float4 form
      const int run8 = run >> 3;
#pragma unroll 2
      for (int i = tid; i < nco * run8; i += nthr) {
        const int co = i / run8, r = (i - co * run8) << 3;
        const long long idx = base + co * co_stride + r;
        float pv[8], gv[8], mv[8], vv[8], o[8];
        ld8f(p + idx, pv);
        ld8f(g + idx, gv);
        ld8f(m + idx, mv);
        ld8f(v + idx, vv);
        float* tl = tile + co * row + r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = adam_one(pv[j], gv[j] + (use_dw ? tl[j] : 0.f), mv[j], vv[j], c);
          tl[j] = o[j];
        }
        st8f(p + idx, o);
        st8f(m + idx, mv);
        st8f(v + idx, vv);
      }
    } else if (vec) {
      const int run4 = run >> 2;
#pragma unroll 2
      for (int i = tid; i < nco * run4; i += nthr) {
        const int co = i / run4, r = (i - co * run4) << 2;
        const long long idx = base + co * co_stride + r;
        const float4 pv = *reinterpret_cast<const float4*>(p + idx), gv = *reinterpret_cast<const float4*>(g + idx);
        float4 mv = *reinterpret_cast<const float4*>(m + idx), vv = *reinterpret_cast<const float4*>(v + idx);
        float* tl = tile + co * row + r;
        float4 o;
        o.x = adam_one(pv.x, gv.x + (use_dw ? tl[0] : 0.f), mv.x, vv.x, c);
        o.y = adam_one(pv.y, gv.y + (use_dw ? tl[1] : 0.f), mv.y, vv.y, c);
        o.z = adam_one(pv.z, gv.z + (use_dw ? tl[2] : 0.f), mv.z, vv.z, c);
        o.w = adam_one(pv.w, gv.w + (use_dw ? tl[3] : 0.f), mv.w, vv.w, c);
        *reinterpret_cast<float4*>(p + idx) = o;
        *reinterpret_cast<float4*>(m + idx) = mv;
        *reinterpret_cast<float4*>(v + idx) = vv;
        tl[0] = o.x;
        tl[1] = o.y;
        tl[2] = o.z;
        tl[3] = o.w;
      }
    } else {
      for (int i = tid; i < total; i += nthr) {
        const int co = i / run, r = i - co * run;
        const long long idx = base + co * co_stride + r;
        float mi = m[idx], vi = v[idx];
        float* tl = tile + co * row + r;
        const float o = adam_one(p[idx], g[idx] + (use_dw ? *tl : 0.f), mi, vi, c);
        p[idx] = o;
        m[idx] = mi;
        v[idx] = vi;
        *tl = o;
      }
    }
  }
  __syncthreads();
  // bf16 packs from the tile
  __nv_bfloat16* pf = packs + d.dst;                                             // [t][co][ci]
  __nv_bfloat16* pd = pf + static_cast<long long>(cout) * cin * k3;              // [k3-1-t][ci][co]
  if (FULL && TCI != 32) {
    // two bf16 per store: pairs along ci (fprop pack), then pairs along co (dgrad pack)
#pragma unroll 4
    for (int i = tid; i < k3 * 32 * (TCI / 2); i += nthr) {
      const int j = (i % (TCI / 2)) << 1, o = (i / (TCI / 2)) & 31, t = i / (TCI / 2 * 32);
      const float a = tile[o * row + j * k3 + t], b = tile[o * row + (j + 1) * k3 + t];
      *reinterpret_cast<__nv_bfloat162*>(pf + (static_cast<long long>(t) * cout + co0 + o) * cin + ci0 + j) =
          __floats2bfloat162_rn(a, b);
    }
#pragma unroll 4
    for (int i = tid; i < k3 * TCI * 16; i += nthr) {
      const int j = (i & 15) << 1, o = (i >> 4) % TCI, t = (i >> 4) / TCI;
      const int tt = k3 - 1 - t;
      const float a = tile[j * row + o * k3 + tt], b = tile[(j + 1) * row + o * k3 + tt];
      *reinterpret_cast<__nv_bfloat162*>(pd + (static_cast<long long>(t) * cin + ci0 + o) * cout + co0 + j) =
          __floats2bfloat162_rn(a, b);
    }
  } else if (FULL) {
    // two bf16 per store: pairs along ci (fprop pack) / along co (dgrad pack)
#pragma unroll 4
    for (int i = tid; i < k3 * 32 * 16; i += nthr) {
      const int j = (i & 15) << 1, o = (i >> 4) & 31, t = i >> 9;
      {   // fprop: o = co, j = ci
        const float a = tile[o * row + j * k3 + t], b = tile[o * row + (j + 1) * k3 + t];
        *reinterpret_cast<__nv_bfloat162*>(pf + (static_cast<long long>(t) * cout + co0 + o) * cin + ci0 + j) =
            __floats2bfloat162_rn(a, b);
      }
      {   // dgrad: o = ci, j = co, taps flipped
        const int tt = k3 - 1 - t;
        const float a = tile[j * row + o * k3 + tt], b = tile[(j + 1) * row + o * k3 + tt];
        *reinterpret_cast<__nv_bfloat162*>(pd + (static_cast<long long>(t) * cin + ci0 + o) * cout + co0 + j) =
            __floats2bfloat162_rn(a, b);
      }
    }
  } else {
    for (int i = tid; i < total; i += nthr) {
      {
        const int ci = i % nci, r = i / nci;
        const int co = r % nco, t = r / nco;
        pf[(static_cast<long long>(t) * cout + co0 + co) * cin + ci0 + ci] = __float2bfloat16(tile[co * row + ci * k3 + t]);
      }
      {
        const int co = i % nco, r = i / nco;
        const int ci = r % nci, t = r / nci;
        pd[(static_cast<long long>(t) * cin + ci0 + ci) * cout + co0 + co] =
            __float2bfloat16(tile[co * row + ci * k3 + (k3 - 1 - t)]);
      }
    }
  }
}

__global__ void __launch_bounds__(256, 2)
    adam_fused_kernel(float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ dw,
                      float* __restrict__ m, float* __restrict__ v, __nv_bfloat16* __restrict__ packs,
                      const AdamDesc* __restrict__ descs, int ndesc, int total_tiles, int tile_ci,
                      const float* __restrict__ hyper, int* __restrict__ state, int use_dw) {
  extern __shared__ float tile[];   // [32][tile_ci * max_k3 + 1]
  AdamCoef c;
  {
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2];
    const float step = static_cast<float>(state[0] + 1);
    c.b1 = b1;
    c.b2 = b2;
    c.eps = hyper[3];
    c.wd = hyper[4];
    c.gscale = hyper[5];
    c.lr_bc1 = lr / (1.f - powf(b1, step));
    c.bc2_sqrt = sqrtf(1.f - powf(b2, step));
  }
  for (int gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
    // record owning brick gt: last record with tile0 <= gt
    int lo = 0, hi = ndesc - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].tile0 <= gt) lo = mid;
      else hi = mid - 1;
    }
    const AdamDesc d = descs[lo];
    const int tix = gt - d.tile0;
    if (d.kind == 1) {
      const long long first = d.src + static_cast<long long>(tix) * kPlainChunk;
      const long long last = min(first + kPlainChunk, d.src + static_cast<long long>(d.cout));
      for (long long i = first + threadIdx.x; i < last; i += blockDim.x) {
        float mi = m[i], vi = v[i];
        p[i] = adam_one(p[i], g[i], mi, vi, c);
        m[i] = mi;
        v[i] = vi;
      }
      continue;
    }
    const int tiles_ci = (d.cin + tile_ci - 1) / tile_ci;
    const int co0 = (tix / tiles_ci) * kBrickCo, ci0 = (tix % tiles_ci) * tile_ci;
    const bool full = d.cout - co0 >= 32 && d.cin - ci0 >= tile_ci && ((d.cout | d.cin | d.dst) & 1) == 0;
    const bool udw = use_dw != 0;
#define B200_BRICK(K3_, TCI_)                                                                     \
  do {                                                                                            \
    if (full) adam_brick<K3_, TCI_, true>(d, tix, tile, p, g, dw, m, v, packs, c, udw);            \
    else adam_brick<K3_, TCI_, false>(d, tix, tile, p, g, dw, m, v, packs, c, udw);                \
  } while (0)
    if (tile_ci == 32) {
      if (d.k3 == 27) B200_BRICK(27, 32);
      else if (d.k3 == 8) B200_BRICK(8, 32);
      else B200_BRICK(0, 32);
    } else {   // networks with 5x5x5 weights (vnet3d.py:25): 32 x 8 bricks for every record of the launch
      if (d.k3 == 125) B200_BRICK(125, 8);
      else if (d.k3 == 8) B200_BRICK(8, 8);
      else B200_BRICK(0, 8);
    }
#undef B200_BRICK
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

extern "C" int b200seg_adam_step_fused(float* param, const float* grad, const float* dw, float* exp_avg,
                                       float* exp_avg_sq, void* packs, const void* descs, int ndesc, int max_k3,
                                       int total_tiles, const float* hyper, int* state, int use_dw, void* stream) {
  B200_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && packs && descs && hyper && state && ndesc > 0 &&
                     total_tiles > 0 && max_k3 >= 1 && max_k3 <= 125,
                 "adam_step_fused: bad arguments");
  B200_CHECK_ARG(!use_dw || dw, "adam_step_fused: use_dw needs the packed gradient arena");
  // bricks of 32 x 32 x k^3 floats while they fit shared memory twice per SM (k^3 <= 27), 32 x 8 x k^3 beyond
  const int tile_ci = max_k3 <= 27 ? 32 : 8;
  const size_t smem = static_cast<size_t>(b200::kBrickCo) * (tile_ci * max_k3 + 1) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    if (cudaFuncSetAttribute(b200::adam_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
        cudaSuccess) {
      b200::set_error("adam_step_fused: cannot raise the dynamic shared memory limit to %zu bytes", smem);
      return B200SEG_ERR_CUDA;
    }
    attr = smem;
  }
  const int grid = std::min(total_tiles, b200::kNumSMs * 2);
  b200::adam_fused_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      param, grad, dw, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(packs), static_cast<const b200::AdamDesc*>(descs),
      ndesc, total_tiles, tile_ci, hyper, state, use_dw);
  B200_CHECK_LAUNCH("adam_step_fused");
  return 0;
}
