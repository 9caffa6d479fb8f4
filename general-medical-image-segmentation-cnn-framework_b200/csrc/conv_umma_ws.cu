// tcgen05 implicit-GEMM convolution for the K-HEAVY layers on TINY grids (8 x 8 planes: the U-Net bottleneck 256 -> 512 ->
// 512 at 8^3, unet3d.py:28; V-Net's 256-channel 5x5x5 layers at 8^3, vnet3d.py:25), "weights stationary, split over taps".
//
// Such a layer is 1 024 voxels against 14-16 MB of weights.  The voxel-tiled kernels give every 128-voxel tile its own pass
// over the whole weight slice (8-64 CTAs, each streaming megabytes through L2) and end up at 60-240 TFLOP/s.  Here the
// work is cut the other way: a unit is one (kd, kh) tap row x one N tile of NT output channels.  Its weights -- k kw taps x
// NT x C_in, 48-98 KB -- are loaded into shared memory ONCE and stay; the CTA then walks over all M tiles (128 voxels = two
// whole 8 x 8 planes; a kw tap is a shifted view of the haloed plane box, the (kd, kh) tap is the box origin, borders are
// TMA zero fill) with one TMEM accumulator per M tile (mtiles x NT <= 512 columns).  k^2 x C_out / NT units (144 for
// 512 -> 512) fill the 148 SMs; the k^2 partial sums of an output element meet in an fp32 workspace through vector
// reductions (red.global.add.v4.f32), and a small second kernel adds the bias, takes the BatchNorm statistics (or applies
// the eval-mode scale / shift / activation) and stores bf16.
// Warp roles (192 threads): 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue (one TMEM lane quadrant each).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.h"
#include "ptx.cuh"

namespace b200 {

bool encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int kc);

constexpr int kWsThreads = 192;
constexpr int kWsKC = 64;          // channels per shared-memory row (128 bytes, 128-byte swizzle)

struct WsParams {
  int n, od, cin, cout, k, pad;
  int NT, WB, nchunks, mtiles, tiles_per_n, n_ntiles, units, S;
  int msplit, csplit, mper, cper;  // a unit covers mper = mtiles / msplit M tiles and cper = nchunks / csplit K chunks
  unsigned slotA, bytesA, wblock_bytes, wbytes_total;
  float* ws;                       // [n * od * 64][cout] fp32, zeroed before the launch
};

struct WsUnit {
  int nt, kh, kd, m0, c0;
};
__device__ __forceinline__ WsUnit ws_decode(const WsParams& p, int unit) {
  WsUnit u;
  u.c0 = (unit % p.csplit) * p.cper;
  unit /= p.csplit;
  u.m0 = (unit % p.msplit) * p.mper;
  unit /= p.msplit;
  u.nt = unit % p.n_ntiles;
  unit /= p.n_ntiles;
  u.kh = unit % p.k;
  u.kd = unit / p.k;
  return u;
}

template <int NT>
__global__ void __launch_bounds__(kWsThreads, 1)
    conv_umma_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                            // [chunk][kw tap][NT rows][64 channels]
  uint8_t* sA = smem + p.wbytes_total;                           // [S] plane-pair boxes [2][8][WB][64 channels]
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sA + static_cast<size_t>(p.S) * p.slotA);
  uint64_t* emptyA = fullA + p.S;
  uint64_t* wFull = emptyA + p.S;
  uint64_t* wEmpty = wFull + 1;
  uint64_t* accFull = wEmpty + 1;
  uint64_t* accEmpty = accFull + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accEmpty + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kRowBytes = kWsKC * 2;
  constexpr int KS = kWsKC / 16;

  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) {
      mbar_init(&fullA[i], 1);
      mbar_init(&emptyA[i], 1);
    }
    mbar_init(wFull, 1);
    mbar_init(wEmpty, 1);
    mbar_init(accFull, 1);
    mbar_init(accEmpty, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  const int k = p.k;

  if (warp == 0) {
    // =========================== producer ===========================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int s = 0;
      uint32_t ph = 0, uph = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        const WsUnit u = ws_decode(p, unit);
        mbar_wait(wEmpty, uph ^ 1);           // the previous unit's instructions have finished reading the weights
        mbar_arrive_expect_tx(wFull, p.wbytes_total);
        for (int c = 0; c < p.cper; ++c)
          for (int e = 0; e < k; ++e)         // packed weights are [tap = (kd*k + kh)*k + kw][C_out][C_in]
            tma_load_3d(sW + static_cast<size_t>(c * k + e) * p.wblock_bytes, &tmB, wFull, (u.c0 + c) * kWsKC, u.nt * NT,
                        (u.kd * k + u.kh) * k + e);
        for (int m = u.m0; m < u.m0 + p.mper; ++m) {
          const int nn = m / p.tiles_per_n, d0 = (m % p.tiles_per_n) * 2;
          for (int c = 0; c < p.cper; ++c) {
            mbar_wait(&emptyA[s], ph ^ 1);
            mbar_arrive_expect_tx(&fullA[s], p.bytesA);
            tma_load_5d(sA + static_cast<size_t>(s) * p.slotA, &tmA, &fullA[s], (u.c0 + c) * kWsKC, -p.pad, u.kh - p.pad,
                        d0 + u.kd - p.pad, nn);
            if (++s == p.S) {
              s = 0;
              ph ^= 1;
            }
          }
        }
        uph ^= 1;
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
    const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 16, static_cast<uint32_t>(p.WB) * kRowBytes, SWZ_128B) >> 32);
    const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 16, 8u * kRowBytes, SWZ_128B) >> 32);
    const uint32_t lo_fixed = 1u << 16;
    const uint32_t sA16 = smem_u32(sA) >> 4, sW16 = smem_u32(sW) >> 4;
    const uint32_t slotA16 = p.slotA >> 4, wblock16 = p.wblock_bytes >> 4, row16 = kRowBytes >> 4;
    int s = 0;
    uint32_t ph = 0, uph = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      mbar_wait(accEmpty, uph ^ 1);           // the epilogue has drained the previous unit's accumulators
      mbar_wait(wFull, uph);
      tc_fence_after();
      for (int m = 0; m < p.mper; ++m) {
        const uint32_t d_tmem = tbase + static_cast<uint32_t>(m * NT);
        for (int c = 0; c < p.cper; ++c) {
          mbar_wait(&fullA[s], ph);
          tc_fence_after();
          const uint32_t a_lo0 = __shfl_sync(0xffffffffu, ((sA16 + s * slotA16) & 0x3FFF) | lo_fixed, 0);
          for (int e = 0; e < k; ++e) {
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(e) * row16;
            const uint32_t b_lo = __shfl_sync(0xffffffffu, ((sW16 + (c * k + e) * wblock16) & 0x3FFF) | lo_fixed, 0);
#pragma unroll
            for (int kk = 0; kk < KS; ++kk)
              umma_f16_pred_lohi(d_tmem, a_lo + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc, (c | e | kk) != 0 ? 1u : 0u, leader);
          }
          umma_commit_pred(&emptyA[s], leader);
          if (++s == p.S) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      umma_commit_pred(accFull, leader);
      umma_commit_pred(wEmpty, leader);
      uph ^= 1;
    }
  } else {
    // =========================== epilogue: TMEM -> fp32 workspace (vector reductions) ===========================
    const int q = warp & 3;                   // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;              // tile row = voxel (plane r >> 6, h = (r >> 3) & 7, w = r & 7)
    uint32_t uph = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const WsUnit u = ws_decode(p, unit);
      mbar_wait(accFull, uph);
      tc_fence_after();
      for (int mi = 0; mi < p.mper; ++mi) {
        const int m = u.m0 + mi;
        const int nn = m / p.tiles_per_n, dd = (m % p.tiles_per_n) * 2 + (r >> 6);
        float* dst = p.ws + (((static_cast<long long>(nn) * p.od + dd) * 8 + ((r >> 3) & 7)) * 8 + (r & 7)) * p.cout + u.nt * NT;
        const uint32_t t_lane = tbase + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mi * NT);
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(t_lane + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(__uint_as_float(v[j])),
                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accEmpty);
      uph ^= 1;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

// workspace -> output: + bias, per-channel {sum, sumsq} of the result (or the eval-mode scale / shift / activation), bf16.
// Block = (C_out / 8 channel chunks) x (256 / (C_out / 8) rows).
__global__ void __launch_bounds__(256) conv_ws_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ bias,
                                                               const float* __restrict__ scale, int act, float slope,
                                                               __nv_bfloat16* __restrict__ out, long long out_pitch,
                                                               float* __restrict__ stats, int rows, int cout) {
  extern __shared__ float red[];              // [2][cout]
  const int cv = cout >> 3;
  const int ch = threadIdx.x % cv, ty = threadIdx.x / cv, TY = blockDim.x / cv;
  float b8[8], sc8[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    b8[j] = bias ? bias[ch * 8 + j] : 0.f;
    sc8[j] = scale ? scale[ch * 8 + j] : 1.f;
    s1[j] = s2[j] = 0.f;
  }
  if (stats != nullptr) {
    for (int i = threadIdx.x; i < 2 * cout; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
  }
  for (int row = blockIdx.x * TY + ty; row < rows; row += gridDim.x * TY) {
    const float4 a = *reinterpret_cast<const float4*>(ws + static_cast<long long>(row) * cout + ch * 8);
    const float4 b = *reinterpret_cast<const float4*>(ws + static_cast<long long>(row) * cout + ch * 8 + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (scale) {
        const float z = fmaf(v[j], sc8[j], b8[j]);
        v[j] = act == B200SEG_ACT_NONE ? z : (z > 0.f ? z : (act == B200SEG_ACT_RELU ? 0.f : slope * z));
      } else {
        v[j] += b8[j];
      }
      s1[j] += v[j];
      s2[j] = fmaf(v[j], v[j], s2[j]);
    }
    st8(out + static_cast<long long>(row) * out_pitch + ch * 8, pack8(v));
  }
  if (stats != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[ch * 8 + j], s1[j]);
      atomicAdd(&red[cout + ch * 8 + j], s2[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * cout; i += blockDim.x) atomicAdd(&stats[i], red[i]);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static bool plan_ws(const UmmaConvArgs& a, WsParams& p, size_t& smem_bytes) {
  if (getenv("B200SEG_DISABLE_UMMA") || getenv("B200SEG_DISABLE_WS")) return false;
  if (!(a.k == 3 || a.k == 5) || a.dil != 1 || a.pad != (a.k - 1) / 2) return false;
  if (a.scatter_cout || a.gather2 || a.tapmode) return false;
  if (a.oh != 8 || a.ow != 8 || a.h != 8 || a.w != 8 || a.od != a.d || (a.od & 1)) return false;
  if (a.cin % kWsKC || a.cout % 16 || 256 % (a.cout / 8) || a.in_pitch % 8 || a.out_pitch % 8) return false;
  if (a.cin * a.k * a.k * a.k < 27 * 256) return false;             // K-heavy layers only
  p = WsParams{};
  p.n = a.n; p.od = a.od; p.cin = a.cin; p.cout = a.cout; p.k = a.k; p.pad = a.pad;
  p.tiles_per_n = a.od / 2;
  p.mtiles = a.n * p.tiles_per_n;
  if (p.mtiles < 1 || p.mtiles > 32) return false;
  p.nchunks = a.cin / kWsKC;
  p.WB = 8 + a.k - 1;
  p.bytesA = static_cast<unsigned>(2 * 8 * p.WB) * kWsKC * 2;
  p.slotA = (p.bytesA + 1023) & ~1023u;
  // Unit shape: N tile x M split x K split with the least (rounds x instructions per unit x instruction time) whose
  // accumulators fit TMEM and whose weights leave room for >= 3 input boxes.  An M = 128 instruction costs ~48 cycles up to
  // N = 32 and ~55 at N = 64 (probes/mma_rate.cu), so the widest N tile that still yields >= 148 units wins; splitting M
  // or K re-reads nothing but the (L2-resident) activations and adds partial sums to the workspace.
  long long best = -1;
  for (int nt : {64, 32, 16}) {
    if (a.cout % nt) continue;
    for (int cs : {1, 2, 4}) {
      if (p.nchunks % cs) continue;
      const size_t wbytes = static_cast<size_t>(p.nchunks / cs) * a.k * nt * kWsKC * 2;
      if (wbytes + 3 * p.slotA + 2048 > 225 * 1024) continue;
      for (int ms : {1, 2, 4}) {
        if (p.mtiles % ms || (p.mtiles / ms) * nt > 512) continue;
        const int units = a.k * a.k * (a.cout / nt) * ms * cs;
        const long long per_unit = static_cast<long long>(p.mtiles / ms) * (p.nchunks / cs) * a.k * 4 * (nt == 64 ? 55 : 48) +
                                   2000 + static_cast<long long>(p.mtiles / ms) * nt * 8;   // + fill, epilogue
        const long long cost = static_cast<long long>((units + kNumSMs - 1) / kNumSMs) * per_unit;
        if (best < 0 || cost < best) {
          best = cost;
          p.NT = nt;
          p.msplit = ms;
          p.csplit = cs;
        }
      }
    }
  }
  if (best < 0) return false;
  p.mper = p.mtiles / p.msplit;
  p.cper = p.nchunks / p.csplit;
  p.n_ntiles = a.cout / p.NT;
  p.units = a.k * a.k * p.n_ntiles * p.msplit * p.csplit;
  p.wblock_bytes = static_cast<unsigned>(p.NT) * kWsKC * 2;
  p.wbytes_total = static_cast<unsigned>(p.cper * a.k) * p.wblock_bytes;
  p.S = static_cast<int>(std::min<size_t>(8, (225 * 1024 - p.wbytes_total - 2048) / p.slotA));
  if (p.S < 3) return false;
  smem_bytes = static_cast<size_t>(p.wbytes_total) + static_cast<size_t>(p.S) * p.slotA + 1024 + 1024;
  return smem_bytes <= 227 * 1024;
}

size_t conv_umma_ws_bytes(const UmmaConvArgs& a) {
  WsParams p;
  size_t smem;
  UmmaConvArgs b = a;
  if (!plan_ws(b, p, smem)) return 0;
  return static_cast<size_t>(a.n) * a.od * 64 * a.cout * sizeof(float);
}

bool conv_umma_ws_supported(const UmmaConvArgs& a) {
  const size_t need = conv_umma_ws_bytes(a);
  return need != 0 && a.ws != nullptr && a.ws_bytes >= need && (reinterpret_cast<uintptr_t>(a.ws) & 15) == 0;
}

template <int NT>
static int launch_ws(const CUtensorMap& tmA, const CUtensorMap& tmB, const WsParams& p, size_t smem, int ctas, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_umma_ws_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("conv_umma_ws: cannot raise the dynamic shared memory limit");
      return B200SEG_ERR_CUDA;
    }
    attr_set = true;
  }
  conv_umma_ws_kernel<NT><<<ctas, kWsThreads, smem, st>>>(tmA, tmB, p);
  B200_CHECK_LAUNCH("conv_umma_ws");
  return 0;
}

int conv_umma_ws_run(const UmmaConvArgs& a, cudaStream_t st) {
  WsParams p;
  size_t smem;
  if (!plan_ws(a, p, smem) || !conv_umma_ws_supported(a)) {
    set_error("conv_umma_ws_run: unsupported geometry or missing workspace");
    return B200SEG_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.wpack)) & 15) {
    set_error("conv_umma_ws_run: buffers must be 16-byte aligned");
    return B200SEG_ERR_INVALID;
  }
  p.ws = static_cast<float*>(a.ws);
  const size_t ws_bytes = conv_umma_ws_bytes(a);
  if (cudaMemsetAsync(a.ws, 0, ws_bytes, st) != cudaSuccess) {
    set_error("conv_umma_ws_run: cannot clear the workspace");
    return B200SEG_ERR_CUDA;
  }
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.w), static_cast<uint64_t>(a.h),
                              static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.n)};
    const uint64_t pb = static_cast<uint64_t>(a.in_pitch) * 2;
    const uint64_t str[4] = {pb, pb * a.w, pb * a.w * a.h, pb * a.w * a.h * a.d};
    const uint32_t box[5] = {static_cast<uint32_t>(kWsKC), static_cast<uint32_t>(p.WB), 8u, 2u, 1u};
    if (!encode_bf16_map(&tmA, a.in, 5, dims, str, box, kWsKC)) return B200SEG_ERR_CUDA;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.cout),
                              static_cast<uint64_t>(a.k * a.k * a.k)};
    const uint64_t str[2] = {static_cast<uint64_t>(a.cin) * 2, static_cast<uint64_t>(a.cin) * a.cout * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(kWsKC), static_cast<uint32_t>(p.NT), 1u};
    if (!encode_bf16_map(&tmB, a.wpack, 3, dims, str, box, kWsKC)) return B200SEG_ERR_CUDA;
  }
  const int ctas = std::min(kNumSMs, p.units);
  int rc = B200SEG_ERR_INVALID;
  if (p.NT == 64) rc = launch_ws<64>(tmA, tmB, p, smem, ctas, st);
  else if (p.NT == 32) rc = launch_ws<32>(tmA, tmB, p, smem, ctas, st);
  else if (p.NT == 16) rc = launch_ws<16>(tmA, tmB, p, smem, ctas, st);
  if (rc) return rc;
  ++g_umma_launches;
  const int rows = a.n * a.od * 64, cv = a.cout / 8;
  if (256 % cv) {      // (C_out / 8 must divide the block: C_out in {16, 32, 64, 128, 256, 512, ...})
    set_error("conv_umma_ws_run: unsupported channel count %d", a.cout);
    return B200SEG_ERR_INVALID;
  }
  const int TY = 256 / cv;
  const int grid = std::max(1, std::min(kNumSMs * 2, (rows + TY - 1) / TY));
  conv_ws_finalize_kernel<<<grid, 256, 2 * a.cout * sizeof(float), st>>>(
      static_cast<const float*>(a.ws), a.bias, a.scale, a.act, a.slope, static_cast<__nv_bfloat16*>(a.out), a.out_pitch,
      a.scale ? nullptr : a.stats, rows, a.cout);
  B200_CHECK_LAUNCH("conv_ws_finalize");
  return 0;
}

}  // namespace b200
