// tcgen05 implicit-GEMM convolution for NARROW outputs (C_out = 16 / 32, 3x3x3, stride 1): "kd taps in N, rolling
// accumulators along d".  This is the kernel for the full-resolution U-Net layers, which hold half of the network's FLOPs.
//
// Why: a tcgen05.mma with M = 128 costs ~48 cycles however small N is (probes/mma_rate.cu: N=32 48.6, N=96 61 cycles), so
// the generic plane kernel (one N = C_out instruction per tap) cannot exceed 668 TFLOP/s at C_out = 32.  Here the three
// kd taps are stacked in the N dimension of ONE instruction:
//     D_j[voxel][a*C_out + co] = sum_{kh,kw,ci} in_plane_j[voxel + (kh,kw)][ci] * W[a][kh][kw][co][ci]          (N = 3*C_out)
// i.e. accumulator j belongs to INPUT plane j and holds its contribution to the three output planes q = j + pad - a.
// The output is assembled in the epilogue from three consecutive accumulators -- same TMEM lanes, different columns:
//     out_q = D_{q-pad}[block 0] + D_{q-pad+1}[block 1] + D_{q-pad+2}[block 2] + bias.
// Per 128 voxels of one plane that is 9*KS instructions of N = 96 (61 cycles) instead of 27*KS of N = 32 (48.6 cycles):
// 2.4x less tensor-pipe time, every input plane is fetched exactly once (no d halo), and the whole weight tensor
// (27 * C_out * C_in bf16 <= 110 KB) stays resident in shared memory for the life of the CTA.
//
// A CTA walks a tile column (16 x 8 voxels in h, w) along d through a ring of R = 5 accumulators (5 * 96 = 480 TMEM
// columns): while the epilogue assembles output plane q from accumulators q-1, q, q+1, the MMA warp is already filling
// q+2 and q+3.  Work items are (column, d-segment) pairs dealt round-robin to one persistent CTA per SM; a segment
// recomputes its two halo planes.
// Warp roles (384 threads): 0 = input-plane TMA producer (+ weight preload), 2 = MMA issuer, 3 = TMEM allocator,
// 4..11 = epilogue (warp w: lane quadrant w & 3, channel half (w >> 2) & 1).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.h"
#include "ptx.cuh"

// 1: MMA descriptors built on the uniform datapath (lane-0 broadcast bases, lo/hi halves) as in the plane and weight-
// gradient kernels.  Measured A/B on one box (probes/ab_roll.py) before choosing the default.
#ifndef B200_ROLL_UNIFORM
#define B200_ROLL_UNIFORM 0
#endif

namespace b200 {

bool encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int kc);

constexpr int kRollThreads = 384;
// accumulators in TMEM: KD of them feed the output plane being assembled, the rest are being filled (KD * C_out columns
// each: 5 * 96 = 480 for the 3x3x3 layers, 6 * 80 = 480 for 5x5x5 with 16 output channels per CTA)
__host__ __device__ constexpr int roll_ring(int kd) { return kd == 3 ? 5 : 6; }
constexpr int kRollEpiWarps = 8;

struct RollParams {
  int n, d, od, oh, ow, cout, pad;   // cout = channels of the whole output tensor
  const float* scale;                // AFF kernels: out = act(scale * conv + bias), see UmmaConvArgs::scale
  int act;
  float slope;
  int wide;                          // output rows allow 256-bit stores (pitch % 16 == 0, base 32-byte aligned)
  int halves;                        // 1, or 2: C_out = 64 handled as two independent 32-channel halves (CTA parity)
  long long out_pitch;
  int KC, NTOT;                    // channels per row (= C_in), N = 3 * C_out
  int WB, HB, S;                   // plane box and ring slots
  int tiles_w, tiles_h, segs, L;   // d-segments per column and their length
  long long items;
  unsigned slotA, bytesA, rowbytes, swz, wtile_bytes, wblock_bytes;
  __nv_bfloat16* out;
  const float* bias;
  float* stats;
};

struct RollItem {
  int nn, h0, w0, d0, lq;   // lq = output planes in this segment
};

template <int HV>
__device__ __forceinline__ RollItem roll_decode(const RollParams& p, long long it) {
  RollItem r;
  it /= HV;         // the half is the CTA's parity (grid and item count are even), not part of the item
  const int seg = static_cast<int>(it % p.segs);
  it /= p.segs;
  r.w0 = static_cast<int>(it % p.tiles_w) * 8;
  it /= p.tiles_w;
  r.h0 = static_cast<int>(it % p.tiles_h) * 16;
  r.nn = static_cast<int>(it / p.tiles_h);
  r.d0 = seg * p.L;
  r.lq = min(p.L, p.od - r.d0);
  return r;
}

// KD = kernel size (3, or 5: the V-Net stem and its 32 -> 2 output convolution, vnet3d.py:25,111), KS = K steps of 16
// channels (C_in / 16), CO = C_out per CTA, HV = halves of the output tensor (compile-time: the single-half kernel must
// not pay for the generality -- a run-time `halves` cost the 128^3 layers 20 %)
template <int KD, int KS, int CO, int HV, bool AFF>
__global__ void __launch_bounds__(kRollThreads, 1)
    conv_umma_roll_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const RollParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kRollRing = roll_ring(KD);
  uint8_t* sW = smem;                                          // [KD*KD (kh,kw)][KD (kd)][CO rows][KC]
  uint8_t* sA = smem + static_cast<unsigned>(KD * KD) * p.wtile_bytes;   // [S] plane boxes
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sA + static_cast<size_t>(p.S) * p.slotA);
  uint64_t* emptyA = fullA + p.S;
  uint64_t* accFull = emptyA + p.S;        // [R]
  uint64_t* accEmpty = accFull + kRollRing;
  uint64_t* wFull = accEmpty + kRollRing;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wFull + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 2);      // [CO]
  float* s_stats = s_bias + CO;                                // [2][CO]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // C_out = 64: even CTAs produce channels [0, 32), odd CTAs [32, 64) of the same tile columns at the same time (the
  // partner's input planes are L2 hits); each keeps only its own half of the weights resident.
  const int co_base = HV == 2 ? static_cast<int>(blockIdx.x & 1) * CO : 0;

  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) {
      mbar_init(&fullA[i], 1);
      mbar_init(&emptyA[i], 1);
    }
    for (int i = 0; i < kRollRing; ++i) {
      mbar_init(&accFull[i], 1);
      mbar_init(&accEmpty[i], kRollEpiWarps);
    }
    mbar_init(wFull, 1);
    fence_mbar_init();
  }
  if (tid < CO) s_bias[tid] = p.bias ? p.bias[co_base + tid] : 0.f;
  if (tid < 2 * CO) s_stats[tid] = 0.f;
  if (warp == 3) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  const long long first = blockIdx.x, stride = gridDim.x;
  constexpr int NTOT = KD * CO;

  if (warp == 0) {
    // =========================== producer: resident weights once, then one plane box per step ====================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      mbar_arrive_expect_tx(wFull, static_cast<unsigned>(KD * KD * KD) * p.wblock_bytes);
      for (int be = 0; be < KD * KD; ++be)
        for (int a = 0; a < KD; ++a)    // packed weights are [tap = (a*KD+kh)*KD+kw][C_out][C_in]
          tma_load_3d(sW + be * p.wtile_bytes + a * p.wblock_bytes, &tmB, wFull, 0, co_base, a * KD * KD + be);
      int s = 0;
      uint32_t ph = 0;
      for (long long it = first; it < p.items; it += stride) {
        const RollItem r = roll_decode<HV>(p, it);
        for (int t = 0; t < r.lq + KD - 1; ++t) {
          mbar_wait(&emptyA[s], ph ^ 1);
          mbar_arrive_expect_tx(&fullA[s], p.bytesA);
          tma_load_5d(sA + static_cast<size_t>(s) * p.slotA, &tmA, &fullA[s], 0, r.w0 - p.pad, r.h0 - p.pad,
                      r.d0 - p.pad + t, r.nn);
          if (++s == p.S) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // =========================== MMA issuer ===========================
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, NTOT, 0, 0);
    const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 16, static_cast<uint32_t>(p.WB) * p.rowbytes, p.swz) >> 32);
    const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 16, 8u * p.rowbytes, p.swz) >> 32);
    const uint32_t lo_fixed = 1u << 16;
#if B200_ROLL_UNIFORM
    const uint32_t sA16 = smem_u32(sA) >> 4, sW16 = __shfl_sync(0xffffffffu, smem_u32(sW) >> 4, 0);
#else
    const uint32_t sA16 = smem_u32(sA) >> 4, sW16 = smem_u32(sW) >> 4;
#endif
    const uint32_t slotA16 = p.slotA >> 4, wtile16 = p.wtile_bytes >> 4, row16 = p.rowbytes >> 4;
    mbar_wait(wFull, 0);
    tc_fence_after();
    int s = 0, acc = 0;
    uint32_t ph = 0, accph = 0;
    for (long long it = first; it < p.items; it += stride) {
      const RollItem r = roll_decode<HV>(p, it);
      for (int t = 0; t < r.lq + KD - 1; ++t) {
        mbar_wait(&accEmpty[acc], accph ^ 1);
        mbar_wait(&fullA[s], ph);
        tc_fence_after();
#if B200_ROLL_UNIFORM
        // lane-0 broadcast: marks the per-step base as warp-uniform so the descriptor arithmetic stays on the uniform datapath
        const uint32_t a_lo0 = __shfl_sync(0xffffffffu, ((sA16 + s * slotA16) & 0x3FFF) | lo_fixed, 0);
#else
        const uint32_t a_lo0 = ((sA16 + s * slotA16) & 0x3FFF) | lo_fixed;
#endif
        const uint32_t d_tmem = tbase + acc * NTOT;
#pragma unroll
        for (int b = 0; b < KD; ++b) {
#pragma unroll
          for (int e = 0; e < KD; ++e) {
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(b * p.WB + e) * row16;
            const uint32_t b_lo = ((sW16 + (b * KD + e) * wtile16) & 0x3FFF) | lo_fixed;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
#if B200_ROLL_UNIFORM
              umma_f16_pred_lohi(d_tmem, a_lo + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc, (b | e | kk) != 0 ? 1u : 0u, leader);
#else
              const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + 2u * kk);
              const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 2u * kk);
              umma_f16_pred(d_tmem, ad, bd, idesc, (b | e | kk) != 0 ? 1u : 0u, leader);
#endif
            }
          }
        }
        umma_commit_pred(&emptyA[s], leader);
        umma_commit_pred(&accFull[acc], leader);
        if (++s == p.S) {
          s = 0;
          ph ^= 1;
        }
        if (++acc == kRollRing) {
          acc = 0;
          accph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int q4 = warp & 3;                  // TMEM lane quadrant
    const int half = (warp >> 2) & 1;         // which half of the C_out columns
    constexpr int CH = CO / 2;                // channels per epilogue thread (8 or 16)
    const int m = q4 * 32 + lane;
    const bool want_stats = p.stats != nullptr;
    float s1[CH], s2[CH], bias_r[CH];
    float scale_r[AFF ? CH : 1];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      s1[j] = s2[j] = 0.f;
      bias_r[j] = s_bias[half * CH + j];
      if constexpr (AFF) scale_r[j] = p.scale[co_base + half * CH + j];
    }
    const uint32_t t_lane = tbase + (static_cast<uint32_t>(q4 * 32) << 16) + half * CH;
    int acc = 0;                              // ring slot of the plane about to be waited for
    uint32_t accph = 0;
    for (long long it = first; it < p.items; it += stride) {
      const RollItem r = roll_decode<HV>(p, it);
      const int oh_ = r.h0 + (m >> 3), ow_ = r.w0 + (m & 7);
      const bool hw_ok = oh_ < p.oh && ow_ < p.ow;
      for (int t = 0; t < r.lq + KD - 1; ++t) {
        mbar_wait(&accFull[acc], accph);
        tc_fence_after();
        if (t >= KD - 1) {
          // output plane u = t - (KD-1): block j of input plane u + j, j = 0 .. KD-1 (the newest one is in slot `acc`)
          uint32_t rr[KD][CH];
#pragma unroll
          for (int j = 0; j < KD; ++j) {
            const int slot = (acc + kRollRing - (KD - 1) + j) % kRollRing;
            if constexpr (CH == 16) tmem_ld_32x16(t_lane + slot * NTOT + j * CO, rr[j]);
            else tmem_ld_32x8(t_lane + slot * NTOT + j * CO, rr[j]);
          }
          tmem_ld_wait();
          const int od_ = r.d0 + t - (KD - 1);
          if (hw_ok) {
            float v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              float head = __uint_as_float(rr[0][j]) + __uint_as_float(rr[1][j]);   // (same order as the 3-tap sum)
#pragma unroll
              for (int a = 2; a < KD - 1; ++a) head += __uint_as_float(rr[a][j]);
              if constexpr (AFF) {
                const float z = fmaf(head + __uint_as_float(rr[KD - 1][j]), scale_r[j], bias_r[j]);
                v[j] = p.act == B200SEG_ACT_NONE ? z : (z > 0.f ? z : (p.act == B200SEG_ACT_RELU ? 0.f : p.slope * z));
              } else {
                v[j] = head + (__uint_as_float(rr[KD - 1][j]) + bias_r[j]);
              }
              if (want_stats) {
                s1[j] += v[j];
                s2[j] = fmaf(v[j], v[j], s2[j]);
              }
            }
            const long long vox = ((static_cast<long long>(r.nn) * p.od + od_) * p.oh + oh_) * p.ow + ow_;
            __nv_bfloat16* optr = p.out + vox * p.out_pitch + co_base + half * CH;
            if (CH == 16 && p.wide) {        // 32-byte aligned rows: one 256-bit store
              float t8[8], u8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                t8[i] = v[i];
                u8[i] = v[8 + i];
              }
              st16(optr, pack8(t8), pack8(u8));
            } else {
#pragma unroll
              for (int j = 0; j < CH; j += 8) {
                float t8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) t8[i] = v[j + i];
                st8(optr + j, pack8(t8));
              }
            }
          }
        }
        // plane t-(KD-1) has now fed all KD of its outputs (or lies outside the segment): free its accumulator
        tc_fence_before();
        __syncwarp();
        if (t >= KD - 1 && lane == 0) mbar_arrive(&accEmpty[(acc + kRollRing - (KD - 1)) % kRollRing]);
        if (t == r.lq + KD - 2 && lane == 0) {     // segment end: the last KD-1 planes have no further consumers
#pragma unroll
          for (int j = 0; j < KD - 1; ++j) mbar_arrive(&accEmpty[(acc + kRollRing - j) % kRollRing]);
        }
        if (++acc == kRollRing) {
          acc = 0;
          accph ^= 1;
        }
      }
    }
    if (want_stats) {
      // one reduction per CTA: lanes -> warp (shuffles), warps -> shared memory
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float a1 = warp_sum(s1[j]), a2 = warp_sum(s2[j]);
        if (lane == 0) {
          atomicAdd(&s_stats[half * CH + j], a1);
          atomicAdd(&s_stats[CO + half * CH + j], a2);
        }
      }
    }
  }
  __syncthreads();
  if (p.stats != nullptr && tid < 2 * CO)
    atomicAdd(&p.stats[HV == 1 ? tid : (tid / CO) * p.cout + co_base + tid % CO], s_stats[tid]);
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static bool plan_roll(const UmmaConvArgs& a, RollParams& p, size_t& smem_bytes) {
  if (getenv("B200SEG_DISABLE_UMMA") || getenv("B200SEG_DISABLE_ROLL")) return false;
  if (!(a.k == 3 || a.k == 5) || a.dil != 1 || a.pad < 0 || a.scatter_cout || a.gather2 || a.tapmode) return false;
  const int KD = a.k;
  int halves;
  if (KD == 3) {
    if (!(a.cout == 16 || a.cout == 32 || a.cout == 64) || (getenv("B200SEG_DISABLE_ROLL_HALVES") && a.cout == 64)) return false;
    if (!(a.cin == 16 || a.cin == 32 || a.cin == 64)) return false;
    halves = a.cout == 64 ? 2 : 1;   // two N = 96 instructions beat three N = 64 ones (61 vs 54.6 cycles each)
  } else {
    // 5x5x5 with (padded) 16-channel sides: N = 5 * 16 = 80 per instruction, 25 instead of 125 of them per K step
    if (getenv("B200SEG_DISABLE_ROLL5") || !(a.cout == 16 || a.cout == 32) || !(a.cin == 16 || a.cin == 32)) return false;
    halves = a.cout / 16;
  }
  const int CO = a.cout / halves;
  if (a.in_pitch % 8 || a.out_pitch % 8) return false;
  if (a.od != a.d + 2 * a.pad - (KD - 1) || a.oh != a.h + 2 * a.pad - (KD - 1) || a.ow != a.w + 2 * a.pad - (KD - 1)) return false;
  if (!(a.oh >= 16 && a.ow >= 8) || a.od < 4) return false;
  p = RollParams{};
  p.n = a.n; p.d = a.d; p.od = a.od; p.oh = a.oh; p.ow = a.ow; p.cout = a.cout; p.pad = a.pad;
  p.out_pitch = a.out_pitch;
  p.halves = halves;
  p.KC = a.cin;
  p.NTOT = KD * CO;
  p.rowbytes = p.KC * 2;
  p.swz = p.KC == 64 ? SWZ_128B : (p.KC == 32 ? SWZ_64B : SWZ_32B);
  p.WB = 8 + KD - 1;
  p.HB = 16 + KD - 1;
  p.bytesA = static_cast<unsigned>(p.WB * p.HB) * p.rowbytes;
  p.slotA = (p.bytesA + 1023) & ~1023u;
  p.wblock_bytes = static_cast<unsigned>(CO) * p.rowbytes;          // one kd tap: C_out rows (of this half)
  p.wtile_bytes = static_cast<unsigned>(KD) * p.wblock_bytes;       // one (kh, kw): KD * C_out rows, contiguous
  if (p.wblock_bytes % (8u * p.rowbytes)) return false;             // blocks must start on a swizzle-atom boundary
  // (16 -> 16 layers, HighRes3DNet's first stage: 512-byte blocks = two 32-byte-swizzle atoms; the 5x5x5 form has run
  //  with them all along)
  if (KD == 3 && p.wblock_bytes % 1024 && getenv("B200SEG_ROLL_STRICT_BLOCKS")) return false;
  const size_t fixed = static_cast<size_t>(KD * KD) * p.wtile_bytes + 2048 + 1024;
  const size_t budget = 220 * 1024;
  p.S = static_cast<int>(std::min<size_t>(8, (budget - fixed) / p.slotA));
  if (p.S < 3) return false;
  p.tiles_w = (a.ow + 7) / 8;
  p.tiles_h = (a.oh + 15) / 16;
  const long long cols = static_cast<long long>(a.n) * p.tiles_h * p.tiles_w * halves;
  // Items are dealt round-robin to the persistent CTAs, so the kernel takes ceil(items / SMs) rounds of (L + 2 halo planes
  // + ~1 plane of pipeline fill) plane steps: take the segment count with the fewest steps (ties: the longest segments).
  int L = a.od;
  long long best_cost = -1;
  for (int segs = 1; segs <= std::max(1, a.od / 4); ++segs) {
    const int cand = (a.od + segs - 1) / segs;
    const long long items = cols * ((a.od + cand - 1) / cand);
    const long long cost = ((items + kNumSMs - 1) / kNumSMs) * (cand + KD);
    if (best_cost < 0 || cost < best_cost) best_cost = cost, L = cand;
  }
  p.L = L;
  p.segs = (a.od + L - 1) / L;
  p.items = cols * p.segs;
  smem_bytes = fixed + static_cast<size_t>(p.S) * p.slotA;
  return smem_bytes <= 227 * 1024;
}

bool conv_umma_roll_supported(const UmmaConvArgs& a) {
  RollParams p;
  size_t smem;
  return plan_roll(a, p, smem);
}

template <int KD, int KS, int CO, int HV, bool AFF>
static int launch_roll_(const CUtensorMap& tmA, const CUtensorMap& tmB, const RollParams& p, size_t smem, int ctas,
                       cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_umma_roll_kernel<KD, KS, CO, HV, AFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess) {
      set_error("conv_umma_roll: cannot raise the dynamic shared memory limit");
      return B200SEG_ERR_CUDA;
    }
    attr_set = true;
  }
  conv_umma_roll_kernel<KD, KS, CO, HV, AFF><<<ctas, kRollThreads, smem, st>>>(tmA, tmB, p);
  B200_CHECK_LAUNCH("conv_umma_roll");
  return 0;
}
template <int KD, int KS, int CO, int HV>
static int launch_roll(const CUtensorMap& tmA, const CUtensorMap& tmB, const RollParams& p, size_t smem, int ctas,
                       cudaStream_t st) {
  return p.scale ? launch_roll_<KD, KS, CO, HV, true>(tmA, tmB, p, smem, ctas, st)
                 : launch_roll_<KD, KS, CO, HV, false>(tmA, tmB, p, smem, ctas, st);
}

int conv_umma_roll_run(const UmmaConvArgs& a, cudaStream_t st) {
  RollParams p;
  size_t smem;
  if (!plan_roll(a, p, smem)) {
    set_error("conv_umma_roll_run: unsupported geometry");
    return B200SEG_ERR_INVALID;
  }
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.bias = a.bias;
  p.stats = a.stats;
  p.scale = a.scale;
  p.act = a.act;
  p.slope = a.slope;
  p.wide = (a.out_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 31) == 0 && !getenv("B200SEG_NO_WIDE_STORES")) ? 1 : 0;
  if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.wpack)) & 15) {
    set_error("conv_umma_roll_run: buffers must be 16-byte aligned");
    return B200SEG_ERR_INVALID;
  }
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.w), static_cast<uint64_t>(a.h),
                              static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.n)};
    const uint64_t pb = static_cast<uint64_t>(a.in_pitch) * 2;
    const uint64_t str[4] = {pb, pb * a.w, pb * a.w * a.h, pb * a.w * a.h * a.d};
    const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB), 1u, 1u};
    if (!encode_bf16_map(&tmA, a.in, 5, dims, str, box, p.KC)) return B200SEG_ERR_CUDA;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.cout),
                              static_cast<uint64_t>(a.k * a.k * a.k)};
    const uint64_t str[2] = {static_cast<uint64_t>(a.cin) * 2, static_cast<uint64_t>(a.cin) * a.cout * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(a.cout / p.halves), 1u};
    if (!encode_bf16_map(&tmB, a.wpack, 3, dims, str, box, p.KC)) return B200SEG_ERR_CUDA;
  }
  int ctas = static_cast<int>(std::min<long long>(kNumSMs, p.items));
  if (p.halves == 2) ctas &= ~1;   // CTA parity selects the half: the stride over the items must be even
  const int KS = a.cin / 16;
  const int co = a.cout / p.halves;
  int rc = B200SEG_ERR_INVALID;
  if (a.k == 5) {
    if (p.halves == 2 && KS == 1) rc = launch_roll<5, 1, 16, 2>(tmA, tmB, p, smem, ctas, st);
    else if (p.halves == 2 && KS == 2) rc = launch_roll<5, 2, 16, 2>(tmA, tmB, p, smem, ctas, st);
    else if (p.halves == 1 && KS == 1) rc = launch_roll<5, 1, 16, 1>(tmA, tmB, p, smem, ctas, st);
    else if (p.halves == 1 && KS == 2) rc = launch_roll<5, 2, 16, 1>(tmA, tmB, p, smem, ctas, st);
  } else if (p.halves == 2) {
    if (KS == 1) rc = launch_roll<3, 1, 32, 2>(tmA, tmB, p, smem, ctas, st);
    else if (KS == 2) rc = launch_roll<3, 2, 32, 2>(tmA, tmB, p, smem, ctas, st);
    else if (KS == 4) rc = launch_roll<3, 4, 32, 2>(tmA, tmB, p, smem, ctas, st);
  } else if (co == 32 && KS == 1) rc = launch_roll<3, 1, 32, 1>(tmA, tmB, p, smem, ctas, st);
  else if (co == 32 && KS == 2) rc = launch_roll<3, 2, 32, 1>(tmA, tmB, p, smem, ctas, st);
  else if (co == 32 && KS == 4) rc = launch_roll<3, 4, 32, 1>(tmA, tmB, p, smem, ctas, st);
  else if (co == 16 && KS == 1) rc = launch_roll<3, 1, 16, 1>(tmA, tmB, p, smem, ctas, st);
  else if (co == 16 && KS == 2) rc = launch_roll<3, 2, 16, 1>(tmA, tmB, p, smem, ctas, st);
  else if (co == 16 && KS == 4) rc = launch_roll<3, 4, 16, 1>(tmA, tmB, p, smem, ctas, st);
  if (rc == 0) ++g_umma_launches;
  return rc;
}

}  // namespace b200
