// Persistent tcgen05 implicit-GEMM convolution, "plane mode" (H_out >= 16, W_out >= 8): the workhorse of the full-,
// half- and quarter-resolution U-Net layers (and of every stride-1 conv of the other models at those sizes).
//
// Same GEMM view and shared-memory scheme as conv_umma.cu (halo'd input planes brought in once by TMA, every tap a
// shifted-view UMMA descriptor over them), re-organised around what the first version's ncu profile showed:
//   * ONE persistent CTA per SM walks a static round-robin list of tiles; barriers, TMEM allocation and tensor-map
//     prefetch happen once per CTA instead of once per tile, and the TMA producers run ahead into the next tile.
//   * The 512 TMEM columns are split into two accumulator stages (P planes x NT channels each): the epilogue of tile i
//     (TMEM -> registers -> bias / statistics -> bf16 -> HBM) overlaps the MMAs of tile i+1.
//   * The MMA issuer was instruction-bound (~107 cycles per tcgen05.mma against a 48-cycle hardware floor at N=32,
//     probes/mma_rate.cu).  P (accumulators per tile) and KS (K=16 steps per chunk) are template parameters, so the
//     loop over accumulators is fully unrolled with the descriptors' low words held in registers: one add per MMA.
//   * 8 epilogue warps (two per TMEM lane quadrant) and one 32-column tcgen05.ld per step; per-channel statistics are
//     folded into shared memory per tile and flushed to global memory once per CTA.
//
// Warp roles (384 threads): 0 = input-plane TMA producer, 1 = weight-tile TMA producer, 2 = MMA issuer,
// 3 = TMEM allocator, 4..11 = epilogue (warp w: lane quadrant w & 3, accumulators acc with (acc & 1) == ((w >> 2) & 1)).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "conv_impl.h"
#include "ptx.cuh"

namespace b200 {

bool encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int kc);

struct PlaneParams {
  int n, od, oh, ow, cout;
  long long out_pitch;
  int k, pad, dil;
  int KC, nchunks, NT, n_ntiles;
  int WB, HB, U, S, NB;
  int wide;                   // output rows allow 256-bit stores (pitch and channel offsets % 16 == 0, base 32-byte aligned)
  int G;                      // weight tiles per ring barrier: k (one kw row) or 1, see conv_umma.cu
  int stages;   // TMEM accumulator stages: 2 = epilogue overlaps the next tile, 1 = all 512 columns for one (K-heavy) tile
  int tiles_w, tiles_h, tiles_d;
  long long tiles;
  int scatter_cout, cpm;
  unsigned slotA, slotB, rowbytes, swz, bytesA, bytesB;
  __nv_bfloat16* out;
  const float* bias;
  float* stats;
  // TAP kernels only (stride-2 decompositions, see UmmaConvArgs::tapmode): weight tile per (class, shifted tap), the class
  // of a single-map launch, and explicit output strides (elements) so that `out` may be a sub-lattice of a finer grid
  int cls;
  long long osn, osd, osh, osw;
  unsigned char tapw[8][8];
  // fused epilogue out = act(scale * conv + bias) (UmmaConvArgs::scale); scale == nullptr: conv + bias
  const float* scale;
  int act;
  float slope;
};

struct TensorMaps8P {
  CUtensorMap m[8];
};

constexpr int kEpiWarps = 8;
constexpr int kThreadsP = 32 * (4 + kEpiWarps);
constexpr int kStageCols = 256;

struct TileCoord {
  int nt, w0, h0, d0, nn;
};

__device__ __forceinline__ TileCoord decode_tile(const PlaneParams& p, long long t, int P) {
  TileCoord c;
  c.nt = static_cast<int>(t % p.n_ntiles);
  t /= p.n_ntiles;
  c.w0 = static_cast<int>(t % p.tiles_w) * 8;
  t /= p.tiles_w;
  c.h0 = static_cast<int>(t % p.tiles_h) * 16;
  t /= p.tiles_h;
  c.d0 = static_cast<int>(t % p.tiles_d) * P;
  c.nn = static_cast<int>(t / p.tiles_d);
  return c;
}

// One 32- or 16-column slab of one accumulator: TMEM -> +bias -> statistics -> bf16 -> global.
template <int CW>
__device__ __forceinline__ void epilogue_slab(uint32_t taddr, const float* s_bias_col, bool valid, __nv_bfloat16* optr,
                                              float (&s1)[CW], float (&s2)[CW], bool want_stats, bool wide,
                                              const float* scale_col = nullptr, int act = 0, float slope = 0.f) {
  uint32_t raw[CW];
  if constexpr (CW == 32) tmem_ld_32x32(taddr, raw);
  else tmem_ld_32x16(taddr, raw);
  tmem_ld_wait();
  if (valid) {
    float v[CW];
    if (scale_col != nullptr) {       // inference: folded BatchNorm scale / shift + activation
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const float z = fmaf(__uint_as_float(raw[j]), __ldg(scale_col + j), s_bias_col[j]);
        v[j] = act == B200SEG_ACT_NONE ? z : (z > 0.f ? z : (act == B200SEG_ACT_RELU ? 0.f : slope * z));
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(raw[j]) + s_bias_col[j];
    }
    if (want_stats) {
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        s1[j] += v[j];
        s2[j] = fmaf(v[j], v[j], s2[j]);
      }
    }
    if (wide) {                       // 32-byte aligned rows: 256-bit stores, half the L2 requests
#pragma unroll
      for (int j = 0; j < CW; j += 16) {
        float t8[8], u8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t8[i] = v[j + i];
          u8[i] = v[j + 8 + i];
        }
        st16(optr + j, pack8(t8), pack8(u8));
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; j += 8) {
        float t8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = v[j + i];
        st8(optr + j, pack8(t8));
      }
    }
  }
}

// Column sums over the 32 lanes of a warp for CW columns held one per register: after the call lane l (l < CW) holds
// the total of column l in v[0].
template <int CW>
__device__ __forceinline__ void warp_colsum(float (&v)[CW], int lane) {
  if constexpr (CW == 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
  }
#pragma unroll
  for (int s = (CW == 32 ? 16 : 8); s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (up ? v[i + s] : v[i]) + recv;
    }
  }
}

template <int P, int KS, bool TAP>
__global__ void __launch_bounds__(kThreadsP, 1)
    conv_umma_plane_kernel(const __grid_constant__ TensorMaps8P tmAs, const __grid_constant__ CUtensorMap tmB,
                           const PlaneParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + static_cast<size_t>(p.S) * p.slotA;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + static_cast<size_t>(p.NB) * p.slotB);
  uint64_t* emptyA = fullA + p.S;
  uint64_t* fullB = emptyA + p.S;
  uint64_t* emptyB = fullB + p.NB;
  uint64_t* accFull = emptyB + p.NB;   // [2]
  uint64_t* accEmpty = accFull + 2;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accEmpty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 2);   // [cout]
  float* s_stats = s_bias + p.cout;                          // [2][cout] (only when p.stats)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k = p.k, k2 = k * k;

  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) {
      mbar_init(&fullA[i], 1);
      mbar_init(&emptyA[i], 1);
    }
    for (int i = 0; i < p.NB; ++i) {
      mbar_init(&fullB[i], 1);
      mbar_init(&emptyB[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&accFull[i], 1);
      mbar_init(&accEmpty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < p.cout; i += kThreadsP)
    s_bias[i] = p.bias ? p.bias[p.scatter_cout ? i % p.scatter_cout : i] : 0.f;
  if (p.stats != nullptr)
    for (int i = tid; i < 2 * p.cout; i += kThreadsP) s_stats[i] = 0.f;
  if (warp == 3) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmAs.m[0]);
  if (warp == 1 && lane == 0) tma_prefetch_desc(&tmB);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  const long long first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // =========================== input-plane producer ===========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long t = first; t < p.tiles; t += step) {
        const TileCoord tc = decode_tile(p, t, P);
        for (int c = 0; c < p.nchunks; ++c) {
          const CUtensorMap* tm = &tmAs.m[c / p.cpm];
          const int c0 = (c % p.cpm) * p.KC;
          for (int u = 0; u < p.U; ++u) {
            mbar_wait(&emptyA[s], ph ^ 1);
            mbar_arrive_expect_tx(&fullA[s], p.bytesA);
            tma_load_5d(sA + static_cast<size_t>(s) * p.slotA, tm, &fullA[s], c0, tc.w0 - p.pad, tc.h0 - p.pad,
                        tc.d0 - p.pad + u, tc.nn);
            if (++s == p.S) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== weight-tile producer ===========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int k3 = k2 * k;
      const bool multimap = p.cpm < p.nchunks;   // gather mode: the map index doubles as the weight "tap"
      for (long long t = first; t < p.tiles; t += step) {
        const int nt = static_cast<int>(t % p.n_ntiles);
        for (int c = 0; c < p.nchunks; ++c) {
          if constexpr (TAP) {
            // shifted-tap mode: the taps that exist for this chunk's class, in the issuer's (a, b, e) order; G == 1
            const int cls = multimap ? c / p.cpm : p.cls;
            for (int sh = 0; sh < 8; ++sh) {
              const int wt = p.tapw[cls][sh];
              if (wt == 0xFF) continue;
              mbar_wait(&emptyB[s], ph ^ 1);
              mbar_arrive_expect_tx(&fullB[s], p.bytesB);
              tma_load_3d(sB + static_cast<size_t>(s) * p.slotB, &tmB, &fullB[s], (c % p.cpm) * p.KC, nt * p.NT, wt);
              if (++s == p.NB) {
                s = 0;
                ph ^= 1;
              }
            }
          } else {
            for (int tap = 0; tap < k3; tap += p.G) {     // one barrier pair per group of G tiles (a kw row)
              mbar_wait(&emptyB[s], ph ^ 1);
              mbar_arrive_expect_tx(&fullB[s], p.G * p.bytesB);
              for (int e = 0; e < p.G; ++e)
                tma_load_3d(sB + static_cast<size_t>(s * p.G + e) * p.slotB, &tmB, &fullB[s], (c % p.cpm) * p.KC, nt * p.NT,
                            multimap ? c / p.cpm : tap + e);
              if (++s == p.NB / p.G) {
                s = 0;
                ph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // =========================== MMA issuer ===========================
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, p.NT, 0, 0);
    const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 16, static_cast<uint32_t>(p.WB) * p.rowbytes, p.swz) >> 32);
    const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 16, 8u * p.rowbytes, p.swz) >> 32);
    const uint32_t lo_fixed = 1u << 16;
    const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
    const uint32_t slotA16 = p.slotA >> 4, slotB16 = p.slotB >> 4;
    const uint32_t row16 = p.rowbytes >> 4;
    int bs = 0;
    uint32_t bphase = 0;
    int unit0 = 0;            // ring slot of plane 0 of the current (tile, chunk)
    uint32_t unit0_phase = 0;
    uint32_t it = 0;
    for (long long t = first; t < p.tiles; t += step, ++it) {
      const uint32_t stage = p.stages == 2 ? (it & 1) : 0u;
      const uint32_t use = p.stages == 2 ? (it >> 1) : it;
      mbar_wait(&accEmpty[stage], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_base = tbase + stage * kStageCols;
      uint32_t started = 0;     // TAP: becomes 1 after the first issued tap of this tile
      for (int c = 0; c < p.nchunks; ++c) {
        int waited = 0, wslot = unit0;
        uint32_t wphase = unit0_phase;
        for (int a = 0; a < k; ++a) {
          const int need = min(p.U, P + a * p.dil);
          for (; waited < need; ++waited) {
            mbar_wait(&fullA[wslot], wphase);
            if (++wslot == p.S) {
              wslot = 0;
              wphase ^= 1;
            }
          }
          tc_fence_after();
          // descriptor low words of the P planes this kd iteration reads (plane acc + a*dil of the chunk)
          uint32_t a_lo[P];
          {
            int sl = unit0 + a * p.dil;
            if (sl >= p.S) sl -= p.S;
#pragma unroll
            for (int acc = 0; acc < P; ++acc) {
              // (lane-0 broadcast: warp-uniform hint, keeps the descriptor arithmetic on the uniform datapath)
              a_lo[acc] = __shfl_sync(0xffffffffu, ((sA16 + sl * slotA16) & 0x3FFF) | lo_fixed, 0);
              if (++sl == p.S) sl = 0;
            }
          }
          uint32_t tap16_row = 0;   // (b*dil*WB) rows -> 16-byte units, advanced per kh
          for (int b = 0; b < k; ++b) {
            uint32_t tap16 = tap16_row;
            for (int e = 0; e < k; ++e) {
              if constexpr (TAP) {
                if (p.tapw[p.cpm < p.nchunks ? c / p.cpm : p.cls][a * 4 + b * 2 + e] == 0xFF) {   // tap absent for this class
                  tap16 += p.dil * row16;
                  continue;
                }
              }
              const int ge = p.G == 1 ? 0 : e;
              if (ge == 0) {
                mbar_wait(&fullB[bs], bphase);
                tc_fence_after();
              }
              const uint32_t b_lo = __shfl_sync(0xffffffffu, ((sB16 + (bs * p.G + ge) * slotB16) & 0x3FFF) | lo_fixed, 0);
              const uint32_t fresh = TAP ? started : ((c | a | b | e) == 0 ? 0u : 1u);
              started = 1;
#pragma unroll
              for (int acc = 0; acc < P; ++acc) {
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                  umma_f16_pred_lohi(d_base + acc * p.NT, a_lo[acc] + tap16 + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc,
                                     kk == 0 ? fresh : 1u, leader);
                }
              }
              if (ge == p.G - 1) {
                umma_commit_pred(&emptyB[bs], leader);
                if (++bs == p.NB / p.G) {
                  bs = 0;
                  bphase ^= 1;
                }
              }
              tap16 += p.dil * row16;
            }
            tap16_row += p.dil * p.WB * row16;
          }
          // release the planes whose last reader was this kd iteration
          {
            int rs = unit0;
            for (int j = 0; j < p.U; ++j) {
              if (min(k - 1, j / p.dil) == a) umma_commit_pred(&emptyA[rs], leader);
              if (++rs == p.S) rs = 0;
            }
          }
        }
        unit0 += p.U;
        while (unit0 >= p.S) {
          unit0 -= p.S;
          unit0_phase ^= 1;
        }
      }
      umma_commit_pred(&accFull[stage], leader);
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int half = (warp >> 2) & 1;       // which accumulators (acc & 1 == half); P == 1: split the column slabs
    const int m = q * 32 + lane;            // tile row = voxel (h = m >> 3, w = m & 7)
    const bool want_stats = p.stats != nullptr;
    uint32_t it = 0;
    for (long long t = first; t < p.tiles; t += step, ++it) {
      const uint32_t stage = p.stages == 2 ? (it & 1) : 0u;
      const uint32_t use = p.stages == 2 ? (it >> 1) : it;
      const TileCoord tc = decode_tile(p, t, P);
      mbar_wait(&accFull[stage], use & 1);
      tc_fence_after();
      const int oh_ = tc.h0 + (m >> 3), ow_ = tc.w0 + (m & 7);
      const bool hw_ok = oh_ < p.oh && ow_ < p.ow;
      const uint32_t t_lane = tbase + stage * kStageCols + (static_cast<uint32_t>(q * 32) << 16);
      auto slabs = [&](auto cw_tag) {
        constexpr int CW = decltype(cw_tag)::value;
        const int nslab = p.NT / CW;
        for (int sidx = (P == 1 ? half : 0); sidx < nslab; sidx += (P == 1 ? 2 : 1)) {
          const int c0 = sidx * CW;
          const int col0 = tc.nt * p.NT + c0;
          float s1[CW], s2[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) s1[j] = s2[j] = 0.f;
#pragma unroll
          for (int acc = 0; acc < P; ++acc) {
            if (P > 1 && (acc & 1) != half) continue;
            const int od_ = tc.d0 + acc;
            const bool valid = hw_ok && od_ < p.od;
            __nv_bfloat16* optr;
            if (p.scatter_cout) {
              const int abe = 7 - col0 / p.scatter_cout, co0 = col0 % p.scatter_cout;
              const long long ovox = ((static_cast<long long>(tc.nn) * 2 * p.od + 2 * od_ + (abe >> 2)) * 2 * p.oh +
                                      2 * oh_ + ((abe >> 1) & 1)) * 2 * p.ow + 2 * ow_ + (abe & 1);
              optr = p.out + ovox * p.out_pitch + co0;
            } else if constexpr (TAP) {
              optr = p.out + tc.nn * p.osn + od_ * p.osd + oh_ * p.osh + ow_ * p.osw + col0;
            } else {
              const long long vox = ((static_cast<long long>(tc.nn) * p.od + od_) * p.oh + oh_) * p.ow + ow_;
              optr = p.out + vox * p.out_pitch + col0;
            }
            epilogue_slab<CW>(t_lane + acc * p.NT + c0, s_bias + col0, valid, optr, s1, s2, want_stats, p.wide != 0,
                              p.scale ? p.scale + col0 : nullptr, p.act, p.slope);
          }
          if (want_stats) {
            warp_colsum<CW>(s1, lane);
            warp_colsum<CW>(s2, lane);
            if (lane < CW) {
              atomicAdd(&s_stats[col0 + lane], s1[0]);
              atomicAdd(&s_stats[p.cout + col0 + lane], s2[0]);
            }
          }
        }
      };
      // (a 32-column slab must lie inside one pixel-shuffle offset: scatter_cout % 32, else 16-column slabs)
      if ((p.NT & 31) == 0 && (p.scatter_cout & 31) == 0) slabs(std::integral_constant<int, 32>{});
      else slabs(std::integral_constant<int, 16>{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accEmpty[stage]);
    }
  }
  __syncthreads();
  if (p.stats != nullptr) {
    for (int i = tid; i < 2 * p.cout; i += kThreadsP) {
      const float v = s_stats[i];
      if (v != 0.f) atomicAdd(&p.stats[i], v);
    }
  }
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host side
// kc_cap: largest K chunk (channels per shared-memory row) to try; relaxed: accept short planes (H_out, W_out >= 4; the
// rows of a 16 x 8 tile that fall outside are masked) -- the last resort for geometries no other kernel takes.
static bool plan_plane_kc(const UmmaConvArgs& a, PlaneParams& p, int& P, int& KS, size_t& smem_bytes, int kc_cap, bool relaxed) {
  if (getenv("B200SEG_DISABLE_UMMA") || getenv("B200SEG_DISABLE_PERSISTENT")) return false;
  if (a.cin % 16 || a.cout % 16) return false;
  if (a.in_pitch % 8 || a.out_pitch % 8) return false;
  const int halo = (a.k - 1) * a.dil;
  if (a.tapmode) {
    // 2x2x2 shifted taps (stride-2 decompositions): the caller vouches for the extents; short tiles are masked, so the
    // small deep levels (8^3 outputs) run here too instead of needing a flat-mode variant
    if (a.k != 2 || a.dil != 1 || a.pad < 0 || a.pad > 1 || a.scatter_cout || a.gather2 || (a.in_sub && a.out_sub)) return false;
    if (!(a.oh >= 4 && a.ow >= 4)) return false;
  } else {
    if (!(a.k == 1 || a.k == 3 || a.k == 5) || a.dil < 1 || a.pad < 0) return false;
    if (a.od != a.d + 2 * a.pad - halo || a.oh != a.h + 2 * a.pad - halo || a.ow != a.w + 2 * a.pad - halo) return false;
    if (relaxed ? !(a.oh >= 4 && a.ow >= 4) : !(a.oh >= 16 && a.ow >= 8)) return false;
  }
  p = PlaneParams{};
  p.n = a.n; p.od = a.od; p.oh = a.oh; p.ow = a.ow; p.cout = a.cout; p.out_pitch = a.out_pitch;
  p.k = a.k; p.pad = a.pad; p.dil = a.dil;
  p.scatter_cout = a.scatter_cout;
  if (a.scatter_cout && (a.scatter_cout % 16 || a.cout != 8 * a.scatter_cout || a.k != 1 || a.stats)) return false;
  const int cin_map = (a.gather2 || a.in_sub) ? a.cin / 8 : a.cin;
  if (a.gather2 && (a.cin % 8 || a.k != 1 || cin_map % 16)) return false;
  if (a.in_sub && (a.cin % 8 || cin_map % 16)) return false;
  p.KC = (cin_map % 64 == 0 && kc_cap >= 64) ? 64 : ((cin_map % 32 == 0 && kc_cap >= 32) ? 32 : 16);
  KS = p.KC / 16;
  p.nchunks = a.cin / p.KC;
  p.cpm = cin_map / p.KC;
  p.rowbytes = p.KC * 2;
  p.swz = p.KC == 64 ? SWZ_128B : (p.KC == 32 ? SWZ_64B : SWZ_32B);
  if (a.cout <= 128 && !a.scatter_cout) p.NT = a.cout;
  // pixel-shuffle GEMMs (ConvTranspose k2s2 forward) are bound by their scattered output writes; the widest N tile reads
  // the input once instead of twice and halves the tile count (probes/convt_knobs.py: 0.126 -> 0.115 ms at 2 x 64^3)
  else if (a.scatter_cout && a.cout % 256 == 0 && !getenv("B200SEG_SCATTER_NT")) p.NT = 256;
  else if (a.cout % 128 == 0) p.NT = 128;
  else if (a.cout % 64 == 0) p.NT = 64;
  else if (a.cout % 32 == 0) p.NT = 32;
  else p.NT = 16;
  if (a.scatter_cout && getenv("B200SEG_SCATTER_NT")) {     // experiment knob (probes/convt_knobs.py)
    const int v = atoi(getenv("B200SEG_SCATTER_NT"));
    if (v >= 16 && v <= 256 && a.cout % v == 0) p.NT = v;
  }
  if (p.NT % 16) return false;
  p.n_ntiles = a.cout / p.NT;
  p.WB = 8 + halo;
  p.HB = 16 + halo;
  if (p.WB > 256 || p.HB > 256) return false;
  p.slotA = (static_cast<unsigned>(p.WB * p.HB) * p.rowbytes + 1023) & ~1023u;
  p.bytesA = static_cast<unsigned>(p.WB * p.HB) * p.rowbytes;
  p.slotB = (p.NT * p.rowbytes + 1023) & ~1023u;
  p.bytesB = p.NT * p.rowbytes;
  const size_t fixed_small = 2048 + static_cast<size_t>(a.cout) * sizeof(float) * (a.stats ? 3 : 1) + 1024;
  // accumulators per stage: as many d-planes as fit 256 TMEM columns (<= 8), not more than the depth needs.
  // (A single 512-column stage with twice the planes per tile halves the weight traffic of K-heavy layers but also halves
  // their tile count; measured on the 32^3 / 16^3 U-Net layers it loses -- they are short of tiles, not of L2 bandwidth --
  // so it stays off unless B200SEG_SINGLE_STAGE is set.)
  const long long mmas_per_acc = static_cast<long long>(p.nchunks) * a.k * a.k * a.k * KS;
  p.stages = (getenv("B200SEG_SINGLE_STAGE") && mmas_per_acc >= 256 && p.NT >= 64) ? 1 : 2;
  int pmax = std::min(8, (p.stages == 1 ? 512 : kStageCols) / p.NT);
  while (pmax & (pmax - 1)) pmax &= pmax - 1;      // power of two (N tiles of 48, 80, ... channels: 5 -> 4, 3 -> 2)
  while (pmax > 1 && pmax / 2 >= a.od) pmax /= 2;
  // Wave quantisation: tiles are dealt to 148 persistent CTAs, so the kernel takes ceil(tiles / 148) rounds of P planes
  // each.  Small volumes (16^3: 64 tiles at P = 2) leave SMs idle; halving P doubles the tile count for free.
  {
    const long long per_plane = static_cast<long long>(a.n) * ((a.oh + 15) / 16) * ((a.ow + 7) / 8) * p.n_ntiles;
    long long best_cost = -1;
    int best_p = pmax;
    // (only when the grid is short of tiles: with many rounds the larger P wins through its smaller d halo)
    const bool starved = per_plane * ((a.od + pmax - 1) / pmax) < 2LL * kNumSMs;
    for (int cand = pmax; cand >= 1 && starved; cand /= 2) {
      const long long tiles = per_plane * ((a.od + cand - 1) / cand);
      const long long cost = ((tiles + kNumSMs - 1) / kNumSMs) * cand;
      if (best_cost < 0 || cost < best_cost) best_cost = cost, best_p = cand;   // ties keep the larger P (less weight traffic)
    }
    pmax = best_p;
  }
  const size_t budget = 225 * 1024;
  P = 0;
  for (int cand = pmax; cand >= 1 && !P; cand /= 2) {
    const int U = cand + halo;
    // weight ring: enough tiles in flight to cover the TMA round trip, within what the planes leave over.  With one or
    // two planes per tile a weight tile lasts only 0.14-0.3 us of MMAs (the 16^3 layers: measured weight-latency-bound
    // with 5 slots), so there the ring gets its 8 slots before the input ring gets spare planes.
    const int want_nb = cand <= 2 ? 8 : 3;
    int best_nb = 0, best_extra = 0;
    for (int extra = 3; extra >= 1; --extra) {
      const size_t a_bytes = static_cast<size_t>(U + extra) * p.slotA;
      if (a_bytes + fixed_small + 2 * p.slotB > budget) continue;
      const int nb = static_cast<int>(std::min<size_t>(8, (budget - a_bytes - fixed_small) / p.slotB));
      if (nb < 2) continue;
      if (nb > best_nb) best_nb = nb, best_extra = extra;
      if (nb >= want_nb) break;
    }
    if (best_nb) {
      P = cand;
      p.U = U;
      p.S = U + best_extra;
      p.NB = best_nb;
    }
  }
  if (!P) return false;
  // Weight-ring grouping (few accumulators per tile only: with P x KS >= 8 MMAs per tile the hand-shake is amortised)
  p.G = 1;
  if (a.k > 1 && !a.gather2 && !a.tapmode && P * KS <= 4 && p.NB / a.k >= 2) {
    p.G = a.k;
    p.NB = (p.NB / a.k) * a.k;
  }
  p.tiles_w = (a.ow + 7) / 8;
  p.tiles_h = (a.oh + 15) / 16;
  p.tiles_d = (a.od + P - 1) / P;
  p.tiles = static_cast<long long>(a.n) * p.tiles_d * p.tiles_h * p.tiles_w * p.n_ntiles;
  smem_bytes = static_cast<size_t>(p.S) * p.slotA + static_cast<size_t>(p.NB) * p.slotB + fixed_small;
  return smem_bytes <= 227 * 1024;
}

// Widest K chunk whose input-plane ring fits shared memory: large halos (dilation 4: 16 x 24-voxel planes, 12 of them per
// tile; 5x5x5 kernels) need narrower rows (highresnet.py:62-84 runs 64-channel layers at dilation 4).
static bool plan_plane(const UmmaConvArgs& a, PlaneParams& p, int& P, int& KS, size_t& smem_bytes, bool relaxed = false) {
  for (int cap : {64, 32, 16})
    if (plan_plane_kc(a, p, P, KS, smem_bytes, cap, relaxed)) return true;
  return false;
}

bool conv_umma_plane_supported(const UmmaConvArgs& a) {
  PlaneParams p;
  int P, KS;
  size_t smem;
  return plan_plane(a, p, P, KS, smem);
}
bool conv_umma_plane_relaxed_supported(const UmmaConvArgs& a) {
  PlaneParams p;
  int P, KS;
  size_t smem;
  return plan_plane(a, p, P, KS, smem, true);
}

template <int P, int KS, bool TAP>
static int launch_plane(const TensorMaps8P& tmAs, const CUtensorMap& tmB, const PlaneParams& p, size_t smem, int ctas,
                        cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_umma_plane_kernel<P, KS, TAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess) {
      set_error("conv_umma_plane: cannot raise the dynamic shared memory limit");
      return B200SEG_ERR_CUDA;
    }
    attr_set = true;
  }
  conv_umma_plane_kernel<P, KS, TAP><<<ctas, kThreadsP, smem, st>>>(tmAs, tmB, p);
  B200_CHECK_LAUNCH("conv_umma_plane");
  return 0;
}

int conv_umma_plane_run(const UmmaConvArgs& a, cudaStream_t st) {
  PlaneParams p;
  int P, KS;
  size_t smem;
  if (!plan_plane(a, p, P, KS, smem) && !plan_plane(a, p, P, KS, smem, true)) {
    set_error("conv_umma_plane_run: unsupported geometry");
    return B200SEG_ERR_INVALID;
  }
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.bias = a.bias;
  p.stats = a.stats;
  p.scale = a.scale;
  p.act = a.act;
  p.slope = a.slope;
  if (a.scale && (a.scatter_cout || a.stats)) {
    set_error("conv_umma_plane_run: the fused scale/activation epilogue excludes the pixel-shuffle and statistics epilogues");
    return B200SEG_ERR_INVALID;
  }
  // pixel-shuffle epilogue: column chunk c0 lands at channel c0 % scatter_cout of another voxel -> scatter_cout % 16 too
  p.wide = (a.out_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 31) == 0 &&
            (!a.scatter_cout || a.scatter_cout % 16 == 0) && !getenv("B200SEG_NO_WIDE_STORES")) ? 1 : 0;
  if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.wpack)) & 15) {
    set_error("conv_umma_plane_run: buffers must be 16-byte aligned");
    return B200SEG_ERR_INVALID;
  }
  TensorMaps8P tmAs;
  CUtensorMap tmB;
  const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB), 1u, 1u};
  if (a.tapmode) {
    p.cls = a.cls;
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) p.tapw[i][j] = a.tapw[i][j];
    if (a.out_sub) {
      // `out` = sub-lattice `cls` of the fine grid [n, fd, fh, fw, cout]
      const long long pit = a.out_pitch;
      p.osw = 2 * pit;
      p.osh = 2 * pit * a.fw;
      p.osd = 2 * pit * a.fw * a.fh;
      p.osn = pit * a.fw * a.fh * a.fd;
      p.out += ((static_cast<long long>(a.cls >> 2) * a.fh + ((a.cls >> 1) & 1)) * a.fw + (a.cls & 1)) * pit;
      p.wide = 0;
      if (a.out_pitch % 8) return B200SEG_ERR_INVALID;
    } else {
      p.osw = a.out_pitch;
      p.osh = p.osw * a.ow;
      p.osd = p.osh * a.oh;
      p.osn = p.osd * a.od;
    }
  }
  if (a.in_sub) {
    // class g = parity bits (d<<2 | h<<1 | w) of the fine grid `in` [n, fd, fh, fw, cin/8]: extent ceil((f - r) / 2)
    const int cm = a.cin / 8;
    const uint64_t pb = static_cast<uint64_t>(a.in_pitch) * 2;
    for (int g = 0; g < 8; ++g) {
      const int rd = g >> 2, rh = (g >> 1) & 1, rw = g & 1;
      const uint64_t off = ((static_cast<uint64_t>(rd) * a.fh + rh) * a.fw + rw) * pb;
      const uint64_t dims[5] = {static_cast<uint64_t>(cm), static_cast<uint64_t>((a.fw - rw + 1) / 2),
                                static_cast<uint64_t>((a.fh - rh + 1) / 2), static_cast<uint64_t>((a.fd - rd + 1) / 2),
                                static_cast<uint64_t>(a.n)};
      if (dims[1] == 0 || dims[2] == 0 || dims[3] == 0) return B200SEG_ERR_INVALID;
      const uint64_t str[4] = {2 * pb, 2 * pb * a.fw, 2 * pb * a.fw * a.fh, pb * a.fw * a.fh * a.fd};
      if (!encode_bf16_map(&tmAs.m[g], static_cast<const uint8_t*>(a.in) + off, 5, dims, str, box, p.KC))
        return B200SEG_ERR_CUDA;
    }
  } else if (!a.gather2) {
    const uint64_t dims[5] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.w), static_cast<uint64_t>(a.h),
                              static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.n)};
    const uint64_t pb = static_cast<uint64_t>(a.in_pitch) * 2;
    const uint64_t str[4] = {pb, pb * a.w, pb * a.w * a.h, pb * a.w * a.h * a.d};
    if (!encode_bf16_map(&tmAs.m[0], a.in, 5, dims, str, box, p.KC)) return B200SEG_ERR_CUDA;
    for (int i = 1; i < 8; ++i) tmAs.m[i] = tmAs.m[0];
  } else {
    const int cm = a.cin / 8;
    const uint64_t pb = static_cast<uint64_t>(a.in_pitch) * 2;
    const uint64_t W2 = 2ull * a.w, H2 = 2ull * a.h, D2 = 2ull * a.d;
    for (int g = 0; g < 8; ++g) {
      const uint64_t off = (((g >> 2) * H2 + ((g >> 1) & 1)) * W2 + (g & 1)) * pb;
      const uint64_t dims[5] = {static_cast<uint64_t>(cm), static_cast<uint64_t>(a.w), static_cast<uint64_t>(a.h),
                                static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.n)};
      const uint64_t str[4] = {2 * pb, 2 * pb * W2, 2 * pb * W2 * H2, pb * W2 * H2 * D2};
      if (!encode_bf16_map(&tmAs.m[g], static_cast<const uint8_t*>(a.in) + off, 5, dims, str, box, p.KC))
        return B200SEG_ERR_CUDA;
    }
  }
  {
    const int k3 = a.tapmode ? a.wtaps : (a.gather2 ? 8 : a.k * a.k * a.k);
    const uint64_t cin_w = (a.gather2 || a.in_sub) ? a.cin / 8 : a.cin;
    const uint64_t dims[3] = {cin_w, static_cast<uint64_t>(a.cout), static_cast<uint64_t>(k3)};
    const uint64_t str[2] = {cin_w * 2, cin_w * a.cout * 2};
    const uint32_t boxb[3] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.NT), 1u};
    if (!encode_bf16_map(&tmB, a.wpack, 3, dims, str, boxb, p.KC)) return B200SEG_ERR_CUDA;
  }
  const int ctas = static_cast<int>(std::min<long long>(kNumSMs, p.tiles));
  int rc = B200SEG_ERR_INVALID;
#define B200_PLANE_CASE(PP, KK)                                                                 \
  if (P == PP && KS == KK)                                                                      \
    rc = a.tapmode ? launch_plane<PP, KK, true>(tmAs, tmB, p, smem, ctas, st)                   \
                   : launch_plane<PP, KK, false>(tmAs, tmB, p, smem, ctas, st);
  B200_PLANE_CASE(8, 1) B200_PLANE_CASE(8, 2) B200_PLANE_CASE(8, 4)
  B200_PLANE_CASE(4, 1) B200_PLANE_CASE(4, 2) B200_PLANE_CASE(4, 4)
  B200_PLANE_CASE(2, 1) B200_PLANE_CASE(2, 2) B200_PLANE_CASE(2, 4)
  B200_PLANE_CASE(1, 1) B200_PLANE_CASE(1, 2) B200_PLANE_CASE(1, 4)
#undef B200_PLANE_CASE
  if (rc == B200SEG_ERR_INVALID) set_error("conv_umma_plane_run: no kernel instance for P = %d, KS = %d", P, KS);
  if (rc == 0) ++g_umma_launches;
  return rc;
}

}  // namespace b200
