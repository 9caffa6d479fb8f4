// CTA-pair (tcgen05 cta_group::2) variant of the persistent plane-mode convolution: the K-heavy layers of the deeper U-Net
// levels (C_in >= 64 at 32^3 / 16^3: enc3/enc4/dec4/dec3 forward and data gradient).
//
// Why: those layers are bound by the L2 -> SM weight stream, not by the tensor pipe (ncu, 256 -> 256 at 16^3: 262 MB of
// L2 -> L1 traffic per launch, 86 % of it weights, tensor pipe 42 % active): every 128-voxel tile streams its whole
// C_in x NT x 27 weight slice, and the chip-wide L2 read path saturates near 6.5 TB/s.  Two CTAs of a cluster on the two SMs
// of a TPC now work on two neighbouring voxel tiles (same rows, adjacent 8-wide column blocks) against ONE copy of the
// weight tile: each CTA fetches half of its N rows, and a single tcgen05.mma.cta_group::2 (M = 256: 128 voxels from each
// CTA's shared memory, N = NT: NT/2 weight rows from each) issued by the even CTA feeds both SMs' tensor cores.  Per SM the
// weight traffic halves and so does the number of MMA instructions.
//
// Protocol (rank 0 = leader):
//   * both CTAs run the input-plane and weight producers for their own tile / their half of the weight rows; every TMA is
//     the .cta_group::2 form whose completion bytes land on the LEADER's full barrier (count 1: the leader's
//     arrive.expect_tx armed with both CTAs' bytes; the peer's slot reuse is ordered by the multicast slot release);
//   * only the leader's MMA warp issues; every tcgen05.commit is multicast to the barrier at the same offset in both CTAs
//     (slot release for both producers, accumulator-full for both epilogues);
//   * both CTAs' epilogue warps drain their own TMEM half and arrive on the leader's accumulator-empty barrier.
// Everything else (halo'd planes as shifted UMMA views, plane ring, two accumulator stages, statistics / bias / fused
// scale-shift-activation epilogue) is conv_umma_p.cu's scheme.
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

struct PairParams {
  int n, od, oh, ow, cout;
  long long out_pitch;
  int k, pad, dil;
  int KC, nchunks, NT, n_ntiles;
  int WB, HB, U, S, NB;
  int wide;
  int tiles_wp, tiles_h, tiles_d;     // tiles_wp = PAIRS of 8-wide column blocks
  long long tiles;                    // pair tiles
  unsigned slotA, slotB, rowbytes, swz, bytesA, bytesB;   // slotB / bytesB: HALF a weight tile (NT/2 rows)
  __nv_bfloat16* out;
  const float* bias;
  float* stats;
  const float* scale;
  int act;
  float slope;
  long long* dbg;     // B200SEG_PAIR_TIMELINE: cycles the leader's MMA warp of cluster 0 spent waiting, by barrier kind
};

constexpr int kEpiWarps2 = 8;
constexpr int kThreadsP2 = 32 * (4 + kEpiWarps2);
constexpr int kStageCols2 = 256;

struct PairTile {
  int nt, w0, h0, d0, nn;
};

__device__ __forceinline__ PairTile decode_pair(const PairParams& p, long long t, int P, int rank) {
  PairTile c;
  c.nt = static_cast<int>(t % p.n_ntiles);
  t /= p.n_ntiles;
  c.w0 = (static_cast<int>(t % p.tiles_wp) * 2 + rank) * 8;
  t /= p.tiles_wp;
  c.h0 = static_cast<int>(t % p.tiles_h) * 16;
  t /= p.tiles_h;
  c.d0 = static_cast<int>(t % p.tiles_d) * P;
  c.nn = static_cast<int>(t / p.tiles_d);
  return c;
}

template <int CW>
__device__ __forceinline__ void warp_colsum2(float (&v)[CW], int lane) {
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

template <int P, int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsP2, 1)
    conv_umma_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PairParams p) {
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
  const int k = p.k;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;

  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) {
      mbar_init(&fullA[i], 1);      // the leader's arrive.expect_tx, armed with BOTH CTAs' bytes (the peer only issues its TMA)
      mbar_init(&emptyA[i], 1);     // one multicast tcgen05.commit per release
    }
    for (int i = 0; i < p.NB; ++i) {
      mbar_init(&fullB[i], 1);
      mbar_init(&emptyB[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&accFull[i], 1);
      mbar_init(&accEmpty[i], 2 * kEpiWarps2);   // the epilogue warps of BOTH CTAs arrive on the leader's barrier
    }
    fence_mbar_init();
  }
  for (int i = tid; i < p.cout; i += kThreadsP2) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  if (p.stats != nullptr)
    for (int i = tid; i < 2 * p.cout; i += kThreadsP2) s_stats[i] = 0.f;
  if (warp == 3) {
    tmem_alloc_2sm(tmem_ptr, 512);
    tmem_relinquish_2sm();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
  if (warp == 1 && lane == 0) tma_prefetch_desc(&tmB);
  tc_fence_before();
  cluster_sync_all();               // barriers of both CTAs are initialised before anybody signals across
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  const long long first = blockIdx.x >> 1, step = gridDim.x >> 1;

  if (warp == 0) {
    // =========================== input-plane producer (both CTAs, own tile) ===========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long t = first; t < p.tiles; t += step) {
        const PairTile tc = decode_pair(p, t, P, rank);
        for (int c = 0; c < p.nchunks; ++c) {
          for (int u = 0; u < p.U; ++u) {
            mbar_wait(&emptyA[s], ph ^ 1);
            // (no per-load remote arrive from the peer: measured ~880 cycles each on the producer thread, which made the
            //  weight stream 4x slower than the single-CTA kernel's)
            if (is_leader) mbar_arrive_expect_tx(&fullA[s], 2 * p.bytesA);
            tma_load_5d_2sm(sA + static_cast<size_t>(s) * p.slotA, &tmA, &fullA[s], c * p.KC, tc.w0 - p.pad, tc.h0 - p.pad,
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
    // =========================== weight producer (both CTAs, half of the N rows each) ===========================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int k3 = k * k * k;
      const int half = p.NT / 2;
      for (long long t = first; t < p.tiles; t += step) {
        const int nt = static_cast<int>(t % p.n_ntiles);
        for (int c = 0; c < p.nchunks; ++c) {
          for (int tap = 0; tap < k3; ++tap) {
            mbar_wait(&emptyB[s], ph ^ 1);
            if (is_leader) mbar_arrive_expect_tx(&fullB[s], 2 * p.bytesB);
            tma_load_3d_2sm(sB + static_cast<size_t>(s) * p.slotB, &tmB, &fullB[s], c * p.KC, nt * p.NT + static_cast<int>(rank) * half,
                            tap);
            if (++s == p.NB) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (is_leader) {
      const uint32_t leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(256, p.NT, 0, 0);
      const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 16, static_cast<uint32_t>(p.WB) * p.rowbytes, p.swz) >> 32);
      const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 16, 8u * p.rowbytes, p.swz) >> 32);
      const uint32_t lo_fixed = 1u << 16;
      const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
      const uint32_t slotA16 = p.slotA >> 4, slotB16 = p.slotB >> 4;
      const uint32_t row16 = p.rowbytes >> 4;
      int bs = 0;
      uint32_t bphase = 0;
      int unit0 = 0;
      uint32_t unit0_phase = 0;
      uint32_t it = 0;
      const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
      long long wA = 0, wB = 0, wE = 0, t_begin = clock64();
      for (long long t = first; t < p.tiles; t += step, ++it) {
        const uint32_t stage = it & 1, use = it >> 1;
        long long c0 = dbg ? clock64() : 0;
        mbar_wait(&accEmpty[stage], (use & 1) ^ 1);
        if (dbg) wE += clock64() - c0;
        tc_fence_after();
        const uint32_t d_base = tbase + stage * kStageCols2;
        for (int c = 0; c < p.nchunks; ++c) {
          int waited = 0, wslot = unit0;
          uint32_t wphase = unit0_phase;
          for (int a = 0; a < k; ++a) {
            const int need = min(p.U, P + a * p.dil);
            c0 = dbg ? clock64() : 0;
            for (; waited < need; ++waited) {
              mbar_wait(&fullA[wslot], wphase);
              if (++wslot == p.S) {
                wslot = 0;
                wphase ^= 1;
              }
            }
            if (dbg) wA += clock64() - c0;
            tc_fence_after();
            uint32_t a_lo[P];
            {
              int sl = unit0 + a * p.dil;
              if (sl >= p.S) sl -= p.S;
#pragma unroll
              for (int acc = 0; acc < P; ++acc) {
                a_lo[acc] = __shfl_sync(0xffffffffu, ((sA16 + sl * slotA16) & 0x3FFF) | lo_fixed, 0);
                if (++sl == p.S) sl = 0;
              }
            }
            uint32_t tap16_row = 0;
            for (int b = 0; b < k; ++b) {
              uint32_t tap16 = tap16_row;
              for (int e = 0; e < k; ++e) {
                c0 = dbg ? clock64() : 0;
                mbar_wait(&fullB[bs], bphase);
                if (dbg) wB += clock64() - c0;
                tc_fence_after();
                const uint32_t b_lo = __shfl_sync(0xffffffffu, ((sB16 + bs * slotB16) & 0x3FFF) | lo_fixed, 0);
                const uint32_t fresh = (c | a | b | e) == 0 ? 0u : 1u;
#pragma unroll
                for (int acc = 0; acc < P; ++acc) {
#pragma unroll
                  for (int kk = 0; kk < KS; ++kk) {
                    umma_f16_2sm_pred_lohi(d_base + acc * p.NT, a_lo[acc] + tap16 + 2u * kk, a_hi, b_lo + 2u * kk, b_hi, idesc,
                                           kk == 0 ? fresh : 1u, leader);
                  }
                }
                umma_commit_2sm_pred(&emptyB[bs], leader);
                if (++bs == p.NB) {
                  bs = 0;
                  bphase ^= 1;
                }
                tap16 += p.dil * row16;
              }
              tap16_row += p.dil * p.WB * row16;
            }
            {
              int rs = unit0;
              for (int j = 0; j < p.U; ++j) {
                if (min(k - 1, j / p.dil) == a) umma_commit_2sm_pred(&emptyA[rs], leader);
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
        umma_commit_2sm_pred(&accFull[stage], leader);
      }
      if (dbg && lane == 0) {
        p.dbg[0] = clock64() - t_begin;
        p.dbg[1] = wA;
        p.dbg[2] = wB;
        p.dbg[3] = wE;
        p.dbg[4] = it;
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue (both CTAs, own TMEM half) ===========================
    const int q = warp & 3;
    const int half = (warp >> 2) & 1;
    const int m = q * 32 + lane;
    const bool want_stats = p.stats != nullptr;
    uint32_t it = 0;
    for (long long t = first; t < p.tiles; t += step, ++it) {
      const uint32_t stage = it & 1, use = it >> 1;
      const PairTile tc = decode_pair(p, t, P, rank);
      mbar_wait(&accFull[stage], use & 1);
      tc_fence_after();
      const int oh_ = tc.h0 + (m >> 3), ow_ = tc.w0 + (m & 7);
      const bool hw_ok = oh_ < p.oh && ow_ < p.ow;
      const uint32_t t_lane = tbase + stage * kStageCols2 + (static_cast<uint32_t>(q * 32) << 16);
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
            const long long vox = ((static_cast<long long>(tc.nn) * p.od + od_) * p.oh + oh_) * p.ow + ow_;
            __nv_bfloat16* optr = p.out + vox * p.out_pitch + col0;
            uint32_t raw[CW];
            if constexpr (CW == 32) tmem_ld_32x32(t_lane + acc * p.NT + c0, raw);
            else tmem_ld_32x16(t_lane + acc * p.NT + c0, raw);
            tmem_ld_wait();
            if (valid) {
              float v[CW];
              if (p.scale != nullptr) {
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                  const float z = fmaf(__uint_as_float(raw[j]), __ldg(p.scale + col0 + j), s_bias[col0 + j]);
                  v[j] = p.act == B200SEG_ACT_NONE ? z : (z > 0.f ? z : (p.act == B200SEG_ACT_RELU ? 0.f : p.slope * z));
                }
              } else {
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(raw[j]) + s_bias[col0 + j];
              }
              if (want_stats) {
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                  s1[j] += v[j];
                  s2[j] = fmaf(v[j], v[j], s2[j]);
                }
              }
              if (p.wide) {
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
          if (want_stats) {
            warp_colsum2<CW>(s1, lane);
            warp_colsum2<CW>(s2, lane);
            if (lane < CW) {
              atomicAdd(&s_stats[col0 + lane], s1[0]);
              atomicAdd(&s_stats[p.cout + col0 + lane], s2[0]);
            }
          }
        }
      };
      if ((p.NT & 31) == 0) slabs(std::integral_constant<int, 32>{});
      else slabs(std::integral_constant<int, 16>{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (is_leader) mbar_arrive(&accEmpty[stage]);
        else mbar_arrive_remote(&accEmpty[stage], 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();      // nobody leaves (or frees TMEM) while the peer may still signal its barriers / use the pair's TMEM
  tc_fence_after();
  if (p.stats != nullptr) {
    for (int i = tid; i < 2 * p.cout; i += kThreadsP2) {
      const float v = s_stats[i];
      if (v != 0.f) atomicAdd(&p.stats[i], v);
    }
  }
  if (warp == 3) tmem_dealloc_2sm(tbase, 512);
}

// ------------------------------------------------------------------------------------------------ host side
static bool plan_pair_kc(const UmmaConvArgs& a, PairParams& p, int& P, int& KS, size_t& smem_bytes, int kc_cap) {
  if (getenv("B200SEG_DISABLE_UMMA") || getenv("B200SEG_DISABLE_PERSISTENT") || getenv("B200SEG_DISABLE_PAIR")) return false;
  if (a.tapmode || a.scatter_cout || a.gather2 || a.in_sub || a.out_sub) return false;
  if (a.cin % 16 || a.cout % 16) return false;
  if (a.in_pitch % 8 || a.out_pitch % 8) return false;
  if (!(a.k == 1 || a.k == 3 || a.k == 5) || a.dil < 1 || a.pad < 0) return false;
  const int halo = (a.k - 1) * a.dil;
  if (a.od != a.d + 2 * a.pad - halo || a.oh != a.h + 2 * a.pad - halo || a.ow != a.w + 2 * a.pad - halo) return false;
  if (!(a.oh >= 16 && a.ow >= 16 && a.ow % 16 == 0)) return false;          // pairs of 8-wide column blocks
  // K-heavy layers: that is where the weight stream, not the tensor pipe, is the bound.  Narrow outputs (C_out 16 / 32) only
  // for 5x5x5 kernels (vnet3d.py:25), which the kd-stacking roll kernel does not take: an M = 128 MMA costs ~48 cycles
  // however small N is, so one M = 256 instruction per SM pair halves the issue-bound time of those layers.
  if (static_cast<long long>(a.cin) * a.k * a.k * a.k < 27 * 64) return false;
  if (a.cout < 64 && a.k != 5) return false;
  p = PairParams{};
  p.n = a.n; p.od = a.od; p.oh = a.oh; p.ow = a.ow; p.cout = a.cout; p.out_pitch = a.out_pitch;
  p.k = a.k; p.pad = a.pad; p.dil = a.dil;
  p.KC = (a.cin % 64 == 0 && kc_cap >= 64) ? 64 : ((a.cin % 32 == 0 && kc_cap >= 32) ? 32 : 16);
  KS = p.KC / 16;
  p.nchunks = a.cin / p.KC;
  p.rowbytes = p.KC * 2;
  p.swz = p.KC == 64 ? SWZ_128B : (p.KC == 32 ? SWZ_64B : SWZ_32B);
  if (a.cout <= 128) p.NT = a.cout;
  else if (a.cout % 128 == 0) p.NT = 128;
  else if (a.cout % 64 == 0) p.NT = 64;
  else if (a.cout % 32 == 0) p.NT = 32;
  else return false;
  p.n_ntiles = a.cout / p.NT;
  p.WB = 8 + halo;
  p.HB = 16 + halo;
  if (p.WB > 256 || p.HB > 256) return false;
  p.slotA = (static_cast<unsigned>(p.WB * p.HB) * p.rowbytes + 1023) & ~1023u;
  p.bytesA = static_cast<unsigned>(p.WB * p.HB) * p.rowbytes;
  p.bytesB = (p.NT / 2) * p.rowbytes;
  p.slotB = (p.bytesB + 1023) & ~1023u;
  const size_t fixed_small = 2048 + static_cast<size_t>(a.cout) * sizeof(float) * (a.stats ? 3 : 1) + 1024;
  int pmax = std::min(4, kStageCols2 / p.NT);     // kernel instances exist for P = 4, 2, 1
  while (pmax & (pmax - 1)) pmax &= pmax - 1;
  while (pmax > 1 && pmax / 2 >= a.od) pmax /= 2;
  {
    // wave quantisation over the 74 CTA pairs
    const long long per_plane = static_cast<long long>(a.n) * ((a.oh + 15) / 16) * (a.ow / 16) * p.n_ntiles;
    const int pairs = kNumSMs / 2;
    long long best_cost = -1;
    int best_p = pmax;
    const bool starved = per_plane * ((a.od + pmax - 1) / pmax) < 2LL * pairs;
    for (int cand = pmax; cand >= 1 && starved; cand /= 2) {
      const long long tiles = per_plane * ((a.od + cand - 1) / cand);
      const long long cost = ((tiles + pairs - 1) / pairs) * cand;
      if (best_cost < 0 || cost < best_cost) best_cost = cost, best_p = cand;
    }
    pmax = best_p;
  }
  const size_t budget = 225 * 1024;
  P = 0;
  for (int cand = pmax; cand >= 1 && !P; cand /= 2) {
    const int U = cand + halo;
    const int want_nb = cand <= 2 ? 10 : 4;
    int best_nb = 0, best_extra = 0;
    for (int extra = 3; extra >= 1; --extra) {
      const size_t a_bytes = static_cast<size_t>(U + extra) * p.slotA;
      if (a_bytes + fixed_small + 2 * p.slotB > budget) continue;
      const int nb = static_cast<int>(std::min<size_t>(12, (budget - a_bytes - fixed_small) / p.slotB));
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
  p.tiles_wp = a.ow / 16;
  p.tiles_h = (a.oh + 15) / 16;
  p.tiles_d = (a.od + P - 1) / P;
  p.tiles = static_cast<long long>(a.n) * p.tiles_d * p.tiles_h * p.tiles_wp * p.n_ntiles;
  smem_bytes = static_cast<size_t>(p.S) * p.slotA + static_cast<size_t>(p.NB) * p.slotB + fixed_small;
  return smem_bytes <= 227 * 1024;
}

static bool plan_pair(const UmmaConvArgs& a, PairParams& p, int& P, int& KS, size_t& smem_bytes) {
  for (int cap : {64, 32, 16})
    if (plan_pair_kc(a, p, P, KS, smem_bytes, cap)) return true;
  return false;
}

bool conv_umma_pair_supported(const UmmaConvArgs& a) {
  PairParams p;
  int P, KS;
  size_t smem;
  return plan_pair(a, p, P, KS, smem);
}

template <int P, int KS>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const PairParams& p, size_t smem, int ctas, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_umma_pair_kernel<P, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("conv_umma_pair: cannot raise the dynamic shared memory limit");
      return B200SEG_ERR_CUDA;
    }
    attr_set = true;
  }
  PairParams q = p;
  q.dbg = nullptr;
  if (getenv("B200SEG_PAIR_TIMELINE")) {
    cudaMalloc(&q.dbg, 8 * sizeof(long long));
    cudaMemset(q.dbg, 0, 8 * sizeof(long long));
  }
  conv_umma_pair_kernel<P, KS><<<ctas, kThreadsP2, smem, st>>>(tmA, tmB, q);   // __cluster_dims__(2,1,1): ctas is even
  B200_CHECK_LAUNCH("conv_umma_pair");
  if (q.dbg) {
    long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, q.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(q.dbg);
    fprintf(stderr, "[pair timeline] P %d KS %d NT %d KC %d chunks %d S %d U %d NB %d tiles %lld ctas %d | issuer total %lld cycles: "
            "wait A %lld, wait B %lld, wait accEmpty %lld over %lld tiles\n", P, KS, p.NT, p.KC, p.nchunks, p.S, p.U, p.NB, p.tiles,
            ctas, h[0], h[1], h[2], h[3], h[4]);
  }
  return 0;
}

int conv_umma_pair_run(const UmmaConvArgs& a, cudaStream_t st) {
  PairParams p;
  int P, KS;
  size_t smem;
  if (!plan_pair(a, p, P, KS, smem)) {
    set_error("conv_umma_pair_run: unsupported geometry");
    return B200SEG_ERR_INVALID;
  }
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.bias = a.bias;
  p.stats = a.stats;
  p.scale = a.scale;
  p.act = a.act;
  p.slope = a.slope;
  if (a.scale && a.stats) {
    set_error("conv_umma_pair_run: the fused scale/activation epilogue excludes the statistics epilogue");
    return B200SEG_ERR_INVALID;
  }
  p.wide = (a.out_pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 31) == 0 && !getenv("B200SEG_NO_WIDE_STORES")) ? 1 : 0;
  if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.wpack)) & 15) {
    set_error("conv_umma_pair_run: buffers must be 16-byte aligned");
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
    const uint64_t dims[3] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.cout), static_cast<uint64_t>(a.k * a.k * a.k)};
    const uint64_t str[2] = {static_cast<uint64_t>(a.cin) * 2, static_cast<uint64_t>(a.cin) * a.cout * 2};
    const uint32_t boxb[3] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.NT / 2), 1u};
    if (!encode_bf16_map(&tmB, a.wpack, 3, dims, str, boxb, p.KC)) return B200SEG_ERR_CUDA;
  }
  const int ctas = 2 * static_cast<int>(std::min<long long>(kNumSMs / 2, p.tiles));
  int rc = B200SEG_ERR_INVALID;
#define B200_PAIR_CASE(PP, KK) \
  if (P == PP && KS == KK) rc = launch_pair<PP, KK>(tmA, tmB, p, smem, ctas, st);
  B200_PAIR_CASE(4, 1) B200_PAIR_CASE(4, 2) B200_PAIR_CASE(4, 4)
  B200_PAIR_CASE(2, 1) B200_PAIR_CASE(2, 2) B200_PAIR_CASE(2, 4)
  B200_PAIR_CASE(1, 1) B200_PAIR_CASE(1, 2) B200_PAIR_CASE(1, 4)
#undef B200_PAIR_CASE
  if (rc == B200SEG_ERR_INVALID) set_error("conv_umma_pair_run: no kernel instance for P = %d, KS = %d", P, KS);
  if (rc == 0) ++g_umma_launches;
  return rc;
}

}  // namespace b200
