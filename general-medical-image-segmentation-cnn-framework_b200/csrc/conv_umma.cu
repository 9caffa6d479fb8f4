// tcgen05 implicit-GEMM 3-D convolution (stride 1, cubic kernel k in {1,3,5}, dilation, zero padding).
//
//   out[voxel][n] = bias[n] + sum_{tap,c} in[voxel - pad + tap*dil][c] * W[tap][n][c]          (fp32 accumulate)
//
// GEMM view: M = output voxels, N = C_out, K = k^3 * C_in.  One CTA owns P accumulators of 128 voxels x NT channels in
// TMEM.  The input is NOT re-fetched per tap: a halo'd box of the channels-last input is brought into shared memory
// once per K-chunk by TMA (out-of-bounds coordinates give the zero padding) and every tap reads a *shifted view* of
// it -- the UMMA shared-memory descriptor starts (tap offset) rows later, with the 8-row group stride set to the box
// row pitch.  The 128B/64B/32B swizzle is a function of the absolute shared-memory address on both the TMA write and
// the UMMA read (verified on hardware by probes/probe.cu), which is what makes unaligned tap views legal.
//
//   plane mode (H_out >= 16): tile = 8 (w) x 16 (h) voxels of one d-plane per accumulator, P consecutive d-planes per
//       CTA; input planes live in a ring of slots, each plane is loaded once and used by up to k accumulators.
//   flat  mode (small H/W):   the whole halo'd (d,h,w) box of a sample is one slot; accumulator p covers box rows
//       [128p, 128p+128) in flattened order (rows that fall in the halo are computed and discarded).
//
// Warp roles (256 threads): 0 = TMA producer for input boxes, 1 = TMA producer for weight tiles, 2 = MMA issuer,
// 3 = TMEM allocator, 4..7 = epilogue (TMEM -> registers -> +bias, per-channel sum / sum-of-squares -> bf16 -> HBM).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.h"
#include "ptx.cuh"

namespace b200 {

struct UmmaConvParams {
  int n, od, oh, ow, cout;
  long long out_pitch;
  int k, pad, dil;
  int KC, nchunks, NT, n_ntiles, P, flat;
  int G;                 // weight tiles per ring barrier (the k kw-taps of one (kd, kh) pair, or 1)
  int WB, HB, U, UP, S, NB, DT;
  int tiles_w, tiles_h, tiles_d;
  int scatter_cout, cpm;  // pixel-shuffle epilogue channel count (0 = off); K-chunks per input tensor map
  unsigned slotA, slotB, rowbytes, swz, bytesA_unit, bytesB, tmem_cols;
  __nv_bfloat16* out;
  const float* bias;
  float* stats;
  long long* dbg;  // optional timeline of CTA 0 (clock64 stamps), enabled by B200SEG_DEBUG_TIMELINE
};

struct TensorMaps8 {
  CUtensorMap m[8];
};

__device__ __forceinline__ void butterfly_colsum(float (&v)[32], int lane) {
  // after the call, v[0] on lane l holds sum over the 32 lanes of the original v[l]
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (up ? v[i + s] : v[i]) + recv;
    }
  }
}

__global__ void __launch_bounds__(256, 2)
    conv_umma_kernel(const __grid_constant__ TensorMaps8 tmAs, const __grid_constant__ CUtensorMap tmB,
                     const UmmaConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + static_cast<size_t>(p.S) * p.slotA;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + static_cast<size_t>(p.NB) * p.slotB);
  uint64_t* emptyA = fullA + p.S;
  uint64_t* fullB = emptyA + p.S;
  uint64_t* emptyB = fullB + p.NB;
  uint64_t* accFull = emptyB + p.NB;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accFull + 1);
  float* s_stats = reinterpret_cast<float*>(tmem_ptr + 2);  // [2][NT]
  float* s_bias = s_stats + 2 * p.NT;                        // [NT]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which tile ------------------------------------------------------------------------------------------
  int bid = blockIdx.x;
  const int nt = bid % p.n_ntiles;
  bid /= p.n_ntiles;
  const int tw = bid % p.tiles_w;
  bid /= p.tiles_w;
  const int th = bid % p.tiles_h;
  bid /= p.tiles_h;
  const int td = bid % p.tiles_d;
  const int nn = bid / p.tiles_d;
  const int w0 = p.flat ? 0 : tw * 8, h0 = p.flat ? 0 : th * 16, d0 = td * p.DT;
  const int k = p.k, k3 = k * k * k;

  if (tid == 0) {
    for (int i = 0; i < p.S; ++i) {
      mbar_init(&fullA[i], 1);
      mbar_init(&emptyA[i], 1);
    }
    for (int i = 0; i < p.NB; ++i) {
      mbar_init(&fullB[i], 1);
      mbar_init(&emptyB[i], 1);
    }
    mbar_init(accFull, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 2 * p.NT; i += blockDim.x) s_stats[i] = 0.f;
  for (int i = tid; i < p.NT; i += blockDim.x)
    s_bias[i] = p.bias ? p.bias[p.scatter_cout ? (nt * p.NT + i) % p.scatter_cout : nt * p.NT + i] : 0.f;
  if (warp == 3) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmAs.m[0]);
  if (warp == 1 && lane == 0) tma_prefetch_desc(&tmB);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
  if (dbg && tid == 0) p.dbg[0] = clock64();

  if (warp == 0) {
    // =========================== input-box producer ===========================
    if (lane == 0) {
      int L = 0;
      for (int c = 0; c < p.nchunks; ++c) {
        for (int u = 0; u < p.U; ++u, ++L) {
          const int s = L % p.S;
          const uint32_t ph = (L / p.S) & 1;
          mbar_wait(&emptyA[s], ph ^ 1);
          mbar_arrive_expect_tx(&fullA[s], p.bytesA_unit);
          tma_load_5d(sA + static_cast<size_t>(s) * p.slotA, &tmAs.m[c / p.cpm], &fullA[s], (c % p.cpm) * p.KC,
                      w0 - p.pad, h0 - p.pad, d0 - p.pad + u * p.UP, nn);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== weight-tile producer ===========================
    if (lane == 0) {
      // One full/empty barrier pair covers G consecutive weight tiles (G = k: the kw taps of one (kd, kh) pair): the
      // consumer then waits and commits once per G tiles (measured -12 % on the P = 1 bottleneck layers, which issue
      // only 4 MMAs per weight tile).
      const int ngroups = p.NB / p.G;
      int L = 0;
      for (int c = 0; c < p.nchunks; ++c) {
        for (int t = 0; t < k3; t += p.G, ++L) {
          const int s = L % ngroups;
          const uint32_t ph = (L / ngroups) & 1;
          mbar_wait(&emptyB[s], ph ^ 1);
          mbar_arrive_expect_tx(&fullB[s], p.G * p.bytesB);
          for (int e = 0; e < p.G; ++e)
            // multi-map (gather) mode: the map index doubles as the weight "tap"
            tma_load_3d(sB + static_cast<size_t>(s * p.G + e) * p.slotB, &tmB, &fullB[s], (c % p.cpm) * p.KC, nt * p.NT,
                        p.cpm < p.nchunks ? c / p.cpm : t + e);
        }
      }
    }
  } else if (warp == 2) {
    // =========================== MMA issuer ===========================
    // One thread issues every tcgen05.mma of the CTA, so this loop is written to cost a handful of integer
    // instructions per MMA: descriptors are (constant high word | running low word), ring slots advance by
    // add-and-wrap instead of modulo, and nothing is multiplied inside the accumulator loop.
    {
      const uint32_t leader = elect_one();  // the one lane whose tcgen05 instructions take effect
      const uint32_t idesc = make_idesc_bf16(128, p.NT, 0, 0);
      const uint32_t a_sbo = (p.flat ? 8u : static_cast<uint32_t>(p.WB)) * p.rowbytes;
      const uint32_t b_sbo = 8u * p.rowbytes;
      const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 16, a_sbo, p.swz) >> 32);
      const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 16, b_sbo, p.swz) >> 32);
      const uint32_t lo_fixed = 1u << 16;  // leading byte offset field (16 B), ignored for swizzled K-major
      const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
      const int ksteps = p.KC / 16;
      const uint32_t acc_stride_flat = 128u * p.rowbytes;
      int bs = 0;                 // weight ring slot
      uint32_t bphase = 0;
      int unit0_slot = 0;         // ring slot of unit 0 of the current chunk
      uint32_t unit0_phase = 0;   // its parity
      for (int c = 0; c < p.nchunks; ++c) {
        int waited = 0;           // units of this chunk whose full barrier has been observed
        int wslot = unit0_slot;
        uint32_t wphase = unit0_phase;
        for (int a = 0; a < k; ++a) {
          const int need = p.flat ? 1 : min(p.U, p.P + a * p.dil);
          for (; waited < need; ++waited) {
            mbar_wait(&fullA[wslot], wphase);
            if (++wslot == p.S) {
              wslot = 0;
              wphase ^= 1;
            }
          }
          tc_fence_after();
          if (dbg && c == 0 && lane == 0) p.dbg[1 + a] = clock64();
          // address of the slot holding plane (a*dil) of this chunk (plane mode) / of the box (flat mode)
          int aslot = unit0_slot + (p.flat ? 0 : a * p.dil);
          if (aslot >= p.S) aslot -= p.S;
          const uint32_t a_first = sA_addr + aslot * p.slotA;
          uint32_t a_lo_acc[8];
          {
            int sl = aslot;
#pragma unroll
            for (int acc = 0; acc < 8; ++acc) {
              a_lo_acc[acc] = (((sA_addr + sl * p.slotA) >> 4) & 0x3FFF) | lo_fixed;
              if (++sl == p.S) sl = 0;
            }
          }
          for (int b = 0; b < k; ++b) {
            for (int e = 0; e < k; ++e) {
              const int ge = p.G == 1 ? 0 : e;          // tile inside the group (groups are whole kw rows)
              if (ge == 0) {
                mbar_wait(&fullB[bs], bphase);
                tc_fence_after();
              }
              const uint32_t b_lo = (((sB_addr + (bs * p.G + ge) * p.slotB) >> 4) & 0x3FFF) | lo_fixed;
              const uint32_t tapoff = p.flat
                  ? static_cast<uint32_t>((a * p.dil * p.HB + b * p.dil) * p.WB + e * p.dil) * p.rowbytes
                  : static_cast<uint32_t>((b * p.dil) * p.WB + e * p.dil) * p.rowbytes;
              const uint32_t accum0 = (c | a | b | e) != 0 ? 1u : 0u;
              if (!p.flat) {
                // plane mode: P <= 8 accumulators, descriptors precomputed per kd iteration (a_lo_acc), fully unrolled
                const uint32_t tap16 = tapoff >> 4;
#pragma unroll
                for (int acc = 0; acc < 8; ++acc) {
                  if (acc < p.P) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                      if (kk < ksteps) {
                        const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo_acc[acc] + tap16 + 2u * kk);
                        const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 2u * kk);
                        umma_f16_pred(tbase + acc * p.NT, ad, bd, idesc, accum0 | (kk != 0 ? 1u : 0u), leader);
                      }
                    }
                  }
                }
              } else {
                // flat mode: these are the K-heavy 8^3 layers, where the single issuing thread was the bottleneck (576
                // cycles per weight tile for 4 MMAs of 55): bases broadcast from lane 0 (uniform datapath), descriptors
                // as (lo, hi) halves, K steps unrolled
                uint32_t a_lo = __shfl_sync(0xffffffffu, (((a_first + tapoff) >> 4) & 0x3FFF) | lo_fixed, 0);
                const uint32_t b_lo_u = __shfl_sync(0xffffffffu, b_lo, 0);
                uint32_t d_tmem = tbase;
                for (int acc = 0; acc < p.P; ++acc) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) {
                    if (kk < ksteps)
                      umma_f16_pred_lohi(d_tmem, a_lo + 2u * kk, a_hi, b_lo_u + 2u * kk, b_hi, idesc,
                                         accum0 | (kk != 0 ? 1u : 0u), leader);
                  }
                  d_tmem += p.NT;
                  a_lo += acc_stride_flat >> 4;
                }
              }
              if (ge == p.G - 1) {
                umma_commit_pred(&emptyB[bs], leader);  // weight tiles of the group consumed once these MMAs retire
                if (++bs == p.NB / p.G) {
                  bs = 0;
                  bphase ^= 1;
                }
              }
            }
          }
          // release input units whose last use was this kd iteration
          if (p.flat) {
            if (a == k - 1) umma_commit_pred(&emptyA[unit0_slot], leader);
          } else {
            int rs = unit0_slot;
            for (int j = 0; j < p.U; ++j) {
              const int a_last = min(k - 1, j / p.dil);
              if (a_last == a) umma_commit_pred(&emptyA[rs], leader);
              if (++rs == p.S) rs = 0;
            }
          }
        }
        unit0_slot += p.U;
        while (unit0_slot >= p.S) {
          unit0_slot -= p.S;
          unit0_phase ^= 1;
        }
      }
      umma_commit_pred(accFull, leader);
      if (dbg && lane == 0) p.dbg[8] = clock64();
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===========================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(accFull, 0);
    tc_fence_after();
    if (dbg && tid == 128) p.dbg[9] = clock64();
    const int plane_rows = p.HB * p.WB;
    for (int c16 = 0; c16 < p.NT; c16 += 16) {
      float bias_r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bias_r[j] = s_bias[c16 + j];
      float s1[16], s2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) s1[j] = s2[j] = 0.f;
      for (int acc = 0; acc < p.P; ++acc) {
        int od_, oh_, ow_;
        bool valid;
        if (p.flat) {
          const int R = acc * 128 + m;
          const int dz = R / plane_rows, rem = R - dz * plane_rows;
          const int hy = rem / p.WB, wx = rem - hy * p.WB;
          od_ = d0 + dz;
          oh_ = hy;
          ow_ = wx;
          valid = dz < p.DT && od_ < p.od && hy < p.oh && wx < p.ow;
        } else {
          od_ = d0 + acc;
          oh_ = h0 + (m >> 3);
          ow_ = w0 + (m & 7);
          valid = od_ < p.od && oh_ < p.oh && ow_ < p.ow;
        }
        __nv_bfloat16* optr;
        if (p.scatter_cout) {
          // 2x2x2 pixel shuffle: this 16-column chunk belongs to one (a,b,e) offset of the up-sampled grid
          const int col0 = nt * p.NT + c16;
          const int abe = 7 - col0 / p.scatter_cout, co0 = col0 % p.scatter_cout;
          const long long ovox = ((static_cast<long long>(nn) * 2 * p.od + 2 * od_ + (abe >> 2)) * 2 * p.oh + 2 * oh_ +
                                  ((abe >> 1) & 1)) * 2 * p.ow + 2 * ow_ + (abe & 1);
          optr = p.out + ovox * p.out_pitch + co0;
        } else {
          const long long vox = ((static_cast<long long>(nn) * p.od + od_) * p.oh + oh_) * p.ow + ow_;
          optr = p.out + vox * p.out_pitch + nt * p.NT + c16;
        }
        uint32_t raw[16];
        tmem_ld_32x16(tbase + (static_cast<uint32_t>(q * 32) << 16) + acc * p.NT + c16, raw);
        tmem_ld_wait();
        if (valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = __uint_as_float(raw[j]) + bias_r[j];
            s1[j] += v[j];
            s2[j] += v[j] * v[j];
          }
          float t8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = v[i];
          st8(optr, pack8(t8));
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = v[8 + i];
          st8(optr + 8, pack8(t8));
        }
      }
      if (p.stats != nullptr) {
        // fold the two half-warps, then a 16-lane butterfly: lane l ends with column (l & 15)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
          s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
        }
#pragma unroll
        for (int st = 8; st >= 1; st >>= 1) {
          const bool up = (lane & st) != 0;
#pragma unroll
          for (int i = 0; i < st; ++i) {
            const float a1 = up ? s1[i] : s1[i + st], a2 = up ? s2[i] : s2[i + st];
            const float r1 = __shfl_xor_sync(0xffffffffu, a1, st), r2 = __shfl_xor_sync(0xffffffffu, a2, st);
            s1[i] = (up ? s1[i + st] : s1[i]) + r1;
            s2[i] = (up ? s2[i + st] : s2[i]) + r2;
          }
        }
        if (lane < 16) {
          atomicAdd(&s_stats[c16 + lane], s1[0]);
          atomicAdd(&s_stats[p.NT + c16 + lane], s2[0]);
        }
      }
    }
    tc_fence_before();
    if (dbg && tid == 128) p.dbg[10] = clock64();
  }
  __syncthreads();
  if (dbg && tid == 0) p.dbg[11] = clock64();
  if (p.stats != nullptr) {
    for (int i = tid; i < p.NT; i += blockDim.x) {
      atomicAdd(&p.stats[nt * p.NT + i], s_stats[i]);
      atomicAdd(&p.stats[p.cout + nt * p.NT + i], s_stats[p.NT + i]);
    }
  }
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tbase, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

bool encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, int kc) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return false;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  // A channel SLICE of a wider buffer (the up-convolution's half of a [up | skip] concat gradient: rows of 64 B every
  // 128 B) must not be promoted to whole 128-byte lines: measured 2x the algorithmic DRAM reads on the memory-bound
  // ConvTranspose dgrad / wgrad launches.  Contiguous rows keep the 128-byte promotion.
  const uint64_t row_bytes = dims[0] * 2;
  const bool sliced = rank > 1 && strides_bytes[0] > row_bytes;
  const CUtensorMapL2promotion promo = !sliced ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                       : (row_bytes % 64 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return false;
  }
  return true;
}

long long g_umma_launches = 0;

static constexpr size_t kSmemBudget = 200 * 1024;

// Chooses the tiling; returns false when the geometry is not handled by this path.
static bool plan(const UmmaConvArgs& a, UmmaConvParams& p, size_t& smem_bytes) {
  if (getenv("B200SEG_DISABLE_UMMA") || getenv("B200SEG_DISABLE_FLAT")) return false;
  if (a.cin % 16 || a.cout % 16) return false;
  if (!(a.k == 1 || a.k == 3 || a.k == 5) || a.dil < 1 || a.pad < 0) return false;
  if (a.in_pitch % 8 || a.out_pitch % 8) return false;
  const int halo = (a.k - 1) * a.dil;
  if (a.od != a.d + 2 * a.pad - halo || a.oh != a.h + 2 * a.pad - halo || a.ow != a.w + 2 * a.pad - halo) return false;
  p = UmmaConvParams{};
  p.n = a.n; p.od = a.od; p.oh = a.oh; p.ow = a.ow; p.cout = a.cout; p.out_pitch = a.out_pitch;
  p.k = a.k; p.pad = a.pad; p.dil = a.dil;
  p.scatter_cout = a.scatter_cout;
  if (a.scatter_cout && (a.scatter_cout % 16 || a.cout != 8 * a.scatter_cout || a.k != 1 || a.stats)) return false;
  const int cin_map = a.gather2 ? a.cin / 8 : a.cin;   // channels behind one input tensor map
  if (a.gather2 && (a.cin % 8 || a.k != 1 || cin_map % 16)) return false;
  p.KC = cin_map % 64 == 0 ? 64 : (cin_map % 32 == 0 ? 32 : 16);
  p.nchunks = a.cin / p.KC;
  p.cpm = cin_map / p.KC;
  p.rowbytes = p.KC * 2;
  p.swz = p.KC == 64 ? SWZ_128B : (p.KC == 32 ? SWZ_64B : SWZ_32B);
  const int k3 = a.k * a.k * a.k;
  // N tile: the whole C_out when it fits one accumulator comfortably, else 128 / 64 / 32 / 16
  if (a.cout <= 128 && !a.scatter_cout) p.NT = a.cout;
  else if (a.cout % 128 == 0) p.NT = 128;
  else if (a.cout % 64 == 0) p.NT = 64;
  else if (a.cout % 32 == 0) p.NT = 32;
  else p.NT = 16;
  if (!(a.oh >= 16 && a.ow >= 8) && !a.scatter_cout && p.NT == 128) {
    // Flat mode on the K-heavy 8^3 layers: a CTA streams taps x C_in x NT weights through shared memory for ONE or two
    // accumulators, so its time is max(MMA issue, weight bytes / ~40 B per clock of L2->SM bandwidth) and N = 128 leaves
    // most SMs without a tile (bottleneck dgrad: 32 CTAs, 3.5 MB of weights each).  Narrower N tiles cost more MMA
    // instructions per FLOP (66.5 / 54.6 / 48.6 cycles at N = 128 / 64 / 32) but spread the weight stream over the SMs.
    const long long rows = static_cast<long long>(a.oh + halo) * (a.ow + halo);
    const long long accs = (rows + 127) / 128;                       // accumulators per d-plane
    const double mma_cyc[3] = {66.5, 54.6, 48.6};
    const int cand[3] = {128, 64, 32};
    double best = -1;
    for (int i = 0; i < 3; ++i) {
      if (a.cout % cand[i]) continue;
      const long long ctas = static_cast<long long>(a.n) * a.od * (a.cout / cand[i]);
      const double waves = static_cast<double>((ctas + kNumSMs - 1) / kNumSMs);
      const double mma = static_cast<double>(k3) * (a.cin / 16) * accs * mma_cyc[i];
      const double wbytes = static_cast<double>(k3) * a.cin * cand[i] * 2 / 40.0;
      const double cost = waves * std::max(mma, wbytes);
      if (best < 0 || cost < best) best = cost, p.NT = cand[i];
    }
  }
  p.n_ntiles = a.cout / p.NT;
  p.slotB = (p.NT * p.rowbytes + 1023) & ~1023u;
  p.NB = static_cast<int>(std::min<size_t>(8, std::max<size_t>(3, 32768 / p.slotB)));  // hide the TMA round trip
  p.bytesB = p.NT * p.rowbytes;
  const size_t fixed_small = 2048 + 3 * p.NT * sizeof(float);
  size_t fixed = static_cast<size_t>(p.NB) * p.slotB + fixed_small;
  if (a.oh >= 16 && a.ow >= 8) {
    // ---- plane mode
    p.flat = 0;
    p.WB = 8 + halo;
    p.HB = 16 + halo;
    if (p.WB > 256 || p.HB > 256) return false;
    p.UP = 1;
    p.slotA = (static_cast<unsigned>(p.WB * p.HB) * p.rowbytes + 1023) & ~1023u;
    p.bytesA_unit = static_cast<unsigned>(p.WB * p.HB) * p.rowbytes;
    // Prefer a footprint that lets two CTAs share an SM (<= 112 KB smem, <= 256 TMEM columns): one CTA's epilogue
    // then overlaps the other's MMA phase.  Otherwise take the largest P that fits one CTA per SM.
    const int pmax = std::min(std::min(512 / p.NT, a.od), 8);
    const int extra = p.nchunks > 1 ? 2 : 0;   // ring slots beyond one chunk's planes (prefetch of the next chunk)
    int best = 0, best_S = 0;
    for (int P = pmax; P >= std::min(4, pmax) && !best; --P) {
      const int S = P + halo + extra;
      if (P * p.NT <= 256 && fixed + static_cast<size_t>(S) * p.slotA + 1024 <= 112 * 1024) best = P, best_S = S;
    }
    for (int P = pmax; P >= 1 && !best; --P) {
      for (int ex = extra; ex >= (p.nchunks > 1 ? 1 : 0) && !best; --ex) {
        const int S = P + halo + ex;
        if (fixed + static_cast<size_t>(S) * p.slotA <= kSmemBudget) best = P, best_S = S;
      }
    }
    if (!best) return false;
    p.P = best;
    p.DT = best;
    p.U = best + halo;
    p.S = best_S;
    p.tiles_w = (a.ow + 7) / 8;
    p.tiles_h = (a.oh + 15) / 16;
    p.tiles_d = (a.od + p.DT - 1) / p.DT;
  } else {
    // ---- flat mode: one halo'd box of the whole (h, w) extent and DT d-planes per CTA
    p.flat = 1;
    p.WB = a.ow + halo;
    p.HB = a.oh + halo;
    if (p.WB > 256 || p.HB > 256) return false;
    const int plane_rows = p.WB * p.HB;
    const int maxoff = halo * (plane_rows + p.WB + 1);
    // Among the depths per CTA that fit, take the one with the fewest rounds x accumulators: these are the small, K-heavy
    // layers (8^3 bottleneck), where the largest box leaves most of the 148 SMs without a tile.
    int best_dt = 0, best_p = 0;
    long long best_cost = -1;
    for (int DT = std::min(a.od, 256 - halo); DT >= 1; --DT) {
      const int P = (DT * plane_rows + 127) / 128;
      if (P * p.NT > 512) continue;
      const size_t slot = ((static_cast<size_t>(P) * 128 + maxoff) * p.rowbytes + 1023) & ~size_t(1023);
      const size_t loaded = static_cast<size_t>(DT + halo) * plane_rows * p.rowbytes;
      if (loaded > slot) continue;
      const int S = p.nchunks > 1 ? 2 : 1;
      if (fixed + S * slot > kSmemBudget) continue;
      const long long ctas = static_cast<long long>(a.n) * ((a.od + DT - 1) / DT) * p.n_ntiles;
      const long long cost = ((ctas + kNumSMs - 1) / kNumSMs) * P;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_dt = DT;
        best_p = P;
      }
    }
    if (!best_dt) return false;
    p.DT = best_dt;
    p.P = best_p;
    p.U = 1;
    p.UP = best_dt + halo;
    p.S = p.nchunks > 1 ? 2 : 1;
    p.slotA = static_cast<unsigned>(((static_cast<size_t>(p.P) * 128 + maxoff) * p.rowbytes + 1023) & ~size_t(1023));
    // These are the K-heavy layers with few accumulators: a weight tile is consumed in P x KC/16 MMAs (0.14 us at P = 1),
    // far less than a TMA round trip, so the weight ring takes all the shared memory the boxes leave (3 slots of 16 KB
    // measured 26 GB/s per SM: latency-bound).
    // (only when the grid is a single wave anyway: larger grids keep the footprint that lets two CTAs share an SM)
    if (static_cast<long long>(a.n) * ((a.od + p.DT - 1) / p.DT) * p.n_ntiles <= kNumSMs)
      p.NB = std::max(p.NB, static_cast<int>(std::min<size_t>(
                                9, (kSmemBudget - fixed_small - static_cast<size_t>(p.S) * p.slotA) / p.slotB)));
    fixed = static_cast<size_t>(p.NB) * p.slotB + fixed_small;
    p.bytesA_unit = static_cast<unsigned>(p.UP) * plane_rows * p.rowbytes;
    p.tiles_w = p.tiles_h = 1;
    p.tiles_d = (a.od + p.DT - 1) / p.DT;
  }
  // weight-ring grouping: whole kw rows per barrier when at least two groups fit (rounds NB down to a multiple of k)
  p.G = 1;
  if (a.k > 1 && !a.gather2 && p.NB / a.k >= 2) {
    p.G = a.k;
    p.NB = (p.NB / a.k) * a.k;
  }
  unsigned cols = 32;
  while (cols < static_cast<unsigned>(p.P * p.NT)) cols <<= 1;
  if (cols > 512) return false;
  p.tmem_cols = cols;
  smem_bytes = static_cast<size_t>(p.S) * p.slotA + fixed + 1024;
  const long long ctas = static_cast<long long>(a.n) * p.tiles_d * p.tiles_h * p.tiles_w * p.n_ntiles;
  if (ctas > 2147483647LL) return false;
  return smem_bytes <= 227 * 1024;
}

bool conv_umma_supported(const UmmaConvArgs& a) {
  if (conv_umma_roll_supported(a) || conv_umma_plane_supported(a)) return true;
  if (a.scale) return conv_umma_plane_relaxed_supported(a);   // the flat kernel has no fused scale / activation epilogue
  UmmaConvParams p;
  size_t smem;
  return plan(a, p, smem) || conv_umma_plane_relaxed_supported(a);
}

int conv_umma_run(const UmmaConvArgs& a, cudaStream_t st) {
  if (conv_umma_ws_supported(a)) return conv_umma_ws_run(a, st);          // K-heavy layers on 8 x 8 planes: weights stationary
  if (conv_umma_roll_supported(a)) return conv_umma_roll_run(a, st);     // narrow outputs: kd taps in N (conv_umma_roll.cu)
  if (conv_umma_pair_supported(a)) return conv_umma_pair_run(a, st);     // K-heavy deep layers: CTA pairs share the weights
  if (conv_umma_plane_supported(a)) return conv_umma_plane_run(a, st);   // persistent kernel (conv_umma_p.cu)
  UmmaConvParams p;
  size_t smem;
  // measured on B200 (probes/flat_vs_plane.py): for K-heavy 3x3x3 layers on 8^3 grids the short-plane kernel (N = 128
  // tiles, half of the M rows idle) is ~10 % faster than the flat kernel's N = 32 tiles over whole haloed boxes
  const bool prefer_plane = a.k == 3 && a.oh <= 8 && a.cin >= 256 && !a.gather2 && conv_umma_plane_relaxed_supported(a);
  if (prefer_plane || a.scale || !plan(a, p, smem)) {
    // small planes with a halo too large for the flat kernel's whole-box slot (5x5x5 at 8^3, vnet3d.py:25)
    if (conv_umma_plane_relaxed_supported(a)) return conv_umma_plane_run(a, st);
    set_error("conv_umma_run: unsupported geometry");
    return B200SEG_ERR_INVALID;
  }
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.bias = a.bias;
  p.stats = a.stats;
  if ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.wpack)) & 15) {
    set_error("conv_umma_run: buffers must be 16-byte aligned");
    return B200SEG_ERR_INVALID;
  }
  TensorMaps8 tmAs;
  CUtensorMap tmB;
  if (!a.gather2) {
    const uint64_t dims[5] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.w), static_cast<uint64_t>(a.h),
                              static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.n)};
    const uint64_t pb = static_cast<uint64_t>(a.in_pitch) * 2;
    const uint64_t str[4] = {pb, pb * a.w, pb * a.w * a.h, pb * a.w * a.h * a.d};
    const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB),
                             static_cast<uint32_t>(p.UP), 1u};
    if (!encode_bf16_map(&tmAs.m[0], a.in, 5, dims, str, box, p.KC)) return B200SEG_ERR_CUDA;
    for (int i = 1; i < 8; ++i) tmAs.m[i] = tmAs.m[0];
  } else {
    // `in` is the fine grid [n, 2d, 2h, 2w, cin/8]; map g = the sub-lattice with offset (a,b,e) = bits of g
    const int cm = a.cin / 8;
    const uint64_t pb = static_cast<uint64_t>(a.in_pitch) * 2;
    const uint64_t W2 = 2ull * a.w, H2 = 2ull * a.h, D2 = 2ull * a.d;
    for (int g = 0; g < 8; ++g) {
      const uint64_t off = (((g >> 2) * H2 + ((g >> 1) & 1)) * W2 + (g & 1)) * pb;
      const uint64_t dims[5] = {static_cast<uint64_t>(cm), static_cast<uint64_t>(a.w), static_cast<uint64_t>(a.h),
                                static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.n)};
      const uint64_t str[4] = {2 * pb, 2 * pb * W2, 2 * pb * W2 * H2, pb * W2 * H2 * D2};
      const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB),
                               static_cast<uint32_t>(p.UP), 1u};
      if (!encode_bf16_map(&tmAs.m[g], static_cast<const uint8_t*>(a.in) + off, 5, dims, str, box, p.KC))
        return B200SEG_ERR_CUDA;
    }
  }
  {
    const int k3 = a.gather2 ? 8 : a.k * a.k * a.k;
    const uint64_t cin_w = a.gather2 ? a.cin / 8 : a.cin;   // weight pack is [taps][cout][cin_w]
    const uint64_t dims[3] = {cin_w, static_cast<uint64_t>(a.cout), static_cast<uint64_t>(k3)};
    const uint64_t str[2] = {cin_w * 2, cin_w * a.cout * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.NT), 1u};
    if (!encode_bf16_map(&tmB, a.wpack, 3, dims, str, box, p.KC)) return B200SEG_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("conv_umma_run: cannot raise the dynamic shared memory limit");
      return B200SEG_ERR_CUDA;
    }
    attr_set = true;
  }
  const int ctas = a.n * p.tiles_d * p.tiles_h * p.tiles_w * p.n_ntiles;
  p.dbg = nullptr;
  if (getenv("B200SEG_DEBUG_TIMELINE")) {
    cudaMalloc(&p.dbg, 16 * sizeof(long long));
    cudaMemset(p.dbg, 0, 16 * sizeof(long long));
  }
  conv_umma_kernel<<<ctas, 256, smem, st>>>(tmAs, tmB, p);
  B200_CHECK_LAUNCH("conv_umma");
  ++g_umma_launches;
  if (p.dbg) {
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    fprintf(stderr, "[timeline] ctas %d P %d NT %d KC %d chunks %d S %d U %d flat %d smem %zu | a0 %lld a1 %lld a2 %lld issued %lld "
            "epi_start %lld epi_end %lld cta_end %lld\n", ctas, p.P, p.NT, p.KC, p.nchunks, p.S, p.U, p.flat, smem,
            h[1] - h[0], h[2] - h[0], h[3] - h[0], h[8] - h[0], h[9] - h[0], h[10] - h[0], h[11] - h[0]);
  }
  return 0;
}

}  // namespace b200
