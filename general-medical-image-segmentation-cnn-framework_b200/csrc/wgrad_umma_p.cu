// Persistent tcgen05 weight-gradient kernel, plane mode (H_out >= 16, W_out >= 8), "tap blocks on both operands".
//
//   dW[a][b][e][ci][co] = sum_{n,d,h,w} x[n, d - pad + a*dil, h - pad + b*dil, w - pad + e*dil, ci] * dy[n, d, h, w, co]
//
// GEMM view: K = voxels, both operands MN-major (a shared-memory row is a voxel, its bytes are that voxel's channels --
// the channels-last layout TMA delivers).  A tcgen05.mma with M = 128 costs ~48 cycles however small N is
// (probes/mma_rate.cu: N=32 48.9, N=96 61.1, N=192 96.1 cycles), so the first version (M = kw taps x C_in, N = C_out,
// one instruction per (kd, kh) pair) was bound by the tensor pipe at a fraction of its rate.  Here the kd taps ride in
// the N dimension as well:
//   A = one halo'd x plane; M block j = kw tap e0+j (the same box shifted by j*dil rows: LBO of the A descriptor);
//       the kh tap is a row offset b*dil*WB of the descriptor start.
//   B = k consecutive dy planes {dx + pad - a*dil}; N block i = dy plane (tap a = k-1-i), LBO = dil plane slots.
// One instruction (M = 128, N = k*NT, K = 16 voxels) therefore covers up to (128/KC) x k taps: for the C=32 layers
// 9 of the 27 taps in 61 cycles instead of 3 in 49.  Each CTA keeps its accumulators in TMEM across its whole share of
// the voxels (persistent split-K) and adds them to the fp32 gradient with red.global at the end.
//
// The dy planes live in a ring whose first (k-1)*dil slots are mirrored after its end, so that any k planes
// `dil` slots apart are also equally spaced in shared memory; walking a tile column along d costs one x plane and one
// new dy plane per step.
// Warp roles (192 threads): 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.h"
#include "ptx.cuh"

namespace b200 {

bool encode_bf16_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int kc);

struct WgradPGroup {
  int b;          // kh tap of this group
  int e0;         // first kw tap (M block 0)
  int nblk;       // kw taps in the group (M blocks used)
};

struct WgradPParams {
  int n, d, od, oh, ow, cin, cout, k, pad, dil;   // d = input depth: one step per x plane
  int KC, NT, NBLK;                 // channels per x row chunk, dy channels per N block, N = NBLK * NT with NBLK = k
  int WB, HB;
  int tiles_w, tiles_h;
  long long steps;                  // (tile column, x plane) pairs in the whole tensor
  int splits;                       // CTAs sharing one (bsel, chunk, ntile) class
  int nchunks, n_ntiles, bsplit;
  int SX, RY, span;                 // x ring slots; dy ring slots (mirror of span-1 slots appended); planes per B operand
  int ngroups;
  WgradPGroup groups[6];
  unsigned slotX, slotY, bytesX, bytesY, rowbytesA, rowbytesB, swzA, swzB, tmem_cols;
  float* dwp;
  float* partial;          // split-K workspace [splits][k^3][cin][cout] (null: atomics into dwp)
  long long slice;         // floats per workspace slice
};

constexpr int kWgradPThreads = 192;

template <int NG>
__global__ void __launch_bounds__(kWgradPThreads, 1)
    wgrad_umma_plane_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                            const WgradPParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = smem + static_cast<size_t>(p.SX) * p.slotX;
  const int ry_total = p.RY + p.span - 1;
  uint64_t* fullX = reinterpret_cast<uint64_t*>(sY + static_cast<size_t>(ry_total) * p.slotY);
  uint64_t* emptyX = fullX + p.SX;
  uint64_t* fullY = emptyX + p.SX;
  uint64_t* emptyY = fullY + p.RY;
  uint64_t* accFull = emptyY + p.RY;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accFull + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which slice of the problem ---------------------------------------------------------------------------
  int bid = blockIdx.x;
  const int split = bid % p.splits;
  bid /= p.splits;
  const int nt = bid % p.n_ntiles;
  bid /= p.n_ntiles;
  const int chunk = bid % p.nchunks;
  const int bsel = bid / p.nchunks;            // kh tap handled by this CTA when the kh taps are split over CTAs
  const long long per = (p.steps + p.splits - 1) / p.splits;
  const long long s0 = split * per;
  const long long s1 = s0 + per < p.steps ? s0 + per : p.steps;

  if (tid == 0) {
    for (int i = 0; i < p.SX; ++i) {
      mbar_init(&fullX[i], 1);
      mbar_init(&emptyX[i], 1);
    }
    for (int i = 0; i < p.RY; ++i) {
      mbar_init(&fullY[i], 1);
      mbar_init(&emptyY[i], 1);
    }
    mbar_init(accFull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;
  const int back = (p.k - 1) * p.dil;   // oldest dy plane of a step is `back` planes behind the newest

  // A step s = (column, dx): column = s / od over (nn, th, tw); dx = s % od.  dy planes are numbered by a running index
  // q (one new plane per step, `back` extra when a column starts): plane q lives in ring slot q % RY.
  if (warp == 0) {
    if (lane == 0 && s0 < s1) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmY);
      int yslot = 0;        // ring slot of the next dy plane, and the parity its empty barrier has to show
      uint32_t yph = 0;
      int sx = 0;
      uint32_t phx = 0;
      auto load_y = [&](int nn, int h0, int w0, int plane) {
        mbar_wait(&emptyY[yslot], yph ^ 1);
        const bool mirror = yslot < p.span - 1;
        mbar_arrive_expect_tx(&fullY[yslot], mirror ? 2 * p.bytesY : p.bytesY);
        tma_load_5d(sY + static_cast<size_t>(yslot) * p.slotY, &tmY, &fullY[yslot], nt * p.NT, w0, h0, plane, nn);
        if (mirror)
          tma_load_5d(sY + static_cast<size_t>(yslot + p.RY) * p.slotY, &tmY, &fullY[yslot], nt * p.NT, w0, h0, plane, nn);
        if (++yslot == p.RY) {
          yslot = 0;
          yph ^= 1;
        }
      };
      // position of the first step: (column, dx) = (s0 / d, s0 % d), column over (nn, th, tw)
      int dx = static_cast<int>(s0 % p.d);
      long long col = s0 / p.d;
      int tw = static_cast<int>(col % p.tiles_w);
      col /= p.tiles_w;
      int th = static_cast<int>(col % p.tiles_h);
      int nn = static_cast<int>(col / p.tiles_h);
      bool fresh = true;
      for (long long s = s0; s < s1; ++s) {
        const int h0 = th * 16, w0 = tw * 8;
        if (fresh) {
          // (re)prime the ring: planes dx + pad - back .. dx + pad - 1 (out-of-range planes arrive as zeros)
          for (int j = back; j >= 1; --j) load_y(nn, h0, w0, dx + p.pad - j);
          fresh = false;
        }
        load_y(nn, h0, w0, dx + p.pad);
        mbar_wait(&emptyX[sx], phx ^ 1);
        mbar_arrive_expect_tx(&fullX[sx], p.bytesX);
        tma_load_5d(sX + static_cast<size_t>(sx) * p.slotX, &tmX, &fullX[sx], chunk * p.KC, w0 - p.pad, h0 - p.pad, dx, nn);
        if (++sx == p.SX) {
          sx = 0;
          phx ^= 1;
        }
        if (++dx == p.d) {
          dx = 0;
          fresh = true;
          if (++tw == p.tiles_w) {
            tw = 0;
            if (++th == p.tiles_h) {
              th = 0;
              ++nn;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (s0 < s1) {
      const uint32_t leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(128, p.NBLK * p.NT, 1, 1);
      const uint32_t rowA16 = p.rowbytesA >> 4, rowB16 = p.rowbytesB >> 4;
      // A: M blocks `dil` rows apart (LBO), 8-row K groups WB rows apart (SBO), 16 voxels = 2 box rows per MMA
      const uint64_t a_desc0 = make_smem_desc(0, static_cast<uint32_t>(p.dil) * p.rowbytesA,
                                              static_cast<uint32_t>(p.WB) * p.rowbytesA, p.swzA);
      // B: N blocks `dil` plane slots apart (LBO), 8-row K groups contiguous, 16 rows per MMA
      const uint64_t b_desc0 = make_smem_desc(0, static_cast<uint32_t>(p.dil) * p.slotY, 8u * p.rowbytesB, p.swzB);
      const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
      const uint32_t a_lbo = static_cast<uint32_t>(a_desc0) & 0x3FFF0000u, b_lbo = static_cast<uint32_t>(b_desc0) & 0x3FFF0000u;
      const uint32_t a_adv16 = 2u * p.WB * rowA16, b_adv16 = 16u * rowB16;
      const uint32_t sX16 = smem_u32(sX) >> 4, sY16 = smem_u32(sY) >> 4;
      const uint32_t slotX16 = p.slotX >> 4, slotY16 = p.slotY >> 4;
      uint32_t g_off16[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int b = p.bsplit > 1 ? bsel : p.groups[g].b;
        g_off16[g] = static_cast<uint32_t>((b * p.dil) * p.WB + p.groups[g].e0 * p.dil) * rowA16;
      }
      const uint32_t NTOT = p.NBLK * p.NT;
      // dy ring bookkeeping without divisions: `oslot` = slot of the OLDEST plane of the current step, `wslot`/`wph` =
      // slot and parity of the next plane whose arrival has not been observed yet
      int oslot = 0, wslot = 0;
      uint32_t wph = 0;
      int sx = 0;
      uint32_t phx = 0, accum = 0;
      int dx = static_cast<int>(s0 % p.d);
      bool fresh = true;
      for (long long s = s0; s < s1; ++s) {
        const int nwait = fresh ? back + 1 : 1;     // planes that arrive for this step
        for (int j = 0; j < nwait; ++j) {
          mbar_wait(&fullY[wslot], wph);
          if (++wslot == p.RY) {
            wslot = 0;
            wph ^= 1;
          }
        }
        fresh = false;
        mbar_wait(&fullX[sx], phx);
        tc_fence_after();
        // (planes that wrapped to slots 0 .. span-2 are read through their mirrors at RY + slot)
        // lane-0 broadcasts: tell ptxas the two per-step bases are warp-uniform, so the 24 descriptor updates below run on
        // the uniform datapath instead of one R2UR per descriptor half per MMA
        const uint32_t b_lo0 = __shfl_sync(0xffffffffu, ((sY16 + oslot * slotY16) & 0x3FFF) | b_lbo, 0);
        const uint32_t x_lo0 = __shfl_sync(0xffffffffu, sX16 + sx * slotX16, 0);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          uint32_t a_lo = ((x_lo0 + g_off16[g]) & 0x3FFF) | a_lbo;
          uint32_t b_lo = b_lo0;
          const uint32_t d_tmem = tbase + g * NTOT;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            umma_f16_pred_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, ks == 0 ? accum : 1u, leader);
            a_lo += a_adv16;
            b_lo += b_adv16;
          }
        }
        accum = 1;
        umma_commit_pred(&emptyX[sx], leader);
        if (++sx == p.SX) {
          sx = 0;
          phx ^= 1;
        }
        // the oldest plane is not needed by the next step of this column; at a column end every plane is released
        const bool last_of_col = (dx == p.d - 1) || (s + 1 == s1);
        const int nrel = last_of_col ? back + 1 : 1;
        for (int j = 0; j < nrel; ++j) {
          umma_commit_pred(&emptyY[oslot], leader);
          if (++oslot == p.RY) oslot = 0;
        }
        if (++dx == p.d) {
          dx = 0;
          fresh = true;
        }
      }
      umma_commit_pred(accFull, leader);
    }
  } else if (s0 < s1 || p.partial != nullptr) {
    // =========================== epilogue: TMEM -> workspace slice (vector stores) or red.global.add.f32 ===========
    const int q4 = warp & 3;            // warps 2..5 -> lane quadrants 2,3,0,1
    const int m = q4 * 32 + lane;       // M row = block * KC + ci
    const int blk = m / p.KC, ci = chunk * p.KC + (m % p.KC);
    const bool have = s0 < s1;          // a split without steps still has to zero its workspace slice
    if (have) {
      mbar_wait(accFull, 0);
      tc_fence_after();
    }
    const uint32_t NTOT = p.NBLK * p.NT;
    float* base = p.partial ? p.partial + static_cast<size_t>(split) * p.slice : p.dwp;
    for (int g = 0; g < NG; ++g) {
      const int b = p.bsplit > 1 ? bsel : p.groups[g].b;
      const int e = p.groups[g].e0 + blk;
      const bool row_ok = blk < p.groups[g].nblk && ci < p.cin;   // ci >= cin: zero-filled rows of a ragged last chunk
      for (int i = 0; i < p.NBLK; ++i) {
        const int a = p.k - 1 - i;      // N block i holds dy plane dx + pad - (k-1-i)*dil
        const int tap = (a * p.k + b) * p.k + e;
        for (int cc = 0; cc < p.NT; cc += 16) {
          uint32_t raw[16];
          if (have) {
            tmem_ld_32x16(tbase + (static_cast<uint32_t>(q4 * 32) << 16) + g * NTOT + i * p.NT + cc, raw);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) raw[j] = 0u;
          }
          if (row_ok) {
            float* dst = base + (static_cast<size_t>(tap) * p.cin + ci) * p.cout + nt * p.NT + cc;
            if (p.partial) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<uint4*>(dst + j) = make_uint4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(raw[j])),
                               "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])),
                               "f"(__uint_as_float(raw[j + 3]))
                               : "memory");
                } else {
                  atomicAdd(dst + j, __uint_as_float(raw[j]));
                  atomicAdd(dst + j + 1, __uint_as_float(raw[j + 1]));
                  atomicAdd(dst + j + 2, __uint_as_float(raw[j + 2]));
                  atomicAdd(dst + j + 3, __uint_as_float(raw[j + 3]));
                }
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tbase, p.tmem_cols);
  }
}

// dwp[i] += sum over splits of partial[s][i]   (coalesced: consecutive threads, consecutive elements of every slice).
// The narrow layers have FEW elements (27 x 32 x 32) in MANY slices (148): one thread per element walking the slices is a
// chain of dependent-latency loads (measured 42 us for 16 MB), so the slices are also dealt over blockIdx.y -- each thread
// sums its share with 8 independent loads in flight and adds the result to dwp with one atomic per element.
__global__ void __launch_bounds__(256)
    wgrad_reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ dwp, long long numel, int splits) {
  const long long nvec = numel / 4;
  const int gy = gridDim.y, y = blockIdx.y;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4* src = reinterpret_cast<const float4*>(partial) + i;
    const long long sstride = numel / 4;   // float4s per slice
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = y;
    for (; s + 7 * gy < splits; s += 8 * gy) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[static_cast<long long>(s + u * gy) * sstride];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc.x += v[u].x;
        acc.y += v[u].y;
        acc.z += v[u].z;
        acc.w += v[u].w;
      }
    }
    for (; s < splits; s += gy) {
      const float4 v = src[static_cast<long long>(s) * sstride];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    float* dst = dwp + 4 * i;
    if (gy == 1) {
      float4 o = *reinterpret_cast<const float4*>(dst);
      o.x += acc.x;
      o.y += acc.y;
      o.z += acc.z;
      o.w += acc.w;
      *reinterpret_cast<float4*>(dst) = o;
    } else {
      // one 4-wide vector reduction (REDG.E.ADD.F32x4; dwp is 16-byte aligned on this path)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w)
                   : "memory");
    }
  }
}

// relaxed: accept short planes (H_out, W_out >= 4): rows of the 16 x 8 dy tile outside the tensor arrive as TMA zero fill
// and contribute nothing -- the last resort for small planes whose halo is too large for the flat kernel (5x5x5 at 8^3)
static bool plan_wgrad_plane(const UmmaWgradArgs& a, WgradPParams& p, size_t& smem_bytes, bool relaxed = false) {
  if (getenv("B200SEG_DISABLE_UMMA") || getenv("B200SEG_DISABLE_UMMA_WGRAD") || getenv("B200SEG_DISABLE_PERSISTENT"))
    return false;
  if (a.gather2) return false;
  // C_in only has to be a multiple of 16 (HighRes3DNet's 16-channel layers, highresnet.py:40-59): the last 32-channel
  // chunk then reaches past the tensor, TMA fills those channels with zeros and the epilogue skips their rows
  if (a.cin % 16 || a.cout % 16) return false;
  if (a.x_pitch % 8 || a.dy_pitch % 8) return false;
  if (!(a.k == 1 || a.k == 3 || a.k == 5) || a.dil < 1 || a.pad < 0) return false;
  const int halo = (a.k - 1) * a.dil;
  if (a.od != a.d + 2 * a.pad - halo || a.oh != a.h + 2 * a.pad - halo || a.ow != a.w + 2 * a.pad - halo) return false;
  if (relaxed ? !(a.oh >= 4 && a.ow >= 4) : !(a.oh >= 16 && a.ow >= 8)) return false;
  p = WgradPParams{};
  p.n = a.n; p.d = a.d; p.od = a.od; p.oh = a.oh; p.ow = a.ow; p.cin = a.cin; p.cout = a.cout;
  p.k = a.k; p.pad = a.pad; p.dil = a.dil;
  // 64-channel rows put two kw taps in one instruction but then the accumulators of all kh taps no longer fit TMEM at
  // N = 96 and the kh taps are split over CTA classes, each re-reading the same x planes: measured L2-bound.  With
  // narrow outputs use 32-channel rows (chunks become the CTA classes): same MMA count, 2.4x less L2 traffic.
  p.KC = (a.cin % 64 == 0 && a.cout > 32) ? 64 : 32;
  const int MB = 128 / p.KC;
  p.nchunks = (a.cin + p.KC - 1) / p.KC;
  p.NBLK = a.k;
  const int egroups = (a.k + MB - 1) / MB;
  // dy channels per N block: N = k*NT <= 256, and the accumulators of one CTA must fit 512 TMEM columns
  int NT = 0, bsplit = 1;
  for (int cand : {64, 32, 16}) {
    if (a.cout % cand || a.k * cand > 256) continue;
    if (a.k * egroups * a.k * cand <= 512 && a.k * egroups <= 6) { NT = cand; bsplit = 1; break; }
    if (egroups * a.k * cand <= 512) { NT = cand; bsplit = a.k; break; }
  }
  if (!NT) return false;
  p.NT = NT;
  p.bsplit = bsplit;
  p.n_ntiles = a.cout / NT;
  p.ngroups = 0;
  for (int b = 0; b < (bsplit > 1 ? 1 : a.k); ++b)
    for (int e0 = 0; e0 < a.k; e0 += MB) {
      WgradPGroup& G = p.groups[p.ngroups++];
      G.b = b;
      G.e0 = e0;
      G.nblk = std::min(MB, a.k - e0);
    }
  if (!(p.ngroups == 1 || p.ngroups == 2 || p.ngroups == 3 || p.ngroups == 6)) return false;
  unsigned cols = 32;
  while (cols < static_cast<unsigned>(p.ngroups * a.k * NT)) cols <<= 1;
  if (cols > 512) return false;
  p.tmem_cols = cols;
  p.rowbytesA = p.KC * 2;
  p.rowbytesB = NT * 2;
  p.swzA = p.KC == 64 ? SWZ_128B : SWZ_64B;
  p.swzB = NT == 64 ? SWZ_128B : (NT == 32 ? SWZ_64B : SWZ_32B);
  p.WB = 8 + halo;
  p.HB = 16 + halo;
  // the last M block of a group reads up to (MB-1)*dil rows past the tap it starts at; keep the box tall enough
  const int extra_rows = (MB - 1) * a.dil;
  p.slotX = (static_cast<unsigned>(p.WB * p.HB + extra_rows + 8) * p.rowbytesA + 1023) & ~1023u;
  p.bytesX = static_cast<unsigned>(p.WB * p.HB) * p.rowbytesA;
  p.slotY = (128u * p.rowbytesB + 1023) & ~1023u;
  p.bytesY = 128u * p.rowbytesB;
  p.span = halo + 1;
  if (static_cast<unsigned long long>(a.dil) * p.slotY / 16 > 0x3FFF) return false;
  const size_t budget = 220 * 1024;
  p.SX = 3;
  p.RY = p.span + 2;
  auto total = [&]() {
    return static_cast<size_t>(p.SX) * p.slotX + static_cast<size_t>(p.RY + p.span - 1) * p.slotY + 2048;
  };
  if (total() > budget) {
    p.SX = 2;
    p.RY = p.span + 1;
    if (total() > budget) return false;
  } else {
    while (p.RY < 2 * p.span + 4 && total() + p.slotY <= budget) ++p.RY;
    while (p.SX < 6 && total() + p.slotX <= budget) ++p.SX;
  }
  p.tiles_w = (a.ow + 7) / 8;
  p.tiles_h = (a.oh + 15) / 16;
  p.steps = static_cast<long long>(a.n) * p.tiles_h * p.tiles_w * a.d;
  const long long combos = static_cast<long long>(bsplit) * p.nchunks * p.n_ntiles;
  long long splits = combos >= kNumSMs ? 1 : kNumSMs / combos;
  splits = std::max<long long>(1, std::min(splits, p.steps));
  p.splits = static_cast<int>(splits);
  smem_bytes = total() + 1024;
  return smem_bytes <= 227 * 1024 && combos * splits <= 2147483647LL;
}

bool wgrad_umma_plane_supported(const UmmaWgradArgs& a) {
  WgradPParams p;
  size_t smem;
  return plan_wgrad_plane(a, p, smem);
}
bool wgrad_umma_plane_relaxed_supported(const UmmaWgradArgs& a) {
  WgradPParams p;
  size_t smem;
  return plan_wgrad_plane(a, p, smem, true);
}

size_t wgrad_umma_plane_workspace_bytes(const UmmaWgradArgs& a) {
  WgradPParams p;
  size_t smem;
  if (!plan_wgrad_plane(a, p, smem) || p.splits < 2) return 0;
  const size_t numel = static_cast<size_t>(a.k) * a.k * a.k * a.cin * a.cout;
  if (numel % 4) return 0;
  return static_cast<size_t>(p.splits) * numel * sizeof(float);
}

template <int NG>
static int launch_wgrad_plane(const CUtensorMap& tmX, const CUtensorMap& tmY, const WgradPParams& p, size_t smem,
                              int ctas, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_umma_plane_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess) {
      set_error("wgrad_umma_plane: cannot raise the dynamic shared memory limit");
      return B200SEG_ERR_CUDA;
    }
    attr_set = true;
  }
  wgrad_umma_plane_kernel<NG><<<ctas, kWgradPThreads, smem, st>>>(tmX, tmY, p);
  B200_CHECK_LAUNCH("wgrad_umma_plane");
  return 0;
}

int wgrad_umma_plane_run(const UmmaWgradArgs& a, cudaStream_t st) {
  WgradPParams p;
  size_t smem;
  if (!plan_wgrad_plane(a, p, smem) && !plan_wgrad_plane(a, p, smem, true)) {
    set_error("wgrad_umma_plane_run: unsupported geometry");
    return B200SEG_ERR_INVALID;
  }
  p.dwp = a.dwp;
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.dy)) & 15) {
    set_error("wgrad_umma_plane_run: buffers must be 16-byte aligned");
    return B200SEG_ERR_INVALID;
  }
  const long long numel = static_cast<long long>(a.k) * a.k * a.k * a.cin * a.cout;
  const size_t need = wgrad_umma_plane_workspace_bytes(a);
  p.partial = nullptr;
  p.slice = numel;
  if (need && a.partial && a.partial_bytes >= need && (reinterpret_cast<uintptr_t>(a.partial) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(a.dwp) & 15) == 0)
    p.partial = a.partial;
  CUtensorMap tmX, tmY;
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(a.cin), static_cast<uint64_t>(a.w), static_cast<uint64_t>(a.h),
                              static_cast<uint64_t>(a.d), static_cast<uint64_t>(a.n)};
    const uint64_t pb = static_cast<uint64_t>(a.x_pitch) * 2;
    const uint64_t str[4] = {pb, pb * a.w, pb * a.w * a.h, pb * a.w * a.h * a.d};
    const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB), 1u, 1u};
    if (!encode_bf16_map(&tmX, a.x, 5, dims, str, box, p.KC)) return B200SEG_ERR_CUDA;
  }
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(a.cout), static_cast<uint64_t>(a.ow), static_cast<uint64_t>(a.oh),
                              static_cast<uint64_t>(a.od), static_cast<uint64_t>(a.n)};
    const uint64_t pb = static_cast<uint64_t>(a.dy_pitch) * 2;
    const uint64_t str[4] = {pb, pb * a.ow, pb * a.ow * a.oh, pb * a.ow * a.oh * a.od};
    const uint32_t box[5] = {static_cast<uint32_t>(p.NT), 8u, 16u, 1u, 1u};
    if (!encode_bf16_map(&tmY, a.dy, 5, dims, str, box, p.NT)) return B200SEG_ERR_CUDA;
  }
  const int ctas = p.bsplit * p.nchunks * p.n_ntiles * p.splits;
  int rc = B200SEG_ERR_INVALID;
  switch (p.ngroups) {
    case 1: rc = launch_wgrad_plane<1>(tmX, tmY, p, smem, ctas, st); break;
    case 2: rc = launch_wgrad_plane<2>(tmX, tmY, p, smem, ctas, st); break;
    case 3: rc = launch_wgrad_plane<3>(tmX, tmY, p, smem, ctas, st); break;
    case 6: rc = launch_wgrad_plane<6>(tmX, tmY, p, smem, ctas, st); break;
    default: set_error("wgrad_umma_plane_run: unsupported group count"); break;
  }
  if (rc == 0 && p.partial) {
    const int bx = grid_for(numel / 4, 256, kNumSMs * 8);
    const int by = std::max(1, std::min(p.splits / 4, 2 * kNumSMs / bx));   // few elements, many slices: split the slices too
    wgrad_reduce_partials_kernel<<<dim3(bx, by), 256, 0, st>>>(p.partial, a.dwp, numel, p.splits);
    B200_CHECK_LAUNCH("wgrad_reduce_partials");
  }
  if (rc == 0) ++g_umma_launches;
  return rc;
}

}  // namespace b200
