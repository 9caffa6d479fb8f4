// tcgen05 weight-gradient kernel:  dW[tap][ci][co] += sum_voxels x[voxel - pad + tap*dil][ci] * dy[voxel][co].
//
// GEMM view per tap: M = C_in, N = C_out, K = voxels.  Both operands are "MN-major" for the tensor core: shared
// memory rows are voxels (K), each row holds the channels (M or N) of that voxel -- exactly the channels-last layout
// TMA delivers.  One MMA (M=128, N=NT, K=16) covers 16 voxels = two 8-row groups.
//   * A = x: a halo'd box of KC input channels.  The 128 M rows are 128/KC *taps*: M-block j of the instruction reads
//     the same box shifted by j*LBO bytes, so neighbouring taps (row offsets in arithmetic progression) share one
//     instruction.  KC=64: pairs of consecutive (kh,kw) taps; KC=32: up to four consecutive kw taps.
//   * B = dy: the un-haloed tile of NT output channels.
// Every CTA keeps its accumulators (one per tap group) in TMEM across ALL the voxel tiles it is assigned
// (persistent split-K), then adds them to the fp32 gradient with red.global -- one pass over x and dy per K-chunk.
//   plane mode: tile = 8 (w) x 16 (h) voxels of one d-plane; flat mode (small H/W): a whole zero-padded (d,h,w) box,
//   dy loaded with the same halo'd row pitch so that row r of A and row r of B are the same voxel (halo rows of dy are
//   TMA zero fill and contribute nothing).
// Warp roles (192 threads): 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kMaxGroups = 16;

struct WgradGroup {
  int plane;      // which loaded x plane (index into this CTA's plane list); flat mode: 0
  int base_rows;  // row offset of M-block 0 inside the plane / box
  int lbo_rows;   // row distance between consecutive M-blocks
  int tap[8];     // tap index of each M-block, -1 = unused block
};

struct TensorMaps8W {
  CUtensorMap m[8];
};

struct WgradParams {
  int n, od, oh, ow, cin, cout, k, pad, dil;
  int KC, MB, NT, flat;
  int WB, HB, DT, UPX;         // box geometry; UPX = planes per x box (1 in plane mode)
  int nplanes;                 // x boxes per work item
  int plane_a[8];              // kd index of each loaded plane
  int ngroups;
  WgradGroup groups[kMaxGroups];
  int ksteps;                  // MMAs (K=16 voxels) per group per work item
  int a_sbo_rows, a_kadv_rows; // A: 8-row-group stride and per-MMA advance, in rows
  int tiles_w, tiles_h, tiles_d;
  long long items;             // work items (voxel tiles) in the whole tensor
  int splits;                  // CTAs sharing one (asplit, chunk, ntile) combination
  int nchunks, n_ntiles, asplit;
  int stages, gather2;
  int padd, padh, padw;        // front padding per dimension (== pad except for the stride-2 class launches)
  unsigned slotX, slotY, stage_bytes, rowbytesA, rowbytesB, swzA, swzB, bytesX, bytesY, tmem_cols;
  float* dwp;
};

__global__ void __launch_bounds__(192, 1)
    wgrad_umma_kernel(const __grid_constant__ TensorMaps8W tmXs, const __grid_constant__ CUtensorMap tmY,
                      const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* accFull = empty + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accFull + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which slice of the problem ---------------------------------------------------------------------------
  int bid = blockIdx.x;
  const int split = bid % p.splits;
  bid /= p.splits;
  const int nt = bid % p.n_ntiles;
  bid /= p.n_ntiles;
  const int chunk = bid % p.nchunks;
  const int asel = bid / p.nchunks;  // kd subset when the taps are split over CTAs
  const long long per = (p.items + p.splits - 1) / p.splits;
  const long long it0 = split * per;
  const long long it1 = it0 + per < p.items ? it0 + per : p.items;

  // zero the whole ring once: flat mode reads rows past the loaded boxes (they must be finite, and zero on the dy side)
  for (size_t i = static_cast<size_t>(tid) * 16; i < static_cast<size_t>(p.stages) * p.stage_bytes; i += 192 * 16)
    *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(accFull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  auto decode = [&](long long it, int& nn, int& d0, int& h0, int& w0) {
    const int tw = static_cast<int>(it % p.tiles_w);
    it /= p.tiles_w;
    const int th = static_cast<int>(it % p.tiles_h);
    it /= p.tiles_h;
    const int td = static_cast<int>(it % p.tiles_d);
    nn = static_cast<int>(it / p.tiles_d);
    d0 = td * p.DT;
    h0 = p.flat ? 0 : th * 16;
    w0 = p.flat ? 0 : tw * 8;
  };

  if (warp == 0) {
    if (lane == 0 && it0 < it1) {
      tma_prefetch_desc(&tmXs.m[0]);
      tma_prefetch_desc(&tmY);
      int L = 0;
      for (long long it = it0; it < it1; ++it, ++L) {
        const int s = L % p.stages;
        mbar_wait(&empty[s], ((L / p.stages) & 1) ^ 1);
        int nn, d0, h0, w0;
        decode(it, nn, d0, h0, w0);
        uint8_t* st = smem + static_cast<size_t>(s) * p.stage_bytes;
        mbar_arrive_expect_tx(&full[s], p.bytesX * p.nplanes + p.bytesY);
        for (int j = 0; j < p.nplanes; ++j) {
          const int a = p.asplit > 1 ? asel : p.plane_a[j];
          if (p.gather2)   // sub-lattice j of the fine grid, no halo
            tma_load_5d(st + static_cast<size_t>(j) * p.slotX, &tmXs.m[j], &full[s], chunk * p.KC, w0, h0, d0, nn);
          else
            tma_load_5d(st + static_cast<size_t>(j) * p.slotX, &tmXs.m[0], &full[s], chunk * p.KC, w0 - p.padw,
                        h0 - p.padh, d0 - p.padd + (p.flat ? 0 : a * p.dil), nn);
        }
        tma_load_5d(st + static_cast<size_t>(p.nplanes) * p.slotX, &tmY, &full[s], nt * p.NT, w0, h0, d0, nn);
      }
    }
  } else if (warp == 1) {
    if (it0 < it1) {  // warp-uniform issue loop; only the elected lane's tcgen05 instructions take effect
      const uint32_t leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(128, p.NT, 1, 1);
      const uint32_t a_sbo = static_cast<uint32_t>(p.a_sbo_rows) * p.rowbytesA;
      const uint32_t a_adv = static_cast<uint32_t>(p.a_kadv_rows) * p.rowbytesA;
      const uint32_t b_sbo = 8u * p.rowbytesB, b_adv = 16u * p.rowbytesB;
      const uint32_t b_hi = static_cast<uint32_t>(make_smem_desc(0, 16, b_sbo, p.swzB) >> 32);
      const uint32_t a_adv16 = a_adv >> 4, b_adv16 = b_adv >> 4;
      const uint32_t kd_rows = (p.flat && p.asplit > 1) ? static_cast<uint32_t>(asel * p.dil * p.HB * p.WB) : 0u;
      // per-group constants: descriptor high word (LBO differs per group) and byte offset inside a stage
      uint32_t g_hi[kMaxGroups], g_off[kMaxGroups], g_lbo[kMaxGroups];
      for (int g = 0; g < p.ngroups; ++g) {
        const WgradGroup& G = p.groups[g];
        g_off[g] = G.plane * p.slotX + (static_cast<uint32_t>(G.base_rows) + kd_rows) * p.rowbytesA;
        const uint64_t d = make_smem_desc(0, static_cast<uint32_t>(G.lbo_rows) * p.rowbytesA, a_sbo, p.swzA);
        g_hi[g] = static_cast<uint32_t>(d >> 32);
        g_lbo[g] = static_cast<uint32_t>(d) & 0x3FFF0000u;
      }
      int s = 0;
      uint32_t phase = 0, accum = 0;
      for (long long it = it0; it < it1; ++it) {
        mbar_wait(&full[s], phase);
        tc_fence_after();
        // (lane-0 broadcast: warp-uniform hint so the descriptor arithmetic below runs on the uniform datapath)
        const uint32_t st = __shfl_sync(0xffffffffu, smem_u32(smem + static_cast<size_t>(s) * p.stage_bytes), 0);
        const uint32_t y_lo0 = (((st + p.nplanes * p.slotX) >> 4) & 0x3FFF) | (1u << 16);
        for (int g = 0; g < p.ngroups; ++g) {
          uint32_t a_lo = (((st + g_off[g]) >> 4) & 0x3FFF) | g_lbo[g];
          uint32_t y_lo = y_lo0;
          const uint32_t d_tmem = tbase + g * p.NT;
          for (int ks = 0; ks < p.ksteps; ++ks) {
            umma_f16_pred_lohi(d_tmem, a_lo, g_hi[g], y_lo, b_hi, idesc, accum | (ks != 0 ? 1u : 0u), leader);
            a_lo += a_adv16;
            y_lo += b_adv16;
          }
        }
        accum = 1;
        umma_commit_pred(&empty[s], leader);
        if (++s == p.stages) {
          s = 0;
          phase ^= 1;
        }
      }
      umma_commit_pred(accFull, leader);
    }
  } else if (it0 < it1) {
    // =========================== epilogue: TMEM -> red.global.add.f32 ===========================
    const int q = warp & 3;           // warps 2..5 -> lane quadrants 2,3,0,1
    const int m = q * 32 + lane;      // M row = block * KC + ci
    const int blk = m / p.KC, ci = chunk * p.KC + (m % p.KC);
    mbar_wait(accFull, 0);
    tc_fence_after();
    for (int g = 0; g < p.ngroups; ++g) {
      const int tap_local = p.groups[g].tap[blk];
      // taps split over CTAs: table holds (kh,kw) flattened, add this CTA's kd
      const int tap = tap_local < 0 ? -1 : (p.asplit > 1 ? asel * p.k * p.k + tap_local : tap_local);
      for (int cc = 0; cc < p.NT; cc += 32) {
        const int ncol = min(32, p.NT - cc);
        uint32_t raw[32];
        const uint32_t taddr = tbase + (static_cast<uint32_t>(q * 32) << 16) + g * p.NT + cc;
        if (ncol == 32) {
          tmem_ld_32x32(taddr, raw);
        } else {
          uint32_t r16[16];
          tmem_ld_32x16(taddr, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) raw[j] = r16[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) raw[j] = 0u;
        }
        tmem_ld_wait();
        if (tap >= 0 && ci < p.cin) {
          float* dst = p.dwp + (static_cast<size_t>(tap) * p.cin + ci) * p.cout + nt * p.NT + cc;
          if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (ncol & 3) == 0) {
            // 4-wide vector reductions (REDG.E.ADD.F32x4): a quarter of the L2 atomic operations
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncol)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(raw[j])),
                             "f"(__uint_as_float(raw[j + 1])), "f"(__uint_as_float(raw[j + 2])),
                             "f"(__uint_as_float(raw[j + 3]))
                             : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) atomicAdd(dst + j, __uint_as_float(raw[j]));
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

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_w() {
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

static bool encode5(CUtensorMap* tm, const void* base, int c, int w, int h, int d, int n, long long pitch,
                    const uint32_t* box, int chans_per_row) {
  PFN_encodeTiled enc = get_encode_w();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return false;
  }
  const cuuint64_t gd[5] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                            static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(n)};
  const cuuint64_t pb = static_cast<cuuint64_t>(pitch) * 2;
  const cuuint64_t gs[4] = {pb, pb * w, pb * w * h, pb * w * h * d};
  const cuuint32_t bx[5] = {box[0], box[1], box[2], box[3], box[4]};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = chans_per_row == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : chans_per_row == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gd, gs, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return false;
  }
  return true;
}

// Every second voxel of a fine grid [n, fd, fh, fw, c] seen as a coarse tensor [n, d, h, w, c].
static bool encode5_strided(CUtensorMap* tm, const void* base, int c, int w, int h, int d, int n, long long pitch,
                            int fw, int fh, int fd, const uint32_t* box, int chans_per_row) {
  PFN_encodeTiled enc = get_encode_w();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return false;
  }
  const cuuint64_t gd[5] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                            static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(n)};
  const cuuint64_t pb = static_cast<cuuint64_t>(pitch) * 2;
  const cuuint64_t gs[4] = {2 * pb, 2 * pb * fw, 2 * pb * fw * fh, pb * fw * fh * fd};
  const cuuint32_t bx[5] = {box[0], box[1], box[2], box[3], box[4]};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = chans_per_row == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : chans_per_row == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gd, gs, bx, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return false;
  }
  return true;
}

static bool plan_wgrad(const UmmaWgradArgs& a, WgradParams& p, size_t& smem_bytes) {
  if (getenv("B200SEG_DISABLE_UMMA") || getenv("B200SEG_DISABLE_UMMA_WGRAD")) return false;
  if (a.cin % 16 || a.cout % 16) return false;   // ragged last 32-channel chunk: TMA zero fill + masked epilogue rows
  if (a.x_pitch % 8 || a.dy_pitch % 8) return false;
  const int halo = a.gather2 ? 0 : (a.k - 1) * a.dil;
  if (a.sub) {
    if (a.gather2 || a.dil != 1 || a.kd < 1 || a.kd > 2 || a.kh < 1 || a.kh > 2 || a.kw < 1 || a.kw > 2) return false;
  } else if (a.gather2) {
    if (a.k != 2 || a.pad != 0 || a.dil != 1 || a.d != 2 * a.od || a.h != 2 * a.oh || a.w != 2 * a.ow) return false;
  } else {
    if (!(a.k == 1 || a.k == 3 || a.k == 5) || a.dil < 1 || a.pad < 0) return false;
    if (a.od != a.d + 2 * a.pad - halo || a.oh != a.h + 2 * a.pad - halo || a.ow != a.w + 2 * a.pad - halo) return false;
  }
  p = WgradParams{};
  p.n = a.n; p.od = a.od; p.oh = a.oh; p.ow = a.ow; p.cin = a.cin; p.cout = a.cout;
  p.k = a.k; p.pad = a.pad; p.dil = a.dil;
  p.gather2 = a.gather2;
  p.padd = a.sub ? a.pd : a.pad;
  p.padh = a.sub ? a.ph : a.pad;
  p.padw = a.sub ? a.pw : a.pad;
  p.KC = (a.cin % 64 == 0 && !a.gather2) ? 64 : 32;   // gather mode keeps 8 boxes per stage: use the narrow chunk
  p.MB = 128 / p.KC;
  p.nchunks = (a.cin + p.KC - 1) / p.KC;
  p.NT = a.cout % 64 == 0 ? 64 : (a.cout % 32 == 0 ? 32 : 16);
  p.n_ntiles = a.cout / p.NT;
  p.rowbytesA = p.KC * 2;
  p.rowbytesB = p.NT * 2;
  p.swzA = p.KC == 64 ? SWZ_128B : SWZ_64B;
  p.swzB = p.NT == 64 ? SWZ_128B : (p.NT == 32 ? SWZ_64B : SWZ_32B);
  const int k = a.k;

  // (class launches always tile planes: rows past a short extent are zero-filled dy rows and contribute nothing)
  p.flat = !(a.oh >= 16 && a.ow >= 8) && !a.sub;
  if (!p.flat) {
    p.WB = 8 + (a.sub ? a.kw - 1 : halo);
    p.HB = 16 + (a.sub ? a.kh - 1 : halo);
    p.DT = 1;
    p.UPX = 1;
    p.tiles_w = (a.ow + 7) / 8;
    p.tiles_h = (a.oh + 15) / 16;
    p.tiles_d = a.od;
    p.ksteps = 8;
    p.a_sbo_rows = p.WB;
    p.a_kadv_rows = 2 * p.WB;
  } else {
    p.WB = a.ow + halo;
    p.HB = a.oh + halo;
    p.tiles_w = p.tiles_h = 1;
    p.a_sbo_rows = 8;
    p.a_kadv_rows = 16;
  }
  if (p.WB > 256 || p.HB > 256) return false;
  const int plane_rows = p.WB * p.HB;

  // ---- tap groups for ONE kd plane (kh,kw flattened), then replicated over kd unless the taps are split over CTAs
  struct G2 { int base, lbo, tap[8]; };
  G2 per_plane[16];
  int gpp = 0;
  const int kh_ = a.sub ? a.kh : k, kw_ = a.sub ? a.kw : k, kd_ = a.sub ? a.kd : k;   // taps per dimension
  if (a.gather2) {
    // filled in below: groups are runs of MB consecutive sub-lattice boxes (LBO = one box)
  } else if (p.KC == 64) {
    for (int i = 0; i < kh_ * kw_; i += 2) {
      G2 g{};
      for (int j = 0; j < 8; ++j) g.tap[j] = -1;
      const int b0 = i / kw_, e0 = i % kw_;
      g.base = (b0 * a.dil) * p.WB + e0 * a.dil;
      g.tap[0] = i;
      g.lbo = 1;
      if (i + 1 < kh_ * kw_) {
        const int b1 = (i + 1) / kw_, e1 = (i + 1) % kw_;
        g.lbo = (b1 * a.dil) * p.WB + e1 * a.dil - g.base;
        g.tap[1] = i + 1;
      }
      if (gpp >= 16) return false;
      per_plane[gpp++] = g;
    }
  } else {
    for (int b = 0; b < kh_; ++b)
      for (int e0 = 0; e0 < kw_; e0 += 4) {
        G2 g{};
        for (int j = 0; j < 8; ++j) g.tap[j] = -1;
        g.base = (b * a.dil) * p.WB + e0 * a.dil;
        g.lbo = a.dil;
        for (int j = 0; j < 4 && e0 + j < kw_; ++j) g.tap[j] = b * kw_ + e0 + j;
        if (gpp >= 16) return false;
        per_plane[gpp++] = g;
      }
  }
  p.asplit = (a.gather2 || a.sub) ? 1 : ((k * gpp * p.NT <= 512 && k * gpp <= kMaxGroups) ? 1 : k);
  if (gpp * p.NT > 512 || gpp > kMaxGroups) return false;
  if (a.sub && (kd_ * gpp * p.NT > 512 || kd_ * gpp > kMaxGroups)) return false;
  const int planes_in_cta = a.gather2 ? 0 : (p.asplit == 1 ? kd_ : 1);
  p.ngroups = 0;
  for (int pa = 0; pa < planes_in_cta; ++pa)
    for (int gi = 0; gi < gpp; ++gi) {
      WgradGroup& G = p.groups[p.ngroups++];
      G.plane = p.flat ? 0 : pa;
      G.base_rows = per_plane[gi].base + (p.flat ? pa * a.dil * plane_rows : 0);
      G.lbo_rows = per_plane[gi].lbo;
      for (int j = 0; j < 8; ++j) {
        const int local = per_plane[gi].tap[j] < 0 ? -1 : (p.asplit == 1 ? pa * kh_ * kw_ : 0) + per_plane[gi].tap[j];
        G.tap[j] = (a.sub && local >= 0) ? a.tapmap[local] : local;   // class launches: straight to the 27-tap index
      }
    }
  for (int pa = 0; pa < 8; ++pa) p.plane_a[pa] = pa;
  unsigned cols = 32;
  while (cols < static_cast<unsigned>(p.ngroups * p.NT)) cols <<= 1;
  if (cols > 512) return false;
  p.tmem_cols = cols;

  const size_t budget = 200 * 1024;
  if (a.gather2) {
    if ((8 / p.MB) * p.NT > 512) return false;
    unsigned cols2 = 32;
    while (cols2 < static_cast<unsigned>((8 / p.MB) * p.NT)) cols2 <<= 1;
    p.tmem_cols = cols2;
  }
  if (!p.flat) {
    p.nplanes = a.gather2 ? 8 : planes_in_cta;
    p.slotX = (static_cast<unsigned>(plane_rows) * p.rowbytesA + 1023) & ~1023u;
    p.bytesX = static_cast<unsigned>(plane_rows) * p.rowbytesA;
    p.slotY = (128u * p.rowbytesB + 1023) & ~1023u;
    p.bytesY = 128u * p.rowbytesB;
  } else {
    p.nplanes = a.gather2 ? 8 : 1;
    const int maxoff = halo * (plane_rows + p.WB + 1);
    int best = 0;
    for (int DT = std::min(a.od, 255 - halo); DT >= 1; --DT) {
      const int rows_k = (DT * plane_rows + 15) / 16 * 16;
      const size_t sx = ((static_cast<size_t>(rows_k) + maxoff + 8) * p.rowbytesA + 1023) & ~size_t(1023);
      const size_t sy = (static_cast<size_t>(rows_k) * p.rowbytesB + 1023) & ~size_t(1023);
      if (static_cast<size_t>(DT + halo) * plane_rows * p.rowbytesA > sx) continue;
      if (2 * (p.nplanes * sx + sy) + 1024 <= budget) {
        best = DT;
        p.slotX = static_cast<unsigned>(sx);
        p.slotY = static_cast<unsigned>(sy);
        p.ksteps = rows_k / 16;
        break;
      }
    }
    if (!best) return false;
    p.DT = best;
    p.UPX = best + halo;
    p.tiles_d = (a.od + best - 1) / best;
    p.bytesX = static_cast<unsigned>(p.UPX) * plane_rows * p.rowbytesA;
    p.bytesY = static_cast<unsigned>(p.DT) * plane_rows * p.rowbytesB;
  }
  if (a.gather2) {
    // groups of MB consecutive sub-lattice boxes; M-block stride = one box slot
    for (int g0 = 0; g0 < 8; g0 += p.MB) {
      WgradGroup& G = p.groups[p.ngroups++];
      G.plane = g0;
      G.base_rows = 0;
      G.lbo_rows = static_cast<int>(p.slotX / p.rowbytesA);
      for (int j = 0; j < 8; ++j) G.tap[j] = j < p.MB ? g0 + j : -1;
    }
  }
  p.stage_bytes = p.nplanes * p.slotX + p.slotY;
  p.stages = static_cast<int>(std::min<size_t>(4, budget / p.stage_bytes));
  if (p.stages < 2) return false;
  p.items = static_cast<long long>(a.n) * p.tiles_d * p.tiles_h * p.tiles_w;
  const long long combos = static_cast<long long>(p.asplit) * p.nchunks * p.n_ntiles;
  long long splits = (2LL * kNumSMs + combos - 1) / combos;  // ~2 waves worth of CTAs, 1 CTA per SM resident
  if (combos >= kNumSMs) splits = 1;
  else splits = std::max<long long>(1, kNumSMs / combos);
  splits = std::min(splits, p.items);
  p.splits = static_cast<int>(splits);
  smem_bytes = static_cast<size_t>(p.stages) * p.stage_bytes + 1024 + 256;
  return smem_bytes <= 227 * 1024 && combos * splits <= 2147483647LL;
}

bool wgrad_umma_supported(const UmmaWgradArgs& a) {
  if (!a.sub && wgrad_umma_plane_supported(a)) return true;
  WgradParams p;
  size_t smem;
  return plan_wgrad(a, p, smem) || (!a.sub && !a.gather2 && wgrad_umma_plane_relaxed_supported(a));
}

int wgrad_umma_run(const UmmaWgradArgs& a, cudaStream_t st) {
  if (!a.sub && wgrad_umma_plane_supported(a)) return wgrad_umma_plane_run(a, st);   // wgrad_umma_p.cu
  WgradParams p;
  size_t smem;
  if (!plan_wgrad(a, p, smem)) {
    if (!a.sub && !a.gather2 && wgrad_umma_plane_relaxed_supported(a)) return wgrad_umma_plane_run(a, st);
    set_error("wgrad_umma_run: unsupported geometry");
    return B200SEG_ERR_INVALID;
  }
  p.dwp = a.dwp;
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.dy)) & 15) {
    set_error("wgrad_umma_run: buffers must be 16-byte aligned");
    return B200SEG_ERR_INVALID;
  }
  TensorMaps8W tmXs;
  CUtensorMap tmY;
  if (a.sub) {
    // x = sub-lattice `cls` of the fine grid: its own extents, doubled strides, shifted base
    const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB), 1u, 1u};
    const long long off = ((static_cast<long long>(a.cls >> 2) * a.fh + ((a.cls >> 1) & 1)) * a.fw + (a.cls & 1)) * a.x_pitch;
    if (!encode5_strided(&tmXs.m[0], static_cast<const __nv_bfloat16*>(a.x) + off, a.cin, a.w, a.h, a.d, a.n, a.x_pitch,
                         a.fw, a.fh, a.fd, box, p.KC))
      return B200SEG_ERR_CUDA;
    for (int i = 1; i < 8; ++i) tmXs.m[i] = tmXs.m[0];
  } else if (!a.gather2) {
    const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB),
                             static_cast<uint32_t>(p.UPX), 1u};
    if (!encode5(&tmXs.m[0], a.x, a.cin, a.w, a.h, a.d, a.n, a.x_pitch, box, p.KC)) return B200SEG_ERR_CUDA;
    for (int i = 1; i < 8; ++i) tmXs.m[i] = tmXs.m[0];
  } else {
    // sub-lattice g = (a,b,e) of the fine grid: coarse extents, doubled strides, shifted base
    const uint32_t box[5] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.WB), static_cast<uint32_t>(p.HB),
                             static_cast<uint32_t>(p.flat ? p.DT : 1), 1u};
    for (int g = 0; g < 8; ++g) {
      const long long off = ((static_cast<long long>(g >> 2) * a.h + ((g >> 1) & 1)) * a.w + (g & 1)) * a.x_pitch;
      if (!encode5_strided(&tmXs.m[g], static_cast<const __nv_bfloat16*>(a.x) + off, a.cin, a.ow, a.oh, a.od, a.n,
                           a.x_pitch, a.w, a.h, a.d, box, p.KC))
        return B200SEG_ERR_CUDA;
    }
  }
  {
    uint32_t box[5] = {static_cast<uint32_t>(p.NT), 8u, 16u, 1u, 1u};
    if (p.flat) {
      box[1] = p.WB;
      box[2] = p.HB;
      box[3] = p.DT;
    }
    if (!encode5(&tmY, a.dy, a.cout, a.ow, a.oh, a.od, a.n, a.dy_pitch, box, p.NT)) return B200SEG_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("wgrad_umma_run: cannot raise the dynamic shared memory limit");
      return B200SEG_ERR_CUDA;
    }
    attr_set = true;
  }
  const int ctas = p.asplit * p.nchunks * p.n_ntiles * p.splits;
  wgrad_umma_kernel<<<ctas, 192, smem, st>>>(tmXs, tmY, p);
  B200_CHECK_LAUNCH("wgrad_umma");
  ++g_umma_launches;
  return 0;
}

}  // namespace b200
