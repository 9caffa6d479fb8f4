// Internal interfaces between the convolution dispatcher (conv.cu) and its two back ends.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200seg.h"

namespace b200 {

typedef b200seg_conv_geom ConvGeom;

// direct CUDA-core path (conv_direct.cu)
int conv_direct_fprop(const ConvGeom& g, const void* x, int64_t x_pitch, const void* wp, const float* bias, void* y,
                      int64_t y_pitch, float* stats, cudaStream_t st);
int conv_direct_dgrad(const ConvGeom& g, const void* dy, int64_t dy_pitch, const void* wd, const float* bias, void* dx,
                      int64_t dx_pitch, cudaStream_t st);
int conv_direct_wgrad(const ConvGeom& g, const void* x, int64_t x_pitch, const void* dy, int64_t dy_pitch, float* dwp,
                      cudaStream_t st);

// C_in = 1 stem helpers (conv_direct.cu): taps-as-channels im2col [v][32] bf16 of a 3x3x3 / pad 1 kernel, dst += src
int stem_im2col_k3(const void* x, int64_t x_pitch, void* xcol, int n, int d, int h, int w, cudaStream_t st);
int add_f32(float* dst, const float* src, int n, cudaStream_t st);

// tcgen05 implicit-GEMM path (conv_umma.cu).  "Conv form": out[o][n] = sum_{tap,k} in[o - pad + tap*dil][k] * W[tap][n][k]
// with stride 1; fprop uses it directly, dgrad uses it with the flipped pack and pad' = dil*(k-1) - pad.
struct UmmaConvArgs {
  int n, d, h, w;        // input extents
  int od, oh, ow;        // output extents
  int cin, cout;         // K and N of the implicit GEMM
  int k, pad, dil;
  const void* in;
  int64_t in_pitch;
  const void* wpack;     // [k^3][cout][cin] bf16
  const float* bias;     // may be null
  void* out;
  int64_t out_pitch;
  float* stats;          // may be null: {sum[cout], sumsq[cout]}
  // ConvTranspose3d(k2,s2) forward as a pointwise GEMM with N = 8*cout_t and a 2x2x2 pixel-shuffle scatter epilogue:
  // column n -> flipped tap t' = n / cout_t (offset abe = 7 - t'), channel n % cout_t; out is [n,2od,2oh,2ow,cout_t].
  int scatter_cout;      // 0 = ordinary epilogue
  // ConvTranspose3d(k2,s2) backward-data as a pointwise GEMM with K = 8*cin_each: K-chunk group g reads the strided
  // sub-lattice (2v + abe_g) of `in` through its own tensor map.  0 = single input map.
  int gather2;           // 1 = `in` is [n,2d,2h,2w,cin/8] and K runs over (abe, channel)
  // Stride-2 decompositions of a 3x3x3 / pad 1 convolution (conv.cu: k3s2_*; residual_unet3d.py:29-44).  tapmode = 1:
  // the kernel is 2x2x2 SHIFTED taps (k must be 2, dil 1, `pad` = front padding only) of which only some exist per
  // class; tapw[class][a*4 + b*2 + e] is the index of the weight tile (in a pack of `wtaps` tiles) or 0xFF.
  //   in_sub = 1 (forward): `in` is the fine grid [n, fd, fh, fw, cin/8]; K-chunk group g (class g) reads the sub-lattice
  //       of parity bits g = (d<<2 | h<<1 | w) through its own strided tensor map; d, h, w are the output extents.
  //   out_sub = 1 (data gradient, one launch per class): `out` is the sub-lattice of parity bits `cls` of the fine grid
  //       [n, fd, fh, fw, cout]; od, oh, ow are that sub-lattice's extents and class `cls` selects the row of tapw.
  // Fused epilogue (inference: eval-mode BatchNorm folded into the convolution, unet3d.py:80-101 in eval()):
  //   out = act(scale[c] * conv + bias[c]) with act in {none, ReLU, LeakyReLU(slope)}; `bias` then carries the folded
  //   shift (beta - mean*scale + scale*conv_bias).  scale == nullptr: plain conv + bias.  Not combined with `stats`.
  const float* scale;
  int act;
  float slope;
  int tapmode, in_sub, out_sub, cls, wtaps;
  int fd, fh, fw;
  unsigned char tapw[8][8];
  // caller-provided scratch (b200seg_conv3d_ws_bytes): the weights-stationary kernel for K-heavy layers on 8 x 8 planes
  // accumulates the partial sums of its tap rows in an fp32 copy of the output (conv_umma_ws.cu).  null: not available.
  void* ws;
  size_t ws_bytes;
};
extern long long g_umma_launches;
bool conv_umma_supported(const UmmaConvArgs& a);
int conv_umma_run(const UmmaConvArgs& a, cudaStream_t st);
// narrow-output kernel with the kd taps stacked in N and rolling accumulators along d (conv_umma_roll.cu)
bool conv_umma_roll_supported(const UmmaConvArgs& a);
int conv_umma_roll_run(const UmmaConvArgs& a, cudaStream_t st);
// persistent plane-mode kernel (conv_umma_p.cu); conv_umma_run dispatches to it when the geometry allows
bool conv_umma_plane_supported(const UmmaConvArgs& a);
bool conv_umma_plane_relaxed_supported(const UmmaConvArgs& a);
// CTA-pair (cta_group::2) variant for the K-heavy deep layers (conv_umma_p2.cu): two SMs share one copy of the weight tile
// weights-stationary kernel for K-heavy layers on tiny grids (conv_umma_ws.cu): units = (kd, kh) tap row x N tile
size_t conv_umma_ws_bytes(const UmmaConvArgs& a);     // scratch it needs for this geometry, 0 = not applicable
bool conv_umma_ws_supported(const UmmaConvArgs& a);
int conv_umma_ws_run(const UmmaConvArgs& a, cudaStream_t st);
bool conv_umma_pair_supported(const UmmaConvArgs& a);
int conv_umma_pair_run(const UmmaConvArgs& a, cudaStream_t st);   // short planes (masked tile rows): last resort
int conv_umma_plane_run(const UmmaConvArgs& a, cudaStream_t st);

struct UmmaWgradArgs {
  int n, d, h, w, od, oh, ow, cin, cout, k, pad, dil;
  const void* x;
  int64_t x_pitch;
  const void* dy;
  int64_t dy_pitch;
  float* dwp;            // [k^3][cin][cout] fp32, accumulated into
  // ConvTranspose3d(k2,s2) weight gradient: x is the FINE grid [n,2od,2oh,2ow,cin] (the up-sampled gradient), dy the
  // coarse grid [n,od,oh,ow,cout]; tap abe pairs coarse voxel v with fine voxel 2v+abe.  k must be 2, pad 0.
  int gather2;
  // Optional split-K workspace: when non-null (and large enough, see wgrad_umma_plane_workspace_bytes) every CTA stores
  // its partial tile with plain vector stores into slice `split` of the workspace and a second kernel sums the slices
  // into dwp, instead of every CTA adding its tile to dwp with fp32 atomics.
  float* partial;
  size_t partial_bytes;
  // Stride-2 3x3x3 / pad 1 weight gradient, one launch per input-parity class (conv.cu: k3s2_wgrad): sub = 1 means x is
  // the sub-lattice `cls` (parity bits d<<2 | h<<1 | w) of the fine grid [n, fd, fh, fw, cin]; d, h, w are that
  // sub-lattice's extents.  The class sees a (kd, kh, kw) in {1,2}^3 kernel with front padding (pd, ph, pw) only;
  // tapmap[(a*kh + b)*kw + e] is the tap index inside dwp ([27][cin][cout]).
  int sub, cls, fd, fh, fw, kd, kh, kw, pd, ph, pw;
  int tapmap[8];
};
bool wgrad_umma_supported(const UmmaWgradArgs& a);
int wgrad_umma_run(const UmmaWgradArgs& a, cudaStream_t st);
// persistent plane-mode kernel with tap blocks on both operands (wgrad_umma_p.cu)
bool wgrad_umma_plane_supported(const UmmaWgradArgs& a);
bool wgrad_umma_plane_relaxed_supported(const UmmaWgradArgs& a);
size_t wgrad_umma_plane_workspace_bytes(const UmmaWgradArgs& a);
int wgrad_umma_plane_run(const UmmaWgradArgs& a, cudaStream_t st);

}  // namespace b200
