// Convolution entry points of the C ABI: validate, pick the back end (tcgen05 implicit GEMM or direct), launch.
// ConvTranspose3d(k=2,s=2) is the exact transpose of a k=2 s=2 convolution whose weight tensor has the same memory
// layout ([C_in_T][C_out_T][2][2][2] == [cout][cin][k^3] of the strided conv), so its three passes reuse the conv code.
#include <cstdlib>

#include "common.cuh"
#include "conv_impl.h"

using namespace b200;

static int check_geom(const b200seg_conv_geom* g, const char* who) {
  B200_CHECK_ARG(g != nullptr, "%s: null geometry", who);
  B200_CHECK_ARG(g->n > 0 && g->d > 0 && g->h > 0 && g->w > 0 && g->cin > 0 && g->cout > 0, "%s: empty tensor", who);
  B200_CHECK_ARG(g->k >= 1 && g->k <= 7 && g->stride >= 1 && g->dil >= 1 && g->pad >= 0, "%s: bad kernel geometry", who);
  const int ext = g->dil * (g->k - 1) + 1;
  B200_CHECK_ARG(g->od == (g->d + 2 * g->pad - ext) / g->stride + 1 && g->oh == (g->h + 2 * g->pad - ext) / g->stride + 1 &&
                     g->ow == (g->w + 2 * g->pad - ext) / g->stride + 1,
                 "%s: output extents (%d,%d,%d) do not match the geometry", who, g->od, g->oh, g->ow);
  B200_CHECK_ARG(g->od > 0 && g->oh > 0 && g->ow > 0, "%s: empty output", who);
  return 0;
}

static UmmaConvArgs fprop_args(const b200seg_conv_geom* g, const void* x, int64_t xp, const void* w, const float* bias,
                               void* y, int64_t yp, float* stats) {
  UmmaConvArgs a{};
  a.n = g->n; a.d = g->d; a.h = g->h; a.w = g->w; a.od = g->od; a.oh = g->oh; a.ow = g->ow;
  a.cin = g->cin; a.cout = g->cout; a.k = g->k; a.pad = g->pad; a.dil = g->dil;
  a.in = x; a.in_pitch = xp; a.wpack = w; a.bias = bias; a.out = y; a.out_pitch = yp; a.stats = stats;
  return a;
}
static UmmaConvArgs dgrad_args(const b200seg_conv_geom* g, const void* dy, int64_t dyp, const void* wd, void* dx,
                               int64_t dxp, float* stats) {
  UmmaConvArgs a{};
  a.n = g->n; a.d = g->od; a.h = g->oh; a.w = g->ow; a.od = g->d; a.oh = g->h; a.ow = g->w;
  a.cin = g->cout; a.cout = g->cin; a.k = g->k; a.pad = g->dil * (g->k - 1) - g->pad; a.dil = g->dil;
  a.in = dy; a.in_pitch = dyp; a.wpack = wd; a.bias = nullptr; a.out = dx; a.out_pitch = dxp; a.stats = stats;
  return a;
}

// Weight gradient of the C_in = 1 stem (unet3d.py:80) on the tensor cores: im2col of the 27 taps into 32 channels, then
// the 1x1x1 / 32-channel weight gradient.  Workspace: [xcol bf16 rows x 32][tmp fp32 32 x cout][inner split-K partials].
struct StemPlan {
  size_t xcol_bytes, tmp_off, inner_off, inner_bytes, total;
  int64_t rows;
};
static UmmaWgradArgs stem_wgrad_args(const b200seg_conv_geom* g, const void* xcol, const void* dy, int64_t dyp, float* tmp,
                                     float* inner, size_t inner_bytes) {
  return UmmaWgradArgs{g->n, g->d, g->h, g->w, g->od, g->oh, g->ow, 32, g->cout, 1, 0, 1,
                       xcol, 32, dy, dyp, tmp, 0, inner, inner_bytes};
}
static bool stem_plan(const b200seg_conv_geom* g, StemPlan& sp) {
  if (getenv("B200SEG_DISABLE_STEM_IM2COL")) return false;
  if (!(g->cin == 1 && g->k == 3 && g->stride == 1 && g->pad == 1 && g->dil == 1 && g->cout % 16 == 0 && g->cout <= 64))
    return false;
  sp.rows = static_cast<int64_t>(g->n) * g->d * g->h * g->w;
  if (sp.rows < (1 << 16)) return false;     // small volumes: the CUDA-core kernel is launch-bound either way
  const UmmaWgradArgs a = stem_wgrad_args(g, nullptr, nullptr, g->cout, nullptr, nullptr, 0);
  if (!wgrad_umma_plane_supported(a)) return false;
  sp.xcol_bytes = (static_cast<size_t>(sp.rows) * 32 * 2 + 255) & ~size_t(255);
  sp.tmp_off = sp.xcol_bytes;
  sp.inner_off = sp.tmp_off + ((static_cast<size_t>(32) * g->cout * sizeof(float) + 255) & ~size_t(255));
  sp.inner_bytes = wgrad_umma_plane_workspace_bytes(a);
  sp.total = sp.inner_off + sp.inner_bytes;
  return true;
}

// ---- Conv3d(k=2, s=2, p=0) (V-Net down-convolutions, vnet3d.py:65) on the tensor cores: the same three GEMMs as
// ConvTranspose3d(k2,s2), with the roles of the passes exchanged (see the ConvT entry points below).
static bool is_k2s2(const b200seg_conv_geom* g) {
  return g->k == 2 && g->stride == 2 && g->pad == 0 && g->dil == 1 && g->d == 2 * g->od && g->h == 2 * g->oh &&
         g->w == 2 * g->ow && !getenv("B200SEG_DISABLE_K2S2_UMMA");
}
// fprop: K runs over (tap abe, C_in); K-chunk group abe reads the sub-lattice x[2v + abe] through its own strided tensor
// map; the fprop pack [abe][C_out][C_in] is the K-major weight matrix.
static UmmaConvArgs k2s2_fprop_args(const b200seg_conv_geom* g, const void* x, int64_t xp, const void* w, const float* bias,
                                    void* y, int64_t yp, float* stats) {
  UmmaConvArgs a{};
  a.n = g->n; a.d = g->od; a.h = g->oh; a.w = g->ow; a.od = g->od; a.oh = g->oh; a.ow = g->ow;
  a.cin = 8 * g->cin; a.cout = g->cout; a.k = 1; a.pad = 0; a.dil = 1;
  a.in = x; a.in_pitch = xp; a.wpack = w; a.bias = bias; a.out = y; a.out_pitch = yp; a.stats = stats;
  a.scatter_cout = 0; a.gather2 = 1;
  return a;
}
// dgrad: pointwise GEMM [coarse voxels x C_out] . [C_out x 8*C_in] with the pixel-shuffle scatter epilogue; the dgrad pack
// [7 - abe][C_in][C_out] is an [8*C_in][C_out] K-major matrix.
static UmmaConvArgs k2s2_dgrad_args(const b200seg_conv_geom* g, const void* dy, int64_t dyp, const void* wd, void* dx,
                                    int64_t dxp) {
  UmmaConvArgs a{};
  a.n = g->n; a.d = g->od; a.h = g->oh; a.w = g->ow; a.od = g->od; a.oh = g->oh; a.ow = g->ow;
  a.cin = g->cout; a.cout = 8 * g->cin; a.k = 1; a.pad = 0; a.dil = 1;
  a.in = dy; a.in_pitch = dyp; a.wpack = wd; a.bias = nullptr; a.out = dx; a.out_pitch = dxp; a.stats = nullptr;
  a.scatter_cout = g->cin; a.gather2 = 0;
  return a;
}
static UmmaWgradArgs k2s2_wgrad_args(const b200seg_conv_geom* g, const void* x, int64_t xp, const void* dy, int64_t dyp,
                                     float* dwp) {
  return UmmaWgradArgs{g->n, g->d, g->h, g->w, g->od, g->oh, g->ow, g->cin, g->cout, 2, 0, 1, x, xp, dy, dyp, dwp, 1,
                       nullptr, 0};
}

// ---- Conv3d(k=3, s=2, p=1) (residual U-Net context down-steps, residual_unet3d.py:29-44) on the tensor cores.
// Input coordinate i = 2o - 1 + t: tap t = 1 reads the EVEN sub-lattice at index o, taps t = 0 / 2 read the ODD
// sub-lattice at indices o - 1 / o.  Per dimension the parity r of the sub-lattice therefore fixes which taps exist, and
// on every sub-lattice the convolution has stride 1 -- "2x2x2 shifted taps, some absent" (UmmaConvArgs::tapmode):
//   forward: ONE launch; K runs over (parity class, C_in), each class reads its sub-lattice through a strided tensor map;
//   data gradient: one launch per OUTPUT parity class (dx[2m + r] gathers dy[m] (t = 1 | 2) and dy[m + 1] (t = 0));
//   weight gradient: one launch per INPUT parity class of the first-generation kernel with a (1|2)^3 tap table.
static bool is_k3s2(const b200seg_conv_geom* g) {
  return g->k == 3 && g->stride == 2 && g->pad == 1 && g->dil == 1 && g->cin % 16 == 0 && g->cout % 16 == 0 &&
         !getenv("B200SEG_DISABLE_K3S2_UMMA");
}
static UmmaConvArgs k3s2_fprop_args(const b200seg_conv_geom* g, const void* x, int64_t xp, const void* w, const float* bias,
                                    void* y, int64_t yp, float* stats) {
  UmmaConvArgs a{};
  a.n = g->n; a.d = g->od; a.h = g->oh; a.w = g->ow; a.od = g->od; a.oh = g->oh; a.ow = g->ow;
  a.cin = 8 * g->cin; a.cout = g->cout; a.k = 2; a.pad = 1; a.dil = 1;
  a.in = x; a.in_pitch = xp; a.wpack = w; a.bias = bias; a.out = y; a.out_pitch = yp; a.stats = stats;
  a.tapmode = 1; a.in_sub = 1; a.wtaps = 27; a.fd = g->d; a.fh = g->h; a.fw = g->w;
  for (int cls = 0; cls < 8; ++cls)
    for (int sh = 0; sh < 8; ++sh) {
      const int r[3] = {cls >> 2, (cls >> 1) & 1, cls & 1}, sft[3] = {sh >> 2, (sh >> 1) & 1, sh & 1};
      int t[3];
      bool ok = true;
      for (int i = 0; i < 3; ++i) {
        ok = ok && (r[i] == 1 || sft[i] == 1);      // even lattice: only the shift that lands on index o
        t[i] = r[i] ? 2 * sft[i] : 1;
      }
      a.tapw[cls][sh] = ok ? static_cast<unsigned char>((t[0] * 3 + t[1]) * 3 + t[2]) : 0xFF;
    }
  return a;
}
static UmmaConvArgs k3s2_dgrad_args(const b200seg_conv_geom* g, int cls, const void* dy, int64_t dyp, const void* wd, void* dx,
                                    int64_t dxp, float* stats) {
  const int r[3] = {cls >> 2, (cls >> 1) & 1, cls & 1};
  UmmaConvArgs a{};
  a.n = g->n; a.d = g->od; a.h = g->oh; a.w = g->ow;
  a.od = (g->d - r[0] + 1) / 2; a.oh = (g->h - r[1] + 1) / 2; a.ow = (g->w - r[2] + 1) / 2;
  a.cin = g->cout; a.cout = g->cin; a.k = 2; a.pad = 0; a.dil = 1;
  a.in = dy; a.in_pitch = dyp; a.wpack = wd; a.bias = nullptr; a.out = dx; a.out_pitch = dxp; a.stats = stats;
  a.tapmode = 1; a.out_sub = 1; a.cls = cls; a.wtaps = 27; a.fd = g->d; a.fh = g->h; a.fw = g->w;
  for (int c2 = 0; c2 < 8; ++c2)
    for (int sh = 0; sh < 8; ++sh) a.tapw[c2][sh] = 0xFF;
  for (int sh = 0; sh < 8; ++sh) {
    const int j[3] = {sh >> 2, (sh >> 1) & 1, sh & 1};
    int t[3];
    bool ok = true;
    for (int i = 0; i < 3; ++i) {
      ok = ok && (r[i] == 1 || j[i] == 0);
      t[i] = r[i] ? 2 - 2 * j[i] : 1;
    }
    // the data-gradient pack is [26 - tap][C_in][C_out]
    if (ok) a.tapw[cls & 7][sh] = static_cast<unsigned char>(26 - ((t[0] * 3 + t[1]) * 3 + t[2]));
  }
  return a;
}
static UmmaWgradArgs k3s2_wgrad_args(const b200seg_conv_geom* g, int cls, const void* x, int64_t xp, const void* dy,
                                     int64_t dyp, float* dwp) {
  const int r[3] = {cls >> 2, (cls >> 1) & 1, cls & 1};
  UmmaWgradArgs a{};
  a.n = g->n; a.d = (g->d - r[0] + 1) / 2; a.h = (g->h - r[1] + 1) / 2; a.w = (g->w - r[2] + 1) / 2;
  a.od = g->od; a.oh = g->oh; a.ow = g->ow; a.cin = g->cin; a.cout = g->cout; a.k = 2; a.pad = 0; a.dil = 1;
  a.x = x; a.x_pitch = xp; a.dy = dy; a.dy_pitch = dyp; a.dwp = dwp;
  a.sub = 1; a.cls = cls; a.fd = g->d; a.fh = g->h; a.fw = g->w;
  a.kd = 1 + r[0]; a.kh = 1 + r[1]; a.kw = 1 + r[2];     // odd sub-lattice: taps 0 and 2 at indices o - 1 and o
  a.pd = r[0]; a.ph = r[1]; a.pw = r[2];
  for (int ta = 0; ta < a.kd; ++ta)
    for (int tb = 0; tb < a.kh; ++tb)
      for (int te = 0; te < a.kw; ++te) {
        const int t0 = r[0] ? 2 * ta : 1, t1 = r[1] ? 2 * tb : 1, t2 = r[2] ? 2 * te : 1;
        a.tapmap[(ta * a.kh + tb) * a.kw + te] = (t0 * 3 + t1) * 3 + t2;
      }
  return a;
}
static bool k3s2_class_empty(const b200seg_conv_geom* g, int cls) {
  return (g->d - (cls >> 2) + 1) / 2 == 0 || (g->h - ((cls >> 1) & 1) + 1) / 2 == 0 || (g->w - (cls & 1) + 1) / 2 == 0;
}

extern "C" {

int64_t b200seg_umma_launch_count(void) { return g_umma_launches; }

int b200seg_conv3d_uses_tensor_cores(const b200seg_conv_geom* g) {
  if (!g) return 0;
  if (is_k2s2(g)) return conv_umma_supported(k2s2_fprop_args(g, nullptr, g->cin, nullptr, nullptr, nullptr, g->cout, nullptr)) ? 1 : 0;
  if (is_k3s2(g)) return conv_umma_supported(k3s2_fprop_args(g, nullptr, g->cin, nullptr, nullptr, nullptr, g->cout, nullptr)) ? 1 : 0;
  if (g->stride != 1) return 0;
  UmmaConvArgs a = fprop_args(g, nullptr, g->cin, nullptr, nullptr, nullptr, g->cout, nullptr);
  return conv_umma_supported(a) ? 1 : 0;
}

size_t b200seg_conv3d_ws_bytes(const b200seg_conv_geom* g, int dgrad) {
  if (!g || g->stride != 1) return 0;
  if (dgrad && g->dil * (g->k - 1) - g->pad < 0) return 0;
  const UmmaConvArgs a = dgrad ? dgrad_args(g, nullptr, g->cout, nullptr, nullptr, g->cin, nullptr)
                               : fprop_args(g, nullptr, g->cin, nullptr, nullptr, nullptr, g->cout, nullptr);
  return conv_umma_ws_bytes(a);
}

size_t b200seg_conv3d_workspace_bytes(const b200seg_conv_geom* g) {
  // only the weight gradient uses scratch memory: split-K partial tiles (0 = not needed / atomics path), or the
  // taps-as-channels copy of a single-channel input (stem_plan)
  if (!g || g->stride != 1) return 0;
  {
    StemPlan sp;
    if (stem_plan(g, sp)) return sp.total;
  }
  UmmaWgradArgs a{g->n, g->d, g->h, g->w, g->od, g->oh, g->ow, g->cin, g->cout, g->k, g->pad, g->dil,
                  nullptr, g->cin, nullptr, g->cout, nullptr, 0, nullptr, 0};
  return wgrad_umma_plane_workspace_bytes(a);
}

int b200seg_conv3d_fprop(const b200seg_conv_geom* g, const void* x, int64_t x_pitch, const void* w_packed,
                         const float* bias, void* y, int64_t y_pitch, float* stats, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (int rc = check_geom(g, "conv3d_fprop")) return rc;
  B200_CHECK_ARG(x && w_packed && y && x_pitch >= g->cin && y_pitch >= g->cout, "conv3d_fprop: bad buffers");
  auto st = static_cast<cudaStream_t>(stream);
  if (g->stride == 1) {
    UmmaConvArgs a = fprop_args(g, x, x_pitch, w_packed, bias, y, y_pitch, stats);
    a.ws = workspace;
    a.ws_bytes = workspace_bytes;
    if (conv_umma_supported(a)) return conv_umma_run(a, st);
  } else if (is_k2s2(g)) {
    UmmaConvArgs a = k2s2_fprop_args(g, x, x_pitch, w_packed, bias, y, y_pitch, stats);
    if (conv_umma_supported(a)) return conv_umma_run(a, st);
  } else if (is_k3s2(g)) {
    UmmaConvArgs a = k3s2_fprop_args(g, x, x_pitch, w_packed, bias, y, y_pitch, stats);
    if (conv_umma_supported(a)) return conv_umma_run(a, st);
  }
  return conv_direct_fprop(*g, x, x_pitch, w_packed, bias, y, y_pitch, stats, st);
}

int b200seg_conv3d_fprop_act_supported(const b200seg_conv_geom* g) {
  if (!g || g->stride != 1 || getenv("B200SEG_DISABLE_FUSED_ACT")) return 0;
  static const float one = 1.f;
  UmmaConvArgs a = fprop_args(g, nullptr, g->cin, nullptr, nullptr, nullptr, g->cout, nullptr);
  a.scale = &one;      // only its presence matters for the support query
  return conv_umma_supported(a) ? 1 : 0;
}

int b200seg_conv3d_fprop_act(const b200seg_conv_geom* g, const void* x, int64_t x_pitch, const void* w_packed,
                             const float* scale, const float* shift, int act, float slope, void* y, int64_t y_pitch,
                             void* stream) {
  if (int rc = check_geom(g, "conv3d_fprop_act")) return rc;
  B200_CHECK_ARG(x && w_packed && y && scale && shift && x_pitch >= g->cin && y_pitch >= g->cout, "conv3d_fprop_act: bad buffers");
  B200_CHECK_ARG(act == B200SEG_ACT_NONE || act == B200SEG_ACT_RELU || act == B200SEG_ACT_LEAKY,
                 "conv3d_fprop_act: the fused epilogue takes no activation, ReLU or LeakyReLU");
  B200_CHECK_ARG(g->stride == 1, "conv3d_fprop_act: stride-1 convolutions only");
  UmmaConvArgs a = fprop_args(g, x, x_pitch, w_packed, shift, y, y_pitch, nullptr);
  a.scale = scale;
  a.act = act;
  a.slope = slope;
  if (!conv_umma_supported(a)) {
    set_error("conv3d_fprop_act: geometry not supported by the fused tensor-core path (query conv3d_fprop_act_supported)");
    return B200SEG_ERR_INVALID;
  }
  return conv_umma_run(a, static_cast<cudaStream_t>(stream));
}

int b200seg_conv3d_dgrad(const b200seg_conv_geom* g, const void* dy, int64_t dy_pitch, const void* w_packed_dgrad,
                         void* dx, int64_t dx_pitch, float* stats, void* workspace, size_t workspace_bytes,
                         void* stream) {
  if (int rc = check_geom(g, "conv3d_dgrad")) return rc;
  B200_CHECK_ARG(dy && w_packed_dgrad && dx && dy_pitch >= g->cout && dx_pitch >= g->cin, "conv3d_dgrad: bad buffers");
  auto st = static_cast<cudaStream_t>(stream);
  if (g->stride == 1 && g->dil * (g->k - 1) - g->pad >= 0) {
    UmmaConvArgs a = dgrad_args(g, dy, dy_pitch, w_packed_dgrad, dx, dx_pitch, stats);
    a.ws = workspace;
    a.ws_bytes = workspace_bytes;
    if (conv_umma_supported(a)) return conv_umma_run(a, st);
  } else if (is_k2s2(g)) {
    UmmaConvArgs a = k2s2_dgrad_args(g, dy, dy_pitch, w_packed_dgrad, dx, dx_pitch);
    if (conv_umma_supported(a)) {
      if (int rc = conv_umma_run(a, st)) return rc;
      if (stats)   // the pixel-shuffle epilogue has no statistics: one extra pass over dx
        return b200seg_channel_stats(dx, dx_pitch, static_cast<int64_t>(g->n) * g->d * g->h * g->w, 1, g->cin, stats, stream);
      return 0;
    }
  } else if (is_k3s2(g) && conv_umma_supported(k3s2_dgrad_args(g, 7, dy, dy_pitch, w_packed_dgrad, dx, dx_pitch, stats))) {
    for (int cls = 0; cls < 8; ++cls) {     // one launch per parity class of dx; the statistics add up over the launches
      if (k3s2_class_empty(g, cls)) continue;
      const UmmaConvArgs a = k3s2_dgrad_args(g, cls, dy, dy_pitch, w_packed_dgrad, dx, dx_pitch, stats);
      if (!conv_umma_supported(a)) {
        set_error("conv3d_dgrad: stride-2 class %d is not supported by the tensor-core path", cls);
        return B200SEG_ERR_INVALID;
      }
      if (int rc = conv_umma_run(a, st)) return rc;
    }
    return 0;
  }
  if (int rc = conv_direct_dgrad(*g, dy, dy_pitch, w_packed_dgrad, nullptr, dx, dx_pitch, st)) return rc;
  // the direct kernels have no statistics epilogue: one extra pass over dx
  if (stats)
    return b200seg_channel_stats(dx, dx_pitch, static_cast<int64_t>(g->n) * g->d * g->h * g->w, 1, g->cin, stats, stream);
  return 0;
}

int b200seg_conv3d_wgrad(const b200seg_conv_geom* g, const void* x, int64_t x_pitch, const void* dy, int64_t dy_pitch,
                         float* dw_packed, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_geom(g, "conv3d_wgrad")) return rc;
  B200_CHECK_ARG(x && dy && dw_packed && x_pitch >= g->cin && dy_pitch >= g->cout, "conv3d_wgrad: bad buffers");
  auto st = static_cast<cudaStream_t>(stream);
  {
    StemPlan sp;
    if (workspace && stem_plan(g, sp) && workspace_bytes >= sp.total && dy_pitch % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(workspace) & 255) == 0) {
      uint8_t* ws = static_cast<uint8_t*>(workspace);
      float* tmp = reinterpret_cast<float*>(ws + sp.tmp_off);
      if (int rc = stem_im2col_k3(x, x_pitch, ws, g->n, g->d, g->h, g->w, st)) return rc;
      if (cudaMemsetAsync(tmp, 0, static_cast<size_t>(32) * g->cout * sizeof(float), st) != cudaSuccess) {
        set_error("conv3d_wgrad: cudaMemsetAsync failed");
        return B200SEG_ERR_CUDA;
      }
      const UmmaWgradArgs a = stem_wgrad_args(g, ws, dy, dy_pitch, tmp, sp.inner_bytes ? reinterpret_cast<float*>(ws + sp.inner_off) : nullptr,
                                              sp.inner_bytes);
      if (int rc = wgrad_umma_run(a, st)) return rc;
      // tmp is [1][32][cout]; its first 27 rows are dW[tap][0][cout], the layout of dw_packed
      return add_f32(dw_packed, tmp, 27 * g->cout, st);
    }
  }
  if (g->stride == 1) {
    UmmaWgradArgs a{g->n, g->d, g->h, g->w, g->od, g->oh, g->ow, g->cin, g->cout, g->k, g->pad, g->dil,
                    x, x_pitch, dy, dy_pitch, dw_packed, 0, static_cast<float*>(workspace), workspace_bytes};
    if (wgrad_umma_supported(a)) return wgrad_umma_run(a, st);
  } else if (is_k2s2(g)) {
    UmmaWgradArgs a = k2s2_wgrad_args(g, x, x_pitch, dy, dy_pitch, dw_packed);
    if (wgrad_umma_supported(a)) return wgrad_umma_run(a, st);
  } else if (is_k3s2(g) && wgrad_umma_supported(k3s2_wgrad_args(g, 7, x, x_pitch, dy, dy_pitch, dw_packed)) &&
             wgrad_umma_supported(k3s2_wgrad_args(g, 0, x, x_pitch, dy, dy_pitch, dw_packed))) {
    for (int cls = 0; cls < 8; ++cls) {     // one launch per parity class of x; every tap belongs to exactly one class
      if (k3s2_class_empty(g, cls)) continue;
      const UmmaWgradArgs a = k3s2_wgrad_args(g, cls, x, x_pitch, dy, dy_pitch, dw_packed);
      if (!wgrad_umma_supported(a)) {
        set_error("conv3d_wgrad: stride-2 class %d is not supported by the tensor-core path", cls);
        return B200SEG_ERR_INVALID;
      }
      if (int rc = wgrad_umma_run(a, st)) return rc;
    }
    return 0;
  }
  return conv_direct_wgrad(*g, x, x_pitch, dy, dy_pitch, dw_packed, st);
}

// ---- ConvTranspose3d k2 s2 == transpose of the strided conv S: [n,2d,2h,2w,cout_T] -> [n,d,h,w,cin_T] ------------
static b200seg_conv_geom convt_as_conv(int n, int d, int h, int w, int cin_t, int cout_t) {
  b200seg_conv_geom g{};
  g.n = n; g.d = 2 * d; g.h = 2 * h; g.w = 2 * w; g.cin = cout_t;
  g.od = d; g.oh = h; g.ow = w; g.cout = cin_t;
  g.k = 2; g.stride = 2; g.pad = 0; g.dil = 1;
  return g;
}

int b200seg_pack_convt_weight(const float* w, void* packed, int cin, int cout, int dgrad, void* stream) {
  // ConvT weight [cin_T][cout_T][8] is S's weight [cout_S = cin_T][cin_S = cout_T][8].
  // ConvT forward = S dgrad -> needs S's dgrad pack; ConvT dgrad = S fprop -> needs S's fprop pack.
  return b200seg_pack_conv_weight(w, packed, cin, cout, 2, 0, cout, dgrad ? 0 : 1, stream);
}

int b200seg_convt_k2s2_fwd(const void* x, int64_t x_pitch, const void* w_packed, const float* bias, void* y,
                           int64_t y_pitch, int n, int d, int h, int w, int cin, int cout, void* stream) {
  B200_CHECK_ARG(x && w_packed && y && n > 0 && d > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && x_pitch >= cin &&
                     y_pitch >= cout, "convt_k2s2_fwd: bad arguments");
  {
    // tensor-core path: pointwise GEMM [voxels x cin] . [cin x 8*cout] with a pixel-shuffle scatter epilogue.
    // The dgrad pack of the equivalent strided conv is [7 - abe][cout][cin] == a [8*cout][cin] K-major matrix.
    UmmaConvArgs a{};
    a.n = n; a.d = d; a.h = h; a.w = w; a.od = d; a.oh = h; a.ow = w;
    a.cin = cin; a.cout = 8 * cout; a.k = 1; a.pad = 0; a.dil = 1;
    a.in = x; a.in_pitch = x_pitch; a.wpack = w_packed; a.bias = bias; a.out = y; a.out_pitch = y_pitch;
    a.stats = nullptr; a.scatter_cout = cout; a.gather2 = 0;
    if (conv_umma_supported(a)) return conv_umma_run(a, static_cast<cudaStream_t>(stream));
  }
  const b200seg_conv_geom g = convt_as_conv(n, d, h, w, cin, cout);
  return conv_direct_dgrad(g, x, x_pitch, w_packed, bias, y, y_pitch, static_cast<cudaStream_t>(stream));
}

int b200seg_convt_k2s2_dgrad(const void* dy, int64_t dy_pitch, const void* w_packed_dgrad, void* dx, int64_t dx_pitch,
                             int n, int d, int h, int w, int cin, int cout, void* stream) {
  B200_CHECK_ARG(dy && w_packed_dgrad && dx && n > 0 && d > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 &&
                     dy_pitch >= cout && dx_pitch >= cin, "convt_k2s2_dgrad: bad arguments");
  {
    // tensor-core path: K runs over (abe, cout); K-chunk group abe reads the sub-lattice dy[2v + abe] through its
    // own strided tensor map.  The fprop pack of the equivalent strided conv is [abe][cin][cout].
    UmmaConvArgs a{};
    a.n = n; a.d = d; a.h = h; a.w = w; a.od = d; a.oh = h; a.ow = w;
    a.cin = 8 * cout; a.cout = cin; a.k = 1; a.pad = 0; a.dil = 1;
    a.in = dy; a.in_pitch = dy_pitch; a.wpack = w_packed_dgrad; a.bias = nullptr; a.out = dx; a.out_pitch = dx_pitch;
    a.stats = nullptr; a.scatter_cout = 0; a.gather2 = 1;
    if (conv_umma_supported(a)) return conv_umma_run(a, static_cast<cudaStream_t>(stream));
  }
  const b200seg_conv_geom g = convt_as_conv(n, d, h, w, cin, cout);
  return conv_direct_fprop(g, dy, dy_pitch, w_packed_dgrad, nullptr, dx, dx_pitch, nullptr,
                           static_cast<cudaStream_t>(stream));
}

int b200seg_convt_k2s2_wgrad(const void* x, int64_t x_pitch, const void* dy, int64_t dy_pitch, float* dw_packed,
                             int n, int d, int h, int w, int cin, int cout, void* stream) {
  B200_CHECK_ARG(x && dy && dw_packed && n > 0 && d > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && x_pitch >= cin &&
                     dy_pitch >= cout, "convt_k2s2_wgrad: bad arguments");
  // S wgrad: dW_S[tap][cin_S = cout_T][cout_S = cin_T] = sum x_S(*)dy_S with x_S = dy_T, dy_S = x_T.
  const b200seg_conv_geom g = convt_as_conv(n, d, h, w, cin, cout);
  {
    UmmaWgradArgs a{g.n, g.d, g.h, g.w, g.od, g.oh, g.ow, g.cin, g.cout, 2, 0, 1, dy, dy_pitch, x, x_pitch, dw_packed, 1,
                    nullptr, 0};
    if (wgrad_umma_supported(a)) return wgrad_umma_run(a, static_cast<cudaStream_t>(stream));
  }
  return conv_direct_wgrad(g, dy, dy_pitch, x, x_pitch, dw_packed, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
