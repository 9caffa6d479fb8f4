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

extern "C" {

int64_t b200seg_umma_launch_count(void) { return g_umma_launches; }

int b200seg_conv3d_uses_tensor_cores(const b200seg_conv_geom* g) {
  if (!g || g->stride != 1) return 0;
  UmmaConvArgs a = fprop_args(g, nullptr, g->cin, nullptr, nullptr, nullptr, g->cout, nullptr);
  return conv_umma_supported(a) ? 1 : 0;
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
  (void)workspace; (void)workspace_bytes;
  if (int rc = check_geom(g, "conv3d_fprop")) return rc;
  B200_CHECK_ARG(x && w_packed && y && x_pitch >= g->cin && y_pitch >= g->cout, "conv3d_fprop: bad buffers");
  auto st = static_cast<cudaStream_t>(stream);
  if (g->stride == 1) {
    UmmaConvArgs a = fprop_args(g, x, x_pitch, w_packed, bias, y, y_pitch, stats);
    if (conv_umma_supported(a)) return conv_umma_run(a, st);
  }
  return conv_direct_fprop(*g, x, x_pitch, w_packed, bias, y, y_pitch, stats, st);
}

int b200seg_conv3d_dgrad(const b200seg_conv_geom* g, const void* dy, int64_t dy_pitch, const void* w_packed_dgrad,
                         void* dx, int64_t dx_pitch, float* stats, void* workspace, size_t workspace_bytes,
                         void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (int rc = check_geom(g, "conv3d_dgrad")) return rc;
  B200_CHECK_ARG(dy && w_packed_dgrad && dx && dy_pitch >= g->cout && dx_pitch >= g->cin, "conv3d_dgrad: bad buffers");
  auto st = static_cast<cudaStream_t>(stream);
  if (g->stride == 1 && g->dil * (g->k - 1) - g->pad >= 0) {
    UmmaConvArgs a = dgrad_args(g, dy, dy_pitch, w_packed_dgrad, dx, dx_pitch, stats);
    if (conv_umma_supported(a)) return conv_umma_run(a, st);
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
