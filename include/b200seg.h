/*
 * b200seg -- C ABI of the B200-native volumetric segmentation hot path.
 *
 * The reference (QingYunA/General-Medical-Image-Segmentation-CNN-Framework) is 100 % Python over torch.nn: it has no
 * FFI.  Its seam is the nn.Module protocol (train.py:324-373 builds the model, train.py:203-214 runs
 * forward / loss / backward).  Each entry point below replaces the library kernel(s) PyTorch launches for one
 * reference call site; the Python package binds them with ctypes.
 *
 * Conventions
 *   - activations: channels-last NDHWC, bf16, innermost dim contiguous; `*_pitch` = elements between voxels.
 *   - parameters / statistics / gradients of parameters: fp32.
 *   - every function enqueues on `stream` (a cudaStream_t), never synchronises, never allocates persistent memory;
 *     all buffers (incl. workspaces) are owned by the caller.
 *   - return 0 on success, negative b200seg_status otherwise; b200seg_last_error() gives a thread-local message.
 */
#ifndef B200SEG_H_
#define B200SEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200SEG_OK = 0,
  B200SEG_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  B200SEG_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  B200SEG_ERR_WORKSPACE = -3    /* workspace too small */
} b200seg_status;

typedef enum { B200SEG_ACT_NONE = 0, B200SEG_ACT_RELU = 1, B200SEG_ACT_LEAKY = 2, B200SEG_ACT_ELU = 3,
               B200SEG_ACT_PRELU = 4 } b200seg_act;

/* Cubic-kernel 3-D convolution geometry (nn.Conv3d: unet3d.py:80-98, vnet3d.py:25,47,65,111,
 * residual_unet3d.py:29-44, convolution.py:57-63).  Input [n,d,h,w,cin], output [n,od,oh,ow,cout]. */
typedef struct {
  int32_t n, d, h, w, cin;
  int32_t od, oh, ow, cout;
  int32_t k, stride, pad, dil;
} b200seg_conv_geom;

const char* b200seg_version(void);
const char* b200seg_last_error(void);
/* Number of tcgen05 kernels launched by this process so far (diagnostic: proves which path ran). */
int64_t b200seg_umma_launch_count(void);
/* 1 if the tcgen05 implicit-GEMM path handles this geometry, 0 if the direct CUDA-core path is used. */
int b200seg_conv3d_uses_tensor_cores(const b200seg_conv_geom* g);

/* ---- layout / packing ------------------------------------------------------------------------------------- */
/* fp32 NCDHW -> bf16 NDHWC (model input, train.py:195) and back (logits handed to the caller as NCDHW fp32). */
int b200seg_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int n, int c, int64_t spatial, void* stream);
int b200seg_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int n, int c, int64_t spatial, void* stream);
/* The same with a row pitch (elements) on the channels-last side, so that `dst` / `src` can be a channel slice of a wider
 * buffer: Double_Unet.py:90 concatenates the image with the coarse network's class scores without a separate cat pass. */
int b200seg_ncdhw_f32_to_ndhwc_bf16_pitched(const float* src, void* dst, int64_t dst_pitch, int n, int c, int64_t spatial,
                                            void* stream);
int b200seg_ndhwc_bf16_to_ncdhw_f32_pitched(const void* src, int64_t src_pitch, float* dst, int n, int c, int64_t spatial,
                                            void* stream);
/* Conv3d weight [cout][cin][k^3] fp32 -> fprop pack [k^3][cout][cin] bf16 (flip=0) or dgrad pack
 * [k^3 flipped][cin][cout] bf16 (flip=1).  cin_off/cin_cnt select an input-channel slice (concat-free decoders). */
int b200seg_pack_conv_weight(const float* w, void* packed, int cout, int cin, int k, int cin_off, int cin_cnt,
                             int dgrad, void* stream);
/* (cin_off + cin_cnt may exceed cin: the excess input channels are packed as zeros -- K-dimension padding.) */
/* The same two packs with BOTH channel counts widened to cout_pad x cin_pad (zero rows / columns): layers whose channel
 * counts are not multiples of 16 run on the tensor-core kernels this way (densevoxelnet3d.py:22, growth rate 12). */
int b200seg_pack_conv_weight_padded(const float* w, void* packed, int cout, int cin, int k, int cout_pad, int cin_pad,
                                    int dgrad, void* stream);
/* y[rows][cpad] (contiguous) = x[rows][0:c] followed by zeros; cpad a multiple of 8.  Used to widen the 1-channel network
 * input to 16 channels so that the stem convolution (unet3d.py:80, C_in = 1) runs on the tensor-core path. */
int b200seg_pad_channels(const void* x, int64_t x_pitch, int c, void* y, int cpad, int64_t rows, void* stream);
/* Both packs of many weight tensors in one launch.  descs: DEVICE array of ndesc 32-byte records
 * {int64 src (float offset into arena), int64 dst (bf16 element offset into packs), int32 cout, cin, k^3, dgrad};
 * max_k3 = the largest k^3 in the table (sizes the shared-memory brick); total_tiles = sum over the records of
 * ceil(cout / 32) * ceil(cin / 8), the number of 32 x 8 x k^3 bricks dealt round-robin to the thread blocks. */
int b200seg_pack_weights_batched(const float* arena, void* packs, const void* descs, int ndesc, int max_k3,
                                 int total_tiles, void* stream);
/* The reverse for the weight gradients of many layers in one launch: grads[dst + (co*cin + ci)*k^3 + t] +=
 * packed[src + (t*cin + ci)*cout + co]; same 32-byte records (dgrad ignored), k^3 <= 125. */
int b200seg_unpack_wgrads_batched(const float* packed, float* grads, const void* descs, int ndesc, int max_k3,
                                  int total_tiles, void* stream);
/* wgrad result [k^3][cin_cnt][cout] fp32 -> torch-layout grad [cout][cin][k^3] fp32 slice.  accumulate bit 0: add to
 * grad_w instead of overwriting; bit 1: clear dw_packed while reading it (a persistent, always-zero accumulator needs no
 * memset between steps; dw_packed is written in that case despite the const). */
int b200seg_unpack_conv_wgrad(const float* dw_packed, float* grad_w, int cout, int cin, int k, int cin_off,
                              int cin_cnt, int accumulate, void* stream);

/* ---- Conv3d (nn.Conv3d fwd / autograd bwd) ------------------------------------------------------------------ */
/* y = conv(x, w) + bias; optionally also per-channel sum / sum-of-squares of the fp32 result (BatchNorm statistics
 * fused into the epilogue, unet3d.py:88,100).  w_packed from b200seg_pack_conv_weight(dgrad=0).
 * stats (may be NULL): float[2*cout] = {sum, sumsq}, accumulated into (caller zeroes). */
size_t b200seg_conv3d_workspace_bytes(const b200seg_conv_geom* g);
/* Scratch bytes b200seg_conv3d_fprop (dgrad = 0) / b200seg_conv3d_dgrad (dgrad = 1) can use for this geometry, 0 = none.
 * Only the K-heavy layers on 8 x 8 planes (the U-Net bottleneck at 8^3, unet3d.py:28; V-Net's deepest 5x5x5 layers) have
 * one: with it they run weights-stationary, split over tap rows, and meet in an fp32 copy of the output (the library
 * clears it); without it they take the voxel-tiled kernels.  Pass it as `workspace`, `workspace_bytes`. */
size_t b200seg_conv3d_ws_bytes(const b200seg_conv_geom* g, int dgrad);
int b200seg_conv3d_fprop(const b200seg_conv_geom* g, const void* x, int64_t x_pitch, const void* w_packed,
                         const float* bias, void* y, int64_t y_pitch, float* stats, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Inference: convolution with the eval-mode BatchNorm scale / shift and the activation applied in the epilogue
 * (unet3d.py:80-101 under model.eval(); north-star "scale-shift and ReLU fused into the epilogue"):
 *   y = act(scale[c] * conv(x, w)[c] + shift[c]),  act in {B200SEG_ACT_NONE, _RELU, _LEAKY(slope)}.
 * scale / shift: fp32 [cout] (rows 2 and 3 of b200seg_norm_eval_coef, which folds the conv bias into the shift).
 * Tensor-core geometries only: ask b200seg_conv3d_fprop_act_supported first (1 = supported). */
int b200seg_conv3d_fprop_act_supported(const b200seg_conv_geom* g);
int b200seg_conv3d_fprop_act(const b200seg_conv_geom* g, const void* x, int64_t x_pitch, const void* w_packed,
                             const float* scale, const float* shift, int act, float slope, void* y, int64_t y_pitch,
                             void* stream);
/* dx = conv_transpose(dy, w).  w_packed from b200seg_pack_conv_weight(dgrad=1).  g describes the FORWARD conv.
 * stats (may be NULL): float[2*cin] = {sum, sumsq} of dx per channel, accumulated into (caller zeroes): the column sums
 * of a decoder conv's input gradient are the bias gradient of the ConvTranspose3d that produced that input
 * (unet3d.py:58-59), so they come out of the same epilogue instead of a separate pass over dx. */
int b200seg_conv3d_dgrad(const b200seg_conv_geom* g, const void* dy, int64_t dy_pitch, const void* w_packed_dgrad,
                         void* dx, int64_t dx_pitch, float* stats, void* workspace, size_t workspace_bytes,
                         void* stream);
/* dw_packed[k^3][cin][cout] (fp32, accumulated into; caller zeroes) = sum_voxels x (*) dy.  With a workspace of
 * b200seg_conv3d_workspace_bytes(g) bytes the split-K partial tiles are stored there with vector stores and summed by
 * a second kernel (one atomic per element and slice group); without one every CTA adds its tile to dw_packed with fp32
 * atomics.  For a single-channel input (the U-Net stem) the workspace also holds the taps-as-channels copy of x. */
int b200seg_conv3d_wgrad(const b200seg_conv_geom* g, const void* x, int64_t x_pitch, const void* dy,
                         int64_t dy_pitch, float* dw_packed, void* workspace, size_t workspace_bytes, void* stream);

/* ---- ConvTranspose3d k=2 s=2 (unet3d.py:29-43) ---------------------------------------------------------------- */
/* weight [cin][cout][2][2][2] fp32 -> pack [8][cout][cin] bf16 (fwd) / [8][cin][cout] bf16 (dgrad). */
int b200seg_pack_convt_weight(const float* w, void* packed, int cin, int cout, int dgrad, void* stream);
int b200seg_convt_k2s2_fwd(const void* x, int64_t x_pitch, const void* w_packed, const float* bias, void* y,
                           int64_t y_pitch, int n, int d, int h, int w, int cin, int cout, void* stream);
int b200seg_convt_k2s2_dgrad(const void* dy, int64_t dy_pitch, const void* w_packed_dgrad, void* dx,
                             int64_t dx_pitch, int n, int d, int h, int w, int cin, int cout, void* stream);
/* dw_packed [8][cout][cin] fp32 is accumulated into (caller zeroes); b200seg_unpack_conv_wgrad(cout=cin, cin=cout,
 * k=2) turns it into the [cin][cout][2][2][2] layout.  The bias gradient is b200seg_channel_stats(dy) row 0. */
int b200seg_convt_k2s2_wgrad(const void* x, int64_t x_pitch, const void* dy, int64_t dy_pitch, float* dw_packed,
                             int n, int d, int h, int w, int cin, int cout, void* stream);

/* ---- normalisation + activation (nn.BatchNorm3d / InstanceNorm3d / SynchronizedBatchNorm3d + ReLU family) ---- */
/* stats[g][2][c] += {sum, sumsq} over the rows of group g; rows = voxels, groups = 1 (batch norm) or n (instance). */
int b200seg_channel_stats(const void* x, int64_t pitch, int64_t rows_per_group, int groups, int c, float* stats,
                          void* stream);
/* From (all-reduced) sums: mean, inv_std, fused scale/shift; updates running stats when non-NULL (momentum, unbiased
 * variance).  clamp_eps=1 reproduces sync_batchnorm/batchnorm.py:125 (var.clamp(eps)^-1/2), 0 = nn.BatchNorm.
 * out: float[groups][4][c] = {mean, inv_std, scale, shift}.  gamma/beta may be NULL (affine=False). */
int b200seg_norm_finalize(const float* stats, double count, int groups, int c, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                          int clamp_eps, float* out, void* stream);
/* F.batch_norm(training=False) constants from the running statistics: out = float[4][c] = {mean, inv_std,
 * scale = gamma*inv_std, shift = beta - mean*scale (+ scale*conv_bias when conv_bias != NULL)}; gamma/beta may be NULL. */
int b200seg_norm_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                           const float* conv_bias, float eps, int c, float* out, void* stream);
/* z = act(y*scale + shift [+ residual]).  coef = the [groups][4][c] block from norm_finalize (NULL: identity).
 * act_param: leaky slope (scalar, host) ; prelu_w: per-channel slope (device) for B200SEG_ACT_PRELU. */
int b200seg_norm_act_fwd(const void* y, int64_t y_pitch, const float* coef, int64_t rows_per_group, int groups,
                         int c, int act, float act_param, const float* prelu_w, const void* residual,
                         int64_t res_pitch, void* z, int64_t z_pitch, void* stream);
/* Training-mode BatchNorm3d + activation WITHOUT the separate b200seg_norm_finalize launch: every block derives the
 * constants of all c <= 512 channels from stats = {sum[c], sumsq[c]} in its prologue (same arithmetic, bit-identical);
 * block 0 stores them to coef_out [4][c] for the backward pass and updates running_mean / running_var (may be NULL). */
int b200seg_norm_act_fwd_stats(const void* y, int64_t y_pitch, const float* stats, double count, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                               int clamp_eps, float* coef_out, int64_t rows, int c, int act, float act_param,
                               const float* prelu_w, const void* residual, int64_t res_pitch, void* z, int64_t z_pitch,
                               void* stream);
/* Backward, pass 1: sums[g][2][c] += {sum(dpre), sum(dpre * xhat)}, dpre = dz * act'(pre).  Also accumulates the
 * PReLU slope gradient when dprelu != NULL, and -- when grad_affine != NULL -- adds the same two sums (over all groups)
 * to grad_affine[0..c) = d(gamma) and grad_affine[c..2c) = d(beta): the affine parameter gradients of the norm layer
 * (unet3d.py:88,100) land in the optimiser's gradient arena without a separate accumulation launch. */
int b200seg_norm_act_bwd_reduce(const void* dz, int64_t dz_pitch, const void* y, int64_t y_pitch,
                                const float* coef, int64_t rows_per_group, int groups, int c, int act,
                                float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                                float* sums, float* dprelu, float* grad_affine, void* stream);
/* Backward, pass 2: dy = scale * (dpre - sum_dpre/count - xhat * sum_dpre_xhat/count) (training statistics), or
 * dy = scale * dpre when sums == NULL (eval / no norm).  dres (may be NULL) = dpre (gradient of the residual). */
int b200seg_norm_act_bwd_apply(const void* dz, int64_t dz_pitch, const void* y, int64_t y_pitch, const float* coef,
                               const float* sums, double count, int64_t rows_per_group, int groups, int c, int act,
                               float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                               void* dy, int64_t dy_pitch, void* dres, int64_t dres_pitch, void* stream);
/* The same with dy = acc + (that gradient); acc (bf16 rows, may be NULL, may alias dy) is the gradient the tensor already
 * received through another consumer -- DenseVoxelNet's dense blocks (densevoxelnet3d.py:36-42): the input of layer i also is
 * the head of the concatenation, and both gradients are accumulated in one buffer instead of being summed by a third pass. */
int b200seg_norm_act_bwd_apply_acc(const void* dz, int64_t dz_pitch, const void* y, int64_t y_pitch, const float* coef,
                                   const float* sums, double count, int64_t rows_per_group, int groups, int c, int act,
                                   float act_param, const float* prelu_w, const void* residual, int64_t res_pitch,
                                   void* dy, int64_t dy_pitch, void* dres, int64_t dres_pitch, const void* acc,
                                   int64_t acc_pitch, void* stream);

/* ---- MaxPool3d(2,2) (unet3d.py:19-25) -------------------------------------------------------------------------- */
/* idx: uint8 per output element, local argmax 0..7 = (a*2+b)*2+e in (d,h,w) scan order; ties -> first, NaN wins. */
int b200seg_maxpool2_fwd(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, uint8_t* idx, int n, int d,
                         int h, int w, int c, void* stream);
/* dx = scatter(dy by idx) [+ addend]: `addend` (may be NULL) is a second gradient of the pooled tensor -- in a U-Net the
 * encoder output feeds both the pool and the skip connection (unet3d.py:52-68) -- summed here instead of in a separate
 * pass. */
int b200seg_maxpool2_bwd(const void* dy, int64_t dy_pitch, const uint8_t* idx, void* dx, int64_t dx_pitch,
                         const void* addend, int64_t addend_pitch, int n, int d, int h, int w, int c, void* stream);
/* local indices -> torch's int64 flat D*H*W indices in NCDHW order (for the bit-exact check). */
int b200seg_maxpool2_idx_to_torch(const uint8_t* idx, int64_t* out, int n, int d, int h, int w, int c, void* stream);

/* ---- nearest x2 upsample (residual_unet3d.py:19,103), elementwise add ----------------------------------------- */
int b200seg_upsample2_fwd(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int n, int d, int h, int w,
                          int c, void* stream);
int b200seg_upsample2_bwd(const void* dy, int64_t dy_pitch, void* dx, int64_t dx_pitch, int n, int d, int h, int w,
                          int c, void* stream);
int b200seg_add(const void* a, int64_t a_pitch, const void* b, int64_t b_pitch, void* out, int64_t out_pitch,
                int64_t rows, int c, void* stream);

/* ---- dropout (vnet3d.py:72,90,94 nn.Dropout3d; residual_unet3d.py:18; densevoxelnet3d.py:25-32 nn.Dropout) ------- */
/* y = keep ? x / (1 - p) : 0 with a counter-based mask hash(*seed, salt, index): calling it again on the upstream
 * gradient with the same (seed, salt) is the backward pass, no mask is stored.  channel_mode 1 draws once per
 * (sample, channel) (Dropout3d; rows_per_sample = voxels of one sample), 0 once per element.  seed: device pointer. */
int b200seg_dropout(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int64_t rows, int64_t rows_per_sample,
                    int c, float p, const unsigned long long* seed, unsigned long long salt, int channel_mode,
                    void* stream);
/* The same with (a) a second independent mask salt2 (0 = none) applied in the same pass, the value rounded to bf16 in
 * between -- _DenseLayer runs its dropout twice in train mode (densevoxelnet3d.py:25-32) -- and (b) channels [c, c_out) of
 * y written as zeros (y_pitch >= c_out): a gradient handed to the 16-channel tensor-core tiles needs no extra pad pass. */
int b200seg_dropout2(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int64_t rows, int64_t rows_per_sample,
                     int c, int c_out, float p, const unsigned long long* seed, unsigned long long salt,
                     unsigned long long salt2, int channel_mode, void* stream);

/* ---- fp32 NCDHW class-score maps: deep-supervision sum (residual_unet3d.py:196-202) ----------------------------- */
/* out[planes][2d][2h][2w] = nearest_x2(coarse[planes][d][h][w]) + fine (fine may be NULL); planes = n * classes. */
int b200seg_classmap_up2_add(const float* coarse, const float* fine, float* out, int64_t planes, int d, int h, int w,
                             void* stream);
/* backward of the up-sampling: dcoarse = sum over each 2x2x2 cell of dfine. */
int b200seg_classmap_down2_sum(const float* dfine, float* dcoarse, int64_t planes, int d, int h, int w, void* stream);

/* ---- non-overlapping strided convolutions as GEMMs (csrnet.py:115-154: Conv3d(k3, s4), ConvTranspose3d(k4, s4)) -------- */
/* y[n][od][oh][ow][k^3 * c] (contiguous) = the k^3 taps of every s^3 cell of x [n][d][h][w][c] side by side in the channel
 * dimension, tap-major (k <= s, no padding): the strided convolution becomes a 1x1x1 convolution with k^3 * c inputs. */
int b200seg_space_to_depth(const void* x, int64_t x_pitch, void* y, int n, int d, int h, int w, int c, int k, int s, int od,
                           int oh, int ow, void* stream);
/* The inverse / adjoint: x[n][d][h][w][c] from y, zero at positions no tap covers.  With k == s it is the pixel shuffle that
 * turns a 1x1x1 convolution with s^3 * c outputs into ConvTranspose3d(kernel = stride = s). */
int b200seg_depth_to_space(const void* y, void* x, int64_t x_pitch, int n, int d, int h, int w, int c, int k, int s, int od,
                           int oh, int ow, void* stream);

/* ---- attention gates of ER-Net / RE-Net / Double-UNet (models/three_d/ER_net.py, RE_net.py, Double_Unet.py, SE.py) --- */
/* ConvTranspose3d(1, 1, kernel 2, stride 2) on fp32 single-channel maps [planes][d][h][w] -> [planes][2d][2h][2w]
 * (ER_net.py:166-168: the reverse-attention map is projected to one channel, then up-sampled).  w: 8 floats, bias: 1. */
int b200seg_convt1_k2s2_fwd(const float* in, const float* w, const float* bias, float* out, int64_t planes, int d, int h,
                            int wd, void* stream);
/* din (may be NULL) and sums[9] += {dw[8], dbias} (caller zeroes sums). */
int b200seg_convt1_k2s2_bwd(const float* dout, const float* in, const float* w, float* din, float* sums, int64_t planes,
                            int d, int h, int wd, void* stream);
/* out[v][c] = fine[v][c] * (2 - sigmoid(g[v])): `x = -1 * sigmoid(g) + 1; x = x.expand(..).mul(enc); x = x + enc`
 * (ER_net.py:184-187).  fine / out bf16 NDHWC rows, g fp32 [rows]. */
int b200seg_reverse_gate_fwd(const void* fine, int64_t fine_pitch, const float* g, void* out, int64_t out_pitch,
                             int64_t rows, int c, void* stream);
/* dfine = dout * (2 - s), dg[v] = -s (1 - s) sum_c dout[v][c] fine[v][c] with s = sigmoid(g[v]).  C: power of two. */
int b200seg_reverse_gate_bwd(const void* dout, int64_t dout_pitch, const void* fine, int64_t fine_pitch, const float* g,
                             void* dfine, int64_t dfine_pitch, float* dg, int64_t rows, int c, void* stream);
/* out[v][c] = x1[v][c] * w1[n][c] (+ x2[v][c] * w2[n][c]); w fp32 [n][c].  Squeeze-and-excitation `x + x * y`
 * (SE.py:41-49, w1 = 1 + y) and the selective fusion of two branches (ER_net.py:86-105, w = softmax over the branches). */
int b200seg_channel_blend_fwd(const void* x1, int64_t x1_pitch, const float* w1, const void* x2, int64_t x2_pitch,
                              const float* w2, void* out, int64_t out_pitch, int64_t rows_per_sample, int n, int c,
                              void* stream);
/* dots1[n][c] += sum_v dout * x1, dots2[n][c] += sum_v dout * x2 (caller zeroes them): the gradients of w1 / w2. */
int b200seg_channel_blend_bwd_reduce(const void* dout, int64_t dout_pitch, const void* x1, int64_t x1_pitch, const void* x2,
                                     int64_t x2_pitch, float* dots1, float* dots2, int64_t rows_per_sample, int n, int c,
                                     void* stream);
/* dx1 = dout * w1 + add, dx2 = dout * w2 + add; add[n][c] (may be NULL) = gradient reaching the inputs through the global
 * average pool that produced the weights, already divided by the voxel count. */
int b200seg_channel_blend_bwd_apply(const void* dout, int64_t dout_pitch, const float* w1, const float* w2, const float* add,
                                    void* dx1, int64_t dx1_pitch, void* dx2, int64_t dx2_pitch, int64_t rows_per_sample,
                                    int n, int c, void* stream);
/* y = sigmoid(x) on fp32 maps (RE_net.py:158 `F.sigmoid(final)`), dx = dy * y * (1 - y). */
int b200seg_f32_sigmoid_fwd(const float* x, float* y, int64_t numel, void* stream);
int b200seg_f32_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t numel, void* stream);

/* ---- head + loss (unet3d.py:46-48,70; loss_function.py:8-16,102-130,148-185; train.py:115,204) ---------------- */
/* logits[n][classes][spatial] fp32 (NCDHW, what the module returns) = 1x1x1 conv of NDHWC bf16 features. */
int b200seg_head_conv1x1_fwd(const void* x, int64_t x_pitch, const float* w, const float* b, float* logits, int n,
                             int64_t spatial, int cin, int classes, void* stream);
/* dx (bf16 NDHWC) = dlogits * w ; grad_w[classes][cin], grad_b[classes] accumulated into. */
int b200seg_head_conv1x1_bwd(const float* dlogits, const void* x, int64_t x_pitch, const float* w, void* dx,
                             int64_t dx_pitch, float* grad_w, float* grad_b, int n, int64_t spatial, int cin,
                             int classes, void* stream);
/* argmax over classes, ties -> lowest index; labels uint8 [n][spatial]. */
int b200seg_argmax_labels(const float* logits, uint8_t* labels, int n, int64_t spatial, int classes, void* stream);
/* class_weights (may be NULL): float[2*classes] = {cross-entropy weight per class (nll_loss(weight=), loss_function.py:13),
 * Dice weight per class (DiceLossss.forward(weight=), loss_function.py:176-183)}.
 * One pass over logits (fp32 NCDHW) and labels (uint8): partial[0] += sum CE nll (times the label's CE weight); partial[1+3k..] += per class
 * {sum p*t, sum p*p, sum t}; partial[1+3*classes..] += sigmoid-dice / BCE sums {sum s*t, sum s, sum t, sum bce}.
 * partial: double[1 + 3*classes + 4] (caller zeroes).  terms: 1 = the soft-max sums (CE, DiceLossss) are needed,
 * 2 = the sigmoid sums (DiceLoss, BCE), 3 = both; sums that are not requested may be left untouched. */
int b200seg_loss_reduce(const float* logits, const uint8_t* labels, int n, int64_t spatial, int classes, int terms,
                        const float* class_weights, double* partial, void* stream);
/* dlogits = w_ce * dCE + w_dice * dDiceLossss(softmax) + w_sdice * dDiceLoss(sigmoid) + w_bce * dBCE, scaled by the
 * upstream gradient *gscale (a device scalar; NULL = 1), using the sums produced by loss_reduce. */
int b200seg_loss_grad(const float* logits, const uint8_t* labels, int n, int64_t spatial, int classes,
                      const double* partial, float w_ce, float w_dice, float w_sdice, float w_bce,
                      const float* gscale, const float* class_weights, float* dlogits, void* stream);
/* Dice on caller-supplied PROBABILITIES: BinaryDiceLoss (loss_function.py:61-99; one loss per sample, by_class = 0) and
 * DiceLossss(softmax=False) (loss_function.py:172-184; one loss per class, by_class = 1).  pred: fp32 [n][classes][spatial];
 * the target is EITHER target_f (fp32, same layout) OR target_l (uint8 labels [n][spatial]; target of class k is
 * label == k, loss_function.py:154-160).  partial: double[rows][3] += {sum x*t, sum x^p, sum t^p}, rows = n or classes. */
int b200seg_dice_sums(const float* pred, const float* target_f, const uint8_t* target_l, int n, int classes, int64_t spatial,
                      int by_class, float p_exp, double* partial, void* stream);
/* dpred = *gscale * (coef_t[row] * t + coef_x[row] * p * x^(p-1)); coefficients computed by the caller from dice_sums. */
int b200seg_dice_grad(const float* pred, const float* target_f, const uint8_t* target_l, int n, int classes, int64_t spatial,
                      int by_class, float p_exp, const float* coef_t, const float* coef_x, const float* gscale, float* dpred,
                      void* stream);

/* ---- training data path (dataloader.py:52-67: torchio ZNormalization + UniformSampler) ------------------------- */
/* sums: double[2] += {sum x, sum x^2} over n floats (a whole image, all channels: ZNormalization without a mask). */
int b200seg_volume_stats(const float* x, int64_t n, double* sums, void* stream);
/* mean_inv_std: float[2] = {mean, 1 / unbiased std} (torch.std semantics, like torchio's ZNormalization.znorm). */
int b200seg_znorm_finalize(const double* sums, int64_t n, float* mean_inv_std, void* stream);
/* One patch of a device-resident volume [c][w][h][d] starting at (x0,y0,z0): fp32 images are written z-normalised
 * ((v - mean) * inv_std; mean_inv_std may be NULL = copy), uint8 label maps (is_label = 1) are copied.  out: [c][pw][ph][pd]. */
int b200seg_crop_patch(const void* vol, int is_label, int c, int w, int h, int d, int x0, int y0, int z0, int pw, int ph, int pd,
                       const float* mean_inv_std, void* out, void* stream);

/* ---- 95th-percentile Hausdorff distance (metric.py:29-32 -> monai.metrics.compute_hausdorff_distance) ------------- */
/* Edge voxels of a binary mask [w][h][d] (uint8, nonzero = foreground): seg ^ binary_erosion(seg), 6-neighbourhood, zero
 * border -- as physical coordinates (index * spacing).  count (device uint64, caller zeroes) receives the number of edge
 * voxels; points (float[capacity][3], may be NULL for a counting pass) the first `capacity` of them, in no particular order. */
int b200seg_mask_edge_points(const uint8_t* mask, int w, int h, int d, float sx, float sy, float sz, float* points,
                             unsigned long long capacity, unsigned long long* count, void* stream);
/* out[i] = min_j |a_i - b_j| (Euclidean) for point sets a[na][3], b[nb][3]: the distance transform of b's complement sampled
 * at a.  Both sets must be non-empty. */
int b200seg_min_distances(const float* a, int64_t na, const float* b, int64_t nb, float* out, void* stream);

/* ---- Pad3d (utils/convolution.py:78-86: F.pad(x, 6*[pad], mode)) ------------------------------------------------- */
/* mode 1 = 'reflect', 2 = 'replicate' ('constant' is folded into the convolution kernels).  x: NDHWC bf16 [n,d,h,w,c],
 * y: [n, d+2p, h+2p, w+2p, c].  bwd is the adjoint: dx[i] = sum of dy over the padded positions that read voxel i. */
int b200seg_pad3d_fwd(const void* x, int64_t x_pitch, void* y, int64_t y_pitch, int n, int d, int h, int w, int c, int pad, int mode,
                      void* stream);
int b200seg_pad3d_bwd(const void* dy, int64_t dy_pitch, void* dx, int64_t dx_pitch, int n, int d, int h, int w, int c, int pad,
                      int mode, void* stream);

/* ---- metric (metric.py:20-75) ---------------------------------------------------------------------------------- */
/* counts: uint64[4] += {sum gt, sum pred, |gt & pred| nonzero, |gt | pred| nonzero} over uint8 label volumes. */
int b200seg_seg_counts(const uint8_t* gt, const uint8_t* pred, int64_t numel, unsigned long long* counts,
                       void* stream);

/* ---- sliding-window aggregation (predict.py:100-147; torchio GridAggregator) ----------------------------------- */
/* crop mode (torchio's default, the one predict.py uses): each patch contributes its interior, i.e. the patch minus
 * overlap/2 on every face that is not on the volume border; where two interiors still overlap the patch that comes later
 * in sampler order wins.  keys: int32 [W][H][D], zero-initialised by the caller; every voxel receives
 * max(key, (patch id + 1) << 8 | label) -- order-independent, so batches, streams and ranks (all-reduce MAX) can add in
 * any order.  patch ids: patch_ids[b] (device int64) when non-NULL, else first_id + b.
 * patches: uint8 [batch][pw][ph][pd]; locations: int64 [batch][6] (i0,j0,k0,i1,j1,k1) on the device. */
int b200seg_window_accumulate_crop(const uint8_t* patches, const int64_t* locations, const int64_t* patch_ids,
                                   int64_t first_id, int batch, int pw, int ph, int pd, int ow, int oh, int od,
                                   int32_t* keys, int vw, int vh, int vd, void* stream);
/* labels[v] = keys[v] & 255 (voxels no patch covered stay 0). */
int b200seg_window_keys_to_labels(const int32_t* keys, uint8_t* labels, int64_t voxels, void* stream);
/* average mode: sum fp32 patches [batch][c][pw][ph][pd] into acc [c][W][H][D] and count [W][H][D]. */
int b200seg_window_accumulate_average(const float* patches, const int64_t* locations, int batch, int c, int pw,
                                      int ph, int pd, float* acc, float* count, int vw, int vh, int vd,
                                      void* stream);
/* acc /= max(count,1), then argmax over c -> labels uint8 [W][H][D] (labels may be NULL). */
int b200seg_window_finalize(float* acc, const float* count, int c, int64_t voxels, uint8_t* labels, void* stream);

/* ---- cross-GPU exchange over NVLink peer memory (models/sync_batchnorm/batchnorm.py:90-111, comm.py:56-137) ------ */
/* cudaMalloc'ed, zeroed buffer of `bytes` bytes plus its CUDA IPC handle (64 bytes): peers map it with b200seg_p2p_open. */
int b200seg_p2p_alloc_bytes(size_t bytes, void** dev_ptr, void* ipc_handle_out);
/* In-place SUM over ranks of the `live` 64-float chunks of a gradient buffer, over NVLink peer memory, capturable in a CUDA
 * graph (replaces the DDP gradient all-reduce of accelerator.backward, train.py:211).  bufs / reds / flags: `world` device
 * pointers each (this rank's own allocation and its peers' IPC mappings, indexed by rank): the gradient buffer, a staging
 * buffer of >= ceil(n_live / world) * 64 floats, and a zero-initialised block of 16 uint32 flags.  live: sorted chunk indices
 * (int32, device), identical on all ranks.  seq: device uint32, zero-initialised, owned by the library afterwards.  Every
 * rank must issue the same sequence of calls; all ranks end with bit-identical sums (rank-ordered addition). */
int b200seg_p2p_grad_allreduce(const void* const* bufs, const void* const* reds, const void* const* flags, const int32_t* live,
                               int n_live, int rank, int world, uint32_t* seq, void* stream);
/* Every rank allocates one mailbox (b200seg_p2p_alloc: cudaMalloc + CUDA-IPC handle, 64 bytes), ships the handle to its
 * peers (any side channel; the Python layer uses torch.distributed.all_gather_object) and maps theirs
 * (b200seg_p2p_open).  b200seg_p2p_close(ptr, opened): opened=1 unmaps a peer's mailbox, 0 frees one's own. */
size_t b200seg_p2p_mailbox_bytes(void);
int b200seg_p2p_alloc(void** dev_ptr, void* ipc_handle_out);
int b200seg_p2p_open(const void* ipc_handle, void** dev_ptr);
int b200seg_p2p_close(void* dev_ptr, int opened);
/* In-place sum over `world` ranks of vec[n] (n <= 2112 fp32) in ONE single-CTA launch per rank: store into every
 * peer's mailbox, system fence, flag, bounded spin for all peers, sum in rank order (bit-identical on every rank).
 * mailboxes: HOST array of `world` device pointers (mailboxes[rank] = own).  seq: device uint32 call counter advanced
 * by the kernel (all ranks must issue the same sequence of calls; graph-replay safe).
 * finalize_channels = C > 0: vec = {sum[C], sumsq[C], count} (the kernel stores local_count into vec[2C] first, so the
 * element count travels with the sums like sum_size in batchnorm.py:58-62); after the reduction the same launch writes
 * coef[4][C] = {mean, inv_std, scale, shift} and updates the running statistics exactly like b200seg_norm_finalize
 * (this is _compute_mean_std, batchnorm.py:113-125, executed identically on every rank instead of on a master).  * phase: 0 = the whole exchange in one launch; 1 = send only (store + publish), 2 = receive only (wait + sum
 * [+ finalize]): a 1 ... 2 pair with unrelated kernels in between hides the NVLink round trip behind them. */
int b200seg_p2p_allreduce(float* vec, int n, const void* const* mailboxes, int rank, int world, uint32_t* seq, int phase,
                          int finalize_channels, double local_count, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, float momentum, float eps, int clamp_eps, float* coef, void* stream);

/* ---- optimiser (train.py:109,214) ------------------------------------------------------------------------------- */
/* Fused Adam over a flat fp32 parameter arena (torch.optim.Adam semantics, no amsgrad, weight_decay as L2). */
int b200seg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                      void* stream);

/* Same update with the step counter and hyper-parameters in device memory, so that a launch captured in a CUDA graph
 * stays correct on replay: hyper = float[6] {lr, beta1, beta2, eps, weight_decay, grad_scale}; state = int32[2]
 * {completed steps (incremented by the kernel), scratch (must be 0)}. */
int b200seg_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                          const float* hyper, int32_t* state, void* stream);
/* The whole optimiser step of a flat parameter arena in one launch, including the two layout changes around it: the
 * packed weight gradients dw ([k^3][cin][cout] fp32, what b200seg_conv3d_wgrad accumulates into; added to grad when
 * use_dw != 0), Adam on param / exp_avg / exp_avg_sq (torch layout), and both bf16 weight packs of every conv weight
 * (as b200seg_pack_weights_batched).  descs: DEVICE array of ndesc 40-byte records {int64 src (element offset in the
 * arenas and in dw), int64 dst (bf16 offset of the fprop pack in packs; the dgrad pack follows it), int32 cout, cin, k^3,
 * kind (0 = cubic conv weight, 1 = plain range of `cout` elements), tile0 (prefix sum of the records' brick counts:
 * ceil(cout/32)*ceil(cin/T) with T = 32 if max_k3 <= 27 else 8 for kind 0, ceil(cout/4096) for kind 1), pad};
 * total_tiles = the sum of all brick counts.  hyper / state as in b200seg_adam_step_dev. */
int b200seg_adam_step_fused(float* param, const float* grad, const float* dw, float* exp_avg, float* exp_avg_sq,
                            void* packs, const void* descs, int ndesc, int max_k3, int total_tiles, const float* hyper,
                            int* state, int use_dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H_ */
