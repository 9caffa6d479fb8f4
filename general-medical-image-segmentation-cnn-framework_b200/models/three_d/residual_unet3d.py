"""Residual (Isensee-style) 3-D U-Net with the reference's constructor and state_dict
(models/three_d/residual_unet3d.py:11-204) on b200seg kernels.

Bias-free 3x3x3 convs (stride-2 for the context down-steps), non-affine InstanceNorm3d + LeakyReLU(0.01) as one
normalise+activate pass, nearest x2 up-sampling, Dropout3d(0.6), residual sums, and the deep-supervision sum of
class-score maps (fp32, :196-202).  `norm_lrelu_conv_c{2..5}` are applied twice per level with shared weights exactly
as the reference does (:126-128).
"""
import torch
import torch.nn as nn

from .._common import OpsMixin, conv_args


class UNet(nn.Module, OpsMixin):
    """Implementations based on the Unet3D paper: https://arxiv.org/pdf/1706.00120.pdf"""

    def __init__(self, in_channels, n_classes, base_n_filter=8):
        super(UNet, self).__init__()
        self.in_channels = in_channels
        self.n_classes = n_classes
        self.base_n_filter = base_n_filter
        b = base_n_filter

        self.lrelu = nn.LeakyReLU()
        self.dropout3d = nn.Dropout3d(p=0.6)
        self.upsacle = nn.Upsample(scale_factor=2, mode='nearest')
        self.softmax = nn.Softmax(dim=1)

        def conv(cin, cout, k=3, stride=1):
            return nn.Conv3d(cin, cout, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=False)

        self.conv3d_c1_1 = conv(in_channels, b)
        self.conv3d_c1_2 = conv(b, b)
        self.lrelu_conv_c1 = self.lrelu_conv(b, b)
        self.inorm3d_c1 = nn.InstanceNorm3d(b)

        self.conv3d_c2 = conv(b, b * 2, stride=2)
        self.norm_lrelu_conv_c2 = self.norm_lrelu_conv(b * 2, b * 2)
        self.inorm3d_c2 = nn.InstanceNorm3d(b * 2)

        self.conv3d_c3 = conv(b * 2, b * 4, stride=2)
        self.norm_lrelu_conv_c3 = self.norm_lrelu_conv(b * 4, b * 4)
        self.inorm3d_c3 = nn.InstanceNorm3d(b * 4)

        self.conv3d_c4 = conv(b * 4, b * 8, stride=2)
        self.norm_lrelu_conv_c4 = self.norm_lrelu_conv(b * 8, b * 8)
        self.inorm3d_c4 = nn.InstanceNorm3d(b * 8)

        self.conv3d_c5 = conv(b * 8, b * 16, stride=2)
        self.norm_lrelu_conv_c5 = self.norm_lrelu_conv(b * 16, b * 16)
        self.norm_lrelu_upscale_conv_norm_lrelu_l0 = self.norm_lrelu_upscale_conv_norm_lrelu(b * 16, b * 8)

        self.conv3d_l0 = conv(b * 8, b * 8, k=1)
        self.inorm3d_l0 = nn.InstanceNorm3d(b * 8)

        self.conv_norm_lrelu_l1 = self.conv_norm_lrelu(b * 16, b * 16)
        self.conv3d_l1 = conv(b * 16, b * 8, k=1)
        self.norm_lrelu_upscale_conv_norm_lrelu_l1 = self.norm_lrelu_upscale_conv_norm_lrelu(b * 8, b * 4)

        self.conv_norm_lrelu_l2 = self.conv_norm_lrelu(b * 8, b * 8)
        self.conv3d_l2 = conv(b * 8, b * 4, k=1)
        self.norm_lrelu_upscale_conv_norm_lrelu_l2 = self.norm_lrelu_upscale_conv_norm_lrelu(b * 4, b * 2)

        self.conv_norm_lrelu_l3 = self.conv_norm_lrelu(b * 4, b * 4)
        self.conv3d_l3 = conv(b * 4, b * 2, k=1)
        self.norm_lrelu_upscale_conv_norm_lrelu_l3 = self.norm_lrelu_upscale_conv_norm_lrelu(b * 2, b)

        self.conv_norm_lrelu_l4 = self.conv_norm_lrelu(b * 2, b * 2)
        self.conv3d_l4 = conv(b * 2, n_classes, k=1)

        self.ds2_1x1_conv3d = conv(b * 8, n_classes, k=1)
        self.ds3_1x1_conv3d = conv(b * 4, n_classes, k=1)
        self.sigmoid = nn.Sigmoid()

    # ---- the reference's block builders (module structure defines the state_dict keys) -------------------------
    def conv_norm_lrelu(self, feat_in, feat_out):
        return nn.Sequential(nn.Conv3d(feat_in, feat_out, kernel_size=3, stride=1, padding=1, bias=False),
                             nn.InstanceNorm3d(feat_out), nn.LeakyReLU())

    def norm_lrelu_conv(self, feat_in, feat_out):
        return nn.Sequential(nn.InstanceNorm3d(feat_in), nn.LeakyReLU(),
                             nn.Conv3d(feat_in, feat_out, kernel_size=3, stride=1, padding=1, bias=False))

    def lrelu_conv(self, feat_in, feat_out):
        return nn.Sequential(nn.LeakyReLU(),
                             nn.Conv3d(feat_in, feat_out, kernel_size=3, stride=1, padding=1, bias=False))

    def norm_lrelu_upscale_conv_norm_lrelu(self, feat_in, feat_out):
        return nn.Sequential(nn.InstanceNorm3d(feat_in), nn.LeakyReLU(), nn.Upsample(scale_factor=2, mode='nearest'),
                             nn.Conv3d(feat_in, feat_out, kernel_size=3, stride=1, padding=1, bias=False),
                             nn.InstanceNorm3d(feat_out), nn.LeakyReLU())

    # ---- kernels -------------------------------------------------------------------------------------------------
    def _in_lrelu(self, x, norm=None, out=None):
        F = self.kernels
        eps = 1e-5 if norm is None else norm.eps
        return F.norm_act(x, F.NormSpec("instance", "leaky_relu", self.lrelu.negative_slope, eps=eps), out=out)

    def _lrelu(self, x, out=None):
        return self.kernels.activation(x, "leaky_relu", self.lrelu.negative_slope, out=out)

    def _conv(self, conv, x, x2=None):
        return self.kernels.conv_norm_act(x, conv.weight, None, x2=x2, **conv_args(conv))

    def _conv_in_lrelu(self, conv, norm, x, x2=None, out=None):
        F = self.kernels
        spec = F.NormSpec("instance", "leaky_relu", self.lrelu.negative_slope, eps=norm.eps)
        return F.conv_norm_act(x, conv.weight, None, x2=x2, spec=spec, out=out, **conv_args(conv))

    def _drop(self, x):
        return self.kernels.dropout(x, self.dropout3d.p, training=self.training, channel=True)

    def _context(self, down_conv, nlc, x):
        """conv(stride 2) -> [IN -> lrelu -> conv] -> dropout -> [same module again] -> + residual (:124-131)."""
        F = self.kernels
        out = self._conv(down_conv, x)
        residual = out
        out = self._conv(nlc[2], self._in_lrelu(out, nlc[0]))
        out = self._drop(out)
        out = self._conv(nlc[2], self._in_lrelu(out, nlc[0]))
        return F.add(out, residual)

    def _up(self, seq, x, out=None):
        """IN -> lrelu -> nearest x2 -> conv -> IN -> lrelu (:98-107)."""
        F = self.kernels
        h = F.upsample_nearest2(self._in_lrelu(x, seq[0]))
        return self._conv_in_lrelu(seq[3], seq[4], h, out=out)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        h = F.to_ndhwc(x)
        b = self.base_n_filter
        n, d, hh, w = F.spatial(h)
        dev = F.device_of(h)
        # localisation-path concat buffers: [up-sampled features | context_k] (torch.cat at :167,175,183,190)
        _, up1, ctx1 = F.alloc_concat(n, d, hh, w, b, b, dev)
        _, up2, ctx2 = F.alloc_concat(n, d // 2, hh // 2, w // 2, 2 * b, 2 * b, dev)
        _, up3, ctx3 = F.alloc_concat(n, d // 4, hh // 4, w // 4, 4 * b, 4 * b, dev)
        _, up4, ctx4 = F.alloc_concat(n, d // 8, hh // 8, w // 8, 8 * b, 8 * b, dev)

        # Level 1 context pathway (:110-122)
        out = self._conv(self.conv3d_c1_1, h)
        residual_1 = out
        out = self._conv(self.conv3d_c1_2, self._lrelu(out))
        out = self._drop(out)
        out = self._conv(self.lrelu_conv_c1[1], self._lrelu(out))
        out = F.add(out, residual_1)
        context_1 = self._lrelu(out, out=ctx1)
        out = self._in_lrelu(out, self.inorm3d_c1)

        # Levels 2-4 (:124-157): context_k is the normalised, activated sum
        out = self._context(self.conv3d_c2, self.norm_lrelu_conv_c2, out)
        context_2 = out = self._in_lrelu(out, self.inorm3d_c2, out=ctx2)
        out = self._context(self.conv3d_c3, self.norm_lrelu_conv_c3, out)
        context_3 = out = self._in_lrelu(out, self.inorm3d_c3, out=ctx3)
        out = self._context(self.conv3d_c4, self.norm_lrelu_conv_c4, out)
        context_4 = out = self._in_lrelu(out, self.inorm3d_c4, out=ctx4)

        # Level 5 (:159-170)
        out = self._context(self.conv3d_c5, self.norm_lrelu_conv_c5, out)
        out = self._up(self.norm_lrelu_upscale_conv_norm_lrelu_l0, out)
        out = self._conv_in_lrelu(self.conv3d_l0, self.inorm3d_l0, out, out=up4)

        # Localisation pathway (:172-194)
        out = self._conv_in_lrelu(self.conv_norm_lrelu_l1[0], self.conv_norm_lrelu_l1[1], out, x2=context_4)
        out = self._conv(self.conv3d_l1, out)
        out = self._up(self.norm_lrelu_upscale_conv_norm_lrelu_l1, out, out=up3)

        out = self._conv_in_lrelu(self.conv_norm_lrelu_l2[0], self.conv_norm_lrelu_l2[1], out, x2=context_3)
        ds2 = out
        out = self._conv(self.conv3d_l2, out)
        out = self._up(self.norm_lrelu_upscale_conv_norm_lrelu_l2, out, out=up2)

        out = self._conv_in_lrelu(self.conv_norm_lrelu_l3[0], self.conv_norm_lrelu_l3[1], out, x2=context_2)
        ds3 = out
        out = self._conv(self.conv3d_l3, out)
        out = self._up(self.norm_lrelu_upscale_conv_norm_lrelu_l3, out, out=up1)

        out = self._conv_in_lrelu(self.conv_norm_lrelu_l4[0], self.conv_norm_lrelu_l4[1], out, x2=context_1)
        out_pred = F.head_conv1x1(out, self.conv3d_l4.weight, None)

        # deep supervision (:196-203): class-score maps, fp32
        ds2_1x1_conv = F.head_conv1x1(ds2, self.ds2_1x1_conv3d.weight, None)
        ds3_1x1_conv = F.head_conv1x1(ds3, self.ds3_1x1_conv3d.weight, None)
        ds_sum = F.classmap_up2_add(ds2_1x1_conv, ds3_1x1_conv)
        return F.classmap_up2_add(ds_sum, out_pred)
