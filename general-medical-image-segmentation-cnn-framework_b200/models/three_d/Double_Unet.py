"""Double_Unet with the reference's constructor and state_dict (models/three_d/Double_Unet.py:15-131) on b200seg kernels.

Two 3-level U-Nets in sequence: a coarse one (width unet_init_features / 2) whose class scores are concatenated with the
image and fed to a fine one (width unet_init_features) whose skip connections pass through squeeze-and-excitation gates
(SE_Residual, SE.py:27-49) before the concatenation (:98-110).  The up-convolutions keep their channel count, so every
decoder block reads 3 x width channels (:33-38).
"""
from collections import OrderedDict

import torch.nn as nn

from .._common import OpsMixin, norm_args, norm_spec
from .SE import SE_Inception, SE_Residual  # noqa: F401


class Double_Unet(nn.Module, OpsMixin):
    def __init__(self, in_channels=1, out_channels=2, unet_init_features=64, Cnn_init_features=64, elu=True):
        super(Double_Unet, self).__init__()
        for prefix, cin, f in (("cu", in_channels, unet_init_features // 2),
                               ("fu", in_channels + out_channels, unet_init_features)):
            widths = (cin, f, 2 * f, 4 * f)
            for lvl in (1, 2, 3):
                setattr(self, "%s_encoder%d" % (prefix, lvl), self._block(widths[lvl - 1], widths[lvl], "%s_enc%d" % (prefix, lvl)))
                setattr(self, "%s_pool%d" % (prefix, lvl), nn.MaxPool3d(kernel_size=2, stride=2))
            setattr(self, prefix + "_bottleneck", self._block(4 * f, 8 * f, prefix + "_bottleneck"))
            for lvl in (3, 2, 1):
                up = f << lvl                       # channels arriving from below: 8f, 4f, 2f
                setattr(self, "%s_upconv%d" % (prefix, lvl), nn.ConvTranspose3d(up, up, kernel_size=2, stride=2))
                setattr(self, "%s_decoder%d" % (prefix, lvl), self._block(up + up // 2, up // 2, "%s_dec%d" % (prefix, lvl)))
            setattr(self, prefix + "_conv", nn.Conv3d(f, out_channels, kernel_size=1))
        f = unet_init_features
        self.SE3, self.SE2, self.SE1 = SE_Residual(4 * f), SE_Residual(2 * f), SE_Residual(f)

    @staticmethod
    def _block(in_channels, features, name):
        layers = OrderedDict()
        for i, c in ((1, in_channels), (2, features)):
            layers["%sconv%d" % (name, i)] = nn.Conv3d(c, features, kernel_size=3, padding=1, bias=True)
            layers["%snorm%d" % (name, i)] = nn.BatchNorm3d(num_features=features)
            layers["%srelu%d" % (name, i)] = nn.ReLU(inplace=True)
        return nn.Sequential(layers)

    # ---- kernels ---------------------------------------------------------------------------------------------------------
    def _run_block(self, seq, x, x2=None, out=None):
        F = self.kernels
        m = list(seq.children())
        h = F.conv_norm_act(x, m[0].weight, m[0].bias, x2=x2, k=3, stride=1, pad=1, dil=1,
                            spec=norm_spec(F, m[1], "relu", 0.0, self.training), **norm_args(m[1]))
        return F.conv_norm_act(h, m[3].weight, m[3].bias, k=3, stride=1, pad=1, dil=1,
                               spec=norm_spec(F, m[4], "relu", 0.0, self.training), out=out, **norm_args(m[4]))

    def _unet(self, prefix, h, gates=None):
        F = self.kernels
        get = lambda name: getattr(self, prefix + "_" + name)      # noqa: E731
        n, d, hh, w = F.spatial(h)
        dev = F.device_of(h)
        f = get("conv").in_channels
        # [up-convolution | (gated) encoder output] halves of one buffer per level (torch.cat at :80-85, :99-110)
        bufs = [F.alloc_concat(n, d >> l, hh >> l, w >> l, (2 * f) << l, f << l, dev) for l in range(3)]
        skips, x = [], h
        for lvl in (1, 2, 3):
            direct = gates is None                  # the fine U-Net gates the skip first: the encoder output stays separate
            enc = self._run_block(get("encoder%d" % lvl), x, out=bufs[lvl - 1][2] if direct else None)
            skips.append(enc)
            x = F.max_pool2(enc)
        x = self._run_block(get("bottleneck"), x)
        for lvl in (3, 2, 1):
            up = get("upconv%d" % lvl)
            x = F.conv_transpose_kxsx(x, up.weight, up.bias, out=bufs[lvl - 1][1])
            skip = skips[lvl - 1]
            if gates is not None:
                skip = gates[lvl - 1](skip, out=bufs[lvl - 1][2])
            x = self._run_block(get("decoder%d" % lvl), x, skip)
        head = get("conv")
        return F.head_conv1x1(x, head.weight, head.bias)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        cu_outputs = self._unet("cu", F.to_ndhwc(x))
        x_ = F.concat_input(x, cu_outputs)                           # torch.cat((x, cu_outputs), dim=1) at :90
        return self._unet("fu", x_, gates=(self.SE1, self.SE2, self.SE3))
