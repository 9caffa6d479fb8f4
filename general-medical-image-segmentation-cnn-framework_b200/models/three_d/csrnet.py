"""CSRNet with the reference's constructor and state_dict (models/three_d/csrnet.py:6-155) on b200seg kernels.

A 3D U-Net (same blocks as unet3d.py) with six cross-scale residual branches: `encoder_r_k` = Conv3d(k3, stride 4, no
padding) -> BatchNorm -> ReLU feeding an encoder level two scales down, `dncoder_r_k` = ConvTranspose3d(k4, stride 4) ->
BatchNorm -> ReLU feeding a decoder level two scales up (:61-80).  The 3x3x3 blocks, pools, k2s2 up-convolutions and the
head run on the U-Net's tensor-core kernels; the stride-4 branches (2 % of the FLOPs) use the generic strided kernels.
"""
from collections import OrderedDict

import torch.nn as nn

from .._common import OpsMixin, norm_args, norm_spec


class CSRNet(nn.Module, OpsMixin):
    def __init__(self, in_channels=1, out_channels=3, init_features=64):
        """
        Implementations based on the Unet3D paper: https://arxiv.org/abs/1606.06650
        """
        super(CSRNet, self).__init__()
        features = init_features
        self.encoder1 = CSRNet._block(in_channels, features, name="enc1")
        self.pool1 = nn.MaxPool3d(kernel_size=2, stride=2)
        self.encoder2 = CSRNet._block(features, features * 2, name="enc2")
        self.pool2 = nn.MaxPool3d(kernel_size=2, stride=2)
        self.encoder3 = CSRNet._block(features * 2, features * 4, name="enc3")
        self.pool3 = nn.MaxPool3d(kernel_size=2, stride=2)
        self.encoder4 = CSRNet._block(features * 4, features * 8, name="enc4")
        self.pool4 = nn.MaxPool3d(kernel_size=2, stride=2)

        self.encoder_r_1 = CSRNet._block_r(features, features * 4, name="enc1_r")
        self.encoder_r_2 = CSRNet._block_r(features * 2, features * 8, name="enc2_r")
        self.encoder_r_3 = CSRNet._block_r(features * 4, features * 16, name="enc3_r")

        self.bottleneck = CSRNet._block(features * 8, features * 16, name="bottleneck")

        self.upconv4 = nn.ConvTranspose3d(features * 16, features * 8, kernel_size=2, stride=2)
        self.decoder4 = CSRNet._block((features * 8) * 2, features * 8, name="dec4")
        self.upconv3 = nn.ConvTranspose3d(features * 8, features * 4, kernel_size=2, stride=2)
        self.decoder3 = CSRNet._block((features * 4) * 2, features * 4, name="dec3")
        self.upconv2 = nn.ConvTranspose3d(features * 4, features * 2, kernel_size=2, stride=2)
        self.decoder2 = CSRNet._block((features * 2) * 2, features * 2, name="dec2")
        self.upconv1 = nn.ConvTranspose3d(features * 2, features, kernel_size=2, stride=2)
        self.decoder1 = CSRNet._block(features * 2, features, name="dec1")

        self.conv = nn.Conv3d(in_channels=features, out_channels=out_channels, kernel_size=1)

        self.dncoder_r_1 = CSRNet._block_rr(features * 16, features * 4, name="dnc1_r")
        self.dncoder_r_2 = CSRNet._block_rr(features * 8, features * 2, name="dnc2_r")
        self.dncoder_r_3 = CSRNet._block_rr(features * 4, features * 1, name="dnc3_r")

    # ---- the reference's block builders (module structure defines the state_dict keys) --------------------------------
    @staticmethod
    def _block(in_channels, features, name):
        return nn.Sequential(OrderedDict([
            (name + "conv1", nn.Conv3d(in_channels=in_channels, out_channels=features, kernel_size=3, padding=1, bias=True)),
            (name + "norm1", nn.BatchNorm3d(num_features=features)),
            (name + "relu1", nn.ReLU(inplace=True)),
            (name + "conv2", nn.Conv3d(in_channels=features, out_channels=features, kernel_size=3, padding=1, bias=True)),
            (name + "norm2", nn.BatchNorm3d(num_features=features)),
            (name + "relu2", nn.ReLU(inplace=True)),
        ]))

    @staticmethod
    def _block_r(in_channels, features, name):
        return nn.Sequential(OrderedDict([
            (name + "conv1", nn.Conv3d(in_channels=in_channels, out_channels=features, kernel_size=3, stride=4, bias=True)),
            (name + "norm1", nn.BatchNorm3d(num_features=features)),
            (name + "relu1", nn.ReLU(inplace=True)),
        ]))

    @staticmethod
    def _block_rr(in_channels, features, name):
        return nn.Sequential(OrderedDict([
            (name + "conv1", nn.ConvTranspose3d(in_channels=in_channels, out_channels=features, kernel_size=4, stride=4,
                                                bias=True)),
            (name + "norm1", nn.BatchNorm3d(num_features=features)),
            (name + "relu1", nn.ReLU(inplace=True)),
        ]))

    # ---- kernels ---------------------------------------------------------------------------------------------------------
    def _conv_bn_relu(self, conv, norm, x, x2=None, out=None, stride=1, pad=1):
        F = self.kernels
        return F.conv_norm_act(x, conv.weight, conv.bias, x2=x2, k=3, stride=stride, pad=pad, dil=1,
                               spec=norm_spec(F, norm, "relu", 0.0, self.training), out=out, **norm_args(norm))

    def _run_block(self, seq, x, x2=None, out=None):
        mods = list(seq.children())
        h = self._conv_bn_relu(mods[0], mods[1], x, x2)
        return self._conv_bn_relu(mods[3], mods[4], h, out=out)

    def _down4(self, seq, x):
        """Conv3d(k3, s4, p0) -> BN -> ReLU (:115-133)."""
        mods = list(seq.children())
        return self._conv_bn_relu(mods[0], mods[1], x, stride=4, pad=0)

    def _up4(self, seq, x):
        """ConvTranspose3d(k4, s4) -> BN -> ReLU (:136-154)."""
        F = self.kernels
        mods = list(seq.children())
        y = F.conv_transpose_kxsx(x, mods[0].weight, mods[0].bias, stride=4)
        return F.norm_act(y, norm_spec(F, mods[1], "relu", 0.0, self.training), **norm_args(mods[1]))

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        h = F.to_ndhwc(x)
        n, d, hh, w = F.spatial(h)
        f = self.encoder1[3].out_channels
        dev = F.device_of(h)
        # [up-convolution (+ cross-scale branch) | encoder output] halves of one buffer per level (torch.cat at :68-78)
        _, up1, skip1 = F.alloc_concat(n, d, hh, w, f, f, dev)
        _, up2, skip2 = F.alloc_concat(n, d // 2, hh // 2, w // 2, 2 * f, 2 * f, dev)
        _, up3, skip3 = F.alloc_concat(n, d // 4, hh // 4, w // 4, 4 * f, 4 * f, dev)
        _, up4, skip4 = F.alloc_concat(n, d // 8, hh // 8, w // 8, 8 * f, 8 * f, dev)

        p1, enc1 = F.max_pool2_skip(self._run_block(self.encoder1, h, out=skip1))
        p2, enc2 = F.max_pool2_skip(self._run_block(self.encoder2, p1, out=skip2))
        enc3 = F.add(self._run_block(self.encoder3, p2), self._down4(self.encoder_r_1, enc1), out=skip3)          # :58
        p3, enc3 = F.max_pool2_skip(enc3)
        enc4 = F.add(self._run_block(self.encoder4, p3), self._down4(self.encoder_r_2, enc2), out=skip4)          # :60
        p4, enc4 = F.max_pool2_skip(enc4)
        bottleneck = F.add(self._run_block(self.bottleneck, p4), self._down4(self.encoder_r_3, enc3))             # :63

        dec4 = F.conv_transpose_kxsx(bottleneck, self.upconv4.weight, self.upconv4.bias, out=up4)
        dec4 = self._run_block(self.decoder4, dec4, enc4)
        dec3 = F.conv_transpose_kxsx(dec4, self.upconv3.weight, self.upconv3.bias)
        dec3 = F.add(dec3, self._up4(self.dncoder_r_1, bottleneck), out=up3)                                      # :69
        dec3 = self._run_block(self.decoder3, dec3, enc3)
        dec2 = F.conv_transpose_kxsx(dec3, self.upconv2.weight, self.upconv2.bias)
        dec2 = F.add(dec2, self._up4(self.dncoder_r_2, dec4), out=up2)                                            # :72
        dec2 = self._run_block(self.decoder2, dec2, enc2)
        dec1 = F.conv_transpose_kxsx(dec2, self.upconv1.weight, self.upconv1.bias)
        dec1 = F.add(dec1, self._up4(self.dncoder_r_3, dec3), out=up1)                                            # :75
        dec1 = self._run_block(self.decoder1, dec1, enc1)
        return F.head_conv1x1(dec1, self.conv.weight, self.conv.bias)
