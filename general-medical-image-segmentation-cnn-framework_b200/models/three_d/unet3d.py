"""3D U-Net with the reference's constructor and state_dict (models/three_d/unet3d.py:10-104), running on b200seg kernels.

The sub-modules are ordinary nn.Conv3d / nn.BatchNorm3d / nn.ConvTranspose3d objects used purely as parameter
containers, so the 136 state_dict keys (`encoder1.enc1conv1.weight`, `encoder1.enc1norm1.running_mean`, ...),
`.apply(weights_init_normal)` (train.py:33-61) and checkpoints written by the reference load unchanged.  forward()
never calls them: each (Conv3d -> BatchNorm3d -> ReLU) triple is one fused conv+statistics kernel followed by one
normalise+ReLU pass, MaxPool3d keeps uint8 arg-max codes, and torch.cat is replaced by writing producers straight into
the two channel halves of one pre-allocated buffer.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from ... import functional as F
from ..sync_batchnorm.batchnorm import _SynchronizedBatchNorm


class UNet3D(nn.Module):
    def __init__(self, in_channels=1, out_channels=3, init_features=64):
        super(UNet3D, self).__init__()
        features = init_features
        self.encoder1 = UNet3D._block(in_channels, features, name="enc1")
        self.pool1 = nn.MaxPool3d(kernel_size=2, stride=2)
        self.encoder2 = UNet3D._block(features, features * 2, name="enc2")
        self.pool2 = nn.MaxPool3d(kernel_size=2, stride=2)
        self.encoder3 = UNet3D._block(features * 2, features * 4, name="enc3")
        self.pool3 = nn.MaxPool3d(kernel_size=2, stride=2)
        self.encoder4 = UNet3D._block(features * 4, features * 8, name="enc4")
        self.pool4 = nn.MaxPool3d(kernel_size=2, stride=2)
        self.bottleneck = UNet3D._block(features * 8, features * 16, name="bottleneck")
        self.upconv4 = nn.ConvTranspose3d(features * 16, features * 8, kernel_size=2, stride=2)
        self.decoder4 = UNet3D._block((features * 8) * 2, features * 8, name="dec4")
        self.upconv3 = nn.ConvTranspose3d(features * 8, features * 4, kernel_size=2, stride=2)
        self.decoder3 = UNet3D._block((features * 4) * 2, features * 4, name="dec3")
        self.upconv2 = nn.ConvTranspose3d(features * 4, features * 2, kernel_size=2, stride=2)
        self.decoder2 = UNet3D._block((features * 2) * 2, features * 2, name="dec2")
        self.upconv1 = nn.ConvTranspose3d(features * 2, features, kernel_size=2, stride=2)
        self.decoder1 = UNet3D._block(features * 2, features, name="dec1")
        self.conv = nn.Conv3d(in_channels=features, out_channels=out_channels, kernel_size=1)

    @staticmethod
    def _block(in_channels, features, name):
        return nn.Sequential(OrderedDict([
            (name + "conv1", nn.Conv3d(in_channels=in_channels, out_channels=features, kernel_size=3, padding=1,
                                       bias=True)),
            (name + "norm1", nn.BatchNorm3d(num_features=features)),
            (name + "relu1", nn.ReLU(inplace=True)),
            (name + "conv2", nn.Conv3d(in_channels=features, out_channels=features, kernel_size=3, padding=1,
                                       bias=True)),
            (name + "norm2", nn.BatchNorm3d(num_features=features)),
            (name + "relu2", nn.ReLU(inplace=True)),
        ]))

    # ------------------------------------------------------------------------------------------------------------
    def _conv_bn_relu(self, conv, norm, x, x2=None, out=None):
        sync = isinstance(norm, (_SynchronizedBatchNorm, nn.SyncBatchNorm))
        spec = F.NormSpec("batch", "relu", eps=norm.eps, momentum=0.1 if norm.momentum is None else norm.momentum,
                          training=self.training or not norm.track_running_stats, sync=sync,
                          clamp_eps=isinstance(norm, _SynchronizedBatchNorm),
                          process_group=getattr(norm, "process_group", None))
        if self.training and norm.track_running_stats and norm.num_batches_tracked is not None:
            F.bump_counter(norm.num_batches_tracked)
        return F.conv_norm_act(x, conv.weight, conv.bias, x2=x2, k=3, stride=1, pad=1, dil=1, spec=spec,
                               gamma=norm.weight, beta=norm.bias, running_mean=norm.running_mean,
                               running_var=norm.running_var, out=out)

    def _run_block(self, seq, x, x2=None, out=None):
        mods = list(seq.children())
        h = self._conv_bn_relu(mods[0], mods[1], x, x2)
        return self._conv_bn_relu(mods[3], mods[4], h, out=out)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        h = F.to_ndhwc(x)
        n, d, hh, w, _ = h.shape
        f = self.encoder1[3].out_channels
        dev = h.device
        # skip buffers: [upconv output | encoder output] share one allocation per level (replaces torch.cat :59-68)
        _, up1, skip1 = F.alloc_concat(n, d, hh, w, f, f, dev)
        _, up2, skip2 = F.alloc_concat(n, d // 2, hh // 2, w // 2, 2 * f, 2 * f, dev)
        _, up3, skip3 = F.alloc_concat(n, d // 4, hh // 4, w // 4, 4 * f, 4 * f, dev)
        _, up4, skip4 = F.alloc_concat(n, d // 8, hh // 8, w // 8, 8 * f, 8 * f, dev)

        # each encoder output feeds the pool AND the skip connection: max_pool2_skip sums both gradients in one kernel
        p1, enc1 = F.max_pool2_skip(self._run_block(self.encoder1, h, out=skip1))
        p2, enc2 = F.max_pool2_skip(self._run_block(self.encoder2, p1, out=skip2))
        p3, enc3 = F.max_pool2_skip(self._run_block(self.encoder3, p2, out=skip3))
        p4, enc4 = F.max_pool2_skip(self._run_block(self.encoder4, p3, out=skip4))
        bottleneck = self._run_block(self.bottleneck, p4)

        dec4 = F.conv_transpose_k2s2(bottleneck, self.upconv4.weight, self.upconv4.bias, out=up4)
        dec4 = self._run_block(self.decoder4, dec4, enc4)
        dec3 = F.conv_transpose_k2s2(dec4, self.upconv3.weight, self.upconv3.bias, out=up3)
        dec3 = self._run_block(self.decoder3, dec3, enc3)
        dec2 = F.conv_transpose_k2s2(dec3, self.upconv2.weight, self.upconv2.bias, out=up2)
        dec2 = self._run_block(self.decoder2, dec2, enc2)
        dec1 = F.conv_transpose_k2s2(dec2, self.upconv1.weight, self.upconv1.bias, out=up1)
        dec1 = self._run_block(self.decoder1, dec1, enc1)
        return F.head_conv1x1(dec1, self.conv.weight, self.conv.bias)
