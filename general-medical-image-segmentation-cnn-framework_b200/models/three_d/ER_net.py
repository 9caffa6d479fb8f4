"""ER-Net with the reference's constructor and state_dict (models/three_d/ER_net.py:37-231) on b200seg kernels.

The encoder, bridge and reverse-attention gates are RE-Net's (RE_net.py); the decoders fuse the up-sampled features and
the gated skip connection with a selective-fusion gate instead of a concatenation (SFConv, :57-105: global average pool of
the sum -> Linear -> one Linear per branch -> softmax over the two branches -> weighted sum), followed by BatchNorm + ReLU
and a residual block (SF_Decoder / ResDecoder, :37-54,108-131).
"""
import torch
import torch.nn as nn

from .._common import OpsMixin, norm_args, norm_spec
from .RE_net import (ResEncoder, _conv_bn_pair, _register_trunk, _register_ups, deconv, downsample,  # noqa: F401
                     initialize_weights, reverse_attention)


class ResDecoder(nn.Module, OpsMixin):
    def __init__(self, in_channels):
        super(ResDecoder, self).__init__()
        _conv_bn_pair(self, in_channels, in_channels)

    forward = ResEncoder.forward       # same graph: 1x1x1 shortcut + two Conv-BN-ReLU, relu(out + residual)


class SFConv(nn.Module, OpsMixin):
    def __init__(self, features, M=2, r=4, L=32):
        super(SFConv, self).__init__()
        d = max(int(features / r), L)
        self.M = M
        self.features = features
        self.fc = nn.Linear(features, d)
        self.fcs = nn.ModuleList([])
        for i in range(M):
            self.fcs.append(nn.Linear(d, features))
        self.softmax = nn.Softmax(dim=1)

    @staticmethod
    def _gate(pooled, w, b, w0, b0, w1, b1):
        z = pooled @ w.t() + b
        att = torch.softmax(torch.stack((z @ w0.t() + b0, z @ w1.t() + b1), dim=1), dim=1)
        return att[:, 0], att[:, 1]

    def forward(self, x1, x2):
        assert self.M == 2, "the reference calls SFConv with its default two branches"
        params = (self.fc.weight, self.fc.bias, self.fcs[0].weight, self.fcs[0].bias, self.fcs[1].weight, self.fcs[1].bias)
        return self.kernels.gated_blend(x1, x2, self._gate, params)


class SF_Decoder(nn.Module, OpsMixin):
    def __init__(self, out_channels):
        super(SF_Decoder, self).__init__()
        self.conv1 = SFConv(out_channels)
        self.bn1 = nn.BatchNorm3d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.ResDecoder = ResDecoder(out_channels)

    def forward(self, x1, x2):
        F = self.kernels
        fused = self.conv1(x1, x2)
        out = F.norm_act(fused, norm_spec(F, self.bn1, "relu", 0.0, self.training), **norm_args(self.bn1))
        return self.ResDecoder(out)


class ER_Net(nn.Module, OpsMixin):
    def __init__(self, classes, channels):
        super(ER_Net, self).__init__()
        _register_trunk(self, channels)
        for i, c in ((3, 128), (2, 64), (1, 32)):
            setattr(self, "decoder%d" % i, SF_Decoder(c))
        _register_ups(self)
        self.final = nn.Conv3d(32, classes, 1)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        h = F.to_ndhwc(x)
        enc1 = self.encoder1(h)
        enc2 = self.encoder2(F.max_pool2(enc1))
        x3 = reverse_attention(F, enc2, enc1, self.conv3_3, self.convTrans3)
        enc3 = self.encoder3(F.max_pool2(enc2))
        x2 = reverse_attention(F, enc3, enc2, self.conv2_2, self.convTrans2)
        bridge = self.bridge(F.max_pool2(enc3))
        x1 = reverse_attention(F, bridge, enc3, self.conv1_1, self.convTrans1)

        dec3 = self.decoder3(F.conv_transpose_kxsx(bridge, self.up3.weight, self.up3.bias), x1)
        dec2 = self.decoder2(F.conv_transpose_kxsx(dec3, self.up2.weight, self.up2.bias), x2)
        dec1 = self.decoder1(F.conv_transpose_kxsx(dec2, self.up1.weight, self.up1.bias), x3)
        return F.head_conv1x1(dec1, self.final.weight, self.final.bias)
