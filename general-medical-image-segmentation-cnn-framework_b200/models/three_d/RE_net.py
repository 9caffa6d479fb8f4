"""RE-Net with the reference's constructor and state_dict (models/three_d/RE_net.py:37-158) on b200seg kernels.

Residual encoders (1x1x1 shortcut + two Conv-BN-ReLU, :37-54), three levels of reverse attention -- the coarser level is
projected to one channel, up-sampled by ConvTranspose3d(1, 1, 2, 2) and gates the finer skip connection as
`enc * (1 - sigmoid(g)) + enc` (:116-139) -- concat decoders (:57-71) and a sigmoid on the 2-class output (:157-158).
The 3x3x3 / 1x1x1 convolutions, pools and up-convolutions are the U-Net's kernels; the gates are csrc/gates.cu.
"""
import torch.nn as nn

from .._common import OpsMixin, norm_args, norm_spec


def downsample():
    return nn.MaxPool3d(2, 2)


def deconv(in_channels, out_channels):
    return nn.ConvTranspose3d(in_channels, out_channels, 2, 2)


def initialize_weights(*models):
    """Kaiming-normal conv / linear weights, zero biases, unit BatchNorm (:26-35)."""
    for m in (mod for model in models for mod in model.modules()):
        if isinstance(m, (nn.Conv3d, nn.Linear)):
            nn.init.kaiming_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm3d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def _conv_bn_pair(owner, cin, cout):
    """Registers conv1 / bn1 / conv2 / bn2 / relu / conv1x1 on `owner` in the reference's order (state_dict keys)."""
    for i, c in ((1, cin), (2, cout)):
        setattr(owner, "conv%d" % i, nn.Conv3d(c, cout, 3, padding=1))
        setattr(owner, "bn%d" % i, nn.BatchNorm3d(cout))
    owner.relu = nn.ReLU(inplace=False)
    owner.conv1x1 = nn.Conv3d(cin, cout, 1)


def _register_trunk(net, in_channels, widths=(32, 64, 128, 256)):
    """Encoders, bridge, the three 1-channel projections and their k2s2 up-samplers (:79-92), in the reference's order."""
    chans = (in_channels,) + tuple(widths)
    for i in range(3):
        setattr(net, "encoder%d" % (i + 1), ResEncoder(chans[i], chans[i + 1]))
    net.bridge = ResEncoder(chans[3], chans[4])
    for name, c in (("conv1_1", widths[3]), ("conv2_2", widths[2]), ("conv3_3", widths[1])):
        setattr(net, name, nn.Conv3d(c, 1, 1))
    for i in (1, 2, 3):
        setattr(net, "convTrans%d" % i, nn.ConvTranspose3d(1, 1, 2, 2))


def _register_ups(net, widths=(32, 64, 128, 256)):
    net.down = downsample()
    for i in (3, 2, 1):
        setattr(net, "up%d" % i, deconv(widths[i], widths[i - 1]))


class ResEncoder(nn.Module, OpsMixin):
    def __init__(self, in_channels, out_channels):
        super(ResEncoder, self).__init__()
        _conv_bn_pair(self, in_channels, out_channels)

    def forward(self, x, out=None):
        F = self.kernels
        residual = F.conv_norm_act(x, self.conv1x1.weight, self.conv1x1.bias, k=1, stride=1, pad=0, dil=1)
        h = F.conv_norm_act(x, self.conv1.weight, self.conv1.bias, k=3, stride=1, pad=1, dil=1,
                            spec=norm_spec(F, self.bn1, "relu", 0.0, self.training), **norm_args(self.bn1))
        h = F.conv_norm_act(h, self.conv2.weight, self.conv2.bias, k=3, stride=1, pad=1, dil=1,
                            spec=norm_spec(F, self.bn2, "relu", 0.0, self.training), **norm_args(self.bn2))
        return F.activation(h, "relu", residual=residual, out=out)       # relu(out + residual)


class Decoder(nn.Module, OpsMixin):
    def __init__(self, in_channels, out_channels):
        super(Decoder, self).__init__()
        layers = []
        for c in (in_channels, out_channels):
            layers += [nn.Conv3d(c, out_channels, 3, padding=1), nn.BatchNorm3d(out_channels), nn.ReLU(inplace=True)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x, x2=None):
        """x2: second half of the channel concatenation (torch.cat((up, skip), 1) at :143-152), never materialised."""
        F = self.kernels
        c = self.conv
        h = F.conv_norm_act(x, c[0].weight, c[0].bias, x2=x2, k=3, stride=1, pad=1, dil=1,
                            spec=norm_spec(F, c[1], "relu", 0.0, self.training), **norm_args(c[1]))
        return F.conv_norm_act(h, c[3].weight, c[3].bias, k=3, stride=1, pad=1, dil=1,
                               spec=norm_spec(F, c[4], "relu", 0.0, self.training), **norm_args(c[4]))


def reverse_attention(F, coarse, fine, proj, up, out=None):
    """`g = up(proj(coarse)); x = -1 * sigmoid(g) + 1; x = x.expand(...).mul(fine); x + fine` (:116-121)."""
    g = F.convt_map_k2s2(F.head_conv1x1(coarse, proj.weight, proj.bias), up.weight, up.bias)
    return F.reverse_gate(fine, g, out=out)


class RE_Net(nn.Module, OpsMixin):
    def __init__(self):
        super(RE_Net, self).__init__()
        _register_trunk(self, 1)
        for i, c in ((3, 128), (2, 64), (1, 32)):
            setattr(self, "decoder%d" % i, Decoder(2 * c, c))
        _register_ups(self)
        self.final = nn.Conv3d(32, 2, 1)
        initialize_weights(self)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        h = F.to_ndhwc(x)
        n, d, hh, w = F.spatial(h)
        dev = F.device_of(h)
        # [up-convolution | gated skip] halves of one buffer per level (the torch.cat of :143-152)
        _, up1, skip1 = F.alloc_concat(n, d, hh, w, 32, 32, dev)
        _, up2, skip2 = F.alloc_concat(n, d // 2, hh // 2, w // 2, 64, 64, dev)
        _, up3, skip3 = F.alloc_concat(n, d // 4, hh // 4, w // 4, 128, 128, dev)

        enc1 = self.encoder1(h)
        enc2 = self.encoder2(F.max_pool2(enc1))
        x3 = reverse_attention(F, enc2, enc1, self.conv3_3, self.convTrans3, out=skip1)
        enc3 = self.encoder3(F.max_pool2(enc2))
        x2 = reverse_attention(F, enc3, enc2, self.conv2_2, self.convTrans2, out=skip2)
        bridge = self.bridge(F.max_pool2(enc3))
        x1 = reverse_attention(F, bridge, enc3, self.conv1_1, self.convTrans1, out=skip3)

        dec3 = self.decoder3(F.conv_transpose_kxsx(bridge, self.up3.weight, self.up3.bias, out=up3), x1)
        dec2 = self.decoder2(F.conv_transpose_kxsx(dec3, self.up2.weight, self.up2.bias, out=up2), x2)
        dec1 = self.decoder1(F.conv_transpose_kxsx(dec2, self.up1.weight, self.up1.bias, out=up1), x3)
        return F.sigmoid_map(F.head_conv1x1(dec1, self.final.weight, self.final.bias))
