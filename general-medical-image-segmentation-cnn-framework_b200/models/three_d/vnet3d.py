"""V-Net with the reference's constructors and state_dict (models/three_d/vnet3d.py:14-157) on b200seg kernels.

Each (Conv3d k5 -> BatchNorm3d -> ELU/PReLU) triple is one conv (+ fused statistics) and one normalise+activate pass;
the residual sums in front of an activation (`relu(add(out, x))`, vnet3d.py:58,79,103) ride in that same pass; the
decoder's `torch.cat((out, skipxdo), 1)` (vnet3d.py:100) is two producers writing the halves of one buffer.
"""
import torch
import torch.nn as nn

from .._common import OpsMixin, act_of, conv_args, norm_args, norm_spec


def passthrough(x, **kwargs):
    return x


def ELUCons(elu, nchan):
    if elu:
        return nn.ELU(inplace=True)
    return nn.PReLU(nchan)


def _conv_bn_act(mod, F, conv, bn, act_mod, x, x2=None, residual=None, out=None):
    act, ap, pw = act_of(act_mod)
    return F.conv_norm_act(x, conv.weight, conv.bias, x2=x2, spec=norm_spec(F, bn, act, ap, mod.training),
                           prelu_weight=pw, residual=residual, out=out, **conv_args(conv), **norm_args(bn))


def _drop(mod, F, do, x, out=None):
    if isinstance(do, nn.Dropout3d):
        return F.dropout(x, do.p, training=mod.training, channel=True, out=out)
    if out is not None:
        out.copy_(x)
        return out
    return x


class LUConv(nn.Module, OpsMixin):
    def __init__(self, nchan, elu):
        super(LUConv, self).__init__()
        self.relu1 = ELUCons(elu, nchan)
        self.conv1 = nn.Conv3d(nchan, nchan, kernel_size=5, padding=2)
        self.bn1 = torch.nn.BatchNorm3d(nchan)

    def forward(self, x, x2=None):
        return _conv_bn_act(self, self.kernels, self.conv1, self.bn1, self.relu1, x, x2=x2)


def _make_nConv(nchan, depth, elu):
    return nn.Sequential(*[LUConv(nchan, elu) for _ in range(depth)])


class InputTransition(nn.Module, OpsMixin):
    def __init__(self, in_channels, elu):
        super(InputTransition, self).__init__()
        self.num_features = 16
        self.in_channels = in_channels
        self.conv1 = nn.Conv3d(self.in_channels, self.num_features, kernel_size=5, padding=2)
        self.bn1 = torch.nn.BatchNorm3d(self.num_features)
        self.relu1 = ELUCons(elu, self.num_features)

    def forward(self, x):
        F = self.kernels
        x16 = F.repeat_channels(x, int(self.num_features / self.in_channels))
        return _conv_bn_act(self, F, self.conv1, self.bn1, self.relu1, x, residual=x16)


class DownTransition(nn.Module, OpsMixin):
    def __init__(self, inChans, nConvs, elu, dropout=False):
        super(DownTransition, self).__init__()
        outChans = 2 * inChans
        self.down_conv = nn.Conv3d(inChans, outChans, kernel_size=2, stride=2)
        self.bn1 = torch.nn.BatchNorm3d(outChans)
        self.do1 = passthrough
        self.relu1 = ELUCons(elu, outChans)
        self.relu2 = ELUCons(elu, outChans)
        if dropout:
            self.do1 = nn.Dropout3d()
        self.ops = _make_nConv(outChans, nConvs, elu)

    def forward(self, x, out=None):
        F = self.kernels
        down = _conv_bn_act(self, F, self.down_conv, self.bn1, self.relu1, x)
        h = _drop(self, F, self.do1, down)
        for layer in self.ops:
            h = layer(h)
        act, ap, pw = act_of(self.relu2)
        return F.activation(h, act, ap, prelu_weight=pw, residual=down, out=out)


class UpTransition(nn.Module, OpsMixin):
    def __init__(self, inChans, outChans, nConvs, elu, dropout=False):
        super(UpTransition, self).__init__()
        self.up_conv = nn.ConvTranspose3d(inChans, outChans // 2, kernel_size=2, stride=2)
        self.bn1 = torch.nn.BatchNorm3d(outChans // 2)
        self.do1 = passthrough
        self.do2 = nn.Dropout3d()
        self.relu1 = ELUCons(elu, outChans // 2)
        self.relu2 = ELUCons(elu, outChans)
        if dropout:
            self.do1 = nn.Dropout3d()
        self.ops = _make_nConv(outChans, nConvs, elu)

    def forward(self, x, skipx):
        F = self.kernels
        h = _drop(self, F, self.do1, x)
        n, d, hh, w = F.spatial(skipx)
        half = self.up_conv.out_channels
        _, first, second = F.alloc_concat(n, d, hh, w, half, F.channels(skipx), F.device_of(skipx))
        skipxdo = _drop(self, F, self.do2, skipx, out=second)
        up = F.conv_transpose_k2s2(h, self.up_conv.weight, self.up_conv.bias)
        act, ap, pw = act_of(self.relu1)
        up = F.norm_act(up, norm_spec(F, self.bn1, act, ap, self.training), prelu_weight=pw, out=first,
                        **norm_args(self.bn1))
        xcat = F.concat_channels(up, skipxdo)
        h = xcat
        for layer in self.ops:
            h = layer(h)
        act, ap, pw = act_of(self.relu2)
        return F.activation(h, act, ap, prelu_weight=pw, residual=xcat)


class OutputTransition(nn.Module, OpsMixin):
    def __init__(self, in_channels, classes, elu):
        super(OutputTransition, self).__init__()
        self.classes = classes
        self.conv1 = nn.Conv3d(in_channels, classes, kernel_size=5, padding=2)
        self.bn1 = torch.nn.BatchNorm3d(classes)
        self.conv2 = nn.Conv3d(classes, classes, kernel_size=1)
        self.relu1 = ELUCons(elu, classes)

    def forward(self, x):
        F = self.kernels
        out = _conv_bn_act(self, F, self.conv1, self.bn1, self.relu1, x)
        return F.head_conv1x1(out, self.conv2.weight, self.conv2.bias)


class VNet(nn.Module, OpsMixin):
    """Implementations based on the Vnet paper: https://arxiv.org/abs/1606.04797 (reference vnet3d.py:124-157)."""

    def __init__(self, elu=True, in_channels=1, classes=2):
        super(VNet, self).__init__()
        self.classes = classes
        self.in_channels = in_channels
        self.in_tr = InputTransition(in_channels, elu=elu)
        self.down_tr32 = DownTransition(16, 1, elu)
        self.down_tr64 = DownTransition(32, 2, elu)
        self.down_tr128 = DownTransition(64, 3, elu, dropout=False)
        self.down_tr256 = DownTransition(128, 2, elu, dropout=False)
        self.up_tr256 = UpTransition(256, 256, 2, elu, dropout=False)
        self.up_tr128 = UpTransition(256, 128, 2, elu, dropout=False)
        self.up_tr64 = UpTransition(128, 64, 1, elu)
        self.up_tr32 = UpTransition(64, 32, 1, elu)
        self.out_tr = OutputTransition(32, classes, elu)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        h = F.to_ndhwc(x)
        out16 = self.in_tr(h)
        out32 = self.down_tr32(out16)
        out64 = self.down_tr64(out32)
        out128 = self.down_tr128(out64)
        out256 = self.down_tr256(out128)
        out = self.up_tr256(out256, out128)
        out = self.up_tr128(out, out64)
        out = self.up_tr64(out, out32)
        out = self.up_tr32(out, out16)
        return self.out_tr(out)
