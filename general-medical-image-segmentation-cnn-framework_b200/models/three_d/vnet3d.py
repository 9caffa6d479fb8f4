"""V-Net with the reference's constructors and state_dict (models/three_d/vnet3d.py:14-157) on b200seg kernels.

Each (Conv3d k5 -> BatchNorm3d -> ELU/PReLU) triple is one conv (+ fused statistics) and one normalise+activate pass;
the residual sums in front of an activation (`relu(add(out, x))`, vnet3d.py:58,79,103) ride in that same pass; the
decoder's `torch.cat((out, skipxdo), 1)` (vnet3d.py:100) is two producers writing the halves of one buffer.
"""
import torch.nn as nn

from .._common import OpsMixin, act_of, conv_args, norm_args, norm_spec


def passthrough(x, **kwargs):
    return x


def ELUCons(elu, nchan):
    """ELU, or a per-channel PReLU (vnet3d.py:14-18)."""
    return nn.ELU(inplace=True) if elu else nn.PReLU(nchan)


def _conv5(cin, cout):
    return nn.Conv3d(cin, cout, 5, padding=2)


def _register(module, **children):
    """Assign sub-modules in the given order (= the reference's registration order, hence its state_dict key order)."""
    for name, child in children.items():
        setattr(module, name, child)


def _conv_bn_act(mod, F, conv, bn, act_mod, x, x2=None, residual=None, out=None):
    act, ap, pw = act_of(act_mod)
    return F.conv_norm_act(x, conv.weight, conv.bias, x2=x2, spec=norm_spec(F, bn, act, ap, mod.training),
                           prelu_weight=pw, residual=residual, out=out, **conv_args(conv), **norm_args(bn))


def _drop(mod, F, do, x, out=None):
    if isinstance(do, nn.Dropout3d):
        return F.dropout(x, do.p, training=mod.training, channel=True, out=out)
    if out is not None:
        return F.activation(x, "none", out=out)        # differentiable copy into the concatenation buffer
    return x


class LUConv(nn.Module, OpsMixin):
    def __init__(self, nchan, elu):
        super().__init__()
        _register(self, relu1=ELUCons(elu, nchan), conv1=_conv5(nchan, nchan), bn1=nn.BatchNorm3d(nchan))

    def forward(self, x, x2=None):
        return _conv_bn_act(self, self.kernels, self.conv1, self.bn1, self.relu1, x, x2=x2)


def _make_nConv(nchan, depth, elu):
    return nn.Sequential(*[LUConv(nchan, elu) for _ in range(depth)])


class InputTransition(nn.Module, OpsMixin):
    def __init__(self, in_channels, elu):
        super().__init__()
        self.num_features, self.in_channels = 16, in_channels
        _register(self, conv1=_conv5(in_channels, 16), bn1=nn.BatchNorm3d(16), relu1=ELUCons(elu, 16))

    def forward(self, x):
        F = self.kernels
        x16 = F.repeat_channels(x, int(self.num_features / self.in_channels))
        return _conv_bn_act(self, F, self.conv1, self.bn1, self.relu1, x, residual=x16)


class DownTransition(nn.Module, OpsMixin):
    def __init__(self, inChans, nConvs, elu, dropout=False):
        super().__init__()
        wide = 2 * inChans
        _register(self, down_conv=nn.Conv3d(inChans, wide, 2, stride=2), bn1=nn.BatchNorm3d(wide),
                  relu1=ELUCons(elu, wide), relu2=ELUCons(elu, wide))
        self.do1 = nn.Dropout3d() if dropout else passthrough
        self.ops = _make_nConv(wide, nConvs, elu)

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
        super().__init__()
        half = outChans // 2
        _register(self, up_conv=nn.ConvTranspose3d(inChans, half, 2, stride=2), bn1=nn.BatchNorm3d(half),
                  do2=nn.Dropout3d(), relu1=ELUCons(elu, half), relu2=ELUCons(elu, outChans))
        self.do1 = nn.Dropout3d() if dropout else passthrough
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
        super().__init__()
        self.classes = classes
        _register(self, conv1=_conv5(in_channels, classes), bn1=nn.BatchNorm3d(classes), conv2=nn.Conv3d(classes, classes, 1),
                  relu1=ELUCons(elu, classes))

    def forward(self, x):
        F = self.kernels
        out = _conv_bn_act(self, F, self.conv1, self.bn1, self.relu1, x)
        return F.head_conv1x1(out, self.conv2.weight, self.conv2.bias)


class VNet(nn.Module, OpsMixin):
    """Implementations based on the Vnet paper: https://arxiv.org/abs/1606.04797 (reference vnet3d.py:124-157)."""

    def __init__(self, elu=True, in_channels=1, classes=2):
        super().__init__()
        self.classes, self.in_channels = classes, in_channels
        self.in_tr = InputTransition(in_channels, elu=elu)
        # (input channels, LUConv count) of the four encoder stages, then (in, out, LUConv count) of the decoder stages
        for cin, depth in ((16, 1), (32, 2), (64, 3), (128, 2)):
            setattr(self, "down_tr%d" % (2 * cin), DownTransition(cin, depth, elu, dropout=False))
        for cin, cout, depth in ((256, 256, 2), (256, 128, 2), (128, 64, 1), (64, 32, 1)):
            setattr(self, "up_tr%d" % cout, UpTransition(cin, cout, depth, elu, dropout=False))
        self.out_tr = OutputTransition(32, classes, elu)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        skips = [self.in_tr(F.to_ndhwc(x))]                      # 16 channels at full resolution
        for stage in (self.down_tr32, self.down_tr64, self.down_tr128, self.down_tr256):
            skips.append(stage(skips[-1]))
        out = skips.pop()
        for stage in (self.up_tr256, self.up_tr128, self.up_tr64, self.up_tr32):
            out = stage(out, skips.pop())
        return self.out_tr(out)
