"""DenseVoxelNet with the reference's constructor and state_dict (models/three_d/densevoxelnet3d.py:17-128).

Dense blocks grow one channels-last buffer in place: layer i normalises + activates the first 16+12i channels, convolves
them to 12 new channels and writes those straight behind (replacing torch.cat, :33).  Reference quirks that are kept:
`_DenseLayer` applies its dropout twice in train mode (:25-32); the network returns only the
auxiliary branch `y2` (:126-128), so `dense_2` / `up_block` receive no gradient -- they are still evaluated in training
mode because their BatchNorm running statistics change, and skipped in eval mode where they cannot affect anything.
"""
import torch
import torch.nn as nn

from .._common import OpsMixin, norm_args, norm_spec


class _DenseLayer(nn.Sequential, OpsMixin):
    def __init__(self, num_input_features, growth_rate, bn_size, drop_rate=0.2):
        super(_DenseLayer, self).__init__()
        self.add_module('norm1', nn.BatchNorm3d(num_input_features))
        self.add_module('relu1', nn.ReLU(inplace=True))
        self.add_module('conv1', nn.Conv3d(num_input_features, bn_size * growth_rate, kernel_size=3, stride=1,
                                           padding=1, bias=False))
        self.drop_rate = drop_rate
        if self.drop_rate > 0:
            self.drop_layer = nn.Dropout(p=self.drop_rate)

    def forward(self, x, out=None, shared_grad=False):
        """Returns only the new features (the caller owns the concatenation buffer)."""
        F = self.kernels
        extra = dict(shared_grad=True) if shared_grad else {}
        h = F.norm_act(x, norm_spec(F, self.norm1, "relu", 0.0, self.training), **norm_args(self.norm1), **extra)
        plain_out = out if (self.drop_rate == 0 or not self.training) else None
        new = F.conv_norm_act(h, self.conv1.weight, None, k=3, stride=1, pad=1, dil=1, out=plain_out)
        if self.drop_rate > 0 and self.training:
            # once inside nn.Sequential.forward (:30), once more by the explicit call (:31-32): two masks, one pass
            new = F.dropout(new, self.drop_rate, training=True, out=out, times=2)
        return new


class _DenseBlock(nn.Sequential, OpsMixin):
    def __init__(self, num_layers, num_input_features, bn_size, growth_rate, drop_rate=0.2):
        super(_DenseBlock, self).__init__()
        for i in range(num_layers):
            layer = _DenseLayer(num_input_features + i * growth_rate, growth_rate, bn_size, drop_rate)
            self.add_module('denselayer%d' % (i + 1), layer)

    def forward(self, x):
        F = self.kernels
        alloc = getattr(F, "alloc_channels", None)
        if alloc is None:                    # reference arithmetic (oracle backend): torch.cat per layer (:41)
            for layer in self:
                x = F.concat_channels(x, layer(x))
            return x
        # one channels-last buffer for the whole block: the input is copied to its head once, every layer writes its
        # `growth` new channels straight behind what exists, and the "concatenation" is a wider view of the same memory
        layers = list(self)
        growth = layers[0].conv1.out_channels
        n, d, h, w = F.spatial(x)
        c = F.channels(x)
        buf = alloc(n, d, h, w, c + len(layers) * growth, F.device_of(x))
        cur = F.activation(x, "none", out=buf[..., :c])
        for layer in layers:
            # `cur` feeds the layer and is the head of the next concatenation: both gradients meet in one buffer
            cur = F.concat_channels_shared(cur, layer(cur, out=buf[..., c:c + growth], shared_grad=True))
            c += growth
        return cur


class _Transition(nn.Module, OpsMixin):
    def __init__(self, num_input_features, num_output_features):
        super().__init__()
        self.conv = nn.Sequential(nn.BatchNorm3d(num_input_features), nn.ReLU(inplace=True),
                                  nn.Conv3d(num_input_features, num_output_features, 1))
        self.max_pool = nn.MaxPool3d(2, 2)

    def forward(self, x):
        F = self.kernels
        norm, conv = self.conv[0], self.conv[2]
        h = F.norm_act(x, norm_spec(F, norm, "relu", 0.0, self.training), **norm_args(norm))
        k = F.conv_norm_act(h, conv.weight, conv.bias, k=1, stride=1, pad=0, dil=1)
        return F.max_pool2(k), k


class _Upsampling(nn.Sequential, OpsMixin):
    def __init__(self, input_features, out_features):
        super().__init__()
        self.tr_conv1_features, self.tr_conv2_features = 128, out_features
        widths = (input_features, self.tr_conv1_features, self.tr_conv2_features)
        children = [('norm', nn.BatchNorm3d(input_features)), ('relu', nn.ReLU(inplace=True)),
                    ('conv', nn.Conv3d(input_features, input_features, 1, bias=False))]
        children += [('transp_conv_%d' % (i + 1), nn.ConvTranspose3d(widths[i], widths[i + 1], 2, stride=2)) for i in (0, 1)]
        for name, child in children:          # names and order of the reference (:63-76): they are the state_dict keys
            self.add_module(name, child)

    def forward(self, x):
        F = self.kernels
        h = F.norm_act(x, norm_spec(F, self.norm, "relu", 0.0, self.training), **norm_args(self.norm))
        h = F.conv_norm_act(h, self.conv.weight, None, k=1, stride=1, pad=0, dil=1)
        h = F.conv_transpose_k2s2(h, self.transp_conv_1.weight, self.transp_conv_1.bias)
        return F.conv_transpose_k2s2(h, self.transp_conv_2.weight, self.transp_conv_2.bias)


class DenseVoxelNet(nn.Module, OpsMixin):
    """Implementation based on https://arxiv.org/abs/1708.00573 (reference densevoxelnet3d.py:90-128)."""

    def __init__(self, in_channels=1, classes=2):
        super().__init__()
        stem, growth, layers = 16, 12, 12
        d1 = stem + layers * growth                 # 160 channels after the first dense block
        d2 = d1 + layers * growth                   # 304 after the second
        self.dense_1_out_features, self.dense_2_out_features, self.up_out_features = d1, d2, 64
        self.classes, self.in_channels = classes, in_channels
        self.conv_init = nn.Conv3d(in_channels, stem, 1, stride=2, bias=False)
        self.dense_1 = _DenseBlock(layers, stem, bn_size=1, growth_rate=growth)
        self.trans = _Transition(d1, d1)
        self.dense_2 = _DenseBlock(layers, d1, bn_size=1, growth_rate=growth)
        self.up_block = _Upsampling(d2, self.up_out_features)
        self.conv_final = nn.Conv3d(self.up_out_features, classes, 1, bias=False)
        self.transpose = nn.ConvTranspose3d(d1, self.up_out_features, 2, stride=2)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        h = F.to_ndhwc(x)
        h = F.conv_norm_act(h, self.conv_init.weight, None, k=1, stride=2, pad=0, dil=1)
        h = self.dense_1(h)
        pooled, t = self.trans(h)
        if self.training:
            # main path: its result (y1) is discarded by the reference (:121-123), only the BatchNorm running
            # statistics of dense_2 / up_block change
            with torch.no_grad():
                self.up_block(self.dense_2(pooled))
        t = F.conv_transpose_k2s2(t, self.transpose.weight, self.transpose.bias)
        return F.head_conv1x1(t, self.conv_final.weight, None)
