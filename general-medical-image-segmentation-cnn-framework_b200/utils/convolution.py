"""ConvolutionalBlock / Pad3d with the reference's signature (utils/convolution.py:12-86) on b200seg kernels.

[norm -> ReLU ->] pad(dilation) -> Conv(k, dilation, bias = not norm) [-> norm -> ReLU].  The explicit `Pad3d` of the
reference is folded into the convolution (TMA out-of-bounds fill), which is exact for the default 'constant' mode;
'reflect' / 'replicate' run an explicit pad kernel (b200seg_pad3d_fwd / _bwd) followed by an unpadded convolution.
"""
import torch.nn as nn

from ..models._common import OpsMixin, norm_args, norm_spec

PADDING_MODES = {
    'reflect': 'Reflection',
    'replicate': 'Replication',
    'constant': 'Zero',
}


class Pad3d(nn.Module):
    """Kept so that the module indices inside `convolutional_block` (and hence the state_dict keys) match."""

    def __init__(self, pad, mode):
        assert mode in PADDING_MODES.keys()
        super().__init__()
        self.pad = 6 * [pad]
        self.mode = mode

    def forward(self, x):
        import torch.nn.functional as TF
        return TF.pad(x, self.pad, self.mode)


class ConvolutionalBlock(nn.Module, OpsMixin):
    def __init__(self, in_channels, out_channels, dilation, dimensions, batch_norm=True, instance_norm=False,
                 norm_affine=True, padding_mode='constant', preactivation=True, kernel_size=3, activation=True):
        assert padding_mode in PADDING_MODES.keys()
        assert not (batch_norm and instance_norm)
        super().__init__()
        if dimensions != 3:
            raise NotImplementedError("b200seg implements the volumetric (dimensions=3) path")
        self.padding_mode = padding_mode
        norm_class = nn.BatchNorm3d if batch_norm else (nn.InstanceNorm3d if instance_norm else None)
        layers = nn.ModuleList()
        pre_norm = post_norm = None
        if preactivation:
            if norm_class is not None:
                pre_norm = norm_class(in_channels, affine=norm_affine)
                layers.append(pre_norm)
            if activation:
                layers.append(nn.ReLU())
        if kernel_size > 1:
            layers.append(Pad3d(dilation, padding_mode))
        use_bias = not (instance_norm or batch_norm)
        conv = nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, dilation=dilation, bias=use_bias)
        layers.append(conv)
        if not preactivation:
            if norm_class is not None:
                post_norm = norm_class(out_channels, affine=norm_affine)
                layers.append(post_norm)
            if activation:
                layers.append(nn.ReLU())
        self.preactivation, self.activation = preactivation, activation
        self.kernel_size, self.dilation = kernel_size, dilation
        self.convolutional_block = nn.Sequential(*layers)

    # shortcuts into `convolutional_block` (the only registered sub-module, so the state_dict keys are the reference's)
    def _find(self, cls, first=True):
        mods = [m for m in self.convolutional_block if isinstance(m, cls)]
        return (mods[0] if first else mods[-1]) if mods else None

    @property
    def _conv(self):
        return self._find(nn.Conv3d)

    @property
    def _pre_norm(self):
        return self._find((nn.modules.batchnorm._BatchNorm, nn.modules.instancenorm._InstanceNorm)) \
            if self.preactivation else None

    @property
    def _post_norm(self):
        return None if self.preactivation else \
            self._find((nn.modules.batchnorm._BatchNorm, nn.modules.instancenorm._InstanceNorm))

    def forward(self, x):
        F = self.kernels
        act = "relu" if self.activation else "none"
        pad = self.dilation * (self.kernel_size - 1) // 2 if self.kernel_size > 1 else 0
        explicit = self.padding_mode != 'constant' and pad > 0     # reflect / replicate: pad kernel + 'valid' convolution
        geom = dict(k=self.kernel_size, stride=1, pad=0 if explicit else pad, dil=self.dilation)
        conv = self._conv
        if self.preactivation:
            if self._pre_norm is not None or self.activation:
                x = F.norm_act(x, norm_spec(F, self._pre_norm, act, 0.0, self.training), **norm_args(self._pre_norm))
            if explicit:
                x = F.pad3d(x, pad, self.padding_mode)
            return F.conv_norm_act(x, conv.weight, conv.bias, **geom)
        if explicit:
            x = F.pad3d(x, pad, self.padding_mode)
        return F.conv_norm_act(x, conv.weight, conv.bias, spec=norm_spec(F, self._post_norm, act, 0.0, self.training),
                               **geom, **norm_args(self._post_norm))
