"""HighRes3DNet with the reference's constructor and state_dict (models/three_d/highresnet.py:13-143): a first
3x3x3 conv, three dilation stages (d = 1, 2, 4; 16 / 32 / 64 channels) of pre-activation residual blocks, a 1x1x1
classifier followed by BatchNorm.  Everything runs at full resolution, so every conv is the persistent tcgen05
plane kernel with dilation folded into the tap offsets."""
import torch
import torch.nn as nn

from ...utils.convolution import ConvolutionalBlock
from ...utils.dilation import DilationBlock
from .._common import OpsMixin

__all__ = ['HighResNet', 'HighRes3DNet']


class HighResNet(nn.Module, OpsMixin):
    def __init__(self, in_channels, out_channels, dimensions=None, initial_out_channels_power=4,
                 layers_per_residual_block=2, residual_blocks_per_dilation=3, dilations=3, batch_norm=True,
                 instance_norm=False, residual=True, padding_mode='constant', add_dropout_layer=False):
        assert dimensions in (2, 3)
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.layers_per_residual_block = layers_per_residual_block
        self.residual_blocks_per_dilation = residual_blocks_per_dilation
        self.dilations = dilations
        norm = dict(dimensions=dimensions, batch_norm=batch_norm, instance_norm=instance_norm)

        def plain(cin, cout, **kw):      # a post-activation conv block at dilation 1 (highresnet.py:39-49, 81-105)
            return ConvolutionalBlock(in_channels=cin, out_channels=cout, dilation=1, preactivation=False, **norm, **kw)

        width = 2 ** initial_out_channels_power
        stages = [plain(in_channels, width, padding_mode=padding_mode)]
        cin = width
        for idx in range(dilations):     # dilation 1, 2, 4, ... at width, 2 width, 4 width, ... (highresnet.py:52-72)
            stages.append(DilationBlock(cin, width << idx, 2 ** idx, dimensions, layers_per_block=layers_per_residual_block,
                                        num_residual_blocks=residual_blocks_per_dilation, batch_norm=batch_norm,
                                        instance_norm=instance_norm, residual=residual, padding_mode=padding_mode))
            cin = width << idx
        self._dropout = None
        if add_dropout_layer:
            stages += [plain(cin, 80, kernel_size=1), nn.Dropout3d()]
            cin = 80
        stages.append(plain(cin, self.out_channels, kernel_size=1, activation=False, padding_mode=padding_mode))
        self.block = nn.Sequential(*stages)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(x.dim()))
        F = self.kernels
        h = F.to_ndhwc(x)
        for blk in self.block:
            if isinstance(blk, nn.Dropout3d):
                h = F.dropout(h, blk.p, training=self.training, channel=True)
            else:
                h = blk(h)
        return F.from_ndhwc(h)

    @property
    def num_parameters(self):
        return sum(p.numel() for p in self.parameters())

    @property
    def receptive_field(self):
        """B conv layers per residual block, N residual blocks per dilation factor, D dilation factors."""
        B, D, N = self.layers_per_residual_block, self.dilations, self.residual_blocks_per_dilation
        d = torch.arange(D)
        return (3 - 1) + torch.sum(B * N * 2 ** (d + 1)) + 1

    def get_receptive_field_world(self, spacing=1):
        return self.receptive_field * spacing


class HighRes3DNet(HighResNet):
    def __init__(self, *args, **kwargs):
        kwargs['dimensions'] = 3
        super().__init__(*args, **kwargs)
