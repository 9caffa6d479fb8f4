"""ResidualBlock with the reference's signature (utils/residual.py:11-85) on b200seg kernels."""
import torch.nn as nn

from ..models._common import OpsMixin
from .convolution import ConvolutionalBlock

BATCH_DIM = 0
CHANNELS_DIM = 1


class ResidualBlock(nn.Module, OpsMixin):
    def __init__(self, in_channels, out_channels, num_layers, dilation, dimensions, batch_norm=True,
                 instance_norm=False, residual=True, residual_type='pad', padding_mode='constant'):
        assert residual_type in ('pad', 'project')
        super().__init__()
        self.residual, self.residual_type, self.dimensions = residual, residual_type, dimensions
        self.change_dimension = in_channels != out_channels
        if self.change_dimension and residual_type == 'project':      # 1x1x1 projection of the shortcut (residual.py:36-43)
            self.change_dim_layer = nn.Conv3d(in_channels, out_channels, kernel_size=1, dilation=dilation, bias=False)
        norm = dict(batch_norm=batch_norm, instance_norm=instance_norm, padding_mode=padding_mode)
        widths = [in_channels] + [out_channels] * num_layers
        self.residual_block = nn.Sequential(*(ConvolutionalBlock(cin, cout, dilation, dimensions, **norm)
                                              for cin, cout in zip(widths[:-1], widths[1:])))

    def forward(self, x):
        F = self.kernels
        out = self.residual_block(x)
        if self.residual:
            if self.change_dimension and self.residual_type == 'project':
                x = F.conv_norm_act(x, self.change_dim_layer.weight, None, k=1, stride=1, pad=0, dil=1)
                return F.add(x, out)
            # 'pad' (residual.py:74-83): the shortcut is zero-padded symmetrically to the wider channel count
            return F.add_channel_padded(out, x)
        return out
