"""DilationBlock with the reference's signature (utils/dilation.py:5-40): `num_residual_blocks` ResidualBlocks at one dilation,
the first of which changes the channel count.  The only registered child is `dilation_block` (state_dict keys)."""
import torch.nn as nn

from .residual import ResidualBlock


class DilationBlock(nn.Module):
    def __init__(self, in_channels, out_channels, dilation, dimensions, layers_per_block=2, num_residual_blocks=3,
                 batch_norm=True, instance_norm=False, residual=True, padding_mode='constant'):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        shared = dict(batch_norm=batch_norm, instance_norm=instance_norm, residual=residual, padding_mode=padding_mode)
        widths = [in_channels] + [out_channels] * num_residual_blocks
        self.dilation_block = nn.Sequential(*(
            ResidualBlock(cin, cout, layers_per_block, dilation, dimensions, **shared)
            for cin, cout in zip(widths[:-1], widths[1:])))

    def forward(self, x):
        return self.dilation_block(x)
