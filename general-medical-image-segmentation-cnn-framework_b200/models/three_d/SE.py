"""Squeeze-and-excitation gates with the reference's constructors and state_dict (models/three_d/SE.py:4-49) on b200seg
kernels: global average pool -> Linear(C, C/r) -> ReLU -> Linear(C/r, C) -> Sigmoid, then `x * y` (SE_Inception) or
`x + x * y` (SE_Residual).  Pool, scaling and their backward passes are kernels (csrc/gates.cu); the two Linear layers act
on an [N, C] tensor and stay torch ops (functional._GatedBlend)."""
import torch
import torch.nn as nn

from .._common import OpsMixin


class _SEBase(nn.Module, OpsMixin):
    residual = False

    def __init__(self, in_channels, reduction=16):
        super().__init__()
        self.gap = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.fc = nn.Sequential(
            nn.Linear(in_channels, in_channels // reduction, bias=False),
            nn.ReLU(),
            nn.Linear(in_channels // reduction, in_channels, bias=False),
            nn.Sigmoid()
        )

    def _gate(self, pooled, w_down, w_up):
        y = torch.sigmoid(torch.relu(pooled @ w_down.t()) @ w_up.t())
        return (1.0 + y if self.residual else y), None

    def forward(self, x, out=None):
        """x: an activation of the ops backend (channels-last bf16 on the CUDA backend)."""
        return self.kernels.gated_blend(x, None, self._gate, (self.fc[0].weight, self.fc[2].weight), out=out)


class SE_Inception(_SEBase):
    residual = False


class SE_Residual(_SEBase):
    residual = True
