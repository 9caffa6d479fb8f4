"""Oracle (test infrastructure): deterministic parameter initialisation shared by the golden-vector generator (which
applies it to the REFERENCE modules) and the tests (which apply it to the mirrors): parameters are visited in
state_dict order, which tests/test_models_cpu.py checks to be identical, so both sides get bit-identical weights
without shipping 45 M-parameter fixtures.  Follows the spirit of weights_init_normal('kaiming') (train.py:33-61):
fan-in scaled normal weights; biases, norm affines and PReLU slopes are perturbed so their gradients are exercised."""
import math

import torch


def init_module_(module, seed=0):
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        with torch.no_grad():
            if p.dim() >= 3:                          # conv / conv-transpose weights
                fan_in = p[0].numel() if p.dim() == 5 else p.numel()
                p.copy_(torch.randn(p.shape, generator=g) * math.sqrt(2.0 / max(fan_in, 1)))
            elif p.dim() == 2:                        # Linear weights (squeeze-and-excitation / selective-fusion gates)
                p.copy_(torch.randn(p.shape, generator=g) * math.sqrt(1.0 / p.shape[1]))
            elif name.endswith("weight") and p.numel() > 1 and ("bn" in name or "norm" in name or ".0.weight" in name):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif name.endswith("weight"):             # PReLU slopes, other 1-D weights
                p.copy_(0.25 + 0.05 * torch.randn(p.shape, generator=g))
            else:                                     # biases
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return module


def disable_dropout_(module):
    """torch's Philox stream cannot be reproduced by another implementation: parity runs use p = 0 everywhere."""
    for m in module.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d, torch.nn.Dropout2d)):
            m.p = 0.0
        if hasattr(m, "drop_rate"):
            m.drop_rate = 0
    return module


MODEL_CASES = {
    # name: (module path, class, kwargs, input size, batch)
    "vnet_elu": ("models.three_d.vnet3d", "VNet", dict(elu=True, in_channels=1, classes=2), 32, 2),
    "vnet_prelu": ("models.three_d.vnet3d", "VNet", dict(elu=False, in_channels=1, classes=2), 32, 2),
    "resunet8": ("models.three_d.residual_unet3d", "UNet", dict(in_channels=1, n_classes=2, base_n_filter=8), 64, 1),
    "highres3d": ("models.three_d.highresnet", "HighRes3DNet", dict(in_channels=1, out_channels=2), 24, 1),
    "densevoxel": ("models.three_d.densevoxelnet3d", "DenseVoxelNet", dict(in_channels=1, classes=2), 32, 2),
    "csrnet": ("models.three_d.csrnet", "CSRNet", dict(in_channels=1, out_channels=2, init_features=8), 32, 2),
    "re_net": ("models.three_d.RE_net", "RE_Net", dict(), 32, 2),
    "er_net": ("models.three_d.ER_net", "ER_Net", dict(classes=2, channels=1), 32, 2),
    "double_unet": ("models.three_d.Double_Unet", "Double_Unet",
                    dict(in_channels=1, out_channels=2, unet_init_features=16), 32, 2),
}


def case_inputs(name, size, batch):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(batch, 1, size, size, size, generator=g)
    lab = (torch.rand(batch, size, size, size, generator=g) > 0.7).long()
    return x, lab
