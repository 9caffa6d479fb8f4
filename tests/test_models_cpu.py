"""CPU: the model mirrors (V-Net, residual U-Net, HighRes3DNet, DenseVoxelNet) have the reference's state_dict, and
their forward graphs -- bound to the torch oracle backend -- reproduce the golden vectors produced by the reference's
own modules (tests/golden/make_golden_models.py): logits, Dice+CE loss, every parameter's gradient norm, running stats.
This pins the graph wiring; tests/test_gpu_models.py then checks the CUDA binding of the same graphs."""
import os

import numpy as np
import pytest
import torch

from oracle import backend_torch, losses as olosses
from oracle.model_init import MODEL_CASES, case_inputs, disable_dropout_, init_module_

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build(name):
    from b200seg.models.three_d.csrnet import CSRNet
    from b200seg.models.three_d.Double_Unet import Double_Unet
    from b200seg.models.three_d.ER_net import ER_Net
    from b200seg.models.three_d.RE_net import RE_Net
    from b200seg.models.three_d.densevoxelnet3d import DenseVoxelNet
    from b200seg.models.three_d.highresnet import HighRes3DNet
    from b200seg.models.three_d.residual_unet3d import UNet
    from b200seg.models.three_d.vnet3d import VNet
    cls = {"VNet": VNet, "UNet": UNet, "HighRes3DNet": HighRes3DNet, "DenseVoxelNet": DenseVoxelNet,
           "CSRNet": CSRNet, "RE_Net": RE_Net, "ER_Net": ER_Net, "Double_Unet": Double_Unet}[MODEL_CASES[name][1]]
    net = cls(**MODEL_CASES[name][2])
    init_module_(net, seed=11)
    disable_dropout_(net)
    return net


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_mirror_graph_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    net = build(name)
    sd = net.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g["shapes"])
    net.set_kernels(backend_torch)
    _, _, _, size, batch = MODEL_CASES[name]
    x, lab = case_inputs(name, size, batch)
    net.train()
    out = net(x)
    ref = torch.from_numpy(g["out_train"])
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-3, atol=2e-4), float((out - ref).abs().max())
    loss = olosses.dice_ce(out, lab)
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    params = dict(net.named_parameters())
    for k, want in zip(g["grad_names"], g["grad_norms"]):
        got = 0.0 if params[k].grad is None else float(params[k].grad.norm())
        assert abs(got - want) <= 1e-2 * max(want, 1e-3) + 1e-6, (k, got, want)   # fp32 round-off through ~20 BN layers
    for k in g.files:
        if k.startswith("grad."):
            a, b = params[k[5:]].grad, torch.from_numpy(g[k])
            # (a conv bias in front of batch statistics has an analytically zero gradient: compare absolutely)
            assert float((a - b).norm()) < 1e-2 * float(b.norm()) + 1e-6, k
        if k.startswith("sd1."):
            assert torch.allclose(net.state_dict()[k[4:]], torch.from_numpy(g[k]), rtol=1e-4, atol=1e-6), k
    net.eval()
    with torch.no_grad():
        out_eval = net(x)
    assert torch.allclose(out_eval, torch.from_numpy(g["out_eval"]), rtol=1e-3, atol=2e-4)


def test_constructor_contracts():
    from b200seg.models.three_d.highresnet import HighRes3DNet, HighResNet
    from b200seg.utils.convolution import ConvolutionalBlock
    from b200seg.utils.residual import ResidualBlock
    with pytest.raises(AssertionError):
        HighResNet(1, 2, dimensions=4)
    with pytest.raises(AssertionError):
        ConvolutionalBlock(4, 4, 1, 3, batch_norm=True, instance_norm=True)
    with pytest.raises(AssertionError):
        ResidualBlock(4, 4, 2, 1, 3, residual_type="concat")
    # reflect / replicate padding: the mirror's graph (pad op + 'valid' conv) equals torch's F.pad + conv on the CPU backend
    for mode in ("reflect", "replicate"):
        blk = ConvolutionalBlock(4, 6, 2, 3, padding_mode=mode, batch_norm=False).set_kernels(backend_torch)
        x = torch.randn(1, 4, 7, 8, 9)
        conv = blk.convolutional_block[-1]
        want = conv(torch.nn.functional.pad(torch.relu(x), 6 * [2], mode))
        assert torch.allclose(blk(x), want, atol=1e-5)
    net = HighRes3DNet(1, 2)
    assert int(net.receptive_field) == 87 and net.num_parameters == 803636
    with pytest.raises(ValueError):
        net(torch.zeros(1, 1, 8, 8))
