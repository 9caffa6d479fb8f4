"""GPU: the CUDA binding of the V-Net / residual U-Net / HighRes3DNet / DenseVoxelNet mirrors against the golden vectors
of the reference's modules (tests/golden/model_*.npz).

Tolerances are calibrated, not guessed: the CUDA path stores activations, their gradients and conv weights in bf16 with
fp32 arithmetic.  The torch oracle backend can emulate exactly that storage (oracle.backend_torch.bf16_storage); the
distance between its fp32 and bf16-storage runs is the error floor of ANY bf16-storage implementation of the graph.
The CUDA results must stay within 3x that floor (plus a small absolute slack) of the reference's golden values."""
import os

import numpy as np
import pytest
import torch

from oracle import backend_torch, losses as olosses
from oracle.model_init import MODEL_CASES, case_inputs
from test_models_cpu import GOLDEN, build

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def _measure(net, x, lab, g, loss_fn, dev):
    """Errors of one implementation against the golden vectors: logits, loss, gradient-norm vector, stored gradients."""
    net.train()
    out = net(x.to(dev))
    e_out = rel(out.detach().float().cpu(), torch.from_numpy(g["out_train"]))
    loss = loss_fn(out, lab.to(dev))
    loss.backward()
    e_loss = abs(float(loss.detach()) - float(g["loss"]))
    params = dict(net.named_parameters())
    sq_err = sq = 0.0
    for k, want in zip(g["grad_names"], g["grad_norms"]):
        got = 0.0 if params[k].grad is None else float(params[k].grad.float().norm())
        sq_err += (got - want) ** 2
        sq += want ** 2
    e_gn = (sq_err / sq) ** 0.5
    e_full = 0.0
    for k in g.files:
        if k.startswith("grad.") and float(np.linalg.norm(g[k])) > 1e-5:
            e_full = max(e_full, rel(params[k[5:]].grad.float().cpu(), torch.from_numpy(g[k])))
    net.eval()
    with torch.no_grad():
        e_eval = rel(net(x.to(dev)).float().cpu(), torch.from_numpy(g["out_eval"]))
    return dict(logits=e_out, eval_logits=e_eval, loss=e_loss, grad_norms=e_gn, grads=e_full)


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_cuda_models_match_reference_golden(name):
    from b200seg.utils.loss_function import DiceCELoss
    g = np.load(os.path.join(GOLDEN, "model_%s.npz" % name))
    _, _, _, size, batch = MODEL_CASES[name]
    x, lab = case_inputs(name, size, batch)
    # error floor of bf16 storage, from the oracle itself (CPU)
    ref_net = build(name).set_kernels(backend_torch)
    with backend_torch.bf16_storage():
        floor = _measure(ref_net, x, lab, g, olosses.dice_ce, torch.device("cpu"))
    cuda = _measure(build(name).to("cuda"), x, lab, g, DiceCELoss(2), torch.device("cuda"))
    torch.cuda.synchronize()
    print("\n%s\n  bf16 floor: %s\n  cuda      : %s" % (name, {k: round(v, 5) for k, v in floor.items()},
                                                         {k: round(v, 5) for k, v in cuda.items()}))
    for k in cuda:
        assert cuda[k] <= 3.0 * floor[k] + 2e-3, (k, cuda[k], floor[k])


def test_dropout_kernels_statistics_and_backward():
    import b200seg.functional as F
    dev = torch.device("cuda")
    x = torch.ones(2, 8, 16, 16, 32, device=dev, dtype=torch.bfloat16, requires_grad=True)
    for channel, p in ((False, 0.2), (True, 0.5)):
        y = F.dropout(x, p, training=True, channel=channel)
        keep = (y != 0).float().mean().item()
        assert abs(keep - (1 - p)) < (0.25 if channel else 0.02)
        vals = torch.unique(y.float())
        assert set(vals.tolist()) <= {0.0, float(torch.tensor(1 / (1 - p)).bfloat16())}
        if channel:   # whole (sample, channel) volumes are dropped together
            per = (y != 0).float().mean(dim=(1, 2, 3))
            assert set(torch.unique(per).tolist()) <= {0.0, 1.0}
        (gx,) = torch.autograd.grad(y.float().sum(), x)
        assert torch.equal(gx != 0, y != 0)       # backward regenerates the same mask
    assert F.dropout(x, 0.3, training=False) is x


def test_double_dropout_in_one_pass_and_widened_gradient():
    """_DenseLayer's two dropout calls (densevoxelnet3d.py:25-32) as one kernel: two independent masks, the same masks in
    the backward, and a 12-channel gradient that arrives at the padded tensor-core convolution without a pad pass."""
    import b200seg.functional as F
    dev = torch.device("cuda")
    p = 0.2
    x = torch.ones(2, 8, 16, 16, 12, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = F.dropout(x, p, training=True, times=2)
    keep = (y != 0).float().mean().item()
    assert abs(keep - (1 - p) ** 2) < 0.02, keep
    s1 = float(torch.tensor(1 / (1 - p)).bfloat16())
    s2 = float((torch.tensor(s1).bfloat16().float() / (1 - p)).bfloat16())
    assert set(torch.unique(y.float()).tolist()) <= {0.0, s2}
    (gx,) = torch.autograd.grad(y.float().sum(), x)
    assert torch.equal(gx != 0, y != 0)
    # conv (C_out = 12, padded path) -> double dropout: gradients equal those of the two-step composition with the same masks
    torch.manual_seed(0)
    h = torch.randn(2, 48, 48, 48, 32, device=dev).bfloat16().requires_grad_()
    w = (torch.randn(12, 32, 3, 3, 3, device=dev) * 0.05).requires_grad_()
    from b200seg.functional import _geom, _use_padded
    assert _use_padded(_geom(h.shape, 32, 12, 3, 1, 1, 1))
    F._SALT[0] = 100
    out = F.dropout(F.conv_norm_act(h, w, None, k=3, stride=1, pad=1, dil=1), p, training=True, times=2)
    gy = torch.randn_like(out)
    F.profile_begin()
    gh, gw = torch.autograd.grad(out, (h, w), gy)
    prof = F.profile_end()
    assert prof["conv_dgrad_padded"]["launches"] == 1 and prof["conv_wgrad_padded"]["launches"] == 1
    assert "b200seg_pad_channels" not in prof, sorted(prof)      # the dropout backward already wrote the wide gradient
    mask = (out != 0)
    conv = F.conv_norm_act(h, w, None, k=3, stride=1, pad=1, dil=1)
    gh2, gw2 = torch.autograd.grad(conv, (h, w), (gy.float() * mask * s2).bfloat16())
    assert float((gh.float() - gh2.float()).norm() / gh2.float().norm()) < 1e-2
    assert float((gw - gw2).norm() / gw2.norm()) < 1e-2
