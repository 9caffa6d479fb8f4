"""GPU: end-to-end parity of the U-Net training path with criteria that can fail.

Oracles: (1) the CPU restatement `oracle/unet3d.py` in fp32 and with bf16 STORAGE emulated at the points where the CUDA
path stores bf16; (2) the reference's own modules (vendored unmodified into oracle/_ref by oracle/build_ref.py) run on
the GPU in strict fp32 (`allow_tf32=False`), the second oracle SURVEY section 8c sanctions -- the oracle port on the GPU
when oracle/_ref is absent.

What bf16 does to a random-init U-Net gradient was measured with the oracle (tests/golden/README-floor in DESIGN.md
section 4): rounding only the stored GRADIENTS to bf16 moves the parameter gradients by 0.7 % (median), rounding only the
forward activations by 37 %, only the conv weights by 26 % -- at random init the parameter gradients are sums without a
coherent component, so the ~0.3 % of ReLU / max-pool decisions that a 2^-9 forward perturbation flips enter as
sqrt(fraction).  A bf16 implementation can therefore NOT match the fp32 per-parameter gradient to 10 % on such a case;
it can and must (a) agree in direction with the fp32 gradient (cosine of the full flattened gradient), (b) match, per
parameter, an fp32 oracle that rounds at the same storage points, and (c) train like the fp32 reference (loss curve)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import losses as olosses
from oracle import metric as ometric
from oracle import unet3d as ounet

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


def shadowed_bias(k):
    """Conv biases directly in front of a BatchNorm (unet3d.py:80-100): analytically zero gradient."""
    return k.endswith(".bias") and "conv" in k.split(".")[-2] and k.count(".") == 2


def structured_batch(batch, size, seed, device="cpu"):
    """Inputs with learnable structure: the label is a threshold of the locally averaged input."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 1, size, size, size, generator=g)
    sm = torch.nn.functional.avg_pool3d(x, 5, 1, 2)
    lab = (sm[:, 0] > 0.8 * sm.std()).long()
    return x.to(device), lab.to(device)


def oracle_grads(sd, x, lab, storage):
    s = {k: v.clone().float().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    out = ounet.forward(s, x, training=True, storage=storage)
    loss = olosses.dice_ce(out, lab)
    loss.backward()
    return out.detach(), float(loss), {k: v.grad for k, v in s.items() if v.requires_grad}


def test_unet_gradient_direction_and_bf16_storage_oracle():
    """UNet3D(1,2,32) on 2 x 1 x 64^3 (bottleneck 4^3), Dice+CE, random init."""
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.utils.loss_function import DiceCELoss
    torch.set_num_threads(os.cpu_count() or 1)
    sd = ounet.init_state_dict(1, 2, 32, seed=0)
    x, lab = structured_batch(2, 64, seed=1)
    net = UNet3D(1, 2, 32).to(DEV)
    net.load_state_dict(sd)
    net.train()
    out = net(x.to(DEV))
    loss = DiceCELoss(2)(out, lab.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    mine = {k: p.grad.detach().float().cpu() for k, p in net.named_parameters()}
    o32, l32, g32 = oracle_grads(sd, x, lab, "fp32")
    o16, l16, g16 = oracle_grads(sd, x, lab, "bf16")
    # forward: logits and loss against the fp32 oracle
    assert rel(out.detach().cpu(), o32) < 6e-2, rel(out.detach().cpu(), o32)   # 40 bf16-stored layers deep
    assert abs(loss.item() - l32) < 5e-3 * max(1.0, abs(l32)), (loss.item(), l32)
    live = [k for k in g32 if float(g32[k].abs().max()) > 1e-7 and not shadowed_bias(k)]
    flat = lambda d: torch.cat([d[k].flatten() for k in live])   # noqa: E731
    cos32, cos16 = cosine(flat(mine), flat(g32)), cosine(flat(mine), flat(g16))
    floor_cos = cosine(flat(g16), flat(g32))
    errs16 = {k: rel(mine[k], g16[k]) for k in live}
    errs32 = {k: rel(mine[k], g32[k]) for k in live}
    floor = {k: rel(g16[k], g32[k]) for k in live}
    worst = sorted(errs16.items(), key=lambda kv: -kv[1])[:5]
    print("\ncos(ours, fp32) %.5f  cos(ours, bf16-storage oracle) %.5f  cos(bf16 oracle, fp32) %.5f" % (cos32, cos16, floor_cos))
    print("per-parameter rel err vs bf16-storage oracle: median %.4f worst %s" % (float(np.median(list(errs16.values()))), worst))
    print("per-parameter rel err vs fp32 oracle: median %.4f ; bf16-storage floor median %.4f" %
          (float(np.median(list(errs32.values()))), float(np.median(list(floor.values())))))
    # (a) direction: the full flattened gradient points where the fp32 gradient points
    assert cos32 >= 0.99, cos32
    # (b) per parameter: the computation is chaotic in the forward decisions (measured on the B200: even against the oracle
    #     that rounds where we do, the encoder parameters differ by 0.30 median -- the oracle's own fp32-vs-bf16 distance is
    #     0.38), so the bound is the oracle's own bf16-storage distance per parameter, with 10 % absolute where that is
    #     small (the decoder's last layers and the head: 0.03), never more than 1.25 x
    assert cos16 >= 0.99, cos16
    for k in live:
        assert errs16[k] <= max(0.1, 1.25 * floor[k]), (k, errs16[k], floor[k])
        assert errs32[k] <= max(0.1, 1.25 * floor[k]), (k, errs32[k], floor[k])
    assert float(np.median(list(errs16.values()))) <= float(np.median(list(floor.values()))), "not closer to the bf16 oracle than fp32 is"
    for k in ("conv.weight", "conv.bias", "decoder1.dec1conv2.weight", "decoder1.dec1norm2.weight", "decoder1.dec1norm2.bias"):
        assert errs32[k] < 0.06, (k, errs32[k])
    # biases in front of a BatchNorm: analytically zero gradient
    for k, p in net.named_parameters():
        if shadowed_bias(k):
            assert p.grad is None or float(p.grad.abs().max()) < 1e-4, k


def _reference_unet(features, sd):
    """fp32 reference network on the GPU: the reference's own UNet3D when oracle/_ref travelled here, else None."""
    if not build_ref.import_ref():
        return None
    from models.three_d.unet3d import UNet3D as RefUNet3D
    net = RefUNet3D(in_channels=1, out_channels=2, init_features=features)
    net.load_state_dict(sd)
    return net.to(DEV).train()


class strict_fp32:
    def __enter__(self):
        self.old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.old


def _ref_loss(out, lab):
    if build_ref.import_ref():
        from utils.loss_function import DiceLossss, cross_entropy_3D
        return cross_entropy_3D(out, lab) + DiceLossss(2)(out, lab, softmax=True)
    return olosses.dice_ce(out, lab)


def test_twenty_optimizer_steps_follow_the_fp32_reference():
    """20 Adam steps from the same weights on the same batches, ours (bf16 kernels, CUDA-graph step) against the
    fp32 reference: loss curve within 2 % (mean relative deviation; 5 % at any single step), final Dice of the predicted
    masks within 4e-2, and Dice computed by our metric kernel == the oracle's metric on equal masks within 1e-4."""
    from b200seg.engine import TrainStep
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.optim import FusedAdam
    from b200seg.utils.loss_function import DiceCELoss
    from b200seg.utils.metric import metric
    import b200seg.functional as F
    feats, size, steps = 16, 64, 20
    sd = ounet.init_state_dict(1, 2, feats, seed=2)
    batches = [structured_batch(2, size, seed=100 + i % 4, device=DEV) for i in range(steps)]
    # ours
    net = UNet3D(1, 2, feats).to(DEV)
    net.load_state_dict(sd)
    net.train()
    opt = FusedAdam(net.parameters(), lr=1e-3)
    step = TrainStep(net, DiceCELoss(2), opt, use_graph=True, warmup=2)
    mine = [float(step(x, y.to(torch.uint8))[0]) for x, y in batches]
    # reference, strict fp32 on the same GPU
    with strict_fp32():
        ref = _reference_unet(feats, sd)
        if ref is None:
            s = {k: v.clone().to(DEV).requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
            params = [v for v in s.values() if v.requires_grad]
        else:
            params = list(ref.parameters())
        topt = torch.optim.Adam(params, lr=1e-3)
        theirs = []
        for x, y in batches:
            topt.zero_grad(set_to_none=True)
            if ref is None:
                stats = {}
                out = ounet.forward(s, x, training=True, new_stats=stats)
            else:
                out = ref(x)
            loss = _ref_loss(out, y)
            loss.backward()
            topt.step()
            if ref is None:
                for k, v in stats.items():
                    s[k] = v
            theirs.append(float(loss))
        # final masks (eval mode) on a held-out batch
        xe, ye = structured_batch(2, size, seed=999, device=DEV)
        with torch.no_grad():
            if ref is None:
                ref_logits = ounet.forward({k: v.detach() for k, v in s.items()}, xe, training=False)
            else:
                ref_logits = ref.eval()(xe)
    net.eval()
    with torch.no_grad():
        my_logits = net(xe)
    dev = [abs(a - b) / abs(b) for a, b in zip(mine, theirs)]
    print("\nloss ours   ", ["%.4f" % v for v in mine])
    print("loss fp32   ", ["%.4f" % v for v in theirs])
    print("mean rel dev %.4f max %.4f" % (float(np.mean(dev)), max(dev)))
    assert theirs[-1] < 0.8 * theirs[0], "the reference did not learn on this task"
    assert float(np.mean(dev)) < 0.02 and max(dev) < 0.05, (float(np.mean(dev)), max(dev))
    my_mask, ref_mask = F.argmax_labels(my_logits), ref_logits.argmax(1, keepdim=True)
    gt = ye.reshape(ye.shape[0], 1, *ye.shape[1:])
    j1, d1 = metric(gt, my_mask)
    j2, d2 = metric(gt, ref_mask)
    print("final Dice ours %.4f reference %.4f ; mask agreement %.4f" % (d1, d2, float((my_mask == ref_mask).float().mean())))
    # (eval mode after only 20 steps: running statistics at momentum 0.1 are far from settled and atomics-ordered sums
    # differ from run to run, so the two masks agree to a few per cent of Dice, not better: 0.015-0.021 observed)
    assert abs(d1 - d2) < 4e-2, (d1, d2)
    # Dice on EQUAL masks: our counts kernel vs the oracle's restatement of metric.py, 1e-4 (it is bit-exact)
    oc = ometric.counts(gt.cpu().numpy(), my_mask.cpu().numpy())
    o_dice = 2 * oc["intersection"] / (oc["gt_sum"] + oc["pred_sum"] + 0.001)
    assert abs(d1 - o_dice) < 1e-4


def test_full_size_config2_step_against_fp32_reference():
    """BASELINE configs[1] at FULL size -- UNet3D(1,2,32), batch 2 x 1 x 128^3, Dice+CE, one forward + backward -- against the
    reference's modules in strict fp32 on the same GPU (oracle port when oracle/_ref is absent)."""
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.utils.loss_function import DiceCELoss
    sd = ounet.init_state_dict(1, 2, 32, seed=0)
    x, lab = structured_batch(2, 128, seed=7, device=DEV)
    net = UNet3D(1, 2, 32).to(DEV)
    net.load_state_dict(sd)
    net.train()
    out = net(x)
    loss = DiceCELoss(2)(out, lab)
    loss.backward()
    torch.cuda.synchronize()
    mine = {k: p.grad.detach().float() for k, p in net.named_parameters()}
    my_stats = {k: v.detach().clone() for k, v in net.state_dict().items() if "running" in k}
    my_out, my_loss = out.detach().clone(), loss.item()
    del net, out, loss
    torch.cuda.empty_cache()
    with strict_fp32():
        ref = _reference_unet(32, sd)
        if ref is not None:
            rout = ref(x)
            rloss = _ref_loss(rout, lab)
            rloss.backward()
            theirs = {k: p.grad.detach() for k, p in ref.named_parameters()}
            ref_stats = {k: v for k, v in ref.state_dict().items() if "running" in k}
        else:
            s = {k: v.clone().to(DEV).requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
            ref_stats = {}
            rout = ounet.forward(s, x, training=True, new_stats=ref_stats)
            rloss = olosses.dice_ce(rout, lab)
            rloss.backward()
            theirs = {k: v.grad for k, v in s.items() if v.requires_grad}
    e_out = rel(my_out, rout.detach())
    live = [k for k in theirs if float(theirs[k].abs().max()) > 1e-7 and not shadowed_bias(k)]
    cos = cosine(torch.cat([mine[k].flatten() for k in live]), torch.cat([theirs[k].flatten() for k in live]))
    errs = {k: rel(mine[k], theirs[k]) for k in live}
    print("\nfull size: logits rel-fro %.4f  loss %.5f vs %.5f  gradient cosine %.5f  per-parameter median %.3f" %
          (e_out, my_loss, float(rloss), cos, float(np.median(list(errs.values())))))
    assert e_out < 6e-2, e_out
    assert abs(my_loss - float(rloss)) < 5e-3 * max(1.0, abs(float(rloss)))
    assert cos >= 0.99, cos
    # the decoder's last layers and the head see the least accumulated bf16 perturbation: tight per-parameter bounds there
    for k in ("conv.weight", "conv.bias", "decoder1.dec1conv2.weight", "decoder1.dec1norm2.weight", "decoder1.dec1norm2.bias"):
        assert errs[k] < 0.05, (k, errs[k])
    for k, v in ref_stats.items():
        if "running" in k and "num_batches" not in k:
            assert rel(my_stats[k].float(), v.float()) < 1e-2, k


@pytest.mark.parametrize("name", ["vnet", "er_net"])
def test_full_size_other_networks_against_fp32_reference(name):
    """BASELINE configs[2] (V-Net) and an extra config.network (ER-Net) at FULL size -- batch 2 x 1 x 128^3, Dice+CE, one
    forward + backward -- against the reference's own modules in strict fp32 on the same GPU.  Exercises at full extent what
    the small parity cases cannot: the kd-stacked 5x5x5 kernel over 128 planes, the zero-padded stem / output layers, the
    reverse-attention and selective-fusion gates."""
    if not build_ref.import_ref():
        pytest.skip("oracle/_ref (vendored reference modules) is not on this box")
    import importlib
    from b200seg.utils.loss_function import DiceCELoss
    from oracle.model_init import disable_dropout_, init_module_
    mod, cls, args = {"vnet": ("vnet3d", "VNet", (True, 1, 2)), "er_net": ("ER_net", "ER_Net", (2, 1))}[name]
    x, lab = structured_batch(2, 128, seed=11, device=DEV)
    mine_net = getattr(importlib.import_module("b200seg.models.three_d." + mod), cls)(*args)
    disable_dropout_(init_module_(mine_net, seed=5))
    sd = {k: v.clone() for k, v in mine_net.state_dict().items()}
    mine_net = mine_net.to(DEV).train()
    out = mine_net(x)
    loss = DiceCELoss(2)(out, lab)
    loss.backward()
    torch.cuda.synchronize()
    mine = {k: p.grad.detach().float() for k, p in mine_net.named_parameters() if p.grad is not None}
    my_out, my_loss = out.detach().clone(), loss.item()
    del mine_net, out, loss
    torch.cuda.empty_cache()
    with strict_fp32():
        ref = getattr(importlib.import_module("models.three_d." + mod), cls)(*args)
        ref.load_state_dict(sd)
        disable_dropout_(ref)
        ref = ref.to(DEV).train()
        rout = ref(x)
        rloss = _ref_loss(rout, lab)
        rloss.backward()
        theirs = {k: p.grad.detach() for k, p in ref.named_parameters() if p.grad is not None}
    e_out = rel(my_out, rout.detach())
    # parameters whose gradient is analytically zero (a conv bias in front of batch statistics) carry only round-off
    live = [k for k in theirs if k in mine and float(theirs[k].norm()) > 1e-4 * float(max(t.norm() for t in theirs.values()))]
    cos = cosine(torch.cat([mine[k].flatten() for k in live]), torch.cat([theirs[k].flatten() for k in live]))
    errs = {k: rel(mine[k], theirs[k]) for k in live}
    print("\n%s full size: logits rel-fro %.4f  loss %.5f vs %.5f  gradient cosine %.5f  per-parameter median %.3f" %
          (name, e_out, my_loss, float(rloss), cos, float(np.median(list(errs.values())))))
    assert e_out < 6e-2, e_out
    assert abs(my_loss - float(rloss)) < 5e-3 * max(1.0, abs(float(rloss)))
    assert cos >= 0.99, cos
    assert float(np.median(list(errs.values()))) < 0.1
