"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the reference's modules from /root/reference, runs them on seeded inputs on the CPU (fp32) and stores inputs
and outputs as small .npz fixtures next to this script.  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
# utils/metric.py imports torchio and monai at module scope; only HD95 needs them (metric.py:29-32).
for name in ("torchio", "monai", "monai.metrics"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["monai.metrics"].compute_hausdorff_distance = None

from models.three_d.unet3d import UNet3D  # noqa: E402
from models.sync_batchnorm.batchnorm import SynchronizedBatchNorm3d  # noqa: E402
from utils.loss_function import (BinaryDiceLoss, DiceLoss, DiceLossss, cross_entropy_3D)  # noqa: E402
from utils.metric import metric  # noqa: E402

torch.set_num_threads(8)


def npd(d):
    return {k: v.detach().numpy() for k, v in d.items()}


def unet_case(name, features, size, batch):
    torch.manual_seed(0)
    net = UNet3D(in_channels=1, out_channels=2, init_features=features)
    # weights_init_normal('kaiming') from train.py:33-61, restated (train.py itself cannot be imported: hydra missing)
    for m in net.modules():
        cn = m.__class__.__name__
        if hasattr(m, "weight") and (cn.find("Conv") != -1 or cn.find("Linear") != -1):
            torch.nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            if m.bias is not None:
                torch.nn.init.constant_(m.bias.data, 0.0)
    # make BN affine / conv bias non-trivial so their gradients are exercised
    g = torch.Generator().manual_seed(1)
    for k, p in net.named_parameters():
        if "norm" in k or k.endswith("bias"):
            p.data += 0.1 * torch.randn(p.shape, generator=g)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(batch, 1, size, size, size, generator=g)
    lab = (torch.rand(batch, size, size, size, generator=g) > 0.7).long()
    net.train()
    out = net(x)
    loss = cross_entropy_3D(out, lab) + DiceLossss(2)(out, lab, softmax=True)
    loss.backward()
    grads = {"grad." + k: p.grad for k, p in net.named_parameters()}
    sd1 = net.state_dict()
    net.eval()
    with torch.no_grad():
        out_eval = net(x)
    fix = {"x": x.numpy(), "lab": lab.numpy(), "out_train": out.detach().numpy(), "loss": loss.detach().numpy(),
           "out_eval": out_eval.numpy(), "argmax_eval": out_eval.argmax(1, keepdim=True).numpy()}
    fix.update({"sd0." + k: v.numpy() for k, v in sd0.items()})
    fix.update(npd(grads))
    fix.update({"sd1." + k: v.numpy() for k, v in sd1.items() if "running" in k})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fix)
    print(name, "loss", float(loss), "params", sum(p.numel() for p in net.parameters()))


def loss_case():
    torch.manual_seed(0)
    pred = torch.randn(2, 2, 8, 8, 8, requires_grad=True)
    lab = (torch.rand(2, 8, 8, 8) > 0.7).long()
    onehot = torch.stack([(lab == 0), (lab == 1)], 1).float()
    fix = {"pred": pred.detach().numpy(), "lab": lab.numpy()}

    def rec(name, fn):
        pred.grad = None
        v = fn()
        v.backward()
        fix[name] = v.detach().numpy()
        fix[name + ".grad"] = pred.grad.numpy().copy()

    rec("cross_entropy_3D", lambda: cross_entropy_3D(pred, lab))
    rec("DiceLoss", lambda: DiceLoss()(pred, onehot))
    rec("DiceLossss_softmax", lambda: DiceLossss(2)(pred, lab, softmax=True))
    rec("DiceLossss_raw", lambda: DiceLossss(2)(pred, lab, softmax=False))
    rec("BinaryDiceLoss", lambda: BinaryDiceLoss()(torch.sigmoid(pred[:, 1]), onehot[:, 1]))
    rec("BCEWithLogits", lambda: torch.nn.BCEWithLogitsLoss()(pred, onehot))
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **fix)
    # round 2: the weighted / non-default variants of the same functions (loss_function.py:8-16, 61-99, 172-184),
    # in a separate file so that losses.npz stays byte-identical
    fix = {"pred": pred.detach().numpy(), "lab": lab.numpy(), "ce_weight": np.float32([0.3, 1.7]),
           "dice_weight": np.float32([0.4, 1.6])}
    probs = lambda: torch.softmax(pred, dim=1)   # noqa: E731
    rec("cross_entropy_3D_weighted", lambda: cross_entropy_3D(pred, lab, weight=torch.tensor([0.3, 1.7])))
    rec("cross_entropy_3D_sum", lambda: cross_entropy_3D(pred, lab, size_average=False))
    rec("DiceLossss_softmax_weighted", lambda: DiceLossss(2)(pred, lab, weight=[0.4, 1.6], softmax=True))
    rec("DiceLossss_raw_weighted", lambda: DiceLossss(2)(pred, lab, weight=[0.4, 1.6], softmax=False))
    rec("DiceLossss_raw_on_probs", lambda: DiceLossss(2)(probs(), lab, softmax=False))
    rec("BinaryDiceLoss_sum", lambda: BinaryDiceLoss(reduction="sum")(torch.sigmoid(pred[:, 1]), onehot[:, 1]))
    rec("BinaryDiceLoss_p1_smooth", lambda: BinaryDiceLoss(smooth=0.5, p=1)(torch.sigmoid(pred), onehot))
    rec("BinaryDiceLoss_p3", lambda: BinaryDiceLoss(p=3)(torch.sigmoid(pred), onehot))
    none = BinaryDiceLoss(reduction="none")(torch.sigmoid(pred[:, 1]), onehot[:, 1])
    fix["BinaryDiceLoss_none"] = none.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "losses_extra.npz"), **fix)
    print("losses", {k: float(v) for k, v in fix.items() if v.ndim == 0})


def metric_case():
    g = torch.Generator().manual_seed(0)
    fix = {}
    for i, shape in enumerate([(2, 1, 16, 16, 16), (1, 1, 8, 12, 20)]):
        gt = (torch.rand(shape, generator=g) > 0.7).float()
        pred = (torch.rand(shape, generator=g) > 0.6).long()
        j, d = metric(gt, pred)
        fix["gt%d" % i], fix["pred%d" % i] = gt.numpy(), pred.numpy()
        fix["jaccard%d" % i], fix["dice%d" % i] = np.float64(j), np.float64(d)
    z = torch.zeros(1, 1, 4, 4, 4)
    j, d = metric(z, z.long())
    fix["jaccard_empty"], fix["dice_empty"] = np.float64(j), np.float64(d)
    np.savez_compressed(os.path.join(OUT, "metric.npz"), **fix)
    print("metric", fix["jaccard0"], fix["dice0"])


def syncbn_case():
    torch.manual_seed(0)
    bn = SynchronizedBatchNorm3d(6)
    bn.weight.data = torch.randn(6)
    bn.bias.data = torch.randn(6)
    xs = [torch.randn(2, 6, 4, 5, 6) * 2 + 1, torch.randn(3, 6, 4, 5, 6) - 0.5]
    # master-side statistics exactly as _data_parallel_master would see them (batchnorm.py:56-62, 113-125)
    sums = [x.view(x.size(0), 6, -1).sum(0).sum(-1) for x in xs]
    ssums = [(x.view(x.size(0), 6, -1) ** 2).sum(0).sum(-1) for x in xs]
    size = sum(x.size(0) * x[0, 0].numel() for x in xs)
    mean, inv_std = bn._compute_mean_std(sum(sums), sum(ssums), size)
    outs = [((x.view(x.size(0), 6, -1) - mean.view(1, -1, 1)) * (inv_std * bn.weight).view(1, -1, 1)
             + bn.bias.view(1, -1, 1)).view(x.shape) for x in xs]  # forward :71-78
    fix = {"x0": xs[0].numpy(), "x1": xs[1].numpy(), "weight": bn.weight.detach().numpy(),
           "bias": bn.bias.detach().numpy(), "mean": mean.numpy(), "inv_std": inv_std.numpy(),
           "running_mean": bn.running_mean.numpy(), "running_var": bn.running_var.numpy(),
           "out0": outs[0].detach().numpy(), "out1": outs[1].detach().numpy()}
    # non-parallel path = F.batch_norm (:50-53)
    bn2 = SynchronizedBatchNorm3d(6)
    y = bn2(xs[0])
    fix["single_out"] = y.detach().numpy()
    fix["single_running_var"] = bn2.running_var.numpy()
    np.savez_compressed(os.path.join(OUT, "syncbn.npz"), **fix)
    print("syncbn ok")


def pool_case():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 8, 6, 10, generator=g)
    x = torch.relu(x).bfloat16().float()  # all-zero windows -> ties; bf16-exact so the bf16 kernels see the same values
    x[0, 0, 0, 0, 1] = float("nan")
    x[1, 2, 3, 2, 4] = float("nan")
    y, idx = torch.nn.functional.max_pool3d(x, 2, 2, return_indices=True)
    lg = torch.randn(2, 3, 4, 4, 4, generator=g).round()  # ties in argmax
    np.savez_compressed(os.path.join(OUT, "pool_argmax.npz"), x=x.numpy(), y=y.numpy(), idx=idx.numpy(),
                        logits=lg.numpy(), argmax=lg.argmax(1, keepdim=True).numpy())
    print("pool ok")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "losses":
    loss_case()
    sys.exit(0)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pool":
        pool_case()
        sys.exit(0)
    unet_case("unet_f4_s32_b2", 4, 32, 2)
    loss_case()
    metric_case()
    syncbn_case()
    pool_case()
