"""Golden vectors for the V-Net / residual U-Net / HighRes3DNet / DenseVoxelNet / CSRNet / RE-Net / ER-Net / Double-UNet mirrors, produced by the UNMODIFIED
reference modules (run in the build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_models.py

Weights come from oracle.model_init.init_module_ (seeded, visited in state_dict order), dropout is disabled (p = 0),
inputs are seeded; stored are the train-mode logits, the Dice+CE loss (reference loss functions), per-parameter gradient
norms, a few full gradients and the updated BatchNorm running statistics, plus the eval-mode logits.
"""
import importlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import types  # noqa: E402

for _name in ("thop", "torchvision"):      # imported (not used) by Double_Unet.py; absent from this image
    try:
        importlib.import_module(_name)
    except Exception:
        sys.modules[_name] = types.ModuleType(_name)
        sys.modules[_name].profile = None

from oracle.model_init import MODEL_CASES, case_inputs, disable_dropout_, init_module_  # noqa: E402
from utils.loss_function import DiceLossss, cross_entropy_3D  # noqa: E402  (reference)

torch.set_num_threads(8)


def run(name):
    mod, cls, kw, size, batch = MODEL_CASES[name]
    net = getattr(importlib.import_module(mod), cls)(**kw)
    init_module_(net, seed=11)
    disable_dropout_(net)
    x, lab = case_inputs(name, size, batch)
    net.train()
    out = net(x)
    loss = cross_entropy_3D(out, lab) + DiceLossss(2)(out, lab, softmax=True)
    loss.backward()
    fix = {"out_train": out.detach().numpy(), "loss": loss.detach().numpy(),
           "keys": np.array(list(net.state_dict().keys())),
           "shapes": np.array([str(tuple(v.shape)) for v in net.state_dict().values()])}
    names, norms = [], []
    for k, p in net.named_parameters():
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.norm()))
    fix["grad_names"], fix["grad_norms"] = np.array(names), np.array(norms, np.float64)
    small = [k for k, p in net.named_parameters() if p.grad is not None and p.numel() <= 20000]
    for k in small[:3] + small[-3:]:
        fix["grad." + k] = dict(net.named_parameters())[k].grad.numpy()
    running = [k for k in net.state_dict() if k.endswith("running_var")]
    for k in running[:2] + running[-2:]:
        fix["sd1." + k] = net.state_dict()[k].numpy()
    net.eval()
    with torch.no_grad():
        fix["out_eval"] = net(x).numpy()
    np.savez_compressed(os.path.join(OUT, "model_%s.npz" % name), **fix)
    print(name, "loss %.6f" % float(loss), "params", sum(p.numel() for p in net.parameters()), "out", tuple(out.shape))


if __name__ == "__main__":
    for n in (sys.argv[1:] or MODEL_CASES):
        run(n)
