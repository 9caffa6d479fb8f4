"""Per-kernel-family time of one eager training step of an in-scope model (CUDA events around every ABI call).
usage: profile_model.py [vnet res_unet highres densevoxel unet]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
from b200seg.engine import TrainStep
from b200seg.optim import FusedAdam
from b200seg.utils.loss_function import DiceCELoss
from bench_models import CASES

dev = torch.device("cuda")
CASES = dict(CASES)
CASES["unet"] = (lambda: __import__("b200seg.models.three_d.unet3d", fromlist=["UNet3D"]).UNet3D(1, 2, 32), 128, 2, 951.3)
CASES["er_net"] = (lambda: __import__("b200seg.models.three_d.ER_net", fromlist=["ER_Net"]).ER_Net(2, 1), 128, 2, 0.0)
CASES["dunet"] = (lambda: __import__("b200seg.models.three_d.Double_Unet", fromlist=["Double_Unet"]).Double_Unet(1, 2), 96, 2, 0.0)
CASES["re_net"] = (lambda: __import__("b200seg.models.three_d.RE_net", fromlist=["RE_Net"]).RE_Net(), 128, 2, 0.0)
CASES["csrnet"] = (lambda: __import__("b200seg.models.three_d.csrnet", fromlist=["CSRNet"]).CSRNet(1, 2, 32), 128, 2, 0.0)

_orig_call = F._call


def _named_call(name, *args, work=0.0, tag=None):
    # tag every launch with its ABI name (+ conv geometry) so the table says where the time goes
    if tag is None:
        tag = name.replace("b200seg_", "")
    elif name.startswith("b200seg_conv3d"):
        g = args[0]._obj
        tag = "%s ci%d co%d k%d s%d d%d %d^3" % (tag, g.cin, g.cout, g.k, g.stride, g.dil, g.oh)
    return _orig_call(name, *args, work=work, tag=tag)


F._call = _named_call

for name in (sys.argv[1:] or ["highres"]):
    make, size, batch, gflop = CASES[name]
    torch.manual_seed(0)
    net = make().to(dev).train()
    opt = FusedAdam(net.parameters(), lr=1e-3)
    step = TrainStep(net, DiceCELoss(2), opt, use_graph=False)
    x = torch.randn(batch, 1, size, size, size, device=dev)
    lab = (torch.rand(batch, size, size, size, device=dev) > 0.9).to(torch.uint8)
    for _ in range(2):
        step(x, lab)
    torch.cuda.synchronize()
    F.profile_begin()
    step(x, lab)
    prof = F.profile_end()
    rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    total = sum(v["ms"] for v in prof.values())
    print("== %s: %.2f ms of kernel time in %d launches" % (name, total, sum(v["launches"] for v in prof.values())))
    for k, v in rows[:int(os.environ.get("ROWS", "28"))]:
        tf = v["work"] / (v["ms"] * 1e-3) / 1e12 if v["work"] and v["ms"] else 0.0
        print("  %-58s %4d x  %9.3f ms  %5.1f %%  %s" % (k, v["launches"], v["ms"], 100 * v["ms"] / total,
                                                       ("%.0f TFLOP/s" % tf) if tf else ""))
    del net, opt, step
    torch.cuda.empty_cache()
