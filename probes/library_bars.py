"""Library bars: the UNMODIFIED reference modules (oracle/_ref, vendored by oracle/build_ref.py) under stock PyTorch / cuDNN
on the same B200, beside our step -- the yardstick VERDICT r01 asks for ("the cuDNN bar measured in the same probe").

For every in-scope model: fwd + Dice/CE + bwd + torch.optim.Adam on the same synthetic batch in
  (a) tf32   : fp32 NCDHW with allow_tf32 = True, how the reference itself would run on this GPU;
  (b) bf16_cl: torch.autocast(bfloat16) + channels_last_3d, cudnn.benchmark = True -- the strongest library configuration;
and our CUDA-graph train step (b200seg).  One JSON line per (model, configuration).
usage: library_bars.py [unet vnet res_unet highres densevoxel]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import build_ref

assert build_ref.import_ref(), "oracle/_ref is missing: run __graft_entry__.build() where /root/reference exists"
from utils.loss_function import DiceLossss, cross_entropy_3D   # noqa: E402  (reference)

dev = torch.device("cuda")


def ref_model(name):
    if name == "unet":
        from models.three_d.unet3d import UNet3D
        return UNet3D(1, 2, 32)
    if name == "vnet":
        from models.three_d.vnet3d import VNet
        return VNet(elu=True, in_channels=1, classes=2)
    if name == "res_unet":
        from models.three_d.residual_unet3d import UNet
        return UNet(1, 2, base_n_filter=32)
    if name == "highres":
        from models.three_d.highresnet import HighRes3DNet
        return HighRes3DNet(1, 2)
    if name == "densevoxel":
        from models.three_d.densevoxelnet3d import DenseVoxelNet
        return DenseVoxelNet(1, 2)
    if name == "csrnet":
        from models.three_d.csrnet import CSRNet
        return CSRNet(1, 2, 32)
    if name == "er_net":
        from models.three_d.ER_net import ER_Net
        return ER_Net(classes=2, channels=1)
    if name == "re_net":
        from models.three_d.RE_net import RE_Net
        return RE_Net()
    if name == "dunet":
        from models.three_d.Double_Unet import Double_Unet
        return Double_Unet(1, 2)
    raise KeyError(name)


def our_model(name):
    import importlib
    mod, cls, args = {"unet": ("unet3d", "UNet3D", (1, 2, 32)), "vnet": ("vnet3d", "VNet", (True, 1, 2)),
                      "res_unet": ("residual_unet3d", "UNet", (1, 2, 32)), "highres": ("highresnet", "HighRes3DNet", (1, 2)),
                      "densevoxel": ("densevoxelnet3d", "DenseVoxelNet", (1, 2)), "csrnet": ("csrnet", "CSRNet", (1, 2, 32)),
                      "er_net": ("ER_net", "ER_Net", (2, 1)), "re_net": ("RE_net", "RE_Net", ()),
                      "dunet": ("Double_Unet", "Double_Unet", (1, 2))}[name]
    return getattr(importlib.import_module("b200seg.models.three_d." + mod), cls)(*args)


SIZES = {"unet": 128, "vnet": 128, "res_unet": 128, "highres": 96, "densevoxel": 96, "csrnet": 128, "er_net": 128, "re_net": 128,
         "dunet": 96}
GFLOP_FWD = {"unet": 951.3, "vnet": 1463.1, "res_unet": 1820.7, "highres": 1419.7, "densevoxel": 144.7}   # others: not counted


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def library(name, mode, iters=5):
    size, batch = SIZES[name], 2
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    net = ref_model(name).to(dev).train()
    x = torch.randn(batch, 1, size, size, size, device=dev)
    lab = (torch.rand(batch, size, size, size, device=dev) > 0.9).long()
    if mode == "bf16_cl":
        net = net.to(memory_format=torch.channels_last_3d)
        x = x.contiguous(memory_format=torch.channels_last_3d)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    dice = DiceLossss(2)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_cl")):
            out = net(x)
        out = out.float()
        loss = cross_entropy_3D(out, lab) + dice(out, lab, softmax=True)
        loss.backward()
        opt.step()
    ms = timed(step, iters)
    mem = torch.cuda.max_memory_allocated() / 1e9
    del net, opt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return ms, mem


def ours(name, iters=10):
    from b200seg.engine import TrainStep
    from b200seg.optim import FusedAdam
    from b200seg.utils.loss_function import DiceCELoss
    size, batch = SIZES[name], 2
    torch.manual_seed(0)
    net = our_model(name).to(dev).train()
    opt = FusedAdam(net.parameters(), lr=1e-3)
    step = TrainStep(net, DiceCELoss(2), opt, use_graph=True)
    x = torch.randn(batch, 1, size, size, size, device=dev)
    lab = (torch.rand(batch, size, size, size, device=dev) > 0.9).to(torch.uint8)
    for _ in range(6):
        step(x, lab)
    ms = timed(lambda: step(x, lab), iters)
    mem = torch.cuda.max_memory_allocated() / 1e9
    del net, opt, step
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return ms, mem


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["unet", "vnet", "res_unet", "highres", "densevoxel"]):
        rec = {"model": name, "batch": 2, "patch": SIZES[name]}
        for mode in ("tf32", "bf16_cl"):
            try:
                ms, mem = library(name, mode)
                rec[mode] = {"ms_per_step": round(ms, 2), "patches_per_s": round(2e3 / ms, 2), "peak_mem_GB": round(mem, 1)}
            except Exception as e:   # out of memory etc.: report, keep going
                rec[mode] = {"error": repr(e)[:200]}
                torch.cuda.empty_cache()
        try:
            ms, mem = ours(name)
            rec["b200seg"] = {"ms_per_step": round(ms, 2), "patches_per_s": round(2e3 / ms, 2), "peak_mem_GB": round(mem, 1),
                              "train_tflops": round(3 * GFLOP_FWD[name] * 2 / ms, 1) if name in GFLOP_FWD else None}
            best = min((rec[m]["ms_per_step"] for m in ("tf32", "bf16_cl") if "ms_per_step" in rec[m]), default=None)
            if best:
                rec["speedup_vs_best_library"] = round(best / ms, 2)
        except Exception as e:
            rec["b200seg"] = {"error": repr(e)[:300]}
        print(json.dumps(rec), flush=True)
