"""Which CUDA kernels of a training step are NOT ours (ATen copies / fills / adds issued by autograd or by host glue)?
torch.profiler over one eager step of each model; kernels are split into b200seg (our library) and everything else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from b200seg.engine import TrainStep
from b200seg.optim import FusedAdam
from b200seg.utils.loss_function import DiceCELoss
from library_bars import our_model, SIZES

dev = torch.device("cuda")
for name in (sys.argv[1:] or ["unet", "vnet", "res_unet", "highres", "densevoxel"]):
    torch.manual_seed(0)
    net = our_model(name).to(dev).train()
    opt = FusedAdam(net.parameters(), lr=1e-3)
    step = TrainStep(net, DiceCELoss(2), opt, use_graph=False)
    size = SIZES[name]
    x = torch.randn(2, 1, size, size, size, device=dev)
    lab = (torch.rand(2, size, size, size, device=dev) > 0.9).to(torch.uint8)
    for _ in range(3):
        step(x, lab)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(x, lab)
        torch.cuda.synchronize()
    ours = other = 0.0
    rows = []
    for e in prof.key_averages():
        t = e.device_time_total / 1e3 if hasattr(e, "device_time_total") else e.cuda_time_total / 1e3
        if t <= 0:
            continue
        if "b200" in e.key:
            ours += t
        else:
            other += t
            rows.append((t, e.count, e.key[:90]))
    print("== %s: our kernels %.2f ms, other kernels %.3f ms" % (name, ours, other))
    for t, n, k in sorted(rows, reverse=True)[:8]:
        print("   %8.3f ms  %4d x  %s" % (t, n, k))
    del net, opt, step
    torch.cuda.empty_cache()
