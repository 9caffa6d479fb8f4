import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from b200seg import parallel
from b200seg.engine import TrainStep
from b200seg.models.sync_batchnorm.batchnorm import convert_model
from b200seg.models.three_d.unet3d import UNet3D
from b200seg.optim import FusedAdam
from b200seg.utils.loss_function import DiceCELoss
from oracle import unet3d as ounet
rank, local, world = parallel.init_from_env("nccl")
dev = torch.device("cuda", local)
sd = ounet.init_state_dict(1, 2, 16, seed=3)
torch.manual_seed(100 + rank)
data = [(torch.randn(2, 1, 32, 32, 32, device=dev), (torch.rand(2, 32, 32, 32, device=dev) > 0.8).to(torch.uint8)) for _ in range(3)]
for reducer in (False, True, False, True):
    net = UNet3D(1, 2, 16).to(dev); net.load_state_dict(sd); convert_model(net); net.train()
    opt = FusedAdam(net.parameters(), lr=1e-3)
    if reducer: opt.attach_reducer()
    step = TrainStep(net, DiceCELoss(2), opt, use_graph=False)
    for it, (x, y) in enumerate(data):
        step(x, y)
        torch.cuda.synchronize()
        if reducer and rank == 1:
            print("  reducer state: buckets", len(opt.reducer.buckets), "stream", opt.reducer._stream, flush=True)
        bad = []
        for k, p in net.named_parameters():
            o = p.detach().clone(); dist.broadcast(o, 0)
            d = float((p.detach() - o).abs().max())
            if d > 0: bad.append((k, d))
        g = opt.grad_arena.clone(); dist.broadcast(g, 0)
        if rank == 1:
            print("reducer", reducer, "step", it, "differing params", len(bad), bad[:6], "grad arena diff", float((opt.grad_arena - g).abs().max()), flush=True)
    if hasattr(opt, "reducer"): opt.reducer.remove()
dist.barrier(); dist.destroy_process_group()
