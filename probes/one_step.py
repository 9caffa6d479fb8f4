"""Two eager warm-up steps, then ONE U-Net(1,2,32) training step between cudaProfilerStart/Stop (target for the
ncu launch list: ncu --profile-from-start off ...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200seg.engine import TrainStep
from b200seg.models.three_d.unet3d import UNet3D
from b200seg.optim import FusedAdam
from b200seg.utils.loss_function import DiceCELoss
dev = torch.device("cuda")
torch.manual_seed(0)
net = UNet3D(1, 2, 32).to(dev).train()
opt = FusedAdam(net.parameters(), lr=1e-3)
step = TrainStep(net, DiceCELoss(2), opt, use_graph=False)
x = torch.randn(2, 1, 128, 128, 128, device=dev)
lab = (torch.rand(2, 128, 128, 128, device=dev) > 0.9).to(torch.uint8)
for _ in range(2):
    step(x, lab)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: covers the autograd thread too
loss, _ = step(x, lab)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
