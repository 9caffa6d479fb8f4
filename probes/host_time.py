"""How long does the host take to enqueue one training step (is the step launch-bound)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
from b200seg.models.three_d.unet3d import UNet3D
from b200seg.optim import FusedAdam
from b200seg.utils.loss_function import DiceCELoss
dev = torch.device("cuda")
net = UNet3D(1, 2, 32).to(dev).train()
opt = FusedAdam(net.parameters(), lr=1e-3)
crit = DiceCELoss(2)
x = torch.randn(2, 1, 128, 128, 128, device=dev)
lab = (torch.rand(2, 128, 128, 128, device=dev) > 0.9).to(torch.uint8)
def step():
    opt.zero_grad()
    loss = crit(net(x), lab)
    loss.backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step, device-complete %.2f ms/step" % ((t1 - t0) * 100, (t2 - t0) * 100))
