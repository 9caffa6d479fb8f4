"""Per-parameter gradient error report: b200seg UNet3D vs the CPU oracle (diagnostic, not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200seg.models.three_d.unet3d import UNet3D
from b200seg.utils.loss_function import DiceCELoss
from oracle import losses as olosses, unet3d as ounet

f, size, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
sd = ounet.init_state_dict(1, 2, f, seed=0)
net = UNet3D(1, 2, f).cuda(); net.load_state_dict(sd); net.train()
x = torch.randn(batch, 1, size, size, size); lab = (torch.rand(batch, size, size, size) > 0.8).long()
out = net(x.cuda()); loss = DiceCELoss(2)(out, lab.cuda()); loss.backward()
rsd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
acts = {}
rout = ounet.forward(rsd, x, training=True, acts=acts); rloss = olosses.dice_ce(rout, lab); rloss.backward()
print("loss", loss.item(), rloss.item(), "logits rel", ((out.cpu() - rout).norm() / rout.norm()).item())
for name, p in net.named_parameters():
    r = rsd[name].grad
    e = ((p.grad.cpu() - r).norm() / (r.norm() + 1e-20)).item()
    print("%-40s ref_norm %.3e  rel_err %.4f" % (name, r.norm().item(), e))
