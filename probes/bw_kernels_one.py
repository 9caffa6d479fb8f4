"""The bandwidth-bound launches of the finest U-Net level, alone, between cudaProfilerStart/Stop (ncu --set full target):
ConvTranspose3d(64 -> 32, k2 s2) forward / data gradient / weight gradient at 64^3 -> 128^3 and MaxPool3d(2, 2) forward /
backward (with the skip-connection gradient added) at 128^3 x 32 channels, batch 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
torch.manual_seed(0)
dev = "cuda"
x = torch.randn(2, 64, 64, 64, 64, device=dev).bfloat16().requires_grad_(True)
w = (torch.randn(64, 32, 2, 2, 2, device=dev) * 0.05).requires_grad_(True)
b = torch.zeros(32, device=dev, requires_grad=True)
e = torch.randn(2, 128, 128, 128, 32, device=dev).bfloat16().requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(profile):
    F.profile_begin()
    y = F.conv_transpose_kxsx(x, w, b)
    p, skip = F.max_pool2_skip(e)
    gy, gp, gs = torch.randn_like(y), torch.randn_like(p), torch.randn_like(skip)
    flush.zero_()
    torch.autograd.backward((y, p, skip), (gy, gp, gs))
    prof = F.profile_end()
    return prof
run(False)
torch.cuda.synchronize()
flush.zero_()
torch.cuda.profiler.start()
prof = run(True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
for k, v in prof.items():
    print("%-40s %d x %.3f ms" % (k, v["launches"], v["ms"]))
