"""One launch each of the two kernels added late in round 2, between cudaProfilerStart/Stop (ncu --set full target):
the weights-stationary kernel (512 -> 512, 3x3x3, 2 x 8^3) and the kd-stacked 5x5x5 kernel (32 -> 32, 2 x 128^3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
torch.manual_seed(0)
cases = []
for cin, cout, k, s in ((512, 512, 3, 8), (32, 32, 5, 128)):
    x = torch.randn(2, s, s, s, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, k, k, k, device="cuda") * 0.02
    b = torch.randn(cout, device="cuda") * 0.1
    cases.append((x, w, b, k))
for x, w, b, k in cases:
    F.conv3d_fprop_raw(x, w, b, k, 1, k // 2, 1, True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for x, w, b, k in cases:
    flush.zero_()
    F.conv3d_fprop_raw(x, w, b, k, 1, k // 2, 1, True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
