"""One weight-gradient launch of a full-resolution U-Net layer between cudaProfilerStart/Stop (ncu --set full target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
cin, cout, s, n = (int(a) for a in (sys.argv[1:5] + ["32", "32", "128", "2"][len(sys.argv) - 1:]))
torch.manual_seed(0)
x = torch.randn(n, s, s, s, cin, device="cuda").bfloat16()
dy = torch.randn(n, s, s, s, cout, device="cuda").bfloat16()
g = F._geom(x.shape, cin, cout, 3, 1, 1, 1)
F.conv3d_wgrad_raw(g, x, dy, (cout, cin, 3, 3, 3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
F.conv3d_wgrad_raw(g, x, dy, (cout, cin, 3, 3, 3))
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("wgrad %d->%d %d^3 n%d: %.3f ms" % (cin, cout, s, n, e0.elapsed_time(e1)))
