"""One conv geometry, one pass kind, a few launches (target for `ncu --set full`).  usage: one_conv.py cin cout s [fprop|dgrad|wgrad] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
cin, cout, s = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
what = sys.argv[4] if len(sys.argv) > 4 else "fprop"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
torch.manual_seed(0)
x = torch.randn(2, s, s, s, cin, device="cuda").bfloat16()
w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
dy = torch.randn(2, s, s, s, cout, device="cuda").bfloat16()
y, stats, g = F.conv3d_fprop_raw(x, w, None, 3, 1, 1, 1, True)
fn = {"fprop": lambda: F.conv3d_fprop_raw(x, w, None, 3, 1, 1, 1, True), "dgrad": lambda: F.conv3d_dgrad_raw(g, dy, w),
      "wgrad": lambda: F.conv3d_wgrad_raw(g, x, dy, w.shape)}[what]
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("%s cin %d cout %d s %d: %.3f ms  %.1f TF/s" % (what, cin, cout, s, ms, 2.0 * 2 * s ** 3 * cin * cout * 27 / ms / 1e9))
