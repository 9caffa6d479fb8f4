"""HighRes3DNet's 16 -> 16 3x3x3 layers at 2 x 96^3: rolling-accumulator kernel vs the plane kernel."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
torch.manual_seed(0)
x = torch.randn(2, 96, 96, 96, 16, device="cuda").bfloat16()
w = torch.randn(16, 16, 3, 3, 3, device="cuda") * 0.05
b = torch.randn(16, device="cuda") * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for mode in ("roll", "plane"):
    if mode == "plane": os.environ["B200SEG_ROLL_STRICT_BLOCKS"] = "1"
    else: os.environ.pop("B200SEG_ROLL_STRICT_BLOCKS", None)
    ts = []
    for it in range(4):
        flush.zero_()
        F.profile_begin()
        y, st, g = F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, True)
        prof = F.profile_end()
        ts.append(sum(v["ms"] for v in prof.values()))
    res[mode] = (min(ts), y.clone(), st.clone())
print("16 -> 16 k3 at 2 x 96^3: roll %.1f us  plane %.1f us  rel %.1e stats rel %.1e" % (res["roll"][0] * 1e3, res["plane"][0] * 1e3,
      float((res["roll"][1].float() - res["plane"][1].float()).norm() / res["plane"][1].float().norm()),
      float((res["roll"][2] - res["plane"][2]).norm() / res["plane"][2].norm())))
