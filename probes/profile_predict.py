"""Per-kernel-family time of sliding-window inference over one 512x512x256 volume (configs[4]), eager, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
from b200seg.inference import sliding_window_predict
from b200seg.models.three_d.unet3d import UNet3D
_orig = F._call
def _named(name, *args, work=0.0, tag=None):
    if tag is None:
        tag = name.replace("b200seg_", "")
    return _orig(name, *args, work=work, tag=tag)
F._call = _named
dev = torch.device("cuda")
torch.manual_seed(0)
net = UNet3D(1, 2, 32).to(dev).eval()
vol = torch.randn(1, 512, 512, 256, device=dev)
for mode in ("crop",):
    sliding_window_predict(net, vol[:, :256, :256, :128], (128,) * 3, (64,) * 3, batch_size=16, overlap_mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    F.profile_begin()
    e0.record()
    out = sliding_window_predict(net, vol, (128,) * 3, (64,) * 3, batch_size=16, overlap_mode=mode)
    e1.record()
    prof = F.profile_end()
    total = sum(v["ms"] for v in prof.values())
    print("== predict (%s): %.1f ms wall (events), %.1f ms in %d of our launches" % (mode, e0.elapsed_time(e1), total, sum(v["launches"] for v in prof.values())))
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:20]:
        print("  %-40s %5d x %9.3f ms  %5.1f %%" % (k, v["launches"], v["ms"], 100 * v["ms"] / total))
