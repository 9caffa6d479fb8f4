"""ConvTranspose3d(64 -> 32, k2 s2) forward at 2 x 64^3 -> 128^3 under a few planner knobs (which N tile, store width)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
torch.manual_seed(0)
x = torch.randn(2, 64, 64, 64, 64, device="cuda").bfloat16()
w = torch.randn(64, 32, 2, 2, 2, device="cuda") * 0.05
b = torch.zeros(32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for knobs in ({}, {"B200SEG_SCATTER_NT": "256"}, {"B200SEG_SCATTER_NT": "64"}, {"B200SEG_SCATTER_NT": "32"}, {"B200SEG_NO_WIDE_STORES": "1"},
              {"B200SEG_SCATTER_NT": "256", "B200SEG_NO_WIDE_STORES": "1"}):
    for k in ("B200SEG_SCATTER_NT", "B200SEG_NO_WIDE_STORES"):
        os.environ.pop(k, None)
    os.environ.update(knobs)
    y = F.conv_transpose_kxsx(x, w, b)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        F.profile_begin()
        y = F.conv_transpose_kxsx(x, w, b)
        prof = F.profile_end()
        ts.append(prof["b200seg_convt_k2s2_fwd"]["ms"])
    if ref is None: ref = y.clone()
    print(knobs, "min %.3f ms  median %.3f ms  (335 MB -> %.2f TB/s)  equal %s" % (min(ts), sorted(ts)[2], 0.335 / min(ts), torch.equal(y, ref)), flush=True)
