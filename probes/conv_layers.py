"""Per-layer timing of the U-Net conv geometries (fprop / dgrad / wgrad), batch 2 x 128^3 (diagnostic)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F

LAYERS = [  # name, cin, cout, spatial
    ("enc1conv2", 32, 32, 128), ("dec1conv1", 64, 32, 128), ("dec1conv2", 32, 32, 128),
    ("enc2conv1", 32, 64, 64), ("enc2conv2", 64, 64, 64), ("dec2conv1", 128, 64, 64),
    ("enc3conv1", 64, 128, 32), ("enc3conv2", 128, 128, 32), ("dec3conv1", 256, 128, 32),
    ("enc4conv1", 128, 256, 16), ("enc4conv2", 256, 256, 16), ("dec4conv1", 512, 256, 16),
    ("bott1", 256, 512, 8), ("bott2", 512, 512, 8),
]
only = sys.argv[1:] if len(sys.argv) > 1 else None
reps = 20
for name, cin, cout, s in LAYERS:
    if only and name not in only:
        continue
    x = torch.randn(2, s, s, s, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
    dy = torch.randn(2, s, s, s, cout, device="cuda").bfloat16()
    y, stats, g = F.conv3d_fprop_raw(x, w, None, 3, 1, 1, 1, True)
    flops = 2.0 * 2 * s ** 3 * cin * cout * 27
    res = {}
    for what, fn in (("fprop", lambda: F.conv3d_fprop_raw(x, w, None, 3, 1, 1, 1, True)),
                     ("dgrad", lambda: F.conv3d_dgrad_raw(g, dy, w)),
                     ("wgrad", lambda: F.conv3d_wgrad_raw(g, x, dy, w.shape))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[what] = (ms, flops / ms / 1e9)
    print("%-10s cin %3d cout %3d s %3d  GF %6.1f | " % (name, cin, cout, s, flops / 1e9) +
          "  ".join("%s %7.3f ms %6.1f TF/s" % (k, v[0], v[1]) for k, v in res.items()), flush=True)
