"""5x5x5 layers of V-Net at full resolution: time of fprop / dgrad with the kd-stacked kernel vs the plane kernel."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
def run(cin, cout, s=128, n=2, reps=10):
    torch.manual_seed(0)
    x = torch.randn(n, s, s, s, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, 5, 5, 5, device="cuda") * 0.02
    g = F._geom(x.shape, cin, cout, 5, 1, 2, 1)
    wp = F.pack_conv_weight(w)
    y = torch.empty((n, s, s, s, cout), dtype=torch.bfloat16, device="cuda")
    def once():
        F._call("b200seg_conv3d_fprop", ctypes.byref(g), F._ptr(x), cin, F._ptr(wp), None, F._ptr(y), cout, None, None, 0, F._stream())
    out = {}
    for mode in ("roll5", "plane"):
        if mode == "plane": os.environ["B200SEG_DISABLE_ROLL5"] = "1"
        else: os.environ.pop("B200SEG_DISABLE_ROLL5", None)
        once(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): once()
        e1.record(); torch.cuda.synchronize()
        out[mode] = (e0.elapsed_time(e1) / reps, y.clone())
    fl = 2.0 * n * s ** 3 * cin * cout * 125
    print("cin %d cout %d %d^3 k5: roll %.3f ms (%.0f TF/s)  plane %.3f ms (%.0f TF/s)  rel %.1e" % (cin, cout, s, out["roll5"][0], fl / out["roll5"][0] / 1e9, out["plane"][0], fl / out["plane"][0] / 1e9, float((out["roll5"][1].float() - out["plane"][1].float()).norm() / out["plane"][1].float().norm())), flush=True)
for c in ((16, 16), (32, 16), (16, 32), (32, 32)):
    run(*c)
