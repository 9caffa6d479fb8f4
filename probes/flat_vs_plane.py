import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
def run(cin, cout, s, k=3, n=2, reps=30):
    torch.manual_seed(0)
    x = torch.randn(n, s, s, s, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, k, k, k, device="cuda") * 0.05
    g = F._geom(x.shape, cin, cout, k, 1, (k - 1) // 2, 1)
    wp = F.pack_conv_weight(w)
    y = torch.empty((n, s, s, s, cout), dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * cout + 1, device="cuda")
    import ctypes
    def once():
        F._call("b200seg_conv3d_fprop", ctypes.byref(g), F._ptr(x), cin, F._ptr(wp), None, F._ptr(y), cout, F._ptr(stats), None, 0, F._stream())
    out = {}
    for mode in ("flat", "plane"):
        if mode == "plane": os.environ["B200SEG_DISABLE_FLAT"] = "1"
        else: os.environ.pop("B200SEG_DISABLE_FLAT", None)
        once(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): once()
        e1.record(); torch.cuda.synchronize()
        out[mode] = (e0.elapsed_time(e1) / reps, y.clone())
    fl = 2.0 * n * s ** 3 * cin * cout * k ** 3
    print("cin %d cout %d %d^3 k%d: flat %.3f ms (%.0f TF/s)  short-plane %.3f ms (%.0f TF/s)  rel %.1e" % (cin, cout, s, k, out["flat"][0], fl / out["flat"][0] / 1e9, out["plane"][0], fl / out["plane"][0] / 1e9, float((out["flat"][1].float() - out["plane"][1].float()).norm() / out["flat"][1].float().norm())), flush=True)
for c in ((256, 512, 8), (512, 512, 8), (512, 256, 8), (256, 256, 8), (128, 128, 8), (256, 256, 8, 5), (256, 256, 4)):
    run(*c)
