"""CTA-pair conv kernel against the single-CTA plane kernel: numerics + timing.  usage: pair_test.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F

def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))

def run(cin, cout, s, n=2, k=3, reps=20):
    torch.manual_seed(0)
    x = torch.randn(n, s, s, s, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, k, k, k, device="cuda") * 0.05
    b = torch.randn(cout, device="cuda") * 0.1
    wp = F.pack_conv_weight(w)
    def once():
        return F.conv3d_fprop_raw(x, w, b, k, 1, (k - 1) // 2, 1, True)
    res = {}
    for mode in ("plane", "pair"):
        if mode == "plane":
            os.environ["B200SEG_DISABLE_PAIR"] = "1"
        else:
            os.environ.pop("B200SEG_DISABLE_PAIR", None)
        y, st, g = once()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            once()
        e1.record()
        torch.cuda.synchronize()
        res[mode] = (y.clone(), st.clone(), e0.elapsed_time(e1) / reps)
    fl = 2.0 * n * s ** 3 * cin * cout * k ** 3
    print("cin %4d cout %4d %3d^3: y rel %.2e  stats rel %.2e | plane %.3f ms (%.0f TF/s)  pair %.3f ms (%.0f TF/s)" % (
        cin, cout, s, rel(res["pair"][0], res["plane"][0]), rel(res["pair"][1][:2 * cout], res["plane"][1][:2 * cout]),
        res["plane"][2], fl / res["plane"][2] / 1e9, res["pair"][2], fl / res["pair"][2] / 1e9), flush=True)

for cin, cout, s in ((256, 256, 16), (128, 128, 32), (512, 256, 16), (128, 256, 16), (256, 128, 32), (64, 128, 32), (128, 64, 64)):
    run(cin, cout, s)
