"""K-heavy layers at 8^3: weights-stationary kernel vs the voxel-tiled kernels (B200SEG_DISABLE_WS=1), cold weights."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(cin, cout, k):
    torch.manual_seed(0)
    x = torch.randn(2, 8, 8, 8, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, k, k, k, device="cuda") * 0.02
    b = torch.randn(cout, device="cuda") * 0.1
    res = {}
    for mode in ("ws", "tiled"):
        if mode == "tiled": os.environ["B200SEG_DISABLE_WS"] = "1"
        else: os.environ.pop("B200SEG_DISABLE_WS", None)
        ts = []
        for it in range(4):
            flush.zero_()
            F.profile_begin()
            y, st, g = F.conv3d_fprop_raw(x, w, b, k, 1, k // 2, 1, True)
            prof = F.profile_end()
            ts.append(sum(v["ms"] for kk, v in prof.items() if "conv_fprop" in kk))
        res[mode] = (min(ts), y.clone(), st.clone())
    os.environ.pop("B200SEG_DISABLE_WS", None)
    fl = 2.0 * 1024 * cin * cout * k ** 3
    print("%d -> %d k%d at 2 x 8^3: ws %.1f us (%.0f TF/s)  tiled %.1f us (%.0f TF/s)  rel %.1e  stats rel %.1e" % (
        cin, cout, k, res["ws"][0] * 1e3, fl / res["ws"][0] / 1e9, res["tiled"][0] * 1e3, fl / res["tiled"][0] / 1e9,
        float((res["ws"][1].float() - res["tiled"][1].float()).norm() / res["tiled"][1].float().norm()),
        float((res["ws"][2] - res["tiled"][2]).norm() / res["tiled"][2].norm())), flush=True)
for c in ((256, 512, 3), (512, 512, 3), (512, 256, 3), (256, 256, 5)):
    run(*c)
