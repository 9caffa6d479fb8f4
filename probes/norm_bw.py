"""Bandwidth of the three normalisation passes per U-Net level (batch 2): achieved TB/s against the 6.55 TB/s copy peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg._lib as _lib
if os.environ.get("B200SEG_LIB_OVERRIDE"):       # probes/norm_bw_variants.sh: a library variant under test
    _lib.LIB_PATH = os.path.abspath(os.environ["B200SEG_LIB_OVERRIDE"])
import b200seg.functional as F
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for s, c in ((128, 32), (64, 64), (32, 128), (16, 256), (8, 512)):
    torch.manual_seed(0)
    y = torch.randn(2, s, s, s, c, device=dev).bfloat16().requires_grad_(True)
    gamma = torch.ones(c, device=dev, requires_grad=True)
    beta = torch.zeros(c, device=dev, requires_grad=True)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    spec = F.NormSpec("batch", "relu", 0.0, training=True)
    best = {}
    for it in range(4):
        flush.zero_()
        F.profile_begin()
        z = F.norm_act(y, spec, gamma=gamma, beta=beta, running_mean=rm, running_var=rv)
        dz = torch.randn_like(z)
        flush.zero_()
        z.backward(dz)
        prof = F.profile_end()
        for k, v in prof.items():
            best[k] = min(best.get(k, 1e9), v["ms"])
    mb = 2 * s ** 3 * c * 2 / 1e6
    passes = {"b200seg_norm_act_fwd": 2, "b200seg_norm_act_bwd_reduce": 2, "b200seg_norm_act_bwd_apply": 3, "b200seg_channel_stats": 1}
    print("%3d^3 x %3d ch (%.0f MB per tensor): " % (s, c, mb) + "  ".join(
        "%s %.1f us (%.2f TB/s)" % (k.replace("b200seg_", ""), best[k] * 1e3, passes[k] * mb / best[k] / 1e3)
        for k in passes if k in best) + "   others: " + ", ".join("%s %.1f us" % (k.replace("b200seg_", ""), v * 1e3) for k, v in best.items() if k not in passes), flush=True)
