"""Timing of the bandwidth-bound kernels at the full-resolution U-Net shape (2 x 128^3 x 32) (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
y = torch.randn(2, S, S, S, C, device="cuda").bfloat16()
dz = torch.randn(2, S, S, S, C, device="cuda").bfloat16()
gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
spec = F.NormSpec("batch", "relu")
elems = y.numel()
def t(fn, bytes_per_elem, name, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-22s %7.3f ms  %7.1f GB/s (algorithmic %d B/elt)" % (name, ms, elems * bytes_per_elem / ms / 1e6, bytes_per_elem), flush=True)
stats = F.channel_stats(y)
z, coef, count, groups = F._norm_forward(y, stats, spec, gamma, beta, rm, rv, None, None, None)
t(lambda: F.channel_stats(y), 2, "channel_stats")
t(lambda: F._norm_forward(y, stats, spec, gamma, beta, rm, rv, None, None, None), 4, "norm_act_fwd(+finalize)")
t(lambda: F._norm_backward(dz, y, coef, count, groups, spec, None, None, False), 10, "norm_act_bwd(red+apply)")
sums = torch.zeros(2 * C, device="cuda")
t(lambda: F._call("b200seg_norm_act_bwd_reduce", F._ptr(dz), C, F._ptr(y), C, F._ptr(coef), 2 * S ** 3, 1, C, spec.act,
                  spec.act_param, None, None, 0, F._ptr(sums), None, None, F._stream()), 4, "norm_act_bwd_reduce", reps=20)
t(lambda: F.max_pool2(y), 2.375, "maxpool_fwd")
w = torch.randn(2, C, 1, 1, 1, device="cuda"); b = torch.zeros(2, device="cuda")
yy = y.clone().requires_grad_(True); ww = w.clone().requires_grad_(True)
lg = F.head_conv1x1(yy, ww, b)
dl = torch.randn_like(lg)
t(lambda: F.head_conv1x1(y, w, b), 2 + 8.0 / C * 2 / 2, "head_fwd")
t(lambda: torch.autograd.grad(F.head_conv1x1(yy, ww, b), (yy, ww), dl), 4, "head_fwd+bwd")
