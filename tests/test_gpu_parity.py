"""GPU parity: every kernel reached through the C ABI is compared with the CPU oracle (and, where they exist, with
the golden vectors produced by the reference itself).  Tolerances: activations are stored as bf16 (relative step
2^-8), accumulation is fp32; integer outputs (pool indices, labels, counts, stitched label maps) are bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import losses as olosses  # noqa: E402
from oracle import metric as ometric  # noqa: E402
from oracle import ops as oops  # noqa: E402
from oracle import unet3d as ounet  # noqa: E402
from oracle import window as owindow  # noqa: E402


@pytest.fixture(scope="module")
def F():
    import b200seg.functional as F
    return F


DEV = "cuda"


def bf(x):
    """round an fp32 tensor to bf16 and back (what the kernels see)."""
    return x.to(torch.bfloat16).float()


def ndhwc(x):
    return x.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).to(DEV)


def ncdhw(x):
    return x.detach().float().cpu().permute(0, 4, 1, 2, 3).contiguous()


def rel_err(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def close(a, b, tol, what=""):
    e = rel_err(a, b)
    m = float((a - b).abs().max() / (b.abs().max() + 1e-12))
    assert e < tol and m < 4 * tol, "%s rel-fro %.4g rel-max %.4g (tol %.3g)" % (what, e, m, tol)


CONV_CASES = [
    # cin, cout, k, stride, pad, dil, (d,h,w), n
    (1, 32, 3, 1, 1, 1, (8, 8, 8), 2),        # U-Net stem (dedicated C_in=1 kernels)
    (1, 32, 3, 1, 1, 1, (9, 20, 33), 1),      # stem, ragged extents
    (1, 32, 3, 1, 1, 1, (32, 48, 48), 1),     # stem large enough for the K-padded tensor-core path (>= 65536 voxels)
    (1, 16, 3, 1, 1, 1, (40, 40, 44), 1),     # HighResNet stem (C_out = 16), K-padded
    (1, 16, 5, 1, 2, 1, (8, 8, 12), 1),       # V-Net stem 5x5x5
    (32, 32, 3, 1, 1, 1, (8, 16, 8), 2),      # full-resolution U-Net layer
    (64, 32, 3, 1, 1, 1, (8, 16, 16), 1),     # decoder conv1 (concat input)
    (32, 64, 3, 1, 1, 1, (4, 16, 8), 2),
    (64, 128, 3, 1, 1, 1, (8, 8, 8), 1),
    (128, 64, 3, 1, 1, 1, (4, 8, 8), 2),
    (16, 16, 3, 1, 2, 2, (10, 12, 16), 1),    # HighResNet dilation 2 ('same')
    (16, 32, 3, 1, 0, 1, (10, 20, 12), 1),    # HighResNet: valid conv on pre-padded input
    (16, 16, 5, 1, 2, 1, (8, 16, 8), 1),      # V-Net 5x5x5
    (16, 32, 2, 2, 0, 1, (8, 8, 8), 2),       # V-Net down conv
    (32, 64, 3, 2, 1, 1, (8, 8, 8), 1),       # residual U-Net strided conv
    (28, 12, 3, 1, 1, 1, (6, 6, 6), 1),       # DenseVoxelNet odd widths
    (32, 2, 1, 1, 0, 1, (4, 4, 8), 2),        # 1x1x1
    (256, 256, 3, 1, 1, 1, (2, 4, 8), 1),     # deep level, tiny spatial extent
    (32, 32, 3, 1, 1, 1, (20, 32, 24), 1),    # several d-tiles: input-plane ring wraps around
    (64, 64, 3, 1, 1, 1, (6, 32, 16), 2),     # multiple h/w tiles, batch 2
    (128, 128, 3, 1, 1, 1, (4, 16, 16), 1),   # two K chunks in plane mode
    (512, 256, 3, 1, 1, 1, (4, 4, 4), 1),     # 8 K chunks, 2 N tiles, flat mode
    (32, 32, 3, 1, 1, 1, (5, 17, 11), 1),     # ragged extents: masked rows in the last tiles
    # round 2: stride-2 / 16-channel geometries of the other model families on the tensor cores
    (32, 64, 3, 2, 1, 1, (32, 32, 32), 1),    # residual U-Net down-step: parity-class decomposition, 16^3 output
    (64, 128, 3, 2, 1, 1, (16, 16, 16), 2),   # ... 8^3 output (short tiles), batch 2
    (32, 64, 3, 2, 1, 1, (18, 34, 22), 1),    # ... ragged tiles
    (32, 32, 3, 2, 1, 1, (17, 33, 19), 1),    # ... odd extents: the parity classes have different sizes
    (16, 32, 2, 2, 0, 1, (32, 32, 32), 1),    # V-Net down conv as gather GEMM / pixel-shuffle GEMM
    (128, 256, 2, 2, 0, 1, (8, 16, 16), 1),   # V-Net deep down conv
    (16, 16, 3, 1, 1, 1, (16, 32, 16), 1),    # HighRes3DNet: C_in = 16 weight gradient (zero-filled half chunk)
    (48, 32, 3, 1, 1, 1, (8, 16, 16), 1),     # ragged last C_in chunk in the weight gradient
    (64, 64, 3, 1, 4, 4, (12, 24, 16), 1),    # HighRes3DNet dilation 4: narrow K chunk so that the 12-plane ring fits
    (32, 64, 3, 1, 4, 4, (10, 16, 24), 2),
    (64, 64, 5, 1, 2, 1, (8, 8, 8), 1),       # V-Net 5x5x5 at 8^3: short planes (masked tile rows)
    (128, 128, 5, 1, 2, 1, (4, 4, 4), 2),
    (32, 32, 5, 1, 2, 1, (12, 24, 20), 2),    # V-Net full-resolution 5x5x5: kd taps stacked in N, two 16-channel halves
    (32, 16, 5, 1, 2, 1, (21, 17, 9), 1),     # ... ragged tiles
    # channel counts that are not multiples of 16, zero-padded onto the tensor cores (>= 65536 voxels)
    (28, 12, 3, 1, 1, 1, (40, 40, 44), 1),    # DenseVoxelNet growth conv
    (32, 2, 5, 1, 2, 1, (40, 40, 44), 1),     # V-Net output conv 5x5x5 to 2 classes
    (1, 16, 5, 1, 2, 1, (40, 40, 44), 1),     # V-Net stem 5x5x5
    (64, 2, 1, 1, 0, 1, (40, 40, 44), 1),     # HighRes3DNet classifier
    (24, 40, 3, 2, 1, 1, (64, 64, 64), 1),    # stride 2 with padded channels
    # K-heavy layers on 8 x 8 planes: weights-stationary kernel, partial sums of the tap rows meet in an fp32 workspace
    (256, 512, 3, 1, 1, 1, (8, 8, 8), 2),     # U-Net bottleneck conv1 (unet3d.py:28)
    (512, 256, 3, 1, 1, 1, (8, 8, 8), 2),     # ... the geometry of its data gradient's twin
    (256, 256, 5, 1, 2, 1, (8, 8, 8), 2),     # V-Net's deepest 5x5x5 layers (vnet3d.py:25)
    (256, 320, 3, 1, 1, 1, (4, 8, 8), 1),     # four planes, an N tile count that does not divide the SM count
    (32, 64, 3, 4, 0, 1, (16, 16, 20), 2),    # CSRNet cross-scale branch: stride 4, no padding (space-to-depth + 1x1x1 GEMM)
    (32, 128, 3, 4, 0, 1, (33, 34, 35), 1),   # ... extents that leave uncovered planes at the far end (zero gradient there)
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "ci%d_co%d_k%d_s%d_p%d_d%d_%s_n%d" % (
    c[0], c[1], c[2], c[3], c[4], c[5], "x".join(map(str, c[6])), c[7]))
def test_conv3d_fprop_dgrad_wgrad(F, case):
    cin, cout, k, stride, pad, dil, size, n = case
    g = torch.Generator().manual_seed(hash(case) % 2 ** 31)
    x = bf(torch.randn(n, cin, *size, generator=g))
    w = torch.randn(cout, cin, k, k, k, generator=g) * (2.0 / (cin * k ** 3)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    wq = bf(w)
    xr = x.clone().requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    y_ref = oops.conv3d(xr, wr, b, stride, pad, dil)
    dy = bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)

    xd = ndhwc(x).requires_grad_(True)
    wd = w.to(DEV).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True)
    n0 = F.umma_launch_count()
    y = F.conv_norm_act(xd, wd, bd, k=k, stride=stride, pad=pad, dil=dil)
    close(ncdhw(y), y_ref.detach(), 8e-3, "fprop")
    y.backward(ndhwc(dy))
    # which path ran: tcgen05 for stride-1 k in {1,3,5} with 16-aligned channels
    if stride == 1 and k in (1, 3, 5) and cin % 16 == 0 and cout % 16 == 0:
        assert F.umma_launch_count() - n0 == 3, "tensor-core path was not taken"
    elif stride > 2 and k <= stride and pad == 0 and (cin * k ** 3) % 16 == 0 and cout % 16 == 0:
        # non-overlapping windows: space-to-depth + a 1x1x1 convolution with k^3 * C_in inputs, all three passes on tcgen05
        assert F.umma_launch_count() - n0 == 3, "space-to-depth GEMM path was not taken"
    elif stride == 2 and cin % 16 == 0 and cout % 16 == 0 and min(size) >= 8:
        # k3s2: 1 forward + 8 data-gradient + 8 weight-gradient class launches; k2s2: one launch per pass
        assert F.umma_launch_count() - n0 == (17 if k == 3 else 3), "strided tensor-core path was not taken"
    elif y_ref.numel() // cout >= 1 << 16 or 2.0 * (y_ref.numel() // cout) * cout * cin * k ** 3 >= 2e8:
        # large volume with odd channel counts: zero-padded channels, every pass on the tensor cores (the C_in = 1 3x3x3
        # stem's weight gradient through its taps-as-channels copy)
        assert F.umma_launch_count() - n0 >= 3, "padded tensor-core path was not taken"
    else:
        assert F.umma_launch_count() == n0
    close(ncdhw(xd.grad), xr.grad, 8e-3, "dgrad")
    close(wd.grad.cpu(), wr.grad, 8e-3, "wgrad")
    close(bd.grad.cpu(), dy.sum((0, 2, 3, 4)), 8e-3, "bias grad")


@pytest.mark.parametrize("k", [3, 5])
def test_conv_stats_epilogue_matches_separate_reduction(F, k):
    g = torch.Generator().manual_seed(3)
    x = bf(torch.randn(2, 32, 8, 16, 8, generator=g))
    w = torch.randn(32, 32, k, k, k, generator=g) * 0.05 * (3.0 / k) ** 1.5
    y, stats, _ = F.conv3d_fprop_raw(ndhwc(x), w.to(DEV), None, k, 1, k // 2, 1, True)
    ref = oops.conv3d(x, bf(w), None, 1, k // 2, 1)
    s = torch.stack((ref.sum((0, 2, 3, 4)), (ref ** 2).sum((0, 2, 3, 4))))
    close(stats[:2 * w.shape[0]].view(2, -1).cpu(), s, 5e-3, "fused stats")   # flat {sum[C], sumsq[C], spare}
    close(F.channel_stats(y)[0].cpu(), s, 1e-2, "stand-alone stats")


@pytest.mark.parametrize("kind,act,c", [("batch", "relu", 32), ("batch", "elu", 16), ("instance", "leaky_relu", 32),
                                        ("batch", "prelu", 16), ("batch", "none", 2), (None, "relu", 12),
                                        ("batch", "relu", 12)])
def test_norm_act_forward_backward(F, kind, act, c):
    g = torch.Generator().manual_seed(7)
    x = bf(torch.randn(2, c, 6, 8, 8, generator=g) * 1.5 + 0.3)
    gamma = (torch.rand(c, generator=g) + 0.5) if kind == "batch" else None
    beta = torch.randn(c, generator=g) * 0.2 if kind == "batch" else None
    slope = torch.full((c,), 0.25) if act == "prelu" else None
    res = bf(torch.randn(x.shape, generator=g)) if act == "elu" else None

    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True) if gamma is not None else None
    br = beta.clone().requires_grad_(True) if beta is not None else None
    sr = slope.clone().requires_grad_(True) if slope is not None else None
    rr = res.clone().requires_grad_(True) if res is not None else None
    if kind == "batch":
        h, mean, var = oops.batch_norm_train(xr, gr, br)
    elif kind == "instance":
        h = oops.instance_norm(xr)
    else:
        h = xr
    if rr is not None:
        h = h + rr
    z_ref = oops.ACTIVATIONS[act](h, sr)
    dz = bf(torch.randn(z_ref.shape, generator=g))
    z_ref.backward(dz)

    xd = ndhwc(x).requires_grad_(True)
    gd = gamma.to(DEV).requires_grad_(True) if gamma is not None else None
    bd = beta.to(DEV).requires_grad_(True) if beta is not None else None
    sd = slope.to(DEV).requires_grad_(True) if slope is not None else None
    rd = ndhwc(res).requires_grad_(True) if res is not None else None
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    spec = F.NormSpec(kind, act, act_param=0.01)
    z = F.norm_act(xd, spec, gd, bd, sd, rd, rm if kind == "batch" else None, rv if kind == "batch" else None)
    close(ncdhw(z), z_ref.detach(), 1e-2, "forward")
    z.backward(ndhwc(dz))
    close(ncdhw(xd.grad), xr.grad, 2e-2, "dx")
    if gamma is not None:
        close(gd.grad.cpu(), gr.grad, 2e-2, "dgamma")
        close(bd.grad.cpu(), br.grad, 2e-2, "dbeta")
        cnt = x.numel() // c
        erm, erv = oops.batch_norm_running_update(torch.zeros(c), torch.ones(c), mean.detach(), var.detach(), cnt)
        close(rm.cpu(), erm, 1e-2, "running_mean")
        close(rv.cpu(), erv, 1e-2, "running_var")
    if slope is not None:
        close(sd.grad.cpu(), sr.grad, 2e-2, "dprelu")
    if res is not None:
        close(ncdhw(rd.grad), rr.grad, 1e-2, "dresidual")


def test_batchnorm_eval_mode(F):
    g = torch.Generator().manual_seed(11)
    c = 16
    x = bf(torch.randn(1, c, 4, 8, 8, generator=g))
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    ref = torch.relu(oops.batch_norm_eval(x, gamma, beta, rm, rv))
    spec = F.NormSpec("batch", "relu", training=False)
    z = F.norm_act(ndhwc(x), spec, gamma.to(DEV), beta.to(DEV), None, None, rm.to(DEV), rv.to(DEV))
    close(ncdhw(z), ref, 1e-2, "eval bn")


@pytest.mark.parametrize("cin,cout,size,act", [(32, 32, (8, 16, 8), "relu"), (64, 128, (4, 16, 16), "leaky_relu"),
                                               (32, 64, (8, 16, 16), "none"), (256, 256, (4, 4, 4), "relu"),
                                               (16, 16, (6, 32, 24), "relu"), (1, 32, (24, 48, 32), "relu")])
def test_eval_mode_conv_bn_act_fused_epilogue(F, cin, cout, size, act):
    """Inference: conv -> BatchNorm(eval) -> activation as ONE launch (scale / shift / activation in the conv epilogue) against
    the oracle's three separate ops, and against our own unfused two-pass path."""
    g = torch.Generator().manual_seed(cin + cout)
    x = bf(torch.randn(2, cin, *size, generator=g))
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (cin * 27)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.3
    rm, rv = torch.randn(cout, generator=g) * 0.2, torch.rand(cout, generator=g) + 0.5
    h = oops.batch_norm_eval(oops.conv3d(x, bf(w), b, 1, 1, 1), gamma, beta, rm, rv)
    ref = oops.ACTIVATIONS[act](h, None) if act != "none" else h
    if act == "leaky_relu":
        ref = torch.nn.functional.leaky_relu(h, 0.01)
    spec = F.NormSpec("batch", act, act_param=0.01, training=False)
    args = dict(k=3, stride=1, pad=1, dil=1, spec=spec, gamma=gamma.to(DEV), beta=beta.to(DEV), running_mean=rm.to(DEV),
                running_var=rv.to(DEV))
    with torch.no_grad():
        k0, u0 = F.launches(), F.umma_launch_count()
        z = F.conv_norm_act(ndhwc(x), w.to(DEV), b.to(DEV), **args)
        fused_launches = F.launches() - k0
        # (pack +) coefficients + conv; the C_in = 1 stem also widens its input to 16 channels first
        assert F.umma_launch_count() - u0 == 1 and fused_launches <= (4 if cin == 1 else 3), fused_launches
    close(ncdhw(z), ref, 1e-2, "fused eval conv+bn+act")
    z2 = F.conv_norm_act(ndhwc(x).requires_grad_(True), w.to(DEV), b.to(DEV), **args)   # grad mode: the unfused passes
    close(ncdhw(z2), ref, 1.2e-2, "unfused eval path")
    assert rel_err(ncdhw(z), ncdhw(z2)) < 8e-3


def test_maxpool_indices_bit_exact_with_ties_and_nan(F, golden):
    gz = golden("pool_argmax")
    x = torch.from_numpy(gz["x"])
    xd = ndhwc(x).requires_grad_(True)
    y, idx = F.max_pool2(xd, return_indices=True)
    tidx = F.maxpool_indices_to_torch(idx, xd.shape).cpu().numpy()
    assert np.array_equal(tidx, gz["idx"])                       # reference (torch) indices, bit-exact
    o_y, o_idx = oops.max_pool3d_k2s2(bf(x).numpy())
    assert np.array_equal(tidx, o_idx)
    yy = ncdhw(y).numpy()
    assert np.array_equal(np.isnan(yy), np.isnan(o_y)) and np.array_equal(np.nan_to_num(yy), np.nan_to_num(o_y))
    # backward: gradient lands exactly on the arg-max voxel
    dy = bf(torch.randn(y.shape))
    y.backward(dy.to(torch.bfloat16).to(DEV))
    xr = torch.nan_to_num(bf(x)).requires_grad_(True)
    yr, ir = torch.nn.functional.max_pool3d(xr, 2, 2, return_indices=True)
    gref = torch.zeros(x.shape).flatten(2)
    gref.scatter_(2, torch.from_numpy(gz["idx"]).flatten(2), ncdhw(dy).flatten(2))
    assert torch.equal(ncdhw(xd.grad), gref.view(x.shape))


def test_conv_transpose_k2s2(F):
    g = torch.Generator().manual_seed(5)
    for cin, cout, size in [(64, 32, (4, 4, 8)), (16, 8, (3, 5, 4)), (32, 32, (2, 2, 2)), (64, 32, (4, 16, 16)),
                            (128, 64, (3, 16, 8)), (512, 256, (2, 2, 2)), (64, 32, (3, 18, 10))]:
        x = bf(torch.randn(2, cin, *size, generator=g))
        w = torch.randn(cin, cout, 2, 2, 2, generator=g) * 0.1
        b = torch.randn(cout, generator=g) * 0.1
        xr, wr = x.clone().requires_grad_(True), bf(w).requires_grad_(True)
        y_ref = oops.conv_transpose3d_k2s2(xr, wr, b)
        dy = bf(torch.randn(y_ref.shape, generator=g))
        y_ref.backward(dy)
        xd, wd, bd = ndhwc(x).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        y = F.conv_transpose_k2s2(xd, wd, bd)
        close(ncdhw(y), y_ref.detach(), 8e-3, "convT fwd")
        y.backward(ndhwc(dy))
        close(ncdhw(xd.grad), xr.grad, 8e-3, "convT dgrad")
        close(wd.grad.cpu(), wr.grad, 8e-3, "convT wgrad")
        close(bd.grad.cpu(), dy.sum((0, 2, 3, 4)), 8e-3, "convT bias grad")


def test_conv_transpose_k4s4(F):
    """nn.ConvTranspose3d(kernel 4, stride 4) (csrnet.py:137-149): a 1x1x1 convolution to 64 * C_out channels on the tensor
    cores followed by the pixel shuffle (C_in % 16 == 0), else the strided convolution's three passes."""
    g = torch.Generator().manual_seed(6)
    for cin, cout, size in [(32, 16, (2, 3, 4)), (64, 8, (2, 2, 2)), (128, 32, (4, 5, 6)), (24, 8, (2, 2, 3))]:
        x = bf(torch.randn(2, cin, *size, generator=g))
        w = torch.randn(cin, cout, 4, 4, 4, generator=g) * 0.1
        b = torch.randn(cout, generator=g) * 0.1
        xr, wr = x.clone().requires_grad_(True), bf(w).requires_grad_(True)
        y_ref = torch.nn.functional.conv_transpose3d(xr, wr, b, stride=4)
        dy = bf(torch.randn(y_ref.shape, generator=g))
        y_ref.backward(dy)
        xd, wd, bd = ndhwc(x).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        y = F.conv_transpose_kxsx(xd, wd, bd, stride=4)
        close(ncdhw(y), y_ref.detach(), 8e-3, "convT k4s4 fwd")
        y.backward(ndhwc(dy))
        close(ncdhw(xd.grad), xr.grad, 8e-3, "convT k4s4 dgrad")
        close(wd.grad.cpu(), wr.grad, 8e-3, "convT k4s4 wgrad")
        close(bd.grad.cpu(), dy.sum((0, 2, 3, 4)), 8e-3, "convT k4s4 bias grad")


@pytest.mark.parametrize("mode", ["reflect", "replicate"])
def test_pad3d_reflect_replicate(F, mode):
    """Pad3d (utils/convolution.py:78-86) forward and adjoint against torch.nn.functional.pad, bit-exact forward."""
    g = torch.Generator().manual_seed(12)
    for pad, shape in ((1, (2, 16, 5, 6, 7)), (2, (1, 12, 6, 5, 9)), (4, (1, 8, 6, 7, 5))):
        x = bf(torch.randn(*shape, generator=g))
        xr = x.clone().requires_grad_(True)
        ref = torch.nn.functional.pad(xr, 6 * [pad], mode)
        dy = bf(torch.randn(ref.shape, generator=g))
        ref.backward(dy)
        xd = ndhwc(x).requires_grad_(True)
        y = F.pad3d(xd, pad, mode)
        assert torch.equal(ncdhw(y), ref.detach())
        y.backward(ndhwc(dy))
        close(ncdhw(xd.grad), xr.grad, 6e-3, "pad3d adjoint (bf16 sums of up to (pad+1)^3 terms)")
    # through ConvolutionalBlock (HighResNet's building block) on the CUDA kernels
    from b200seg.utils.convolution import ConvolutionalBlock
    blk = ConvolutionalBlock(16, 16, 2, 3, padding_mode=mode, batch_norm=False).to(DEV)
    x = bf(torch.randn(1, 16, 10, 12, 16, generator=g))
    conv = blk.convolutional_block[-1]
    want = torch.nn.functional.conv3d(torch.nn.functional.pad(torch.relu(x), 6 * [2], mode), bf(conv.weight.detach().cpu()),
                                      conv.bias.detach().cpu(), dilation=2)
    close(ncdhw(blk(ndhwc(x))), want.detach(), 1e-2, "ConvolutionalBlock " + mode)


def test_concat_free_decoder_input_equals_torch_cat(F):
    g = torch.Generator().manual_seed(9)
    a = bf(torch.randn(1, 16, 4, 8, 8, generator=g))
    b = bf(torch.randn(1, 16, 4, 8, 8, generator=g))
    w = torch.randn(32, 32, 3, 3, 3, generator=g) * 0.05
    buf, va, vb = F.alloc_concat(1, 4, 8, 8, 16, 16, DEV)
    va.copy_(ndhwc(a))
    vb.copy_(ndhwc(b))
    va.requires_grad_(True)
    vb.requires_grad_(True)
    y = F.conv_norm_act(va, w.to(DEV), None, x2=vb, k=3, pad=1)
    cat = torch.cat((a, b), 1).requires_grad_(True)
    ref = oops.conv3d(cat, bf(w), None, 1, 1, 1)
    close(ncdhw(y), ref.detach(), 8e-3, "concat-free fprop")
    dy = bf(torch.randn(ref.shape, generator=g))
    ref.backward(dy)
    y.backward(ndhwc(dy))
    close(ncdhw(va.grad), cat.grad[:, :16], 8e-3, "d(first half)")
    close(ncdhw(vb.grad), cat.grad[:, 16:], 8e-3, "d(second half)")


def test_losses_match_reference_golden(F, golden):
    gz = golden("losses")
    pred = torch.from_numpy(gz["pred"])
    lab = torch.from_numpy(gz["lab"])
    cases = {"cross_entropy_3D": (1, 0, 0, 0), "DiceLossss_softmax": (0, 1, 0, 0), "DiceLoss": (0, 0, 1, 0),
             "BCEWithLogits": (0, 0, 0, 1)}
    for name, wts in cases.items():
        p = pred.to(DEV).requires_grad_(True)
        loss = F.seg_loss(p, lab.to(DEV), *[float(v) for v in wts])
        loss.backward()
        assert abs(loss.item() - float(gz[name])) < 2e-6, name
        assert torch.allclose(p.grad.cpu(), torch.from_numpy(gz[name + ".grad"]), rtol=2e-4, atol=1e-8), name
    # combined criterion with an upstream scale
    p = pred.to(DEV).requires_grad_(True)
    (3.0 * F.seg_loss(p, lab.to(DEV), 1.0, 1.0)).backward()
    ref = 3.0 * (torch.from_numpy(gz["cross_entropy_3D.grad"]) + torch.from_numpy(gz["DiceLossss_softmax.grad"]))
    assert torch.allclose(p.grad.cpu(), ref, rtol=2e-4, atol=1e-8)


def test_reference_named_loss_classes(golden):
    from b200seg.utils.loss_function import (Binary_Loss, DiceCELoss, DiceLoss, DiceLossss, cross_entropy_3D)
    gz = golden("losses")
    pred = torch.from_numpy(gz["pred"]).to(DEV)
    lab = torch.from_numpy(gz["lab"]).to(DEV)
    onehot = torch.stack([(lab == 0), (lab == 1)], 1).float()
    assert abs(cross_entropy_3D(pred, lab).item() - float(gz["cross_entropy_3D"])) < 2e-6
    assert abs(DiceLossss(2)(pred, lab, softmax=True).item() - float(gz["DiceLossss_softmax"])) < 2e-6
    assert abs(DiceLoss()(pred, onehot).item() - float(gz["DiceLoss"])) < 2e-6
    assert abs(Binary_Loss()(pred, onehot).item() - float(gz["BCEWithLogits"])) < 2e-6
    both = float(gz["cross_entropy_3D"]) + float(gz["DiceLossss_softmax"])
    assert abs(DiceCELoss(2)(pred, lab).item() - both) < 4e-6


def test_nondefault_loss_variants_match_reference_golden(golden):
    """DiceLossss(softmax=False) -- the reference's default --, per-class weights, cross_entropy_3D(weight=, size_average=)
    and BinaryDiceLoss (p, smooth, reduction) through the fused kernels, values and gradients against the reference's own
    outputs (tests/golden/losses.npz, losses_extra.npz; loss_function.py:8-16, 61-99, 172-184)."""
    from b200seg.utils.loss_function import BinaryDiceLoss, DiceLossss, cross_entropy_3D
    g0, g = golden("losses"), golden("losses_extra")
    lab = torch.from_numpy(g["lab"]).to(DEV)
    onehot = torch.stack([(lab == 0), (lab == 1)], 1).float()
    cw, dw = g["ce_weight"].tolist(), g["dice_weight"].tolist()
    cases = [
        (g0, "DiceLossss_raw", lambda p: DiceLossss(2)(p, lab, softmax=False)),
        (g0, "BinaryDiceLoss", lambda p: BinaryDiceLoss()(torch.sigmoid(p[:, 1]), onehot[:, 1])),
        (g, "cross_entropy_3D_weighted", lambda p: cross_entropy_3D(p, lab, weight=torch.tensor(cw))),
        (g, "cross_entropy_3D_sum", lambda p: cross_entropy_3D(p, lab, size_average=False)),
        (g, "DiceLossss_softmax_weighted", lambda p: DiceLossss(2)(p, lab, weight=dw, softmax=True)),
        (g, "DiceLossss_raw_weighted", lambda p: DiceLossss(2)(p, lab, weight=dw, softmax=False)),
        (g, "DiceLossss_raw_on_probs", lambda p: DiceLossss(2)(torch.softmax(p, 1), lab, softmax=False)),
        (g, "BinaryDiceLoss_sum", lambda p: BinaryDiceLoss(reduction="sum")(torch.sigmoid(p[:, 1]), onehot[:, 1])),
        (g, "BinaryDiceLoss_p1_smooth", lambda p: BinaryDiceLoss(smooth=0.5, p=1)(torch.sigmoid(p), onehot)),
        (g, "BinaryDiceLoss_p3", lambda p: BinaryDiceLoss(p=3)(torch.sigmoid(p), onehot)),
    ]
    for gz, name, fn in cases:
        p = torch.from_numpy(gz["pred"]).to(DEV).requires_grad_(True)
        v = fn(p)
        v.backward()
        want = float(gz[name])
        assert abs(v.item() - want) < 2e-6 * max(1.0, abs(want)), (name, v.item(), want)
        assert torch.allclose(p.grad.cpu(), torch.from_numpy(gz[name + ".grad"]), rtol=3e-4, atol=1e-8), name
    none = BinaryDiceLoss(reduction="none")(torch.sigmoid(torch.from_numpy(g["pred"]).to(DEV)[:, 1]), onehot[:, 1])
    assert torch.allclose(none.cpu(), torch.from_numpy(g["BinaryDiceLoss_none"]), rtol=1e-6)
    with pytest.raises(Exception):
        BinaryDiceLoss(reduction="bogus")(torch.sigmoid(p[:, 1]), onehot[:, 1])
    # more than two classes takes the generic (non two-class) kernels: compare with the oracle on the CPU
    gen = torch.Generator().manual_seed(4)
    pr = torch.randn(2, 3, 6, 5, 7, generator=gen)
    lb = torch.randint(0, 3, (2, 6, 5, 7), generator=gen)
    w3 = [0.5, 1.0, 2.0]
    for fn_o, fn_d in (
            (lambda p: olosses.cross_entropy_3d(p, lb, weight=w3), lambda p: cross_entropy_3D(p, lb.to(DEV), weight=w3)),
            (lambda p: olosses.dice_loss_per_class(p, lb, 3, softmax=True, weight=w3),
             lambda p: DiceLossss(3)(p, lb.to(DEV), weight=w3, softmax=True)),
            (lambda p: olosses.dice_loss_per_class(torch.softmax(p, 1), lb, 3, weight=w3),
             lambda p: DiceLossss(3)(torch.softmax(p, 1), lb.to(DEV), weight=w3))):
        po = pr.clone().requires_grad_(True)
        vo = fn_o(po)
        vo.backward()
        pd = pr.to(DEV).requires_grad_(True)
        vd = fn_d(pd)
        vd.backward()
        assert abs(vd.item() - vo.item()) < 3e-6 and torch.allclose(pd.grad.cpu(), po.grad, rtol=3e-4, atol=1e-8)


def test_argmax_and_metric_bit_exact(F, golden):
    gz = golden("pool_argmax")
    lab = F.argmax_labels(torch.from_numpy(gz["logits"]).to(DEV))
    assert np.array_equal(lab.cpu().numpy().astype(np.int64), gz["argmax"])
    gm = golden("metric")
    from b200seg.utils.metric import metric
    for i in (0, 1):
        gt, pred = torch.from_numpy(gm["gt%d" % i]), torch.from_numpy(gm["pred%d" % i])
        j, d = metric(gt.to(DEV), pred.to(DEV))
        assert j == float(gm["jaccard%d" % i]) and d == float(gm["dice%d" % i])
        c = F.seg_counts(gt.to(DEV), pred.to(DEV)).tolist()
        oc = ometric.counts(gt.numpy(), pred.numpy())
        assert c == [oc["gt_sum"], oc["pred_sum"], oc["intersection"], oc["union"]]
    # a large ragged size exercises the vector body + scalar tail
    g = torch.Generator().manual_seed(1)
    a = (torch.rand(1000003, generator=g) > 0.5).to(torch.uint8)
    b = (torch.rand(1000003, generator=g) > 0.3).to(torch.uint8)
    c = F.seg_counts(a.to(DEV), b.to(DEV)).tolist()
    assert c == [int(a.sum()), int(b.sum()), int((a & b).sum()), int((a | b).sum())]


def _bf16_storage_floor(sd, x, lab):
    """Per-parameter gradient error of the fp32 oracle when only its STORAGE is bf16 (oracle/unet3d.py storage="bf16").
    At random init the Dice+CE gradient is dominated by a common-mode component that BatchNorm's backward removes, so
    bf16 rounding of stored gradients is amplified layer by layer (0.05 at decoder1, ~0.4 at the encoder); this is a
    property of bf16 storage, not of a kernel, and it is the yardstick for the end-to-end gradient tolerance."""
    x, lab = torch.as_tensor(x), torch.as_tensor(lab)
    grads = []
    for storage in ("fp32", "bf16"):
        s = {k: v.clone().float().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
        out = ounet.forward(s, x, training=True, storage=storage)
        olosses.dice_ce(out, lab).backward()
        grads.append({k: v.grad for k, v in s.items() if v.requires_grad})
    return {k: rel_err(grads[1][k], grads[0][k]) for k in grads[0]}


def test_unet_against_reference_golden(golden):
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.utils.loss_function import DiceCELoss
    gz = golden("unet_f4_s32_b2")
    sd = {k[4:]: torch.from_numpy(gz[k]) for k in gz.files if k.startswith("sd0.")}
    net = UNet3D(1, 2, 4).to(DEV)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net.train()
    x, lab = torch.from_numpy(gz["x"]).to(DEV), torch.from_numpy(gz["lab"]).to(DEV)
    out = net(x)
    ref = torch.from_numpy(gz["out_train"])
    assert out.shape == ref.shape and out.dtype == torch.float32
    close(out.detach().cpu(), ref, 5e-2, "train-mode logits")
    loss = DiceCELoss(2)(out, lab)
    assert abs(loss.item() - float(gz["loss"])) < 5e-3            # Dice within 1e-4 is checked on equal inputs below
    loss.backward()
    floor = _bf16_storage_floor(sd, gz["x"], gz["lab"])
    for name, p in net.named_parameters():
        r = torch.from_numpy(gz["grad." + name])
        if r.abs().max() < 1e-6:                                   # conv biases in front of BN: analytically zero
            assert p.grad.abs().max().item() < 1e-4, name
            continue
        e = rel_err(p.grad.cpu(), r)
        assert e < 2.0 * floor[name] + 0.05, "%s err %.3f vs bf16-storage floor %.3f" % (name, e, floor[name])
    for k in gz.files:
        if k.startswith("sd1."):
            close(net.state_dict()[k[4:]].cpu(), torch.from_numpy(gz[k]), 2e-2, k)
    net.eval()
    with torch.no_grad():
        oe = net(x)
    close(oe.cpu(), torch.from_numpy(gz["out_eval"]), 4e-2, "eval logits")
    import b200seg.functional as F
    agree = (F.argmax_labels(oe).cpu().numpy() == gz["argmax_eval"]).mean()
    assert agree > 0.99, agree


def test_unet_f32_against_live_oracle():
    """Tensor-core-eligible widths (32..512 channels), compared with the oracle evaluated on the box's CPU."""
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.utils.loss_function import DiceCELoss
    torch.manual_seed(0)
    sd = ounet.init_state_dict(1, 2, 32, seed=0)
    net = UNet3D(1, 2, 32).to(DEV)
    net.load_state_dict(sd)
    net.train()
    x = torch.randn(1, 1, 32, 32, 32)
    lab = (torch.rand(1, 32, 32, 32) > 0.8).long()
    out = net(x.to(DEV))
    loss = DiceCELoss(2)(out, lab.to(DEV))
    loss.backward()
    rsd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    rout = ounet.forward(rsd, x, training=True)
    rloss = olosses.dice_ce(rout, lab)
    rloss.backward()
    close(out.detach().cpu(), rout.detach(), 6e-2, "logits")
    assert abs(loss.item() - rloss.item()) < 1e-2
    floor = _bf16_storage_floor(sd, x.numpy(), lab.numpy())
    for name, p in net.named_parameters():
        r = rsd[name].grad
        if r.abs().max() < 1e-6:
            continue
        e = rel_err(p.grad.cpu(), r)
        assert e < 2.0 * floor[name] + 0.05, "%s err %.3f vs bf16-storage floor %.3f" % (name, e, floor[name])


def test_sliding_window_aggregator_bit_exact():
    from b200seg.inference import GridAggregator, GridSampler
    rng = np.random.default_rng(0)
    vol = rng.integers(0, 3, size=(1, 40, 36, 50)).astype(np.int64)
    for mode in ("crop", "average"):
        sampler = GridSampler(vol.shape[1:], (16, 16, 16), (4, 4, 6))
        assert np.array_equal(sampler.locations.numpy(), owindow.grid_locations(vol.shape[1:], (16,) * 3, (4, 4, 6)))
        agg = GridAggregator(sampler, overlap_mode=mode, device=DEV)
        oagg = owindow.Aggregator(vol.shape[1:], (4, 4, 6), mode)
        locs = sampler.locations
        for s in range(0, len(locs), 7):
            lb = locs[s:s + 7]
            patches = np.stack([vol[:, a:d, b:e, c:f] for a, b, c, d, e, f in lb.numpy()])
            agg.add_batch(torch.from_numpy(patches).to(DEV), lb)
            oagg.add_batch(patches, lb.numpy())
        out = agg.get_output_tensor().cpu().numpy()
        assert np.array_equal(out.astype(np.int64), oagg.get_output_tensor().astype(np.int64))
        assert np.array_equal(out.astype(np.int64), vol)
    # patches that DISAGREE where their cropped interiors overlap: the later patch must win, whatever the batching and
    # the order in which a rank's share is added (torchio semantics, restated by the oracle)
    shape, patch, ov = (40, 36, 50), (16, 16, 16), (4, 4, 6)
    sampler = GridSampler(shape, patch, ov)
    locs = sampler.locations
    patches = torch.from_numpy(rng.integers(0, 5, size=(len(locs), 1) + patch).astype(np.uint8))
    oagg = owindow.Aggregator(shape, ov, "crop")
    oagg.add_batch(patches.numpy(), locs.numpy())
    want = oagg.get_output_tensor().astype(np.int64)
    agg = GridAggregator(sampler, "crop", device=DEV)
    agg.add_batch(patches.to(DEV), locs)                         # one batch: overlapping writers inside one launch
    assert np.array_equal(agg.get_output_tensor().cpu().numpy().astype(np.int64), want)
    agg = GridAggregator(sampler, "crop", device=DEV)
    perm = torch.from_numpy(rng.permutation(len(locs)))
    for s in range(0, len(locs), 5):                             # shuffled shares with explicit patch ids
        ids = perm[s:s + 5]
        agg.add_batch(patches[ids].to(DEV), locs[ids], patch_ids=ids)
    assert np.array_equal(agg.get_output_tensor().cpu().numpy().astype(np.int64), want)


@pytest.mark.gpu
def test_graph_captured_train_step_matches_eager():
    """engine.TrainStep: the CUDA-graph replay of (zero_grad, forward, Dice+CE, backward, fused Adam) follows the same
    trajectory as enqueueing every kernel eagerly (train.py:187-214).  Differences come only from the order of
    floating-point atomics in the statistics / weight-gradient reductions."""
    from b200seg.engine import TrainStep
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.optim import FusedAdam
    from b200seg.utils.loss_function import DiceCELoss
    from oracle import unet3d as ounet
    dev = torch.device("cuda")
    sd = ounet.init_state_dict(1, 2, 16, seed=3)
    torch.manual_seed(5)
    xs = [torch.randn(2, 1, 32, 32, 32, device=dev) for _ in range(6)]
    labs = [(torch.rand(2, 32, 32, 32, device=dev) > 0.8).to(torch.uint8) for _ in range(6)]
    results = []
    for use_graph in (False, True):
        net = UNet3D(1, 2, 16).to(dev)
        net.load_state_dict(sd)
        net.train()
        opt = FusedAdam(net.parameters(), lr=1e-3)
        step = TrainStep(net, DiceCELoss(2), opt, use_graph=use_graph, warmup=2)
        losses = [float(step(x, y)[0]) for x, y in zip(xs, labs)]
        assert (step.graph is not None) == use_graph
        if use_graph:
            assert step.kernels_per_step > 100 and step.umma_per_step > 20
        assert int(opt.state_dict()["state"][0]["step"]) == 6
        results.append((losses, {k: v.clone() for k, v in net.state_dict().items()}))
    (l0, p0), (l1, p1) = results
    assert max(abs(a - b) for a, b in zip(l0, l1)) < 6e-3, (l0, l1)
    assert p0["encoder1.enc1norm1.num_batches_tracked"] == p1["encoder1.enc1norm1.num_batches_tracked"] == 6
    # Adam turns the run-to-run noise of the atomically reduced gradients into +-lr steps wherever a gradient is near
    # zero, so compare the well-conditioned full-resolution layers tightly and the 2^3-voxel bottleneck only coarsely
    for k, tol in (("encoder1.enc1conv1.weight", 2e-2), ("decoder1.dec1conv2.weight", 2e-2), ("conv.weight", 2e-2),
                   ("encoder2.enc2norm1.running_var", 2e-2), ("bottleneck.bottleneckconv1.weight", 0.3)):
        close(p1[k].float().cpu(), p0[k].float().cpu(), tol, k)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,classes", [(32, 2), (16, 2), (64, 1), (24, 3), (256, 2), (128, 4)])
def test_head_conv1x1_forward_backward(F, cin, classes):
    """1x1x1 class head (unet3d.py:46-48; residual_unet3d.py:60-66 deep-supervision heads up to 256 channels): fp32 NCDHW
    logits out, gradients w.r.t. features / weight / bias against torch conv3d on the same bf16 features."""
    g = torch.Generator().manual_seed(cin * 10 + classes)
    x = bf(torch.randn(2, cin, 6, 10, 12, generator=g))
    w = torch.randn(classes, cin, 1, 1, 1, generator=g) * 0.2
    b = torch.randn(classes, generator=g) * 0.1
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.conv3d(xr, wr, br)
    dl = torch.randn(ref.shape, generator=g)
    ref.backward(dl)
    xd = ndhwc(x).requires_grad_(True)
    wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    out = F.head_conv1x1(xd, wd, bd)
    assert out.dtype == torch.float32 and tuple(out.shape) == tuple(ref.shape)
    close(out.cpu(), ref.detach(), 1e-5, "logits")
    out.backward(dl.to(DEV))
    close(ncdhw(xd.grad), xr.grad, 8e-3, "dx")       # dx is stored in bf16
    close(wd.grad.cpu(), wr.grad, 1e-4, "dw")
    close(bd.grad.cpu(), br.grad, 1e-4, "db")


@pytest.mark.gpu
def test_fused_optimizer_step_is_bit_identical_to_the_three_kernel_path():
    """b200seg_adam_step_fused (packed weight gradients + Adam + both bf16 packs in one launch) against the separate
    unpack -> adam_step_dev -> pack_weights_batched launches: bit-identical parameters, moments and packs over several steps,
    for full 32 x 32 bricks, ragged bricks, k^3 in {1, 8, 27, 125} and plain ranges; and against torch.optim.Adam
    (train.py:109) within fp32 rounding."""
    import os
    from b200seg.optim import FusedAdam
    dev = torch.device("cuda")
    for shapes in ([(64, 32, 3, 3, 3), (64,), (32, 1, 3, 3, 3), (48, 40, 3, 3, 3), (48,), (64, 32, 2, 2, 2), (2, 32, 1, 1, 1),
                    (2,), (96, 64, 3, 3, 3), (7,)],
                   [(16, 16, 5, 5, 5), (16,), (32, 16, 2, 2, 2), (33, 9, 3, 3, 3), (64, 32, 5, 5, 5), (40, 24, 5, 5, 5),
                    (64, 16, 3, 3, 3)]):
        runs = []
        for fused in (True, False):
            g = torch.Generator(device=dev).manual_seed(11)
            params = [torch.nn.Parameter(torch.randn(*s, device=dev, generator=g) * 0.1) for s in shapes]
            ref = [p.detach().clone().requires_grad_(True) for p in params]
            os.environ.pop("B200SEG_DISABLE_FUSED_ADAM", None)
            if not fused:
                os.environ["B200SEG_DISABLE_FUSED_ADAM"] = "1"
            try:
                opt = FusedAdam(params, lr=1e-2, weight_decay=0.01)
                topt = torch.optim.Adam(ref, lr=1e-2, weight_decay=0.01)
                for it in range(3):
                    opt.zero_grad()
                    for p, r in zip(params, ref):
                        gr = torch.randn(p.shape, device=dev, generator=g)
                        r.grad = gr.clone()
                        if p.dim() == 5 and it != 1:
                            # what the weight-gradient kernels leave behind: [tap][C_in][C_out] in the packed accumulator
                            a, b = p.shape[0], p.shape[1]
                            p._b200_dwp.copy_(gr.reshape(a, b, -1).permute(2, 1, 0).reshape(-1))
                            p._b200_pending[0] = True
                        else:
                            p.grad.copy_(gr)     # the autograd-accumulated form (it == 1: also for conv weights)
                    opt.step()
                    topt.step()
                torch.cuda.synchronize()
            finally:
                os.environ.pop("B200SEG_DISABLE_FUSED_ADAM", None)
            for p, r in zip(params, ref):
                close(p.detach().float().cpu(), r.detach().float().cpu(), 1e-5, "adam vs torch %s" % (tuple(p.shape),))
            runs.append((opt.param_arena.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt._pack_arena.clone(),
                         int(opt.state_dict()["state"][0]["step"])))
            # the packs are the bf16 transposes of the updated weights
            for p, o, n in opt._packed:
                a, b = p.shape[0], p.shape[1]
                w = p.detach().reshape(a, b, -1)
                assert torch.equal(opt._pack_arena[o:o + n].view(-1, a, b), w.permute(2, 0, 1).bfloat16())
                assert torch.equal(opt._pack_arena[o + n:o + 2 * n].view(-1, b, a), w.flip(2).permute(2, 1, 0).bfloat16())
        for x, y in zip(runs[0][:4], runs[1][:4]):
            assert torch.equal(x, y)
        assert runs[0][4] == runs[1][4] == 3


# ---- attention gates of ER-Net / RE-Net / Double-UNet (csrc/gates.cu) against the torch restatement -------------------
def test_reverse_attention_gate_forward_backward(F):
    """`-1 * sigmoid(up(proj(coarse))) + 1` expanded over the channels, times the skip, plus the skip (ER_net.py:184-187)."""
    from oracle import backend_torch as B
    g = torch.Generator().manual_seed(21)
    for c, cc, size in ((32, 64, (4, 6, 8)), (64, 128, (3, 4, 4)), (128, 256, (2, 2, 4))):
        fine = bf(torch.randn(2, c, 2 * size[0], 2 * size[1], 2 * size[2], generator=g))
        coarse = bf(torch.randn(2, cc, *size, generator=g))
        wp = torch.randn(1, cc, 1, 1, 1, generator=g) * 0.2
        bp = torch.randn(1, generator=g) * 0.1
        wt = torch.randn(1, 1, 2, 2, 2, generator=g)
        bt = torch.randn(1, generator=g) * 0.1
        dy = bf(torch.randn(fine.shape, generator=g))
        # reference arithmetic
        r = [t.clone().requires_grad_(True) for t in (fine, coarse, wp, bp, wt, bt)]
        gmap = B.convt_map_k2s2(torch.nn.functional.conv3d(r[1], r[2], r[3]), r[4], r[5])
        ref = B.reverse_gate(r[0], gmap)
        ref.backward(dy)
        # kernels
        d = [ndhwc(fine).requires_grad_(True), ndhwc(coarse).requires_grad_(True)] + \
            [t.to(DEV).requires_grad_(True) for t in (wp, bp, wt, bt)]
        gm = F.convt_map_k2s2(F.head_conv1x1(d[1], d[2], d[3]), d[4], d[5])
        close(gm.detach().cpu(), gmap.detach(), 5e-3, "gate map")
        out = F.reverse_gate(d[0], gm)
        close(ncdhw(out), ref.detach(), 6e-3, "reverse gate")
        out.backward(ndhwc(dy))
        close(ncdhw(d[0].grad), r[0].grad, 8e-3, "d fine")
        close(ncdhw(d[1].grad), r[1].grad, 1.5e-2, "d coarse")
        for i, what in ((2, "d proj weight"), (3, "d proj bias"), (4, "d up weight"), (5, "d up bias")):
            close(d[i].grad.cpu(), r[i].grad, 1.5e-2, what)


@pytest.mark.parametrize("kind", ["se_residual", "se_inception", "selective_fusion"])
def test_gated_blend_forward_backward(F, kind):
    """SE.py:4-49 (x * y, x + x * y) and SFConv (ER_net.py:57-105) through the mirrors' own modules, bound once to the
    kernels and once to the torch restatement."""
    from oracle import backend_torch as B
    from b200seg.models.three_d.ER_net import SFConv
    from b200seg.models.three_d.SE import SE_Inception, SE_Residual
    g = torch.Generator().manual_seed(22)
    c, shape = 64, (2, 64, 6, 5, 8)
    torch.manual_seed(5)
    mod = {"se_residual": lambda: SE_Residual(c), "se_inception": lambda: SE_Inception(c),
           "selective_fusion": lambda: SFConv(c)}[kind]()
    x1 = bf(torch.randn(*shape, generator=g) + 0.3)
    x2 = bf(torch.randn(*shape, generator=g))
    dy = bf(torch.randn(*shape, generator=g))
    two = kind == "selective_fusion"
    mod.set_kernels(B)
    a1, a2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    ref = mod(a1, a2) if two else mod(a1)
    ref.backward(dy)
    ref_grads = {k: p.grad.clone() for k, p in mod.named_parameters()}
    mod.zero_grad()
    mod.set_kernels(None)
    mod = mod.to(DEV)
    d1, d2 = ndhwc(x1).requires_grad_(True), ndhwc(x2).requires_grad_(True)
    out = mod(d1, d2) if two else mod(d1)
    close(ncdhw(out), ref.detach(), 6e-3, kind)
    out.backward(ndhwc(dy))
    close(ncdhw(d1.grad), a1.grad, 8e-3, "dx1")
    if two:
        close(ncdhw(d2.grad), a2.grad, 8e-3, "dx2")
    for k, p in mod.named_parameters():
        close(p.grad.cpu(), ref_grads[k], 2e-2, "d " + k)


def test_sigmoid_map_and_concat_input(F):
    g = torch.Generator().manual_seed(23)
    x = torch.randn(2, 2, 5, 6, 7, generator=g) * 3
    xd = x.to(DEV).requires_grad_(True)
    y = F.sigmoid_map(xd)
    assert float((y.cpu() - torch.sigmoid(x)).abs().max()) < 1e-6
    dy = torch.randn(x.shape, generator=g)
    y.backward(dy.to(DEV))
    s = torch.sigmoid(x)
    assert float((xd.grad.cpu() - dy * s * (1 - s)).abs().max()) < 1e-6
    # torch.cat((image, class scores), 1) -> channels-last bf16 (Double_Unet.py:90), gradient back to the class scores
    img = torch.randn(2, 1, 5, 6, 7, generator=g)
    maps = torch.randn(2, 2, 5, 6, 7, generator=g).to(DEV).requires_grad_(True)
    cat = F.concat_input(img.to(DEV), maps)
    assert torch.equal(ncdhw(cat), bf(torch.cat((img, maps.detach().cpu()), 1)))
    gcat = bf(torch.randn(2, 3, 5, 6, 7, generator=g))
    cat.backward(ndhwc(gcat))
    assert torch.equal(maps.grad.cpu(), gcat[:, 1:])


def test_frozen_parameters_reuse_packs_and_constants(F):
    """Inside functional.frozen_parameters() (sliding-window inference: many batches, fixed weights) a layer's weight pack and
    eval-mode constants are computed once; results are identical, and nothing is cached outside the context or in grad mode."""
    g = torch.Generator().manual_seed(31)
    x = ndhwc(bf(torch.randn(2, 32, 8, 16, 8, generator=g)))
    w = (torch.randn(32, 32, 3, 3, 3, generator=g) * 0.05).to(DEV)
    b = (torch.randn(32, generator=g) * 0.1).to(DEV)
    gamma, beta = (torch.rand(32, generator=g) + 0.5).to(DEV), (torch.randn(32, generator=g) * 0.3).to(DEV)
    rm, rv = (torch.randn(32, generator=g) * 0.2).to(DEV), (torch.rand(32, generator=g) + 0.5).to(DEV)
    args = dict(k=3, stride=1, pad=1, dil=1, spec=F.NormSpec("batch", "relu", training=False), gamma=gamma, beta=beta,
                running_mean=rm, running_var=rv)
    with torch.no_grad():
        k0 = F.launches()
        ref = F.conv_norm_act(x, w, b, **args)
        plain = F.launches() - k0
        with F.frozen_parameters():
            first = F.conv_norm_act(x, w, b, **args)
            k1 = F.launches()
            second = F.conv_norm_act(x, w, b, **args)
            cached = F.launches() - k1
        k2 = F.launches()
        after = F.conv_norm_act(x, w, b, **args)
        assert F.launches() - k2 == plain
    assert plain == 3 and cached == 1, (plain, cached)
    assert torch.equal(ref, first) and torch.equal(ref, second) and torch.equal(ref, after)
    assert F._FROZEN[0] is None


def test_finalisation_in_the_normalise_prologue_is_bit_identical(F, monkeypatch):
    """Training-mode BatchNorm: the constants derived in the prologue of the normalise kernel (b200seg_norm_act_fwd_stats)
    equal the separate b200seg_norm_finalize launch bit for bit -- output, saved constants, running statistics."""
    g = torch.Generator().manual_seed(41)
    for c, shape in ((32, (2, 8, 16, 8)), (512, (2, 4, 4, 4)), (24, (1, 6, 10, 8))):
        y = ndhwc(bf(torch.randn(shape[0], c, *shape[1:], generator=g) * 2 + 0.5))
        gamma, beta = (torch.rand(c, generator=g) + 0.5).to(DEV), (torch.randn(c, generator=g) * 0.3).to(DEV)
        outs = []
        stats = F.channel_stats(y, 1, spare=1)       # one statistics vector for both runs (its atomics are not ordered)
        for separate in (False, True):
            if separate:
                monkeypatch.setenv("B200SEG_SEPARATE_FINALIZE", "1")
            else:
                monkeypatch.delenv("B200SEG_SEPARATE_FINALIZE", raising=False)
            rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
            k0 = F.launches()
            z, coef, count, groups = F._norm_forward(y, stats.clone(), F.NormSpec("batch", "relu", training=True), gamma, beta,
                                                     rm, rv, None, None, None)
            outs.append((z.clone(), coef.clone(), rm.clone(), rv.clone(), F.launches() - k0))
        monkeypatch.delenv("B200SEG_SEPARATE_FINALIZE", raising=False)
        for a, b in zip(outs[0][:4], outs[1][:4]):
            assert torch.equal(a, b)
        assert outs[0][4] == outs[1][4] - 1      # one launch fewer
