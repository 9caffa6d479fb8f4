"""GPU: size-independent properties at BASELINE.json's FULL sizes (2 x 1 x 128^3 training batch; 512 x 512 x 256 sliding-window
volume), where the CPU oracle is too slow to be the checker:
  * independent kernels agree: the narrow-output rolling-accumulator kernel, the generic persistent plane kernel and the
    first-generation kernel compute the same full-resolution convolution (forward, data gradient), and the persistent
    weight-gradient kernel matches the first-generation one;
  * linearity of the convolution in its weights;
  * per-channel statistics fused in the conv epilogue equal a separate reduction of the stored output;
  * samples of a batch do not influence each other in eval mode (bit-exact);
  * max-pool backward routes every gradient to exactly one voxel (sum preserved, bit-exact in fp32 on integers);
  * sliding-window stitching of patches cut from a volume reproduces the volume bit-exactly (crop and average modes),
    and Dice/IoU of a label map with itself is 1."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_full_resolution_conv_kernels_agree():
    import b200seg.functional as F
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(2, 128, 128, 128, 32, device=DEV, generator=g).bfloat16()
    w = torch.randn(32, 32, 3, 3, 3, device=DEV, generator=g) * 0.05
    b = torch.randn(32, device=DEV, generator=g) * 0.1
    dy = torch.randn(2, 128, 128, 128, 32, device=DEV, generator=g).bfloat16()
    y_roll, st_roll, geom = F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, True)
    dx_roll = F.conv3d_dgrad_raw(geom, dy, w)
    dw_new = F.conv3d_wgrad_raw(geom, x, dy, w.shape)
    with env(B200SEG_DISABLE_ROLL="1"):
        y_plane, st_plane, _ = F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, True)
        dx_plane = F.conv3d_dgrad_raw(geom, dy, w)
    with env(B200SEG_DISABLE_ROLL="1", B200SEG_DISABLE_PERSISTENT="1"):
        y_old, _, _ = F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, False)
        dw_old = F.conv3d_wgrad_raw(geom, x, dy, w.shape)
    # same bf16 inputs, fp32 accumulation in a different order, one bf16 rounding at the end
    assert rel(y_roll, y_plane) < 3e-3 and rel(y_roll, y_old) < 3e-3
    assert float((y_roll.float() - y_plane.float()).abs().max()) <= 2 ** -6 * float(y_plane.float().abs().max())
    assert rel(dx_roll, dx_plane) < 3e-3
    assert rel(dw_new, dw_old) < 2e-3
    # fused statistics == separate reduction over the stored (bf16-rounded) output
    c = 32
    sep = F.channel_stats(y_roll)[0]
    assert rel(st_roll[:c], sep[0]) < 2e-3 and rel(st_roll[c:2 * c], sep[1]) < 2e-3
    assert rel(st_roll[:2 * c], st_plane[:2 * c]) < 1e-4
    # linearity in the weights
    w2 = torch.randn(32, 32, 3, 3, 3, device=DEV, generator=g) * 0.05
    ya, _, _ = F.conv3d_fprop_raw(x, w.bfloat16().float(), None, 3, 1, 1, 1, False)
    yb, _, _ = F.conv3d_fprop_raw(x, w2.bfloat16().float(), None, 3, 1, 1, 1, False)
    yab, _, _ = F.conv3d_fprop_raw(x, (w.bfloat16().float() + w2.bfloat16().float()), None, 3, 1, 1, 1, False)
    assert rel(yab, ya.float() + yb.float()) < 8e-3


def test_wide_output_roll_halves_agree_with_plane_kernel_and_torch():
    """C_out = 64 goes through the rolling-accumulator kernel as two 32-channel halves (CTA parity): compare with the
    generic plane kernel at 2 x 64^3 (U-Net level 2), with torch fp32 at a small size, and check the column sums the
    dgrad epilogue publishes (the up-convolution's bias gradient) against a separate reduction."""
    import b200seg.functional as F
    g = torch.Generator(device=DEV).manual_seed(1)
    for cin, cout, s in ((32, 64, 64), (64, 64, 64), (64, 64, 24)):
        x = torch.randn(2, s, s, s, cin, device=DEV, generator=g).bfloat16()
        w = torch.randn(cout, cin, 3, 3, 3, device=DEV, generator=g) * 0.05
        b = torch.randn(cout, device=DEV, generator=g) * 0.1
        dy = torch.randn(2, s, s, s, cout, device=DEV, generator=g).bfloat16()
        y_r, st_r, geom = F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, True)
        dx_r, cs_r = F.conv3d_dgrad_raw(geom, dy, w, colsum=True)
        # dgrad of a 64 -> 32 conv produces a 64-channel gradient: the halves path on the dgrad side
        w_t = torch.randn(32, 64, 3, 3, 3, device=DEV, generator=g) * 0.05
        dy32 = torch.randn(2, s, s, s, 32, device=DEV, generator=g).bfloat16()
        x64 = torch.randn(2, s, s, s, 64, device=DEV, generator=g).bfloat16()
        _, _, geom_t = F.conv3d_fprop_raw(x64, w_t, None, 3, 1, 1, 1, False)
        dx64_r, cs64_r = F.conv3d_dgrad_raw(geom_t, dy32, w_t, colsum=True)
        with env(B200SEG_DISABLE_ROLL_HALVES="1"):
            y_p, st_p, _ = F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, True)
            dx_p = F.conv3d_dgrad_raw(geom, dy, w)
            dx64_p = F.conv3d_dgrad_raw(geom_t, dy32, w_t)
        assert rel(y_r, y_p) < 3e-3 and rel(dx_r, dx_p) < 3e-3 and rel(dx64_r, dx64_p) < 3e-3, (cin, cout, s)
        assert rel(st_r[:2 * cout], st_p[:2 * cout]) < 1e-4
        sep = F.channel_stats(dx64_r)[0]
        assert rel(cs64_r[:64], sep[0]) < 5e-3 and rel(cs64_r[64:], sep[1]) < 5e-3
        sep = F.channel_stats(dx_r)[0]
        assert rel(cs_r[:cin], sep[0]) < 5e-3
        if s == 24:
            ref = torch.nn.functional.conv3d(x.float().permute(0, 4, 1, 2, 3), w.bfloat16().float(), b, padding=1)
            assert rel(y_r.float().permute(0, 4, 1, 2, 3), ref) < 6e-3
            refdx = torch.nn.functional.conv_transpose3d(dy.float().permute(0, 4, 1, 2, 3), w.bfloat16().float(), padding=1)
            assert rel(dx_r.float().permute(0, 4, 1, 2, 3), refdx) < 6e-3


def test_stem_weight_gradient_tensor_core_path():
    """C_in = 1 stem weight gradient (unet3d.py:80): taps-as-channels im2col + the tcgen05 1x1x1 weight-gradient kernel
    against the CUDA-core stem kernel at the full 2 x 128^3 size and against torch fp32 at 2 x 40^3 (ragged tiles:
    40 is not a multiple of the 16 x 8 tile)."""
    import b200seg.functional as F
    g = torch.Generator(device=DEV).manual_seed(2)
    for s, cout in ((128, 32), (40, 32), (40, 16)):
        x = torch.randn(2, s, s, s, 1, device=DEV, generator=g).bfloat16()
        w = torch.randn(cout, 1, 3, 3, 3, device=DEV, generator=g) * 0.2
        dy = torch.randn(2, s, s, s, cout, device=DEV, generator=g).bfloat16()
        _, _, geom = F.conv3d_fprop_raw(x, w, None, 3, 1, 1, 1, False)
        k0 = F.umma_launch_count()
        dw_tc = F.conv3d_wgrad_raw(geom, x, dy, w.shape)
        assert F.umma_launch_count() == k0 + 1, "stem weight gradient did not take the tensor-core path"
        with env(B200SEG_DISABLE_STEM_IM2COL="1"):
            dw_cc = F.conv3d_wgrad_raw(geom, x, dy, w.shape)
        assert rel(dw_tc, dw_cc) < 2e-3, (s, cout, rel(dw_tc, dw_cc))
        if s == 40:
            xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(False)
            wr = w.clone().requires_grad_(True)
            torch.nn.functional.conv3d(xr, wr, None, padding=1).backward(dy.float().permute(0, 4, 1, 2, 3))
            assert rel(dw_tc, wr.grad) < 2e-3, (s, cout, rel(dw_tc, wr.grad))


def test_full_size_unet_sample_independence_and_pool_routing():
    import b200seg.functional as F
    from b200seg.models.three_d.unet3d import UNet3D
    torch.manual_seed(0)
    net = UNet3D(1, 2, 32).to(DEV).eval()
    a, b, c = (torch.randn(1, 1, 128, 128, 128, device=DEV) for _ in range(3))
    with torch.no_grad():
        o1 = net(torch.cat((a, b)))
        o2 = net(torch.cat((a, c)))
    assert torch.equal(o1[0], o2[0]) and not torch.equal(o1[1], o2[1])
    lab = F.argmax_labels(torch.cat((a, -a), dim=1))          # a non-trivial binary map of the full 128^3 size
    assert 0 < int(lab.sum()) < lab.numel()
    counts = F.seg_counts(lab, lab).tolist()
    assert counts[0] == counts[1] == counts[2] == counts[3] == int(lab.sum())
    from b200seg.utils.metric import metric
    j, d = metric(lab, lab)
    assert abs(d - 1) < 1e-6 and abs(j - 1) < 1e-6
    j0, d0 = metric(lab, 1 - lab)
    assert j0 == 0.0 and d0 == 0.0
    # max-pool backward: every output gradient lands on exactly one input voxel
    x = torch.randn(2, 128, 128, 128, 32, device=DEV).bfloat16().requires_grad_(True)
    y = F.max_pool2(x)
    gy = torch.randint(-3, 4, y.shape, device=DEV).bfloat16()
    (gx,) = torch.autograd.grad(y, x, gy)
    assert float(gx.float().sum()) == float(gy.float().sum())
    assert int((gx != 0).sum()) == int((gy != 0).sum())
    assert torch.equal(F.max_pool2(x.detach()), torch.nn.functional.max_pool3d(
        x.detach().permute(0, 4, 1, 2, 3).float(), 2, 2).permute(0, 2, 3, 4, 1).bfloat16())


def test_config5_sliding_window_round_trip():
    from b200seg.inference import GridAggregator, GridSampler
    g = torch.Generator(device=DEV).manual_seed(1)
    vol = torch.randint(0, 4, (1, 512, 512, 256), device=DEV, generator=g, dtype=torch.uint8)
    sampler = GridSampler(vol, (128,) * 3, (64,) * 3)
    assert len(sampler) == 147
    for mode in ("crop", "average"):
        agg = GridAggregator(sampler, mode, device=DEV)
        for data, locs in sampler.batches(16):
            agg.add_batch(data if mode == "crop" else data.float(), locs)
        out = agg.get_output_tensor()
        assert torch.equal(out.to(torch.uint8), vol), mode


@pytest.mark.parametrize("cin,cout", [(32, 32), (32, 2), (1, 16)])
def test_full_resolution_5x5x5_kernels_agree(cin, cout):
    """V-Net's 5x5x5 layers at 2 x 128^3 (vnet3d.py:25,47,111): the kd-stacked rolling-accumulator kernel (five taps in N,
    ring of six accumulators over all 128 planes, two 16-channel halves) against the plane kernel that issues one instruction
    per tap -- forward with fused statistics, data gradient -- and the weight gradient's linearity."""
    import b200seg.functional as F
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(2, 128, 128, 128, cin, device=DEV, generator=g).bfloat16()
    w = torch.randn(cout, cin, 5, 5, 5, device=DEV, generator=g) * (2.0 / (125 * cin)) ** 0.5
    b = torch.randn(cout, device=DEV, generator=g) * 0.1
    dy = torch.randn(2, 128, 128, 128, cout, device=DEV, generator=g).bfloat16()
    n0 = F.umma_launch_count()
    y_roll, st_roll, geom = F.conv3d_fprop_raw(x, w, b, 5, 1, 2, 1, True)
    dx_roll = F.conv3d_dgrad_raw(geom, dy, w) if cin > 1 else None
    assert F.umma_launch_count() - n0 == (2 if cin > 1 else 1)
    with env(B200SEG_DISABLE_ROLL5="1"):
        y_plane, st_plane, _ = F.conv3d_fprop_raw(x, w, b, 5, 1, 2, 1, True)
        dx_plane = F.conv3d_dgrad_raw(geom, dy, w) if cin > 1 else None
    assert rel(y_roll, y_plane) < 3e-3
    assert float((y_roll.float() - y_plane.float()).abs().max()) <= 2 ** -6 * float(y_plane.float().abs().max())
    if cin > 1:
        assert rel(dx_roll, dx_plane) < 3e-3
    sep = F.channel_stats(y_roll.contiguous())[0]
    assert rel(st_roll[:cout], sep[0]) < 2e-3 and rel(st_roll[cout:2 * cout], sep[1]) < 2e-3
    assert rel(st_roll[:2 * cout], st_plane[:2 * cout]) < 1e-4
    # the first and last planes see the zero padding: compare one border plane with cuDNN in strict fp32
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xs = x[:1, :6].permute(0, 4, 1, 2, 3).float()
        ref = torch.nn.functional.conv3d(xs, w.bfloat16().float(), b, padding=2)[:, :, :4]
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert rel(y_roll[:1, :4].permute(0, 4, 1, 2, 3), ref) < 6e-3
