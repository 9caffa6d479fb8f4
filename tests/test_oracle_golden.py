"""CPU: the oracle restatements reproduce the reference's own outputs (fixtures from tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import losses, metric, ops, syncbn, unet3d, window


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_unet_forward_backward_matches_reference(golden):
    g = golden("unet_f4_s32_b2")
    sd = {k[4:]: T(g[k]).clone() for k in g.files if k.startswith("sd0.")}
    params = [k for k in sd if sd[k].dtype.is_floating_point and "running" not in k]
    for k in params:
        sd[k].requires_grad_(True)
    new_stats = {}
    out = unet3d.forward(sd, T(g["x"]), training=True, new_stats=new_stats)
    assert torch.allclose(out, T(g["out_train"]), rtol=1e-3, atol=1e-3)  # fp32 re-association; the 1^3 bottleneck normalises over 2 samples
    loss = losses.dice_ce(out, T(g["lab"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    for k in params:
        ref = T(g["grad." + k])
        # fp32 BN-backward cancellation differs between autograd-through-the-formula and torch's fused backward by
        # ~2e-3 of the tensor's max; conv biases in front of a BatchNorm have an analytically zero gradient (noise).
        assert float((sd[k].grad - ref).abs().max()) <= 1e-6 + 5e-3 * float(ref.abs().max()), k
    for k, v in new_stats.items():
        assert torch.allclose(v, T(g["sd1." + k]), rtol=1e-5, atol=1e-6), k
    sd_eval = {k: v.detach() for k, v in sd.items()}
    sd_eval.update(new_stats)
    out_eval = unet3d.forward(sd_eval, T(g["x"]), training=False)
    assert torch.allclose(out_eval, T(g["out_eval"]), rtol=1e-3, atol=1e-3)
    assert np.array_equal(ops.argmax_labels(out_eval.numpy()), g["argmax_eval"])


def test_init_state_dict_has_reference_keys(golden):
    g = golden("unet_f4_s32_b2")
    ref = {k[4:]: g[k].shape for k in g.files if k.startswith("sd0.")}
    mine = {k: tuple(v.shape) for k, v in unet3d.init_state_dict(1, 2, 4).items()}
    assert mine == ref and len(mine) == 136


def test_losses_match_reference(golden):
    g = golden("losses")
    lab = T(g["lab"])
    onehot = torch.stack([(lab == 0), (lab == 1)], 1).float()
    cases = {
        "cross_entropy_3D": lambda p: losses.cross_entropy_3d(p, lab),
        "DiceLoss": lambda p: losses.dice_loss_sigmoid(p, onehot),
        "DiceLossss_softmax": lambda p: losses.dice_loss_per_class(p, lab, 2, softmax=True),
        "DiceLossss_raw": lambda p: losses.dice_loss_per_class(p, lab, 2, softmax=False),
        "BinaryDiceLoss": lambda p: losses.binary_dice_loss(torch.sigmoid(p[:, 1]), onehot[:, 1]),
        "BCEWithLogits": lambda p: losses.bce_with_logits(p, onehot),
    }
    for name, fn in cases.items():
        p = T(g["pred"]).clone().requires_grad_(True)
        v = fn(p)
        v.backward()
        assert abs(float(v) - float(g[name])) < 2e-6, name
        assert torch.allclose(p.grad, T(g[name + ".grad"]), rtol=1e-4, atol=1e-8), name
    # weighted / non-default variants (losses_extra.npz, produced by the same reference functions)
    g = golden("losses_extra")
    cw, dw = g["ce_weight"].tolist(), g["dice_weight"].tolist()
    cases = {
        "cross_entropy_3D_weighted": lambda p: losses.cross_entropy_3d(p, lab, weight=cw),
        "cross_entropy_3D_sum": lambda p: losses.cross_entropy_3d(p, lab, size_average=False),
        "DiceLossss_softmax_weighted": lambda p: losses.dice_loss_per_class(p, lab, 2, softmax=True, weight=dw),
        "DiceLossss_raw_weighted": lambda p: losses.dice_loss_per_class(p, lab, 2, softmax=False, weight=dw),
        "DiceLossss_raw_on_probs": lambda p: losses.dice_loss_per_class(torch.softmax(p, 1), lab, 2, softmax=False),
        "BinaryDiceLoss_sum": lambda p: losses.binary_dice_loss(torch.sigmoid(p[:, 1]), onehot[:, 1], reduction="sum"),
        "BinaryDiceLoss_p1_smooth": lambda p: losses.binary_dice_loss(torch.sigmoid(p), onehot, smooth=0.5, p=1),
        "BinaryDiceLoss_p3": lambda p: losses.binary_dice_loss(torch.sigmoid(p), onehot, p=3),
    }
    for name, fn in cases.items():
        p = T(g["pred"]).clone().requires_grad_(True)
        v = fn(p)
        v.backward()
        assert abs(float(v) - float(g[name])) < 2e-6 * max(1.0, abs(float(g[name]))), name
        assert torch.allclose(p.grad, T(g[name + ".grad"]), rtol=1e-4, atol=1e-8), name
    none = losses.binary_dice_loss(torch.sigmoid(T(g["pred"])[:, 1]), onehot[:, 1], reduction="none")
    assert torch.allclose(none, T(g["BinaryDiceLoss_none"]), rtol=1e-6)


def test_metric_matches_reference(golden):
    g = golden("metric")
    for i in (0, 1):
        m = metric.metric(g["gt%d" % i], g["pred%d" % i])
        assert m["jaccard"] == float(g["jaccard%d" % i])
        assert m["dice"] == float(g["dice%d" % i])
    z = np.zeros((1, 1, 4, 4, 4))
    m = metric.metric(z, z)
    assert m["jaccard"] == float(g["jaccard_empty"]) and m["dice"] == float(g["dice_empty"])


def test_syncbn_matches_reference(golden):
    g = golden("syncbn")
    xs = [T(g["x0"]), T(g["x1"])]
    outs, mean, inv_std, rm, rv = syncbn.forward_replicas(xs, T(g["weight"]), T(g["bias"]), torch.zeros(6),
                                                          torch.ones(6))
    assert torch.allclose(mean, T(g["mean"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(inv_std, T(g["inv_std"]), rtol=1e-6)
    assert torch.allclose(rm, T(g["running_mean"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(rv, T(g["running_var"]), rtol=1e-6)
    for o, k in zip(outs, ("out0", "out1")):
        assert torch.allclose(o, T(g[k]), rtol=1e-5, atol=1e-6)
    y, m, v = ops.batch_norm_train(xs[0], torch.ones(6), torch.zeros(6))
    assert torch.allclose(y, T(g["single_out"]), rtol=1e-4, atol=1e-5)


def test_maxpool_indices_and_argmax_bit_exact(golden):
    g = golden("pool_argmax")
    y, idx = ops.max_pool3d_k2s2(g["x"])
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(np.isnan(y), np.isnan(g["y"]))
    assert np.array_equal(np.nan_to_num(y), np.nan_to_num(g["y"]))
    assert np.array_equal(ops.argmax_labels(g["logits"]), g["argmax"])


def test_conv_restatement_agrees_with_torch():
    torch.manual_seed(0)
    for (k, s, p, d) in [(3, 1, 1, 1), (3, 2, 1, 1), (5, 1, 2, 1), (2, 2, 0, 1), (1, 1, 0, 1), (3, 1, 2, 2)]:
        x = torch.randn(2, 3, 9, 8, 10)
        w = torch.randn(4, 3, k, k, k)
        b = torch.randn(4)
        a = ops.conv3d(x, w, b, s, p, d).numpy()
        r = ops.conv3d_direct_numpy(x.numpy(), w.numpy(), b.numpy(), s, p, d)
        assert np.allclose(a, r, rtol=1e-4, atol=1e-4), (k, s, p, d)
    x = torch.randn(2, 6, 3, 4, 5)
    w = torch.randn(6, 4, 2, 2, 2)
    b = torch.randn(4)
    assert torch.allclose(ops.conv_transpose3d_k2s2(x, w, b), torch.nn.functional.conv_transpose3d(x, w, b, stride=2),
                          rtol=1e-4, atol=1e-5)


def test_window_sampler_counts_and_roundtrip():
    assert window.grid_locations((512, 512, 256), (128,) * 3, (64,) * 3).shape == (147, 6)
    assert window.grid_locations((512, 512, 256), (128,) * 3, (4, 4, 36)).shape == (75, 6)
    rng = np.random.default_rng(0)
    vol = rng.integers(0, 3, size=(1, 40, 36, 50))
    for mode in ("crop", "average"):
        locs = window.grid_locations(vol.shape[1:], (16, 16, 16), (4, 4, 6))
        agg = window.Aggregator(vol.shape[1:], (4, 4, 6), mode)
        patches = np.stack([vol[:, a:d, b:e, c:f] for a, b, c, d, e, f in locs])
        agg.add_batch(patches, locs)
        assert np.array_equal(agg.get_output_tensor(), vol)


def test_hd95_restatement_equals_brute_force():
    """oracle/metric.py: hausdorff_distance (MONAI's edge + distance-transform algorithm restated with scipy; parity unpinned
    against MONAI itself) against its definition evaluated by brute force: percentile over edge voxels of the distance to the
    nearest edge voxel of the other mask."""
    rng = np.random.default_rng(4)
    for shape, spacing in (((9, 8, 11), None), ((12, 10, 7), (0.5, 2.0, 1.25))):
        a = rng.random(shape) > 0.55
        b = rng.random(shape) > 0.6
        sp = np.array(spacing if spacing else (1, 1, 1), dtype=np.float64)

        def edges(m):
            pad = np.pad(m, 1)
            core = pad[1:-1, 1:-1, 1:-1]
            er = core & pad[:-2, 1:-1, 1:-1] & pad[2:, 1:-1, 1:-1] & pad[1:-1, :-2, 1:-1] & pad[1:-1, 2:, 1:-1] & \
                pad[1:-1, 1:-1, :-2] & pad[1:-1, 1:-1, 2:]
            return np.argwhere(core & ~er) * sp

        ea, eb = edges(a), edges(b)

        def directed(p, q):
            d = np.sqrt(((p[:, None, :] - q[None, :, :]) ** 2).sum(-1)).min(1)
            return np.percentile(d, 95)
        want = max(directed(ea, eb), directed(eb, ea))
        assert abs(metric.hausdorff_distance(a, b, 95, spacing) - want) < 1e-9
    assert np.isnan(metric.hausdorff_distance(np.zeros((4, 4, 4)), np.zeros((4, 4, 4))))


def test_data_oracle_znorm_and_crop():
    from oracle import data as odata
    rng = np.random.default_rng(0)
    v = (rng.random((1, 6, 7, 8)) * 50 + 3).astype(np.float32)
    z = odata.znormalize(v)
    assert abs(z.mean()) < 1e-5 and abs(z.std(ddof=1) - 1) < 1e-5            # torch.std semantics: unbiased
    assert np.array_equal(odata.crop(v, (1, 2, 3), (4, 3, 2)), v[:, 1:5, 2:5, 3:5])
