"""Entry points and config surface: the Hydra-style composer (CPU) and, on the GPU, a short `train` run that must learn,
checkpoint in the reference's schema and resume, and a `predict` run whose stitched volume equals the oracle
aggregator fed with the same per-patch predictions."""
import os

import numpy as np
import pytest
import torch


def test_config_compose_and_overrides():
    from b200seg.config import build_model, compose, weights_init_normal
    c = compose([])
    assert c.network == "unet" and c.patch_size == (64, 64, 64) and c.init_lr == 0.001 and c.ckpt is None
    assert c.patch_overlap == (4, 4, 36) and c.latest_checkpoint_file == "latest_checkpoint.pt"   # predict.py:100
    c = compose(["config=res_unet", "config.batch_size=2", "config.patch_size=128,128,128", "config.ckpt=/a/b.pt"])
    assert (c.network, c.batch_size, c.patch_size, c.ckpt) == ("res_unet", 2, (128, 128, 128), "/a/b.pt")
    with pytest.raises(FileNotFoundError):
        compose(["config=does_not_exist"])
    with pytest.raises(ValueError):
        compose(["batch_size=2"])
    for name, params in (("unet", 22581250), ("vnet", 45600316), ("res_unet", 28495584), ("densevoxelnet", 1783152),
                         ("highresnet", 803636)):
        m = build_model(compose(["config=" + name]))
        assert sum(p.numel() for p in m.parameters()) == params
    # the extra config.network values of train.py:332-339,366-373 (parameter counts of the reference's constructors)
    for name, params in (("er_net", 5110176), ("re_net", 5646560), ("csrnet", None), ("dunet", None)):
        m = build_model(compose(["config=unet", "config.network=" + name]))
        assert params is None or sum(p.numel() for p in m.parameters()) == params
    m = build_model(compose(["config=unet"]))
    m.apply(weights_init_normal("kaiming"))
    assert float(m.encoder1[0].bias.abs().max()) == 0.0 and float(m.upconv1.bias.abs().max()) == 0.0
    assert float(m.encoder1[1].weight.min()) == 1.0          # BatchNorm3d untouched (train.py:39 only matches BatchNorm2d)
    with pytest.raises(NotImplementedError):
        m.apply(weights_init_normal("bogus"))
    with pytest.raises(ValueError):
        build_model(compose(["config=unet", "config.network=unetr"]))


def test_synthetic_batches_have_the_reference_layout():
    from b200seg.data import SyntheticPatches
    b = next(iter(SyntheticPatches((16, 16, 16), 3, 2, pin=False)))
    assert tuple(b["source"]["data"].shape) == (3, 1, 16, 16, 16) and b["source"]["data"].dtype == torch.float32
    assert tuple(b["gt"]["data"].shape) == (3, 1, 16, 16, 16) and set(b["gt"]["data"].unique().tolist()) <= {0.0, 1.0}


@pytest.mark.gpu
def test_train_entry_learns_checkpoints_and_resumes(tmp_path):
    from b200seg import train as T
    args = ["config=unet", "config.batch_size=2", "config.patch_size=32,32,32", "config.epochs=3",
            "config.iters_per_epoch=8", "config.epochs_per_checkpoint=2", "config.output_dir=%s" % tmp_path,
            "config.criterion=dice_ce", "config.init_lr=0.002"]
    hist = T.main(args)
    assert len(hist) == 3 and hist[-1][1] < hist[0][1], hist            # the loss goes down
    ckpt = torch.load(os.path.join(str(tmp_path), "latest_checkpoint.pt"), map_location="cpu")
    assert set(ckpt) == {"model", "optim", "scheduler", "epoch"} and ckpt["epoch"] == 3
    assert len(ckpt["model"]) == 136 and int(ckpt["optim"]["state"][0]["step"]) == 24
    assert os.path.exists(os.path.join(str(tmp_path), "checkpoint_0002.pt"))
    hist2 = T.main(args[:-1] + ["config.epochs=4", "config.load_mode=1",
                                "config.ckpt=%s" % os.path.join(str(tmp_path), "latest_checkpoint.pt")])
    assert [h[0] for h in hist2] == [4]


@pytest.mark.gpu
def test_predict_entry_matches_oracle_stitching(tmp_path):
    import b200seg.functional as F
    from b200seg import predict as P
    from b200seg.config import build_model, compose
    from b200seg.data import synthetic_volume
    from b200seg.inference import GridSampler
    from oracle import window as owindow
    args = ["config=unet", "config.batch_size=4", "config.patch_size=32,32,32", "config.volume_size=72,64,80",
            "config.patch_overlap=4,4,12", "config.output_dir=%s" % tmp_path]
    config = compose(args)
    torch.manual_seed(0)
    model = build_model(config)
    rows = P.predict(config, model)
    assert len(rows) == 1 and 0.0 <= rows[0][2] <= 1.0
    pred = np.load(os.path.join(str(tmp_path), "pred-synthetic-0000.npy"))
    assert pred.shape == (1, 72, 64, 80) and os.path.exists(os.path.join(str(tmp_path), "metrics.csv"))
    # the same per-patch predictions through the oracle aggregator (sequential, last writer wins)
    vol, _ = synthetic_volume(config.volume_size, 1, seed=config.seed)
    sampler = GridSampler(vol, config.patch_size, config.patch_overlap)
    oagg = owindow.Aggregator(vol.shape[1:], config.patch_overlap, "crop")
    model = model.cuda().eval()
    with torch.no_grad():
        for data, locs in sampler.batches(4):
            lab = F.argmax_labels(model(data.cuda().float())).cpu().numpy()
            oagg.add_batch(lab, locs.numpy())
    assert np.array_equal(pred, oagg.get_output_tensor())


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: stdout carries ONE JSON line (library chatter such as NCCL's banner is diverted to stderr by
    duplicating fd 1); the reference arm runs the oracle port on the host cores and needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_patches_per_s_128cubed" and d["unit"] == "patches/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference") and d["e2e"]["h2d_bytes_per_step"] == 0
