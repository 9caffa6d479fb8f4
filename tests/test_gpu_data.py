"""GPU: the device-resident training data path (data.GpuPatchSampler: z-normalisation + uniform patch crop,
dataloader.py:52-67) against the numpy restatement oracle/data.py."""
import numpy as np
import pytest
import torch

from oracle import data as odata

pytestmark = pytest.mark.gpu


def test_gpu_patch_sampler_matches_oracle():
    from b200seg.data import GpuPatchSampler
    g = torch.Generator().manual_seed(0)
    vols = [torch.randn(1, 40, 36, 50, generator=g) * 3 + 7, torch.randn(1, 33, 48, 32, generator=g) * 0.5 - 2,
            torch.randn(1, 64, 64, 64, generator=g) * 100 + 1000]
    labs = [(v > v.mean()).to(torch.uint8) for v in vols]
    sampler = GpuPatchSampler(vols, labs, (32, 32, 32), batch_size=4, samples_per_volume=6, seed=3)
    assert len(sampler) == 3 * 6 // 4
    # z-normalisation constants: mean and UNBIASED std of the whole image
    for v, norm in zip(vols, sampler.norms):
        m, inv = norm.cpu().tolist()
        assert abs(m - float(v.double().mean())) < 1e-5 * max(1.0, abs(float(v.mean())))
        assert abs(inv - 1.0 / float(v.double().std())) < 1e-5 / float(v.std())
    locs = sampler.draw_locations()
    assert len(locs) == 18 and sorted(set(l[0] for l in locs)) == [0, 1, 2]
    for vi, x0, y0, z0 in locs:
        shp = vols[vi].shape[1:]
        assert 0 <= x0 <= shp[0] - 32 and 0 <= y0 <= shp[1] - 32 and 0 <= z0 <= shp[2] - 32
    assert any(l[1] > 0 for l in locs if l[0] == 0) and all(l[1:] == (0, 0, 0)[:0] + (l[1], l[2], l[3]) for l in locs)
    batch = sampler.crop(locs[:5])
    x, gt = batch["source"]["data"], batch["gt"]["data"]
    assert x.shape == (5, 1, 32, 32, 32) and x.dtype == torch.float32 and gt.shape == (5, 1, 32, 32, 32)
    for b, (vi, x0, y0, z0) in enumerate(locs[:5]):
        want = odata.crop(odata.znormalize(vols[vi].numpy()), (x0, y0, z0), (32, 32, 32))
        assert np.allclose(x[b].cpu().numpy(), want, rtol=1e-4, atol=1e-4), b
        assert np.array_equal(gt[b].cpu().numpy(), odata.crop(labs[vi].numpy(), (x0, y0, z0), (32, 32, 32)))
    # an epoch: drop_last batches of the reference's dict layout; same seed -> same stream of patches
    again = GpuPatchSampler(vols, labs, (32, 32, 32), batch_size=4, samples_per_volume=6, seed=3)
    a = [b["source"]["data"].clone() for b in sampler]
    b_ = [b["source"]["data"].clone() for b in GpuPatchSampler(vols, labs, (32, 32, 32), 4, 6, seed=3)]
    assert len(a) == 4 and all(t.shape == (4, 1, 32, 32, 32) for t in a)
    del again
    assert not torch.equal(a[0], a[1])
    # (the first sampler already drew one epoch above, so compare a fresh pair instead)
    c = [b["source"]["data"].clone() for b in GpuPatchSampler(vols, labs, (32, 32, 32), 4, 6, seed=3)]
    assert all(torch.equal(p, q) for p, q in zip(b_, c))
    with pytest.raises(ValueError):
        GpuPatchSampler(vols, labs, (128, 32, 32), 4)
    # the full-size case of BASELINE config 2: 128^3 patches out of a 512x512x256 volume, one read + one write per voxel
    big = torch.randn(1, 512, 512, 256, generator=g)
    s2 = GpuPatchSampler([big], [(big > 0).to(torch.uint8)], (128, 128, 128), batch_size=2, samples_per_volume=4, seed=1)
    bb = next(iter(s2))
    xb = bb["source"]["data"]
    assert xb.shape == (2, 1, 128, 128, 128) and abs(float(xb.mean())) < 0.05 and abs(float(xb.std()) - 1) < 0.05


def test_train_entry_runs_on_device_resident_volumes(tmp_path):
    """config.data=volumes: the training loop of train.py fed by GpuPatchSampler instead of the synthetic host generator."""
    from b200seg import train as T
    args = ["config=unet", "config.batch_size=2", "config.patch_size=32,32,32", "config.epochs=2", "config.data=volumes",
            "config.volume_size=48,64,40", "config.train_volumes=2", "config.samples_per_volume=4",
            "config.output_dir=%s" % tmp_path, "config.criterion=dice_ce", "config.init_lr=0.002"]
    hist = T.main(args)
    assert len(hist) == 2 and hist[-1][1] < hist[0][1], hist
