"""GPU: HD95 (utils/metric.py:29-32) against the scipy restatement of MONAI's algorithm (oracle/metric.py, parity unpinned:
MONAI is not installed), and the prediction writer of predict.py:204-214 (NIfTI-1) round trip."""
import os

import numpy as np
import pytest
import torch

from oracle import metric as ometric

pytestmark = pytest.mark.gpu


def _blobs(shape, seed, thr):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 1, *shape, generator=g)
    for _ in range(3):
        x = torch.nn.functional.avg_pool3d(x, 5, 1, 2)
    return (x[0, 0] > thr * x.std()).to(torch.uint8)


@pytest.mark.parametrize("spacing", [None, (1.0, 1.0, 1.0), (0.7, 1.3, 2.5)])
def test_hd95_matches_scipy_restatement_of_monai(spacing):
    from b200seg.utils.metric import hausdorff_distance, metric
    for shape, seed in (((40, 36, 50), 0), ((64, 64, 32), 1), ((17, 9, 23), 2)):
        gt, pred = _blobs(shape, seed, 0.5), _blobs(shape, seed + 100, 0.4)
        pred = (pred | torch.roll(gt, 2, 0)).to(torch.uint8)            # overlapping but different surfaces
        want = ometric.hausdorff_distance(pred.numpy(), gt.numpy(), 95, spacing)
        got = hausdorff_distance(pred.cuda(), gt.cuda(), 95, spacing)
        assert abs(got - want) <= 1e-4 * max(1.0, want), (shape, spacing, got, want)
        assert abs(hausdorff_distance(pred.cuda(), gt.cuda(), None, spacing) -
                   ometric.hausdorff_distance(pred.numpy(), gt.numpy(), None, spacing)) <= 1e-4 * max(1.0, want)
        d = hausdorff_distance(pred.cuda(), gt.cuda(), 95, spacing, directed=True)
        assert abs(d - ometric.hausdorff_distance(pred.numpy(), gt.numpy(), 95, spacing, directed=True)) <= 1e-4 * max(1.0, want)
    # mask touching the volume border, identical masks, empty masks
    full = torch.ones(8, 8, 8, dtype=torch.uint8)
    assert hausdorff_distance(full.cuda(), full.cuda(), 95, spacing) == 0.0
    empty = torch.zeros(8, 8, 8, dtype=torch.uint8)
    assert np.isinf(hausdorff_distance(full.cuda(), empty.cuda(), 95, spacing)) and \
        np.isinf(ometric.hausdorff_distance(full.numpy(), empty.numpy(), 95, spacing))
    assert np.isnan(hausdorff_distance(empty.cuda(), empty.cuda(), 95, spacing))
    # the reference's metric(gt, pred, spacing) tuple: precision, recall, jaccard, dice, hs95
    gt, pred = _blobs((32, 32, 32), 5, 0.3), _blobs((32, 32, 32), 6, 0.3)
    out = metric(gt[None].cuda(), pred[None].cuda(), spacing or (1, 1, 1))
    o = ometric.metric(gt.numpy(), pred.numpy())
    assert len(out) == 5 and abs(out[0] - o["precision"]) < 1e-9 and abs(out[3] - o["dice"]) < 1e-9
    assert abs(out[4] - ometric.hausdorff_distance(pred.numpy(), gt.numpy(), 95, spacing or (1, 1, 1))) < 1e-4 * max(1.0, out[4])


def test_predict_writes_nifti_and_hd95(tmp_path):
    from b200seg import predict as P
    from b200seg.config import build_model, compose
    from b200seg.utils.nifti import load_nifti
    args = ["config=unet", "config.batch_size=4", "config.patch_size=32,32,32", "config.volume_size=48,40,56",
            "config.patch_overlap=4,4,4", "config.output_dir=%s" % tmp_path, "config.save_format=nii.gz", "config.hd95=True"]
    config = compose(args)
    torch.manual_seed(0)
    rows = P.predict(config, build_model(config))
    path = os.path.join(str(tmp_path), "pred_file", "pred-0000.nii.gz")
    assert os.path.exists(path)
    vol, affine = load_nifti(path)
    assert vol.shape == (48, 40, 56) and vol.dtype == np.uint8 and np.allclose(affine, np.eye(4))
    import csv
    with open(os.path.join(str(tmp_path), "metrics.csv")) as f:
        table = list(csv.reader(f))
    assert table[0] == ["precision", "recall", "jaccard", "dice", "hs95"] and len(table) == 3     # one file + the mean row
    assert len(rows) == 1
