"""Prediction entry point with the reference's command line (`python predict.py config=unet config.ckpt=/abs/path.pt`,
predict.py:217-289) and flow (predict.py:62-183): load checkpoint["model"], sliding-window inference per volume
(GridSampler -> batched eval forward -> argmax -> GridAggregator), Dice/IoU per volume, metrics.csv.  The volumes are
synthetic unless `config.data` points at a directory of `<name>.npy` / `<name>_gt.npy` pairs.  `config.save_format=nii.gz`
writes `pred_file/pred-%04d.nii.gz` like predict.py:209-214 (utils/nifti.py), `config.hd95=True` adds the reference's
precision / recall / HD95 columns and the mean row to metrics.csv (predict.py:154-201; utils/metric.py on the GPU)."""
import csv
import glob
import os
import sys

import numpy as np
import torch

from . import parallel
from .config import build_model, compose
from .data import synthetic_volume
from .inference import sliding_window_predict
from .utils.metric import metric
from .utils.nifti import save_nifti


def volumes(config):
    if config.data == "synthetic":
        for i in range(int(config.get("num_volumes", 1))):
            vol, gt = synthetic_volume(config.volume_size, config.in_classes, seed=config.seed + i)
            yield "synthetic-%04d" % i, vol, gt
    else:
        for path in sorted(glob.glob(os.path.join(config.data, "*.npy"))):
            if path.endswith("_gt.npy"):
                continue
            gt_path = path[:-4] + "_gt.npy"
            vol = torch.from_numpy(np.load(path)).float()
            gt = torch.from_numpy(np.load(gt_path)).to(torch.uint8) if os.path.exists(gt_path) else None
            yield os.path.basename(path)[:-4], (vol if vol.dim() == 4 else vol[None]), gt


def predict(config, model, log=print):
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if config.ckpt:
        ckpt = torch.load(config.ckpt, map_location="cpu")      # predict.py:79-80
        model.load_state_dict(ckpt["model"])
    model = model.to(dev).eval()
    os.makedirs(config.hydra_path, exist_ok=True)
    rows, full = [], []
    want_hd95 = str(config.get("hd95", False)).lower() in ("true", "1")
    fmt = str(config.get("save_format", "npy"))
    spacing = tuple(float(v) for v in str(config.get("spacing", "1,1,1")).replace(" ", "").split(","))
    for i, (name, vol, gt) in enumerate(volumes(config)):
        labels = sliding_window_predict(model, vol, config.patch_size, config.patch_overlap,
                                        batch_size=config.batch_size, overlap_mode=config.overlap_mode)
        if gt is not None and want_hd95:
            precision, recall, jaccard, dice, hs95 = metric(gt.to(dev), labels, spacing)       # predict.py:154
            full.append((precision, recall, jaccard, dice, hs95))
        elif gt is not None:
            jaccard, dice = metric(gt.to(dev), labels)
        else:
            jaccard = dice = float("nan")
        rows.append((name, jaccard, dice))
        if rank == 0:
            if fmt in ("nii", "nii.gz"):                                                       # predict.py:209-214
                os.makedirs(os.path.join(config.hydra_path, "pred_file"), exist_ok=True)
                save_nifti(os.path.join(config.hydra_path, "pred_file", "pred-%04d.%s" % (i, fmt)), labels.cpu().numpy())
            else:
                np.save(os.path.join(config.hydra_path, "pred-%s.npy" % name), labels.cpu().numpy())
            log("%s: jaccard %.4f dice %.4f" % (name, jaccard, dice))
    if rank == 0:
        with open(os.path.join(config.hydra_path, "metrics.csv"), "w", newline="") as f:   # predict.py:186-201
            w = csv.writer(f)
            if want_hd95 and full:
                w.writerow(["precision", "recall", "jaccard", "dice", "hs95"])
                w.writerows(full)
                w.writerow([float(np.mean([r[j] for r in full])) for j in range(5)])           # the mean row of save_csv
            else:
                w.writerow(["name", "jaccard", "dice"])
                w.writerows(rows)
    return rows


def main(argv=None):
    config = compose(sys.argv[1:] if argv is None else argv)
    return predict(config, build_model(config))


if __name__ == "__main__":
    main()
