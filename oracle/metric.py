"""Oracle (test infrastructure): utils/metric.py:20-75 restated in numpy (HD95 / MONAI path excluded)."""
import numpy as np


def counts(gt, pred):
    """Integer side of metric(): casts to int, then &, | and sums (metric.py:26-55). Returns a dict of int counts."""
    g = np.asarray(gt).astype(np.int64).squeeze()
    p = np.asarray(pred).astype(np.int64).squeeze()
    inter = g & p
    union = g | p
    fp = np.where((p - g) < 1, 0, p)
    fn = np.where((g - p) < 1, 0, g)
    tn = 1 - union
    return {"gt_sum": int(g.sum()), "pred_sum": int(p.sum()), "intersection": int(np.count_nonzero(inter)),
            "union": int(np.count_nonzero(union)), "tp": int(inter.sum()), "fp": int(fp.sum()), "fn": int(fn.sum()),
            "tn": int(tn.sum())}


def metric(gt, pred, smooth=0.001):
    """Returns (jaccard, dice) like metric(gt, pred) without spacing, plus precision/recall (metric.py:57-75)."""
    c = counts(gt, pred)
    precision = c["tp"] / (c["pred_sum"] + smooth)
    recall = c["tp"] / (c["gt_sum"] + smooth)
    jaccard = c["intersection"] / (c["union"] + smooth)
    dice = 2 * c["intersection"] / (c["gt_sum"] + c["pred_sum"] + smooth)
    return {"precision": precision, "recall": recall, "jaccard": jaccard, "dice": dice}
