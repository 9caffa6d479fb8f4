"""Oracle (test infrastructure): utils/metric.py:20-75 restated in numpy.

HD95 (metric.py:29-32) delegates to monai==1.3.1 `compute_hausdorff_distance`, which is neither vendored nor installed:
`hausdorff_distance` below restates MONAI's published algorithm with scipy.ndimage (**parity unpinned**): edges =
seg ^ binary_erosion(seg) (6-neighbourhood, border 0), surface distance = distance_transform_edt(~edges_other,
sampling=spacing) sampled at the edges, np.percentile(.., 95), maximum of the two directions."""
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


def _edges(seg):
    from scipy import ndimage
    seg = np.asarray(seg).astype(bool)
    return ndimage.binary_erosion(seg) ^ seg


def hausdorff_distance(pred, gt, percentile=95, spacing=None, directed=False):
    """monai.metrics.hausdorff_distance.compute_hausdorff_distance for one binary mask pair [W, H, D] (see module doc)."""
    from scipy import ndimage
    ep, eg = _edges(np.squeeze(pred) != 0), _edges(np.squeeze(gt) != 0)

    def one(a, b):
        if not a.any():
            return float("nan")
        if not b.any():
            return float("inf")
        dis = ndimage.distance_transform_edt(~b, sampling=spacing)
        d = dis[a]
        return float(d.max()) if not percentile else float(np.percentile(d, percentile))
    d1 = one(ep, eg)
    if directed:
        return d1
    d2 = one(eg, ep)
    return float(np.nanmax([d1, d2])) if not (np.isnan(d1) and np.isnan(d2)) else float("nan")
