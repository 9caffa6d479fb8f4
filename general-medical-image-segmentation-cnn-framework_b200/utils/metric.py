"""Dice / IoU / HD95 metric with the reference's signature (utils/metric.py:20-75); counts, mask edges and surface distances
are computed on the GPU."""
import ctypes

import numpy as np
import torch

from .. import functional as F


def _edge_points(mask, spacing):
    """Physical coordinates [M, 3] (fp32, device) of the edge voxels of a binary mask [W, H, D] (uint8, device)."""
    w, h, d = mask.shape
    count = torch.zeros(1, dtype=torch.int64, device=mask.device)
    sx, sy, sz = (float(s) for s in spacing)
    F._call("b200seg_mask_edge_points", F._ptr(mask), w, h, d, sx, sy, sz, None, 0, F._ptr(count), F._stream())
    m = int(count.item())
    pts = torch.empty((max(m, 1), 3), dtype=torch.float32, device=mask.device)
    if m:
        count.zero_()
        F._call("b200seg_mask_edge_points", F._ptr(mask), w, h, d, sx, sy, sz, F._ptr(pts), m, F._ptr(count), F._stream())
    return pts[:m]


def _directed_percentile(a, b, percentile):
    """np.percentile of the distances from the points of a to the point set b (MONAI's compute_percent_hausdorff_distance):
    nan when a is empty, inf when b is empty."""
    if a.shape[0] == 0:
        return float("nan")
    if b.shape[0] == 0:
        return float("inf")
    out = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    F._call("b200seg_min_distances", F._ptr(a), a.shape[0], F._ptr(b), b.shape[0], F._ptr(out), F._stream())
    dist = out.cpu().numpy().astype(np.float64)
    return float(dist.max()) if not percentile else float(np.percentile(dist, percentile))


def hausdorff_distance(pred, gt, percentile=95, spacing=None, directed=False):
    """monai.metrics.compute_hausdorff_distance(pred[None], gt[None], percentile=..., spacing=...)[0][0] for ONE binary
    mask pair [1, W, H, D] / [W, H, D] (metric.py:29-32): the larger of the two directed percentile surface distances
    between the masks' edge sets; nan when a mask has no foreground at all (MONAI warns and returns nan / inf)."""
    dev = pred.device if pred.is_cuda else (gt.device if gt.is_cuda else torch.device("cuda"))
    p = (pred.to(dev).reshape(pred.shape[-3:]) != 0).to(torch.uint8).contiguous()
    g = (gt.to(dev).reshape(gt.shape[-3:]) != 0).to(torch.uint8).contiguous()
    if spacing is None:
        spacing = (1.0, 1.0, 1.0)
    elif isinstance(spacing, (int, float)):
        spacing = (float(spacing),) * 3
    ep, eg = _edge_points(p, spacing), _edge_points(g, spacing)
    d1 = _directed_percentile(ep, eg, percentile)
    if directed:
        return d1
    d2 = _directed_percentile(eg, ep, percentile)
    return float(np.nanmax([d1, d2])) if not (np.isnan(d1) and np.isnan(d2)) else float("nan")


def metric(gt, pred, spacing=None):
    """(jaccard, dice) like the reference; with `spacing` (precision, recall, jaccard, dice, hs95) (metric.py:69-75)."""
    dev = gt.device if gt.is_cuda else (pred.device if pred.is_cuda else torch.device("cuda"))
    c = F.seg_counts(gt.to(dev), pred.to(dev)).tolist()   # one 32-byte device->host read
    gdth_sum, pred_sum, intersection_sum, union_sum = c
    smooth = 0.001
    jaccard = intersection_sum / (union_sum + smooth)
    dice = 2 * intersection_sum / (gdth_sum + pred_sum + smooth)
    if spacing:
        hs95 = hausdorff_distance(pred, gt, percentile=95, spacing=spacing)
        return intersection_sum / (pred_sum + smooth), intersection_sum / (gdth_sum + smooth), jaccard, dice, hs95
    return jaccard, dice


def metric_full(gt, pred):
    """precision, recall, jaccard, dice (metric.py:57-66) for binary masks."""
    dev = gt.device if gt.is_cuda else (pred.device if pred.is_cuda else torch.device("cuda"))
    gdth_sum, pred_sum, inter, union = F.seg_counts(gt.to(dev), pred.to(dev)).tolist()
    smooth = 0.001
    return (inter / (pred_sum + smooth), inter / (gdth_sum + smooth), inter / (union + smooth),
            2 * inter / (gdth_sum + pred_sum + smooth))
