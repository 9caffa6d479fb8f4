"""Dice / IoU metric with the reference's signature (utils/metric.py:20-75); counts are reduced on the GPU."""
import torch

from .. import functional as F


def metric(gt, pred, spacing=None):
    """Returns (jaccard, dice) like the reference; with `spacing` the reference also returns HD95 via MONAI, which is
    outside this path and raises here."""
    if spacing:
        raise NotImplementedError("HD95 (MONAI) is outside the b200seg hot path")
    dev = gt.device if gt.is_cuda else (pred.device if pred.is_cuda else torch.device("cuda"))
    c = F.seg_counts(gt.to(dev), pred.to(dev)).tolist()   # one 32-byte device->host read
    gdth_sum, pred_sum, intersection_sum, union_sum = c
    smooth = 0.001
    jaccard = intersection_sum / (union_sum + smooth)
    dice = 2 * intersection_sum / (gdth_sum + pred_sum + smooth)
    return jaccard, dice


def metric_full(gt, pred):
    """precision, recall, jaccard, dice (metric.py:57-66) for binary masks."""
    dev = gt.device if gt.is_cuda else (pred.device if pred.is_cuda else torch.device("cuda"))
    gdth_sum, pred_sum, inter, union = F.seg_counts(gt.to(dev), pred.to(dev)).tolist()
    smooth = 0.001
    return (inter / (pred_sum + smooth), inter / (gdth_sum + smooth), inter / (union + smooth),
            2 * inter / (gdth_sum + pred_sum + smooth))
