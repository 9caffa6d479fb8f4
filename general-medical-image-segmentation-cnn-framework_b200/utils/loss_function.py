"""Losses with the reference's names and call signatures (utils/loss_function.py), computed by one fused reduction
kernel and one fused gradient kernel (b200seg_loss_reduce / b200seg_loss_grad)."""
import torch
from torch import nn

from .. import functional as F


def _labels_from(target, logits):
    """Class-index labels from either index maps ([N,*] / [N,1,*]) or one-hot maps shaped like the logits."""
    if target.shape == logits.shape and logits.shape[1] > 1:
        return target.argmax(dim=1)
    return target


def cross_entropy_3D(input, target, weight=None, size_average=True):
    """loss_function.py:8-16: nll_loss(log_softmax, weight, sum) / numel -- with class weights the divisor stays the voxel
    count (not the weight sum F.cross_entropy(reduction='mean') would use)."""
    loss = F.seg_loss(input, target, w_ce=1.0, w_dice=0.0, ce_class_weight=weight)
    if not size_average:
        loss = loss * float(target.numel())
    return loss


class Binary_Loss(nn.Module):
    """loss_function.py:19-41 (nn.BCEWithLogitsLoss on a one-hot target; the criterion train.py:115 trains with)."""

    def forward(self, model_output, targets):
        return F.seg_loss(model_output, _labels_from(targets, model_output), w_ce=0.0, w_dice=0.0, w_bce=1.0)


BCEWithLogitsLoss = Binary_Loss


def make_one_hot(input, num_classes):
    """loss_function.py:44-58: [N,1,*] class indices -> [N,num_classes,*] one-hot (CPU tensor, like the reference)."""
    shape = list(input.shape)
    shape[1] = num_classes
    result = torch.zeros(tuple(shape))
    return result.scatter_(1, input.cpu().long(), 1)


class BinaryDiceLoss(nn.Module):
    """loss_function.py:61-99: per-sample 1 - (sum x*t + smooth) / (sum x^p + sum t^p + smooth) on probabilities supplied
    by the caller, reduced over the batch ('mean' | 'sum' | 'none').  One reduction and one gradient kernel."""

    def __init__(self, smooth=1, p=2, reduction='mean'):
        super(BinaryDiceLoss, self).__init__()
        self.smooth, self.p, self.reduction = smooth, p, reduction

    def forward(self, predict, target):
        assert predict.shape[0] == target.shape[0], "predict & target batch size don't match"
        n = predict.shape[0]
        loss = F.prob_dice_rows(predict.reshape(n, 1, -1), target=target.reshape(n, 1, -1), by_class=False, p=self.p,
                                smooth=self.smooth, intersect_scale=1.0)
        if self.reduction == 'mean':
            return loss.mean()
        elif self.reduction == 'sum':
            return loss.sum()
        elif self.reduction == 'none':
            return loss
        raise Exception('Unexpected reduction {}'.format(self.reduction))


class DiceLoss(nn.Module):
    """loss_function.py:102-130: global sigmoid Dice against a one-hot target."""

    def __init__(self, weight=None, ignore_index=None, **kwargs):
        super(DiceLoss, self).__init__()
        self.kwargs, self.weight, self.eplison, self.ignore_index = kwargs, weight, 1e-5, ignore_index

    def forward(self, predict, target):
        assert predict.shape == target.shape, 'predict & target shape do not match'
        return F.seg_loss(predict, _labels_from(target, predict), w_ce=0.0, w_dice=0.0, w_sdice=1.0)


class DiceLossss(nn.Module):
    """loss_function.py:148-185: per-class Dice on softmax probabilities, averaged over classes."""

    def __init__(self, n_classes):
        super(DiceLossss, self).__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, weight=None, softmax=False):
        assert inputs.shape[1] == self.n_classes, 'predict & target shape do not match'
        if softmax:       # logits: soft-max, per-class sums and (weighted) class average in one fused pass
            return F.seg_loss(inputs, target, w_ce=0.0, w_dice=1.0, dice_class_weight=weight)
        # the reference's default: `inputs` already are probabilities (:172-184)
        per_class = F.prob_dice_rows(inputs, labels=target, by_class=True, p=2.0, smooth=1e-5, intersect_scale=2.0)
        if weight is not None:
            per_class = per_class * torch.as_tensor(weight, dtype=per_class.dtype, device=per_class.device)
        return per_class.sum() / self.n_classes


class DiceCELoss(nn.Module):
    """cross_entropy_3D + DiceLossss(softmax=True): the criterion BASELINE.json names, in a single pass."""

    def __init__(self, n_classes=2):
        super(DiceCELoss, self).__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target):
        return F.seg_loss(inputs, target, w_ce=1.0, w_dice=1.0)
