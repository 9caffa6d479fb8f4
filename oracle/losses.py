"""Oracle (test infrastructure): the reference's loss functions restated as pure functions (utils/loss_function.py)."""
import torch
import torch.nn.functional as F


def cross_entropy_3d(logits, target, weight=None, size_average=True):
    """cross_entropy_3D (loss_function.py:8-16): log_softmax over channels, NLL (times the label's class weight) summed,
    then / numel -- the voxel count, also when weights are given."""
    logp = F.log_softmax(logits, dim=1)
    t = target.reshape(target.shape[0], 1, *logits.shape[2:]).long()
    nll = -(logp.gather(1, t))
    if weight is not None:
        nll = nll * torch.as_tensor(weight, dtype=nll.dtype)[t]
    return nll.sum() / float(t.numel()) if size_average else nll.sum()


def dice_loss_sigmoid(logits, onehot, eps=1e-5):
    """DiceLoss.forward (loss_function.py:121-130): sigmoid, global sums over batch and classes."""
    pre = torch.sigmoid(logits)
    inter = (pre * onehot).sum()
    union = (pre + onehot).sum()
    return 1 - 2 * (inter + eps) / (union + eps)


def dice_loss_per_class(inputs, target, n_classes, softmax=False, weight=None, smooth=1e-5):
    """DiceLossss.forward (loss_function.py:148-185): per-class 1 - (2*sum(p*t)+s)/(sum(p^2)+sum(t^2)+s), mean."""
    if softmax:
        inputs = torch.softmax(inputs, dim=1)
    if weight is None:
        weight = [1] * n_classes
    loss = 0.0
    for i in range(n_classes):
        t = (target == i).float()
        s = inputs[:, i]
        loss = loss + (1 - (2 * (s * t).sum() + smooth) / ((s * s).sum() + (t * t).sum() + smooth)) * weight[i]
    return loss / n_classes


def binary_dice_loss(predict, target, smooth=1, p=2, reduction="mean"):
    """BinaryDiceLoss.forward (loss_function.py:83-99)."""
    predict = predict.reshape(predict.shape[0], -1)
    target = target.reshape(target.shape[0], -1)
    num = (predict * target).sum(1) + smooth
    den = (predict.pow(p) + target.pow(p)).sum(1) + smooth
    loss = 1 - num / den
    return {"mean": loss.mean, "sum": loss.sum, "none": lambda: loss}[reduction]()


def bce_with_logits(logits, target):
    """nn.BCEWithLogitsLoss() -- the criterion train.py:115 actually trains with; mean over all elements."""
    return (torch.clamp(logits, min=0) - logits * target + torch.log1p(torch.exp(-logits.abs()))).mean()


def dice_ce(logits, labels, n_classes=2):
    """BASELINE.json config 1/2 criterion: cross_entropy_3D + DiceLossss(n)(., ., softmax=True)."""
    return cross_entropy_3d(logits, labels) + dice_loss_per_class(logits, labels, n_classes, softmax=True)
