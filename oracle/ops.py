"""Oracle (test infrastructure): fp32 CPU restatement of the layer arithmetic the reference gets from torch.nn.

The reference contains no layer math of its own; its modules call nn.Conv3d / nn.BatchNorm3d / nn.MaxPool3d /
nn.ConvTranspose3d (models/three_d/unet3d.py:19-48, 73-104).  The floating-point oracle for those layers is therefore
torch's fp32 CPU implementation (the same third-party code, torch pinned at 1.13.1 in requirements.txt:12, 2.11 here),
wrapped so tests can call it layer by layer.  The integer side conditions (max-pool argmax, arg-max label maps) are
restated in numpy because they must match bit-exactly.
"""
import numpy as np
import torch
import torch.nn.functional as F


def conv3d(x, w, b=None, stride=1, padding=0, dilation=1):
    """nn.Conv3d forward (unet3d.py:80-98; vnet3d.py:25,47,65,111; residual_unet3d.py:29-44)."""
    return F.conv3d(x, w, b, stride=stride, padding=padding, dilation=dilation)


def conv3d_direct_numpy(x, w, b=None, stride=1, padding=0, dilation=1):
    """Plain-loop cross-correlation in float64 for tiny cases; checks that F.conv3d means what we think it means."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, ci, d, h, ww = x.shape
    co, _, kd, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (padding,) * 2, (padding,) * 2, (padding,) * 2))
    do = (d + 2 * padding - dilation * (kd - 1) - 1) // stride + 1
    ho = (h + 2 * padding - dilation * (kh - 1) - 1) // stride + 1
    wo = (ww + 2 * padding - dilation * (kw - 1) - 1) // stride + 1
    y = np.zeros((n, co, do, ho, wo))
    for a in range(kd):
        for bb in range(kh):
            for c in range(kw):
                patch = xp[:, :, a * dilation:a * dilation + (do - 1) * stride + 1:stride,
                           bb * dilation:bb * dilation + (ho - 1) * stride + 1:stride,
                           c * dilation:c * dilation + (wo - 1) * stride + 1:stride]
                y += np.einsum("ncdhw,oc->nodhw", patch, w[:, :, a, bb, c])
    if b is not None:
        y += np.asarray(b, np.float64)[None, :, None, None, None]
    return y


def conv_transpose3d_k2s2(x, w, b=None):
    """nn.ConvTranspose3d(kernel_size=2, stride=2) (unet3d.py:29-43): non-overlapping, so a GEMM + pixel shuffle.
    w is [C_in, C_out, 2, 2, 2]."""
    n, ci, d, h, ww = x.shape
    co = w.shape[1]
    y = torch.einsum("ncdhw,coabk->nodahbwk", x, w).reshape(n, co, 2 * d, 2 * h, 2 * ww)
    if b is not None:
        y = y + b.view(1, -1, 1, 1, 1)
    return y


def batch_norm_train(x, gamma, beta, eps=1e-5):
    """nn.BatchNorm3d in training mode (unet3d.py:88,100): biased variance over (N,D,H,W).
    Returns y, batch mean, biased var."""
    dims = (0, 2, 3, 4)
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    shp = (1, -1, 1, 1, 1)
    y = (x - mean.view(shp)) / torch.sqrt(var.view(shp) + eps)
    if gamma is not None:
        y = y * gamma.view(shp) + beta.view(shp)
    return y, mean, var


def batch_norm_running_update(running_mean, running_var, mean, var_biased, count, momentum=0.1):
    """Running statistics as nn.BatchNorm updates them: unbiased variance, momentum 0.1."""
    unbiased = var_biased * (count / max(count - 1, 1))
    return ((1 - momentum) * running_mean + momentum * mean, (1 - momentum) * running_var + momentum * unbiased)


def batch_norm_eval(x, gamma, beta, running_mean, running_var, eps=1e-5):
    shp = (1, -1, 1, 1, 1)
    y = (x - running_mean.view(shp)) / torch.sqrt(running_var.view(shp) + eps)
    if gamma is not None:
        y = y * gamma.view(shp) + beta.view(shp)
    return y


def instance_norm(x, eps=1e-5):
    """nn.InstanceNorm3d default (affine=False, no running stats) (residual_unet3d.py:21-27 etc.)."""
    dims = (2, 3, 4)
    mean = x.mean(dims, keepdim=True)
    var = x.var(dims, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def max_pool3d_k2s2(x):
    """nn.MaxPool3d(2, 2) with torch's index convention (unet3d.py:19-25): int64 flat index into the input D*H*W
    plane of that (n, c); ties -> first element in (d, h, w) scan order; NaN propagates with its index.
    numpy restatement; x: [N, C, D, H, W] float array.  Returns (values, indices)."""
    x = np.asarray(x)
    n, c, d, h, w = x.shape
    do, ho, wo = d // 2, h // 2, w // 2
    out = np.empty((n, c, do, ho, wo), x.dtype)
    idx = np.empty((n, c, do, ho, wo), np.int64)
    first = True
    for a in range(2):
        for b in range(2):
            for e in range(2):
                v = x[:, :, a:2 * do:2, b:2 * ho:2, e:2 * wo:2]
                dd = (np.arange(do) * 2 + a)[:, None, None]
                hh = (np.arange(ho) * 2 + b)[None, :, None]
                ww = (np.arange(wo) * 2 + e)[None, None, :]
                flat = np.broadcast_to((dd * h + hh) * w + ww, v.shape)
                if first:
                    out[...] = v
                    idx[...] = flat
                    first = False
                else:
                    # strictly greater replaces; a NaN candidate replaces a non-NaN best (and then sticks)
                    take = (v > out) | (np.isnan(v) & ~np.isnan(out))
                    out = np.where(take, v, out)
                    idx = np.where(take, flat, idx)
    return out, idx


def argmax_labels(logits):
    """pred.argmax(dim=1, keepdim=True) (train.py:204, predict.py:139): ties -> lowest class index."""
    x = np.asarray(logits)
    best = x[:, 0]
    lab = np.zeros(best.shape, np.int64)
    for k in range(1, x.shape[1]):
        take = (x[:, k] > best) | (np.isnan(x[:, k]) & ~np.isnan(best))
        best = np.where(take, x[:, k], best)
        lab = np.where(take, k, lab)
    return lab[:, None]


ACTIVATIONS = {
    "none": lambda x, p=None: x,
    "relu": lambda x, p=None: torch.relu(x),
    "leaky_relu": lambda x, p=None: F.leaky_relu(x, 0.01),
    "elu": lambda x, p=None: F.elu(x, 1.0),
    "prelu": lambda x, p: F.prelu(x, p),
}


def upsample_nearest2(x):
    """nn.Upsample(scale_factor=2, mode='nearest') (residual_unet3d.py:19,103)."""
    return x.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
