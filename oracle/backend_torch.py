"""Oracle (test infrastructure): the ops interface of b200seg.functional restated with torch CPU fp32 arithmetic.

The model mirrors in the product package write their forward passes against a small ops interface
(`models/_common.py`).  Binding that interface to this module runs the very same graph with the reference's own
arithmetic: torch.nn.functional on NCDHW fp32 tensors, i.e. what the reference's nn.Modules execute
(models/three_d/vnet3d.py, residual_unet3d.py, highresnet.py, densevoxelnet3d.py; utils/convolution.py, residual.py).
The graph wiring itself is pinned by golden vectors produced by the reference's modules (tests/golden/make_golden.py).

Layout: activations stay [N, C, D, H, W]; `out=` hints (concat-free buffers) are ignored.
"""
import contextlib

import torch
import torch.nn.functional as TF

# storage emulation: with bf16_storage() active, every activation the CUDA path stores as bf16 (and its gradient) and
# every conv weight is rounded to bf16 while the arithmetic stays fp32.  The distance between the two modes is the error
# floor any bf16-storage implementation of the same graph has; tests calibrate their tolerances with it.
_BF16 = [False]


@contextlib.contextmanager
def bf16_storage():
    _BF16[0] = True
    try:
        yield
    finally:
        _BF16[0] = False


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _ra(x):
    return _RoundBoth.apply(x) if _BF16[0] else x


def _rw(w):
    return _RoundFwd.apply(w) if _BF16[0] else w


class NormSpec:
    def __init__(self, kind=None, act="none", act_param=0.0, eps=1e-5, momentum=0.1, training=True, sync=False,
                 clamp_eps=False, process_group=None):
        assert kind in (None, "batch", "instance")
        self.kind, self.act, self.act_param = kind, act, float(act_param)
        self.eps, self.momentum, self.training = eps, momentum, training


def _act(x, act, param, prelu_weight):
    if act in (None, "none"):
        return x
    if act == "relu":
        return torch.relu(x)
    if act == "leaky_relu":
        return TF.leaky_relu(x, param)
    if act == "elu":
        return TF.elu(x, param)
    if act == "prelu":
        return TF.prelu(x, prelu_weight)
    raise ValueError(act)


def to_ndhwc(x):
    return _ra(x.float())


def from_ndhwc(x):
    return x


def spatial(x):
    return (x.shape[0],) + tuple(x.shape[2:])


def channels(x):
    return x.shape[1]


def device_of(x):
    return x.device


def channel_slice(x, lo, hi):
    return x[:, lo:hi]


def alloc_concat(n, d, h, w, c_first, c_second, device):
    return None, None, None


def concat_channels(a, b):
    return torch.cat((a, b), dim=1)


merge_channels = concat_channels


def repeat_channels(x, times):
    return x.repeat(1, times, 1, 1, 1)


def norm_act(y, spec, gamma=None, beta=None, prelu_weight=None, residual=None, running_mean=None, running_var=None,
             out=None):
    if spec.kind == "batch":
        y = TF.batch_norm(y, running_mean, running_var, gamma, beta, spec.training, spec.momentum, spec.eps)
    elif spec.kind == "instance":
        y = TF.instance_norm(y, eps=spec.eps)
    if residual is not None:
        y = y + residual
    return _ra(_act(y, spec.act, spec.act_param, prelu_weight))


def conv_norm_act(x, weight, bias=None, *, x2=None, k=3, stride=1, pad=1, dil=1, spec=None, gamma=None, beta=None,
                  prelu_weight=None, residual=None, running_mean=None, running_var=None, out=None):
    if x2 is not None:
        x = torch.cat((x, x2), dim=1)
    y = _ra(TF.conv3d(x, _rw(weight), bias, stride=stride, padding=pad, dilation=dil))
    spec = spec or NormSpec()
    if spec.kind is None and spec.act in (None, "none") and residual is None:
        return y
    return norm_act(y, spec, gamma, beta, prelu_weight, residual, running_mean, running_var)


def activation(x, act, act_param=0.0, prelu_weight=None, residual=None, out=None):
    return norm_act(x, NormSpec(None, act, act_param), prelu_weight=prelu_weight, residual=residual)


def max_pool2(x, return_indices=False):
    return TF.max_pool3d(x, 2, 2, return_indices=return_indices)


def conv_transpose_k2s2(x, weight, bias=None, out=None):
    return _ra(TF.conv_transpose3d(x, _rw(weight), bias, stride=2))


def conv_transpose_kxsx(x, weight, bias=None, stride=2, out=None):
    return _ra(TF.conv_transpose3d(x, _rw(weight), bias, stride=stride))


def max_pool2_skip(x):
    return TF.max_pool3d(x, 2, 2), x


def upsample_nearest2(x):
    return TF.interpolate(x, scale_factor=2, mode="nearest")


def pad3d(x, pad, mode):
    return TF.pad(x, 6 * [pad], mode) if pad else x


def add(a, b, out=None):
    return _ra(a + b)


def add_channel_padded(out, x):
    diff = out.shape[1] - x.shape[1]
    if diff:
        z = x.new_zeros((x.shape[0], diff // 2) + tuple(x.shape[2:]))
        x = torch.cat((z, x, z), dim=1)
    return _ra(x + out)


def dropout(x, p, training=True, channel=False, out=None, times=1):
    if not training or p == 0.0:
        return x
    for _ in range(times):
        x = TF.dropout3d(x, p, True) if channel else TF.dropout(x, p, True)
    return x


def head_conv1x1(x, weight, bias=None):
    return TF.conv3d(x, weight, bias)


def classmap_up2_add(coarse, fine=None):
    up = TF.interpolate(coarse, scale_factor=2, mode="nearest")
    return up if fine is None else up + fine


# ---- attention gates (ER_net.py, RE_net.py, Double_Unet.py + SE.py) ----------------------------------------------------
def convt_map_k2s2(gmap, weight, bias=None):
    return TF.conv_transpose3d(gmap, weight, bias, stride=2)


def reverse_gate(fine, g, out=None):
    x = -1 * torch.sigmoid(g) + 1
    return _ra(x.expand(-1, fine.shape[1], -1, -1, -1).mul(fine) + fine)


def gated_blend(x1, x2, gate_fn, params, out=None):
    pooled = x1.mean((2, 3, 4)) if x2 is None else (x1 + x2).mean((2, 3, 4))
    w1, w2 = gate_fn(pooled, *params)
    y = x1 * w1[:, :, None, None, None]
    if x2 is not None:
        y = y + x2 * w2[:, :, None, None, None]
    return _ra(y)


def sigmoid_map(x):
    return torch.sigmoid(x)


def concat_input(x, maps):
    return _ra(torch.cat((x.float(), maps), dim=1))
