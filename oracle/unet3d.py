"""Oracle (test infrastructure): functional 3D U-Net driven by a reference-format state_dict.

Follows models/three_d/unet3d.py: `_block` :73-104 = (Conv3d k3 p1 bias -> BatchNorm3d -> ReLU) x 2,
`forward` :50-71 = 4 encoder levels with MaxPool3d(2,2), bottleneck, 4 x (ConvTranspose3d k2 s2 -> cat -> block),
1x1x1 head :46-48.  Running statistics are updated the way nn.BatchNorm3d does (momentum 0.1, unbiased variance).
"""
import torch
import torch.nn.functional as F

from . import ops



class _RoundBoth(torch.autograd.Function):
    """bf16 storage of an activation: value rounded in forward, its gradient rounded in backward."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    """bf16 copy of an fp32 master weight: rounded value, fp32 gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _ident(x):
    return x


LEVELS = (("encoder1", "enc1"), ("encoder2", "enc2"), ("encoder3", "enc3"), ("encoder4", "enc4"))
DECODERS = (("decoder4", "dec4", "upconv4"), ("decoder3", "dec3", "upconv3"), ("decoder2", "dec2", "upconv2"),
            ("decoder1", "dec1", "upconv1"))


def _block(sd, x, prefix, name, training, new_stats, acts, ra=_ident, rw=_ident):
    for i in (1, 2):
        k = "%s.%sconv%d" % (prefix, name, i)
        nk = "%s.%snorm%d" % (prefix, name, i)
        x = ra(ops.conv3d(x, rw(sd[k + ".weight"]), sd[k + ".bias"], padding=1))
        if acts is not None:
            acts[k] = x
        if training:
            cnt = x.numel() // x.shape[1]
            x, mean, var = ops.batch_norm_train(x, sd[nk + ".weight"], sd[nk + ".bias"])
            if new_stats is not None:
                rm, rv = ops.batch_norm_running_update(sd[nk + ".running_mean"], sd[nk + ".running_var"],
                                                       mean.detach(), var.detach(), cnt)
                new_stats[nk + ".running_mean"], new_stats[nk + ".running_var"] = rm, rv
        else:
            x = ops.batch_norm_eval(x, sd[nk + ".weight"], sd[nk + ".bias"], sd[nk + ".running_mean"],
                                    sd[nk + ".running_var"])
        x = ra(torch.relu(x))
        if acts is not None:
            acts[nk] = x
    return x


def forward(sd, x, training=True, new_stats=None, acts=None, storage="fp32"):
    """sd: mapping with the 136 reference keys (tensors; parameters may require grad). x: [N,C,D,H,W] fp32.

    storage="fp32" is the reference arithmetic.  storage="bf16" keeps fp32 arithmetic but rounds every stored
    activation (and its gradient) and every conv weight to bf16 at the points where the CUDA path stores bf16; the
    distance between the two is the error floor any bf16-storage implementation of this network has, and tests use
    it to calibrate their end-to-end tolerances."""
    ra, rw = (_RoundBoth.apply, _RoundFwd.apply) if storage == "bf16" else (_ident, _ident)
    skips = []
    h = ra(x)
    for li, (prefix, name) in enumerate(LEVELS):
        if li:
            h = F.max_pool3d(h, 2, 2)
        h = _block(sd, h, prefix, name, training, new_stats, acts, ra, rw)
        skips.append(h)
    h = _block(sd, F.max_pool3d(h, 2, 2), "bottleneck", "bottleneck", training, new_stats, acts, ra, rw)
    for (prefix, name, up), skip in zip(DECODERS, reversed(skips)):
        h = ra(ops.conv_transpose3d_k2s2(h, rw(sd[up + ".weight"]), sd[up + ".bias"]))
        if acts is not None:
            acts[up] = h
        h = torch.cat((h, skip), dim=1)
        h = _block(sd, h, prefix, name, training, new_stats, acts, ra, rw)
    return ops.conv3d(h, sd["conv.weight"], sd["conv.bias"])


def init_state_dict(in_channels=1, out_channels=2, features=32, seed=0):
    """Random-init parameters with the reference's key names and shapes (weights_init_normal 'kaiming' semantics,
    train.py:33-61: kaiming_normal_ fan_in on Conv*/ConvTranspose* weights, zero bias; BatchNorm3d left at 1/0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, co, ci, k, transpose=False):
        shape = (ci, co, k, k, k) if transpose else (co, ci, k, k, k)
        fan_in = shape[1] * k ** 3
        sd[key + ".weight"] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        sd[key + ".bias"] = torch.zeros(co)

    def norm(key, c):
        sd[key + ".weight"] = torch.ones(c)
        sd[key + ".bias"] = torch.zeros(c)
        sd[key + ".running_mean"] = torch.zeros(c)
        sd[key + ".running_var"] = torch.ones(c)
        sd[key + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def block(prefix, name, ci, f):
        conv("%s.%sconv1" % (prefix, name), f, ci, 3)
        norm("%s.%snorm1" % (prefix, name), f)
        conv("%s.%sconv2" % (prefix, name), f, f, 3)
        norm("%s.%snorm2" % (prefix, name), f)

    f = features
    block("encoder1", "enc1", in_channels, f)
    block("encoder2", "enc2", f, 2 * f)
    block("encoder3", "enc3", 2 * f, 4 * f)
    block("encoder4", "enc4", 4 * f, 8 * f)
    block("bottleneck", "bottleneck", 8 * f, 16 * f)
    conv("upconv4", 8 * f, 16 * f, 2, transpose=True)
    block("decoder4", "dec4", 16 * f, 8 * f)
    conv("upconv3", 4 * f, 8 * f, 2, transpose=True)
    block("decoder3", "dec3", 8 * f, 4 * f)
    conv("upconv2", 2 * f, 4 * f, 2, transpose=True)
    block("decoder2", "dec2", 4 * f, 2 * f)
    conv("upconv1", f, 2 * f, 2, transpose=True)
    block("decoder1", "dec1", 2 * f, f)
    conv("conv", out_channels, f, 1)
    return sd
