"""Oracle (test infrastructure): cross-replica batch-norm statistics (models/sync_batchnorm/batchnorm.py)."""
import torch


def local_sums(x):
    """forward :56-62: per-replica sum and square-sum over (B, L) of input viewed as [B, C, L], plus B*L."""
    v = x.reshape(x.shape[0], x.shape[1], -1)
    return v.sum(0).sum(-1), (v ** 2).sum(0).sum(-1), v.shape[0] * v.shape[2]


def compute_mean_std(sum_, ssum, size, running_mean, running_var, eps=1e-5, momentum=0.1):
    """_compute_mean_std :113-125.  Note clamp(eps)**-0.5, not (var+eps)**-0.5."""
    assert size > 1
    mean = sum_ / size
    sumvar = ssum - sum_ * mean
    unbias_var = sumvar / (size - 1)
    bias_var = sumvar / size
    new_rm = (1 - momentum) * running_mean + momentum * mean
    new_rv = (1 - momentum) * running_var + momentum * unbias_var
    return mean, bias_var.clamp(eps) ** -0.5, new_rm, new_rv


def forward_replicas(xs, weight, bias, running_mean, running_var, eps=1e-5, momentum=0.1):
    """N replicas in one process: reduce (master :90-111), then normalise each shard (:71-78)."""
    parts = [local_sums(x) for x in xs]
    s = sum(p[0] for p in parts)
    ss = sum(p[1] for p in parts)
    n = sum(p[2] for p in parts)
    mean, inv_std, rm, rv = compute_mean_std(s, ss, n, running_mean, running_var, eps, momentum)
    shp = (1, -1) + (1,) * (xs[0].dim() - 2)
    outs = []
    for x in xs:
        if weight is not None:
            outs.append((x - mean.view(shp)) * (inv_std * weight).view(shp) + bias.view(shp))
        else:
            outs.append((x - mean.view(shp)) * inv_std.view(shp))
    return outs, mean, inv_std, rm, rv
