"""Helpers shared by the model mirrors.

Every model keeps the reference's sub-module tree (plain torch.nn modules used as parameter containers, so state_dict
keys, `.apply(weights_init_normal)` and reference checkpoints work unchanged) and writes its forward pass against a
small ops interface.  The product binds that interface to `b200seg.functional` (the CUDA kernels).  tests/ can bind it
to `oracle.backend_torch` to run the very same graph with the reference's torch CPU arithmetic.
"""
import torch.nn as nn


class OpsMixin:
    """`self.kernels` = the ops backend (default: the CUDA kernels); `set_kernels` rebinds a whole module tree."""
    _kernels = None

    @property
    def kernels(self):
        if self._kernels is None:
            from .. import functional
            return functional
        return self._kernels

    def set_kernels(self, backend):
        for m in self.modules():
            if isinstance(m, OpsMixin):
                m._kernels = backend
        return self


def bn_kind(norm):
    if norm is None:
        return None
    if isinstance(norm, (nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)):
        return "instance"
    return "batch"


def norm_spec(F, norm, act="none", act_param=0.0, module_training=True):
    """NormSpec for a torch.nn normalisation module (BatchNorm3d / SynchronizedBatchNorm3d / InstanceNorm3d / None)."""
    if norm is None:
        return F.NormSpec(None, act, act_param)
    if bn_kind(norm) == "instance":
        return F.NormSpec("instance", act, act_param, eps=norm.eps)
    from .sync_batchnorm.batchnorm import _SynchronizedBatchNorm
    sync = isinstance(norm, (_SynchronizedBatchNorm, nn.SyncBatchNorm))
    training = module_training or not norm.track_running_stats
    if module_training and norm.track_running_stats and norm.num_batches_tracked is not None:
        bump = getattr(F, "bump_counter", None)    # the CUDA backend batches these increments; other backends add now
        if bump is not None:
            bump(norm.num_batches_tracked)
        else:
            norm.num_batches_tracked += 1
    return F.NormSpec("batch", act, act_param, eps=norm.eps, momentum=0.1 if norm.momentum is None else norm.momentum,
                      training=training, sync=sync, clamp_eps=isinstance(norm, _SynchronizedBatchNorm),
                      process_group=getattr(norm, "process_group", None))


def norm_args(norm):
    """gamma, beta, running_mean, running_var keyword arguments for conv_norm_act / norm_act."""
    if norm is None or bn_kind(norm) == "instance":
        return {}
    return dict(gamma=norm.weight, beta=norm.bias, running_mean=norm.running_mean, running_var=norm.running_var)


def act_of(module):
    """(name, scalar parameter, per-channel PReLU weight) of a torch.nn activation module."""
    if module is None:
        return "none", 0.0, None
    if isinstance(module, nn.ReLU):
        return "relu", 0.0, None
    if isinstance(module, nn.LeakyReLU):
        return "leaky_relu", module.negative_slope, None
    if isinstance(module, nn.ELU):
        return "elu", module.alpha, None
    if isinstance(module, nn.PReLU):
        return "prelu", 0.0, module.weight
    raise TypeError("unsupported activation %r" % (module,))


def conv_args(conv):
    k = conv.kernel_size[0]
    assert conv.kernel_size == (k, k, k) and conv.stride[0] == conv.stride[1] == conv.stride[2], "cubic kernels only"
    return dict(k=k, stride=conv.stride[0], pad=conv.padding[0], dil=conv.dilation[0])
