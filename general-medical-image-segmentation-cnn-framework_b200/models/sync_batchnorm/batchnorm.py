"""Cross-GPU batch normalisation with the reference's class names and numerics
(models/sync_batchnorm/batchnorm.py:38-125), re-based on one process per GPU.

The reference's version runs inside a single process (nn.DataParallel + threads + queues, comm.py) and reduces
(sum, sum-of-squares) on a master replica.  Here every rank all-reduces the same two vectors over NCCL
(torch.distributed) and finalises locally, which yields the identical mean / clamp(var, eps)^-1/2 on every rank
(_compute_mean_std :113-125) without the master/slave hand-shake.  Off-distributed (world size 1) and in eval mode
it behaves like nn.BatchNorm, exactly as the reference falls back to F.batch_norm (:50-53).
"""
import torch
from torch.nn.modules.batchnorm import _BatchNorm

from ... import functional as F

__all__ = ["SynchronizedBatchNorm1d", "SynchronizedBatchNorm2d", "SynchronizedBatchNorm3d", "convert_model"]


class _SynchronizedBatchNorm(_BatchNorm):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, process_group=None):
        super(_SynchronizedBatchNorm, self).__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self.process_group = process_group

    def _spec(self):
        import torch.distributed as dist
        parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1
        # the clamp(eps) variant only applies on the synchronised path; otherwise F.batch_norm semantics (:50-53)
        return F.NormSpec("batch", "none", eps=self.eps, momentum=self.momentum, training=self.training,
                          sync=parallel, clamp_eps=parallel and self.training, process_group=self.process_group)

    def forward(self, input):
        self._check_input_dim(input)
        shape = input.shape
        if self.training and self.num_batches_tracked is not None:
            self.num_batches_tracked += 1
        # view as [N, D, H, W, C] rows for the kernels: channels-last bf16
        x5 = input.reshape(shape[0], shape[1], -1, 1, 1) if input.dim() < 5 else input
        h = F.to_ndhwc(x5)
        z = F.norm_act(h, self._spec(), self.weight, self.bias, running_mean=self.running_mean,
                       running_var=self.running_var)
        return F.from_ndhwc(z).reshape(shape)


class SynchronizedBatchNorm1d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 2 and input.dim() != 3:
            raise ValueError('expected 2D or 3D input (got {}D input)'.format(input.dim()))


class SynchronizedBatchNorm2d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))


class SynchronizedBatchNorm3d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 5:
            raise ValueError('expected 5D input (got {}D input)'.format(input.dim()))


def convert_model(module, process_group=None):
    """Replace every nn.BatchNorm{1,2,3}d in `module` by its synchronised counterpart (parameters shared)."""
    mapping = {torch.nn.BatchNorm1d: SynchronizedBatchNorm1d, torch.nn.BatchNorm2d: SynchronizedBatchNorm2d,
               torch.nn.BatchNorm3d: SynchronizedBatchNorm3d}
    for name, child in list(module.named_children()):
        cls = mapping.get(type(child))
        if cls is not None:
            new = cls(child.num_features, child.eps, child.momentum, child.affine, process_group)
            if child.affine:
                new.weight, new.bias = child.weight, child.bias
            new.running_mean, new.running_var = child.running_mean, child.running_var
            new.num_batches_tracked = child.num_batches_tracked
            new.training = child.training
            setattr(module, name, new)
        else:
            convert_model(child, process_group)
    return module
