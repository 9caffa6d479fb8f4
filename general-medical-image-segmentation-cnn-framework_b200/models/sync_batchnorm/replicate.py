"""API shims for models/sync_batchnorm/replicate.py:27-94.

The reference needs DataParallelWithCallback / patch_replication_callback so that its thread-based SyncBN can find
its master after nn.DataParallel.replicate().  With one process per GPU there is nothing to replicate: the
synchronised layers talk through torch.distributed.  The names are kept so code written against the reference
imports and runs; wrapping a model returns it unchanged (single device) and patching is a no-op.
"""
import torch.nn as nn

__all__ = ["DataParallelWithCallback", "patch_replication_callback", "execute_replication_callbacks"]


def execute_replication_callbacks(modules):
    for m in modules[0].modules():
        if hasattr(m, "__data_parallel_replicate__"):
            m.__data_parallel_replicate__(None, 0)


class DataParallelWithCallback(nn.Module):
    """One-process-per-GPU stand-in: forwards to the wrapped module on its own device."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super(DataParallelWithCallback, self).__init__()
        if device_ids is not None and len(device_ids) > 1:
            raise ValueError("b200seg runs one process per GPU (torchrun); DataParallel over %d devices in one "
                             "process is not supported" % len(device_ids))
        self.module = module

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)


def patch_replication_callback(data_parallel):
    return data_parallel
