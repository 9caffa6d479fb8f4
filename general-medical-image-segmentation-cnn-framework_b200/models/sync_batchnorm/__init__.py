from .batchnorm import SynchronizedBatchNorm1d, SynchronizedBatchNorm2d, SynchronizedBatchNorm3d  # noqa: F401
from .replicate import DataParallelWithCallback, patch_replication_callback  # noqa: F401
