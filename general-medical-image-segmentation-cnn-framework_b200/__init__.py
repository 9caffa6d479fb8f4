"""b200seg: B200-native volumetric segmentation hot path (3D U-Net family training / sliding-window inference).

Host side mirrors the reference's module layout (models/three_d, models/sync_batchnorm, utils); every tensor
operation on the path is a hand-written sm_100a kernel reached through the C ABI in include/b200seg.h.
There is no CPU or library fallback: importing the kernels without the built extension raises.
"""
__version__ = "0.1.0"
