"""Oracle (test infrastructure): the training data transforms of dataloader.py:52-67 restated with numpy.

torchio 0.20.3 (requirements.txt:13) is neither vendored nor installed: **parity unpinned**, restated from its documented
behaviour -- ZNormalization without a masking method subtracts the mean and divides by torch.std (UNBIASED) of all voxels
of the image; UniformSampler draws the patch's start index uniformly from [0, size - patch] per axis."""
import numpy as np


def znormalize(volume):
    v = np.asarray(volume, dtype=np.float64)
    return ((v - v.mean()) / v.std(ddof=1)).astype(np.float32)


def crop(volume, start, patch):
    x0, y0, z0 = start
    pw, ph, pd = patch
    return np.asarray(volume)[..., x0:x0 + pw, y0:y0 + ph, z0:z0 + pd]
