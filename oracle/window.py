"""Oracle (test infrastructure): sliding-window sampler / aggregator.  PARITY UNPINNED.

predict.py:100,117-118,141-147 delegates to torchio.inference.GridSampler / GridAggregator (torchio==0.20.3,
requirements.txt:13).  torchio is neither vendored in the reference nor installed in this image and there is no
network, so this file restates torchio 0.20.x's published algorithm (torchio/data/sampler/grid.py,
torchio/data/inference/aggregator.py) from its documentation; it cannot be checked against torchio here.  The
reference's own call sites fix the usage: patch_overlap=(4, 4, 36), default overlap_mode 'crop', integer label maps
fed to add_batch, get_output_tensor() -> [C, W, H, D].
"""
import numpy as np


def grid_locations(shape, patch, overlap):
    """GridSampler._get_patches_locations: per axis, starts 0, step, 2*step ... with step = patch - overlap, and a
    final start at size - patch when the last window does not end on the border.  Returns [P, 6] int64 rows
    (i0, j0, k0, i1, j1, k1), first axis slowest (itertools.product order)."""
    axes = []
    for size, p, o in zip(shape, patch, overlap):
        if p > size:
            raise ValueError("patch larger than volume")
        if o % 2 or o >= p:
            raise ValueError("overlap must be even and smaller than the patch")
        step = p - o
        starts = list(range(0, size - p + 1, step))
        if starts[-1] != size - p:
            starts.append(size - p)
        axes.append(starts)
    locs = [(i, j, k, i + patch[0], j + patch[1], k + patch[2]) for i in axes[0] for j in axes[1] for k in axes[2]]
    return np.asarray(sorted(set(locs)), np.int64)


class Aggregator:
    """GridAggregator with overlap_mode in {'crop', 'average'}; volume shape [C, W, H, D]."""

    def __init__(self, shape, patch_overlap, overlap_mode="crop"):
        self.shape = tuple(shape)
        self.overlap = np.asarray(patch_overlap, np.int64)
        self.mode = overlap_mode
        self.out = None
        self.count = None

    def _crop(self, patch, loc):
        """Trim overlap//2 from every patch face that is not on the volume border (aggregator.py:_crop_patch)."""
        half = self.overlap // 2
        i0 = np.asarray(loc[:3])
        i1 = np.asarray(loc[3:])
        lo = np.where(i0 > 0, half, 0)
        hi = np.where(i1 < np.asarray(self.shape), half, 0)
        new0 = i0 + lo
        new1 = i1 - hi
        sl = tuple(slice(int(a), int(patch.shape[1 + d] - b)) for d, (a, b) in enumerate(zip(lo, hi)))
        return patch[(slice(None),) + sl], new0, new1

    def add_batch(self, batch, locations):
        batch = np.asarray(batch)
        if self.out is None:
            dt = batch.dtype if self.mode == "crop" else np.float32
            self.out = np.zeros((batch.shape[1],) + self.shape, dt)
            if self.mode == "average":
                self.count = np.zeros((batch.shape[1],) + self.shape, np.float32)
        for patch, loc in zip(batch, np.asarray(locations)):
            if self.mode == "crop":
                c, a, b = self._crop(patch, loc)
                self.out[:, a[0]:b[0], a[1]:b[1], a[2]:b[2]] = c
            else:
                i0, j0, k0, i1, j1, k1 = [int(v) for v in loc]
                self.out[:, i0:i1, j0:j1, k0:k1] += patch
                self.count[:, i0:i1, j0:j1, k0:k1] += 1

    def get_output_tensor(self):
        if self.mode == "average":
            return self.out / np.maximum(self.count, 1)
        return self.out
