"""Sliding-window patch sampler and GPU aggregator (predict.py:100-147).

The reference delegates to torchio.inference.GridSampler / GridAggregator (torchio 0.20.3, not vendored); the same
names, call order (`add_batch(tensor, locations)`, `get_output_tensor()`) and the [C, W, H, D] output convention are
kept.  Patch enumeration is host logic; stitching runs on the device: 'crop' mode overwrites the cropped interior of
each label patch (the mode predict.py uses), 'average' mode overlap-adds fp32 logits with a visit-count volume.
"""
import ctypes

import torch

from . import functional as F
from .functional import _call, _ptr, _stream


def grid_locations(shape, patch, overlap):
    """Per axis: starts 0, step, ... with step = patch - overlap, plus size - patch if the last window falls short."""
    axes = []
    for size, p, o in zip(shape, patch, overlap):
        if p > size:
            raise ValueError("patch size %d is larger than the volume extent %d" % (p, size))
        if o % 2 or o >= p:
            raise ValueError("patch overlap must be even and smaller than the patch size")
        starts = list(range(0, size - p + 1, p - o))
        if starts[-1] != size - p:
            starts.append(size - p)
        axes.append(starts)
    locs = sorted({(i, j, k, i + patch[0], j + patch[1], k + patch[2]) for i in axes[0] for j in axes[1] for k in axes[2]})
    return torch.tensor(locs, dtype=torch.int64)


class GridSampler:
    def __init__(self, volume, patch_size, patch_overlap=(0, 0, 0)):
        """volume: a [C, W, H, D] tensor or just its spatial shape."""
        self.volume = volume if torch.is_tensor(volume) else None
        self.spatial_shape = tuple(volume.shape[-3:]) if torch.is_tensor(volume) else tuple(volume)
        self.patch_size = tuple(int(v) for v in patch_size)
        self.patch_overlap = tuple(int(v) for v in patch_overlap)
        self.locations = grid_locations(self.spatial_shape, self.patch_size, self.patch_overlap)

    def __len__(self):
        return len(self.locations)

    def __getitem__(self, i):
        a, b, c, d, e, f = self.locations[i].tolist()
        return {"data": self.volume[..., a:d, b:e, c:f], "location": self.locations[i]}

    def batches(self, batch_size):
        for s in range(0, len(self), batch_size):
            locs = self.locations[s:s + batch_size]
            data = torch.stack([self.volume[..., a:d, b:e, c:f] for a, b, c, d, e, f in locs.tolist()])
            yield data, locs


class GridAggregator:
    def __init__(self, sampler, overlap_mode="crop", device="cuda"):
        if overlap_mode not in ("crop", "average"):
            raise ValueError("overlap_mode must be 'crop' or 'average'")
        self.sampler, self.mode, self.device = sampler, overlap_mode, torch.device(device)
        self.out = self.count = None

    def add_batch(self, batch, locations):
        vw, vh, vd = self.sampler.spatial_shape
        ow, oh, od = self.sampler.patch_overlap
        locations = locations.to(self.device, torch.int64).contiguous()
        b, c, pw, ph, pd = batch.shape
        if self.mode == "crop":
            if c != 1:
                raise ValueError("crop mode stitches single-channel label maps")
            patches = batch.to(self.device).to(torch.uint8).contiguous()
            if self.out is None:
                self.out = torch.zeros((1, vw, vh, vd), dtype=torch.uint8, device=self.device)
            _call("b200seg_window_accumulate_crop", _ptr(patches), _ptr(locations), b, pw, ph, pd, ow, oh, od,
                  _ptr(self.out), vw, vh, vd, _stream())
        else:
            patches = batch.to(self.device).float().contiguous()
            if self.out is None:
                self.out = torch.zeros((c, vw, vh, vd), dtype=torch.float32, device=self.device)
                self.count = torch.zeros((vw, vh, vd), dtype=torch.float32, device=self.device)
            _call("b200seg_window_accumulate_average", _ptr(patches), _ptr(locations), b, c, pw, ph, pd, _ptr(self.out),
                  _ptr(self.count), vw, vh, vd, _stream())

    def get_output_tensor(self, return_labels=False):
        """[C, W, H, D].  Average mode divides by the visit count (and can also return the arg-max label map)."""
        if self.mode == "crop":
            return self.out
        acc = self.out.clone()
        labels = torch.empty(self.count.shape, dtype=torch.uint8, device=self.device) if return_labels else None
        _call("b200seg_window_finalize", _ptr(acc), _ptr(self.count), acc.shape[0], self.count.numel(), _ptr(labels),
              _stream())
        return (acc, labels.unsqueeze(0)) if return_labels else acc
