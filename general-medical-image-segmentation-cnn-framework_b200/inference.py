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
        self.out = self.count = self.keys = None
        self._added = 0

    def add_batch(self, batch, locations, patch_ids=None):
        """batch: [B, C, pw, ph, pd] (labels for 'crop', class scores for 'average'); locations: [B, 6] int64.
        patch_ids: position of each patch in sampler order (defaults to the order of the add_batch calls); needed when
        a rank only sees a share of the patches, so that overlapping interiors resolve as in a sequential run."""
        vw, vh, vd = self.sampler.spatial_shape
        ow, oh, od = self.sampler.patch_overlap
        locations = locations.to(self.device, torch.int64).contiguous()
        b, c, pw, ph, pd = batch.shape
        if self.mode == "crop":
            if c != 1:
                raise ValueError("crop mode stitches single-channel label maps")
            patches = batch.to(self.device).to(torch.uint8).contiguous()
            self._ensure_buffers(1)
            last = int(patch_ids.max()) if patch_ids is not None and patch_ids.numel() else self._added + b - 1
            if last >= (1 << 23) - 1 or (patch_ids is not None and patch_ids.numel() and int(patch_ids.min()) < 0):
                raise ValueError("crop-mode keys hold 23-bit patch indices: patch id %d is out of range" % last)
            ids = patch_ids.to(self.device, torch.int64).contiguous() if patch_ids is not None else None
            _call("b200seg_window_accumulate_crop", _ptr(patches), _ptr(locations), _ptr(ids), self._added, b, pw, ph, pd,
                  ow, oh, od, _ptr(self.keys), vw, vh, vd, _stream())
        else:
            patches = batch.to(self.device).float().contiguous()
            self._ensure_buffers(c)
            if self.out.shape[0] != c:
                raise ValueError("average mode: %d channels, the aggregator holds %d" % (c, self.out.shape[0]))
            _call("b200seg_window_accumulate_average", _ptr(patches), _ptr(locations), b, c, pw, ph, pd, _ptr(self.out),
                  _ptr(self.count), vw, vh, vd, _stream())
        self._added += b

    def _ensure_buffers(self, channels):
        vw, vh, vd = self.sampler.spatial_shape
        if self.mode == "crop":
            if self.keys is None:
                self.keys = torch.zeros((vw, vh, vd), dtype=torch.int32, device=self.device)
        elif self.out is None:
            self.out = torch.zeros((channels, vw, vh, vd), dtype=torch.float32, device=self.device)
            self.count = torch.zeros((vw, vh, vd), dtype=torch.float32, device=self.device)

    def all_reduce(self, group=None, channels=None):
        """Merge the volumes of ranks that each aggregated a share of the patches (parallel.shard_patches).  EVERY rank
        joins the collective, also one whose share was empty (more ranks than patches): its volumes are all-zero.
        channels: class-score channels of the average mode, needed only by a rank that never called add_batch."""
        from . import parallel
        if self.mode == "crop":
            self._ensure_buffers(1)
            parallel.reduce_volume(self.keys, group, op="max")
        else:
            if self.out is None:
                if channels is None:
                    raise ValueError("all_reduce on an empty average-mode aggregator needs the channel count")
                self._ensure_buffers(channels)
            parallel.reduce_volume(self.out, group)
            parallel.reduce_volume(self.count, group)

    def get_output_tensor(self, return_labels=False):
        """[C, W, H, D].  Average mode divides by the visit count (and can also return the arg-max label map)."""
        if self.mode == "crop":
            labels = torch.empty(self.keys.shape, dtype=torch.uint8, device=self.device)
            _call("b200seg_window_keys_to_labels", _ptr(self.keys), _ptr(labels), self.keys.numel(), _stream())
            return labels.unsqueeze(0)
        acc = self.out.clone()
        labels = torch.empty(self.count.shape, dtype=torch.uint8, device=self.device) if return_labels else None
        _call("b200seg_window_finalize", _ptr(acc), _ptr(self.count), acc.shape[0], self.count.numel(), _ptr(labels),
              _stream())
        return (acc, labels.unsqueeze(0)) if return_labels else acc


def _out_channels(model):
    """Class-score channels of a segmentation model = output channels of its last convolution."""
    last = None
    for m in model.modules():
        if isinstance(m, torch.nn.Conv3d):
            last = m
    return last.out_channels if last is not None else None


@torch.no_grad()
def sliding_window_predict(model, volume, patch_size, patch_overlap, batch_size=16, overlap_mode="crop", group=None):
    """predict.py:98-147 for one volume: grid patches -> batched eval-mode forward -> argmax -> stitched label volume.

    volume: [C, W, H, D] tensor (host or device).  Returns uint8 labels [1, W, H, D] on the device ('crop': stitched
    argmax patches, the reference's behaviour; 'average': argmax of the overlap-averaged class scores).  Under
    torch.distributed every rank runs a round-robin share of the patches and the volumes are merged with one all-reduce
    (the reference shards the same way, predict.py:111, but never merges)."""
    from . import parallel
    dev = next(model.parameters()).device
    sampler = GridSampler(volume, patch_size, patch_overlap)
    agg = GridAggregator(sampler, overlap_mode, device=dev)
    mine = parallel.shard_patches(len(sampler), group)
    vol = volume.to(dev, non_blocking=True)
    was_training = model.training
    model.eval()
    with F.frozen_parameters():     # weight packs and eval-mode normalisation constants: once per volume, not per batch
        for s in range(0, len(mine), batch_size):
            ids = torch.tensor(mine[s:s + batch_size], dtype=torch.int64)
            locs = sampler.locations[ids]
            x = torch.stack([vol[..., a:d, b:e, c:f] for a, b, c, d, e, f in locs.tolist()]).float()
            logits = model(x)
            if overlap_mode == "crop":
                agg.add_batch(F.argmax_labels(logits), locs, patch_ids=ids)
            else:
                agg.add_batch(logits, locs)
    agg.all_reduce(group, channels=getattr(model, "out_channels", None) or _out_channels(model))
    model.train(was_training)
    if overlap_mode == "crop":
        return agg.get_output_tensor()
    return agg.get_output_tensor(return_labels=True)[1]
