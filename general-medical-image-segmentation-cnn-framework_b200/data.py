"""Synthetic patch source with the batch layout the reference's loop reads (train.py:189-190):
batch["source"]["data"] fp32 [B, C, W, H, D] and batch["gt"]["data"] [B, 1, W, H, D] binary labels.  Replaces the
torchio SubjectsDataset -> Queue -> UniformSampler chain (dataloader.py:52-112), which is real-data IO outside the path."""
import torch


class SyntheticPatches:
    def __init__(self, patch_size, batch_size, iters, in_channels=1, seed=0, pin=True):
        self.patch_size, self.batch_size, self.iters = tuple(patch_size), batch_size, iters
        self.in_channels, self.seed, self.pin = in_channels, seed, pin and torch.cuda.is_available()

    def __len__(self):
        return self.iters

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.iters):
            x = torch.randn((self.batch_size, self.in_channels) + self.patch_size, generator=g)
            # a blob-like foreground: threshold of a smoothed field correlated with the image
            gt = (torch.nn.functional.avg_pool3d(x[:, :1], 5, 1, 2) > 0.25).float()
            if self.pin:
                x, gt = x.pin_memory(), gt.pin_memory()
            yield {"source": {"data": x}, "gt": {"data": gt}}


class DevicePrefetcher:
    """Iterate over host batches with ONE batch of look-ahead on the device: the pinned-memory -> HBM copy of batch k+1 is
    issued on a side stream right before batch k is handed out, so it overlaps step k instead of preceding step k+1
    (the reference's loop copies synchronously, train.py:189-196; its Queue prefetches on the host only).

    `batches` yields tensors or (nested) tuples / dicts of tensors; the same structure comes back on `device`."""

    def __init__(self, batches, device):
        self.it, self.device = iter(batches), torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._next = self._ready = None
        self._preload()

    def _move(self, obj):
        if torch.is_tensor(obj):
            return obj.to(self.device, non_blocking=True)
        if isinstance(obj, dict):
            return {k: self._move(v) for k, v in obj.items()}
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._move(v) for v in obj)
        return obj

    def _record(self, obj, stream):
        if torch.is_tensor(obj):
            obj.record_stream(stream)      # allocated on the copy stream, consumed on the compute stream
        elif isinstance(obj, dict):
            for v in obj.values():
                self._record(v, stream)
        elif isinstance(obj, (tuple, list)):
            for v in obj:
                self._record(v, stream)

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            self._next = self._move(host)
            self._ready = torch.cuda.Event()
            self._ready.record(self.stream)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready)
        batch = self._next
        self._record(batch, cur)
        self._preload()
        return batch


def synthetic_volume(size, in_channels=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    vol = torch.randn((in_channels,) + tuple(size), generator=g)
    gt = (torch.nn.functional.avg_pool3d(vol[None, :1], 5, 1, 2)[0] > 0.25).to(torch.uint8)
    return vol, gt
