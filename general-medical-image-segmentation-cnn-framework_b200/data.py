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

    Two sets of staging tensors are allocated once and alternate (no allocator traffic, no cudaMalloc in the loop); the
    copy stream waits until the work that was enqueued on the compute stream while a set was handed out has finished
    before overwriting it.  A batch is therefore valid until the next-but-one `next()`.
    `batches` yields tensors or (nested) tuples / dicts of tensors; the same structure comes back on `device`."""

    def __init__(self, batches, device):
        self.it, self.device = iter(batches), torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._home = torch.cuda.current_stream(self.device)
        self._slots = ([], [])             # staging tensors of the two sets, in traversal order
        self._released = [None, None]      # event after which a set may be overwritten
        self._k = 0
        self._next = self._ready = self._out_slot = None
        self._preload()

    def _move(self, obj, bufs, pos):
        if torch.is_tensor(obj):
            i = pos[0]
            pos[0] += 1
            if i == len(bufs):
                bufs.append(None)
            if bufs[i] is None or bufs[i].shape != obj.shape or bufs[i].dtype != obj.dtype:
                # allocated under the CALLER's stream (long-lived, event-guarded by hand): a fresh side stream would miss
                # the caching allocator's pools and cudaMalloc -- a device-wide synchronisation -- inside the loop
                with torch.cuda.stream(self._home):
                    bufs[i] = torch.empty(obj.shape, dtype=obj.dtype, device=self.device)
            bufs[i].copy_(obj, non_blocking=True)
            return bufs[i]
        if isinstance(obj, dict):
            return {k: self._move(v, bufs, pos) for k, v in obj.items()}
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._move(v, bufs, pos) for v in obj)
        return obj

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        slot = self._k & 1
        self._k += 1
        with torch.cuda.stream(self.stream):
            if self._released[slot] is not None:
                self.stream.wait_event(self._released[slot])
            self._next = (self._move(host, self._slots[slot], [0]), slot)
            self._ready = torch.cuda.Event()
            self._ready.record(self.stream)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        if self._out_slot is not None:     # everything that read the previous batch has been enqueued by now
            ev = torch.cuda.Event()
            ev.record(cur)
            self._released[self._out_slot] = ev
        cur.wait_event(self._ready)
        batch, self._out_slot = self._next
        self._preload()
        return batch


def synthetic_volume(size, in_channels=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    vol = torch.randn((in_channels,) + tuple(size), generator=g)
    gt = (torch.nn.functional.avg_pool3d(vol[None, :1], 5, 1, 2)[0] > 0.25).to(torch.uint8)
    return vol, gt
