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


class GpuPatchSampler:
    """Training batches cut on the GPU from volumes that stay resident in HBM: torchio's ZNormalization +
    UniformSampler(patch_size) + Queue(samples_per_volume) (dataloader.py:52-67) without the host in the loop.

    volumes: list of fp32 [C, W, H, D] tensors (host or device); labels: list of [1, W, H, D] label maps.  Each volume is
    uploaded once and its z-normalisation constants (mean, unbiased std over all voxels) are computed once on the device;
    an epoch draws `samples_per_volume` uniformly placed patches per volume (start index uniform in [0, size - patch]
    per axis), shuffles them like the Queue does, and yields the reference's batch dict
    {"source": {"data": fp32 [B, C, pw, ph, pd]}, "gt": {"data": [B, 1, pw, ph, pd]}} -- already on the device, normalised
    inside the cropping kernel.  drop_last like the reference's DataLoader (train.py:158)."""

    def __init__(self, volumes, labels, patch_size, batch_size, samples_per_volume=10, device="cuda", seed=0, shuffle=True):
        from .functional import _call, _ptr, _stream
        assert len(volumes) == len(labels) and len(volumes) > 0
        self.device = torch.device(device)
        self.patch_size, self.batch_size = tuple(int(p) for p in patch_size), int(batch_size)
        self.samples_per_volume, self.shuffle = int(samples_per_volume), shuffle
        self.gen = torch.Generator().manual_seed(seed)
        self.volumes, self.labels, self.norms = [], [], []
        for v, lab in zip(volumes, labels):
            v = v.to(self.device, torch.float32).contiguous()
            lab = lab.to(self.device).to(torch.uint8).contiguous()
            if v.dim() != 4 or lab.dim() != 4 or tuple(lab.shape[1:]) != tuple(v.shape[1:]):
                raise ValueError("expected [C, W, H, D] images and [1, W, H, D] label maps of the same extent")
            if any(p > s for p, s in zip(self.patch_size, v.shape[1:])):
                raise ValueError("patch size %s is larger than a volume of extent %s" % (self.patch_size, tuple(v.shape[1:])))
            sums = torch.zeros(2, dtype=torch.float64, device=self.device)
            norm = torch.empty(2, dtype=torch.float32, device=self.device)
            _call("b200seg_volume_stats", _ptr(v), v.numel(), _ptr(sums), _stream())
            _call("b200seg_znorm_finalize", _ptr(sums), v.numel(), _ptr(norm), _stream())
            self.volumes.append(v)
            self.labels.append(lab)
            self.norms.append(norm)

    def __len__(self):
        return len(self.volumes) * self.samples_per_volume // self.batch_size

    def draw_locations(self):
        """[(volume index, x0, y0, z0)] of one epoch: samples_per_volume uniform locations per volume, shuffled."""
        locs = []
        for vi, v in enumerate(self.volumes):
            for _ in range(self.samples_per_volume):
                locs.append((vi,) + tuple(int(torch.randint(0, s - p + 1, (1,), generator=self.gen))
                                          for s, p in zip(v.shape[1:], self.patch_size)))
        if self.shuffle:
            order = torch.randperm(len(locs), generator=self.gen).tolist()
            locs = [locs[i] for i in order]
        return locs

    def crop(self, picks):
        from .functional import _call, _ptr, _stream
        c = self.volumes[0].shape[0]
        pw, ph, pd = self.patch_size
        x = torch.empty((len(picks), c, pw, ph, pd), dtype=torch.float32, device=self.device)
        gt = torch.empty((len(picks), 1, pw, ph, pd), dtype=torch.uint8, device=self.device)
        for b, (vi, x0, y0, z0) in enumerate(picks):
            v, lab = self.volumes[vi], self.labels[vi]
            _call("b200seg_crop_patch", _ptr(v), 0, v.shape[0], v.shape[1], v.shape[2], v.shape[3], x0, y0, z0, pw, ph, pd,
                  _ptr(self.norms[vi]), _ptr(x[b]), _stream())
            _call("b200seg_crop_patch", _ptr(lab), 1, 1, v.shape[1], v.shape[2], v.shape[3], x0, y0, z0, pw, ph, pd, None,
                  _ptr(gt[b]), _stream())
        return {"source": {"data": x}, "gt": {"data": gt}}

    def __iter__(self):
        locs = self.draw_locations()
        for s in range(0, len(locs) - self.batch_size + 1, self.batch_size):
            yield self.crop(locs[s:s + self.batch_size])


def synthetic_volume(size, in_channels=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    vol = torch.randn((in_channels,) + tuple(size), generator=g)
    gt = (torch.nn.functional.avg_pool3d(vol[None, :1], 5, 1, 2)[0] > 0.25).to(torch.uint8)
    return vol, gt
