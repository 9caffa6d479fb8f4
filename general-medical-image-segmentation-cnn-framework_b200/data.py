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


def synthetic_volume(size, in_channels=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    vol = torch.randn((in_channels,) + tuple(size), generator=g)
    gt = (torch.nn.functional.avg_pool3d(vol[None, :1], 5, 1, 2)[0] > 0.25).to(torch.uint8)
    return vol, gt
