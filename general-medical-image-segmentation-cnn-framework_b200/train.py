"""Training entry point with the reference's command line (`python train.py config=unet config.key=value`,
train.py:310-388) and loop (train.py:90-307) on the b200seg hot path.

    python -m b200seg.train config=unet config.batch_size=2 config.patch_size=128,128,128 config.epochs=1
    torchrun --nproc-per-node 8 -m b200seg.train config=unet ...       (one process per GPU)

Kept: the model factory, `weights_init_normal`, Adam + StepLR, the BCE-on-one-hot criterion, the per-iteration Dice,
the checkpoint dict {"model", "optim", "scheduler", "epoch"} and file names.  Replaced: accelerate/DDP by
b200seg.parallel, the loop body by engine.TrainStep (one CUDA graph per step), the CPU numpy metric by GPU counts whose
4 integers are all-reduced over ranks (the TODO at train.py:220-224), torchio data by data.SyntheticPatches.
"""
import os
import sys
import time

import torch

from . import parallel
from .config import build_model, compose, weights_init_normal
from .data import DevicePrefetcher, GpuPatchSampler, SyntheticPatches, synthetic_volume
from .engine import TrainStep
from .models.sync_batchnorm.batchnorm import convert_model
from .optim import FusedAdam
from .utils import loss_function as L


class StepLR:
    """torch.optim.lr_scheduler.StepLR for FusedAdam (train.py:119-120, 259-261)."""

    def __init__(self, optimizer, step_size, gamma):
        self.optimizer, self.step_size, self.gamma = optimizer, step_size, gamma
        self.base_lr, self.last_epoch = optimizer.lr, 0

    def step(self):
        self.last_epoch += 1
        self.optimizer.lr = self.base_lr * self.gamma ** (self.last_epoch // self.step_size)

    def get_last_lr(self):
        return [self.optimizer.lr]

    def state_dict(self):
        """Same keys as torch.optim.lr_scheduler.StepLR.state_dict() (what the reference checkpoints, train.py:288-294)."""
        return {"step_size": self.step_size, "gamma": self.gamma, "base_lrs": [self.base_lr], "last_epoch": self.last_epoch,
                "_step_count": self.last_epoch + 1, "_last_lr": [self.optimizer.lr]}

    def load_state_dict(self, sd):
        """Accepts torch's StepLR state ({"base_lrs": [..], ...}, a reference checkpoint) and the round-1 form."""
        self.step_size, self.gamma, self.last_epoch = sd["step_size"], sd["gamma"], sd["last_epoch"]
        self.base_lr = sd["base_lrs"][0] if "base_lrs" in sd else sd["base_lr"]
        self.optimizer.lr = self.base_lr * self.gamma ** (self.last_epoch // self.step_size)


def make_criterion(name, classes):
    if name == "bce":
        return L.Binary_Loss()
    if name == "dice_ce":
        return L.DiceCELoss(classes)
    if name == "ce":
        return lambda p, t: L.cross_entropy_3D(p, t)
    if name == "dice":
        return lambda p, t: L.DiceLossss(classes)(p, t, softmax=True)
    raise ValueError("unknown criterion %r" % name)


def train(config, model, log=print):
    from . import functional as F
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    model = model.to(dev)
    if world > 1:
        parallel.broadcast_parameters(model)
        if config.sync_bn:
            convert_model(model)
    optimizer = FusedAdam(model.parameters(), lr=config.init_lr, peer_grads=world > 1)
    if world > 1 and not optimizer.peer_grads:
        optimizer.attach_reducer()
    criterion = make_criterion(config.criterion, config.out_classes)
    scheduler = StepLR(optimizer, config.scheduler_step_size, config.scheduler_gamma) if config.use_scheduler else None
    elapsed_epochs = 0
    if config.load_mode == 1:
        ckpt = torch.load(config.ckpt, map_location="cpu")
        model.load_state_dict(ckpt["model"])
        optimizer.load_state_dict(ckpt["optim"])
        if scheduler is not None and ckpt.get("scheduler"):
            scheduler.load_state_dict(ckpt["scheduler"])
        elapsed_epochs = ckpt["epoch"]
    model.train()
    step = TrainStep(model, criterion, optimizer, use_graph=bool(config.cuda_graph))
    if str(config.get("data", "synthetic")) == "volumes":
        # device-resident volumes: z-normalisation + uniform patch crop on the GPU (dataloader.py:52-67)
        vols = [synthetic_volume(config.volume_size, config.in_classes, seed=config.seed + 17 * rank + i)
                for i in range(int(config.get("train_volumes", 2)))]
        loader = GpuPatchSampler([v for v, _ in vols], [g for _, g in vols], config.patch_size, config.batch_size,
                                 int(config.get("samples_per_volume", 10)), device=dev, seed=config.seed + rank)
    else:
        loader = SyntheticPatches(config.patch_size, config.batch_size, config.iters_per_epoch, config.in_classes,
                                  seed=config.seed + rank)
    os.makedirs(config.hydra_path, exist_ok=True)
    history = []
    for epoch in range(elapsed_epochs + 1, config.epochs + 1):
        t0, loss_sum, dice_sum, n = time.time(), 0.0, 0.0, 0
        batches = loader if isinstance(loader, GpuPatchSampler) else DevicePrefetcher(loader, dev)
        for i, batch in enumerate(batches):                         # host batches: the copy of batch i+1 overlaps step i
            x = batch["source"]["data"]
            gt = batch["gt"]["data"]
            labels = gt.reshape(gt.shape[0], *gt.shape[2:]).to(torch.uint8)   # the one-hot of train.py:191-193, as indices
            loss, pred = step(x, labels)
            mask = F.argmax_labels(pred)                                      # train.py:204
            counts = parallel.all_reduce_counts(F.seg_counts(labels, mask))   # train.py:221 (+ the TODO at :220)
            gsum, psum, inter, _ = counts.tolist()
            dice = 2 * inter / (gsum + psum + 0.001)
            loss_sum += float(loss)
            dice_sum += dice
            n += 1
        if scheduler is not None:
            scheduler.step()
        history.append((epoch, loss_sum / max(n, 1), dice_sum / max(n, 1)))
        if rank == 0:
            log("Epoch %d used time: %.3f s  Loss Avg: %.5f  Dice Avg: %.4f  lr %.6f" %
                (epoch, time.time() - t0, loss_sum / max(n, 1), dice_sum / max(n, 1), optimizer.lr))
            state = {"model": model.state_dict(), "optim": optimizer.state_dict(),
                     "scheduler": scheduler.state_dict() if scheduler is not None else None, "epoch": epoch}
            torch.save(state, os.path.join(config.hydra_path, config.latest_checkpoint_file))
            if epoch % config.epochs_per_checkpoint == 0:
                torch.save(state, os.path.join(config.hydra_path, "checkpoint_%04d.pt" % epoch))
    return history


def main(argv=None):
    config = compose(sys.argv[1:] if argv is None else argv)
    model = build_model(config)
    model.apply(weights_init_normal(config.init_type))
    return train(config, model)


if __name__ == "__main__":
    main()
