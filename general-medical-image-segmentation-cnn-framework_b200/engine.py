"""The training-step loop body of the reference (train.py:182-214) as one replayable unit.

    optimizer.zero_grad(); outputs = model(x); loss = criterion(outputs, y); accelerator.backward(loss);
    optimizer.step()                                                     (train.py:187-214)

A U-Net step is ~2 000 kernel launches of 5-500 us each; enqueueing them from Python costs about as much wall time as
running them.  `TrainStep` therefore captures the whole body -- forward, loss, backward, the bucketed NCCL gradient
all-reduce on its side stream, and the fused Adam update -- into ONE CUDA graph after a few eager warm-up steps and
replays it: the host's per-step work becomes two async copies into the static input buffers and one graph launch.
Everything inside is stream-ordered device work (no .item(), no host-side shape logic that depends on data), the
optimiser keeps its step counter and hyper-parameters in device memory, and every buffer comes from the capture's
private memory pool, so the TMA descriptors baked into the captured conv launches stay valid.
"""
import torch

from . import functional as F
from . import parallel


class TrainStep:
    def __init__(self, model, criterion, optimizer, use_graph=True, warmup=3):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.use_graph, self.warmup = use_graph, warmup
        self.graph = None
        self._x = self._y = self._loss = self._out = None
        self._seen = 0
        self.kernels_per_step = self.umma_per_step = 0   # b200seg kernels inside the captured graph

    # the reference's loop body: forward + loss + backward (captured) ...
    def _fwd_bwd(self, x, y):
        for seed in F._SEEDS.values():    # fresh dropout masks every step (device-side counter: graph-replay safe)
            seed.add_(1)
        self.optimizer.zero_grad()
        F.begin_step(x.device)            # one memset for every small accumulator of the step
        try:
            out = self.model(x)
            loss = self.criterion(out, y)
            loss.backward()
        finally:
            F.end_step()
        return loss.detach(), out.detach()

    # ... + gradient all-reduce + Adam.  On one GPU everything is captured.  On several, the SyncBatchNorm exchanges run
    # over NVLink peer memory inside the graph (csrc/p2p.cu) while the NCCL gradient all-reduce -- which cannot be
    # captured on this stack -- and the Adam launch follow the replay eagerly.
    def _update(self):
        scale = self.optimizer.all_reduce_grads()
        self.optimizer.step(grad_scale=scale)

    def _body(self, x, y):
        res = self._fwd_bwd(x, y)
        self._update()
        return res

    def _capture(self, x, y):
        self._x, self._y = x.clone(), y.clone()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        k0, u0 = F.launches(), F.umma_launch_count()
        # several GPUs: everything is captured when the optimiser exchanges gradients over NVLink peer memory
        # (optim.PeerGradExchange); with NCCL the exchange and the update follow the replay eagerly ("split")
        peer = getattr(self.optimizer, "peer_grads", False)
        self._split = parallel.is_parallel() and not peer
        reducer = getattr(self.optimizer, "reducer", None)
        if self._split and reducer is not None:
            reducer.enabled = False           # no NCCL launches from autograd hooks while capturing / replaying
        self._grad_scale = 1.0 / parallel.world_size() if (peer and parallel.is_parallel()) else 1.0
        if not self._split:
            self.optimizer._sync_hyper(self._grad_scale)   # a scheduler may have moved lr since the last eager step
        with torch.cuda.graph(self.graph):
            if self._split:
                # the packed weight gradients are transposed into the gradient arena INSIDE the graph: the host-side
                # "some packed gradient is pending" flag is only ever set by the Python backward, which a replay skips
                self._loss, self._out = self._fwd_bwd(self._x, self._y)
                self.optimizer.finalize_grads()
            else:
                self._loss, self._out = self._body(self._x, self._y)
        self.kernels_per_step, self.umma_per_step = F.launches() - k0, F.umma_launch_count() - u0
        torch.cuda.synchronize()

    def __call__(self, x, y):
        """x: fp32 NCDHW batch, y: labels; both already on the device (any memory).  Returns (loss, logits) as device
        tensors that are overwritten by the next call."""
        if not self.use_graph:
            return self._body(x, y)
        if self.graph is None:
            if not parallel.graph_safe():
                return self._body(x, y)           # SyncBatchNorm over NCCL: stay eager
            if self._seen < self.warmup:          # eager steps first: lazy initialisation must not be captured
                self._seen += 1
                return self._body(x, y)
            self._capture(x, y)                   # capture does not execute: fall through to the first replay
        elif x.shape != self._x.shape or y.shape != self._y.shape:
            return self._body(x, y)               # odd-sized last batch: run it eagerly
        self._x.copy_(x, non_blocking=True)
        self._y.copy_(y, non_blocking=True)
        if not self._split:
            # the captured Adam launch reads lr / betas / eps / weight decay from device memory: push host-side changes
            # (StepLR, train.py:119-120) before the replay -- a no-op when nothing changed
            self.optimizer._sync_hyper(self._grad_scale)
        self.graph.replay()
        if self._split:
            self._update()
        return self._loss, self._out
