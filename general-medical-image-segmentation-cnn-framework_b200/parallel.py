"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the exchanges.

The reference gets data parallelism from HuggingFace accelerate -> DistributedDataParallel (train.py:167-169, 211:
the gradient all-reduce runs inside `accelerator.backward(loss)`), and cross-device batch-norm statistics from the
vendored thread/queue SyncBN (models/sync_batchnorm/batchnorm.py:90-111, comm.py:56-137).  Here:

  * `init_from_env()` joins the process group the way torchrun launches it (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
  * `GradBucketReducer` all-reduces the flat gradient arena of optim.FusedAdam in contiguous buckets.  A bucket is
    launched on a side stream as soon as the last gradient inside it has been produced (parameters are registered in
    forward order, backward produces them roughly in reverse, so buckets are cut from the END of the arena), which
    overlaps the exchange with the rest of the backward pass; `finish()` joins the side stream before the optimiser.
  * `all_reduce_stats()` is the SyncBatchNorm exchange: one all-reduce of [sum, sum-of-squares] (forward) or
    [sum dy, sum dy*xhat] (backward) per layer; every rank finalises locally, so the reference's master -> replica
    broadcast (batchnorm.py:105) disappears.
  * `shard_patches()` deals sliding-window patches round-robin to ranks and `reduce_volume()` merges the per-rank
    output volumes (predict.py:111 shards the loader the same way but never merges -- SURVEY.md section 2d).

Everything here is device-agnostic host logic (it is exercised with the gloo backend on CPU tensors in
tests/test_parallel_cpu.py); the arithmetic stays in the kernels.
"""
import os

import torch
import torch.distributed as dist


def is_parallel(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def world_size(group=None):
    return dist.get_world_size(group) if is_parallel(group) else 1


def rank(group=None):
    return dist.get_rank(group) if is_parallel(group) else 0


def init_from_env(backend=None):
    """Join the default process group from torchrun's environment.  Returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rk, world_size=world, **kwargs)
    return rk, local, world


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s weights and buffers (what DDP does at construction)."""
    if not is_parallel(group):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)


class PeerExchange:
    """Small all-reduce over NVLink peer memory (csrc/p2p.cu): one mailbox per rank, mapped by its peers through CUDA
    IPC; each exchange is a single one-CTA kernel per rank, capturable in the training step's CUDA graph, optionally
    fused with the batch-norm finalisation.  Replaces the master/slave pipes of sync_batchnorm/comm.py."""

    def __init__(self, group=None):
        import ctypes
        from ._lib import call
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        call("b200seg_p2p_alloc", ctypes.byref(own), handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(own.value)
            else:
                p = ctypes.c_void_p()
                call("b200seg_p2p_open", ctypes.create_string_buffer(h, 64), ctypes.byref(p))
                ptrs.append(p.value)
        self._ptrs = ptrs
        self.boxes = (ctypes.c_void_p * self.world)(*ptrs)
        self.seq = torch.zeros(1, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        dist.barrier(group=group)     # every mailbox is mapped (and zeroed) before anybody's first exchange

    def _call(self, vec, n, finalize_c=0, count=0.0, gamma=None, beta=None, running_mean=None, running_var=None,
              momentum=0.1, eps=1e-5, clamp_eps=False, coef=None, phase=0):
        import ctypes
        from .functional import _call, _ptr, _stream
        _call("b200seg_p2p_allreduce", _ptr(vec), int(n), self.boxes, self.rank, self.world, _ptr(self.seq), int(phase),
              int(finalize_c), float(count), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
              float(momentum), float(eps), int(clamp_eps), _ptr(coef), _stream())

    def all_reduce_(self, vec):
        """In-place sum over ranks of a contiguous fp32 vector of at most 2112 elements."""
        assert vec.is_contiguous() and vec.dtype == torch.float32
        self._call(vec, vec.numel())
        return vec

    def all_reduce_split_(self, vec, between=None):
        """all_reduce_ as a send launch and a receive launch with `between()` (kernels that do not depend on the result)
        enqueued in the middle: the NVLink round trip and the wait for the slowest rank hide behind them."""
        assert vec.is_contiguous() and vec.dtype == torch.float32
        self._call(vec, vec.numel(), phase=1)
        if between is not None:
            between()
        self._call(vec, vec.numel(), phase=2)
        return vec

    def reduce_and_finalize(self, stats, count, c, gamma, beta, running_mean, running_var, momentum, eps, clamp_eps):
        """stats: flat fp32 [2c + 1] whose first 2c entries are this rank's {sum, sumsq}.  Returns coef [1][4][c]."""
        coef = torch.empty((1, 4, c), dtype=torch.float32, device=stats.device)
        self._call(stats, 2 * c + 1, c, count, gamma, beta, running_mean, running_var, momentum, eps, clamp_eps, coef)
        return coef


_PEER = {}


def peer_exchange(group=None):
    """The process group's PeerExchange (created on first use), or None when the NVLink path is unavailable / disabled
    (B200SEG_SYNCBN=nccl) -- callers then fall back to an NCCL all-reduce, which cannot be graph-captured here."""
    if not is_parallel(group) or not torch.cuda.is_available() or os.environ.get("B200SEG_SYNCBN", "p2p") == "nccl":
        return None
    key = id(group) if group is not None else 0
    if key not in _PEER:
        try:
            _PEER[key] = PeerExchange(group)
        except Exception as e:     # no peer access between the GPUs of this box
            import warnings
            warnings.warn("b200seg: NVLink peer exchange unavailable (%s); SyncBatchNorm falls back to NCCL" % e)
            _PEER[key] = None
    return _PEER[key]


def graph_safe(group=None):
    """True when every collective inside the training step can be captured into a CUDA graph."""
    return not is_parallel(group) or peer_exchange(group) is not None


def all_reduce_stats(stats, group=None, between=None):
    """Sum a small fp32 statistics tensor over ranks, in place, in stream order.  Returns the number of ranks.
    between: optional callable enqueuing independent kernels; over NVLink peer memory it runs between the send and the
    receive launch of the exchange (PeerExchange.all_reduce_split_), otherwise before the collective."""
    if not is_parallel(group):
        if between is not None:
            between()
        return 1
    px = peer_exchange(group) if stats.is_cuda else None
    if px is not None and stats.numel() <= 2112 and stats.is_contiguous():
        px.all_reduce_split_(stats, between) if between is not None else px.all_reduce_(stats)
    else:
        if between is not None:
            between()
        dist.all_reduce(stats, group=group)
    return dist.get_world_size(group)


def all_reduce_counts(counts, group=None):
    """Integer metric counts (metric.py:36-46) summed over ranks: the TODO at train.py:220-224."""
    if is_parallel(group):
        dist.all_reduce(counts, group=group)
    return counts


class GradBucketReducer:
    """Bucketed, backward-overlapped all-reduce of a flat gradient arena.

    arena:    1-D tensor holding every gradient (optim.FusedAdam.grad_arena).
    params:   the parameters, in arena order; offsets[i] / numels[i] locate parameter i inside the arena.
    Buckets are contiguous arena ranges of at least `bucket_bytes`, cut from the last parameter backwards.
    """

    def __init__(self, arena, params, offsets, bucket_bytes=32 << 20, group=None, pre_launch=None):
        """pre_launch(bucket index, start, end): called on the caller's stream right before a bucket is all-reduced
        (optim.FusedAdam transposes that bucket's packed conv weight gradients into the arena there).
        NOTE: torch fires post-accumulate-grad hooks for a parameter also when its backward returned None, i.e. for the
        conv weights whose gradient is accumulated outside autograd -- every parameter counts, every step."""
        self.arena, self.group, self.pre_launch = arena, group, pre_launch
        self.params = list(params)
        self.offsets = list(offsets)
        self.enabled = is_parallel(group)
        self.buckets = []          # (start, end, [param indices])
        self._bucket_of = {}
        esz = arena.element_size()
        end, members, size = arena.numel(), [], 0
        for i in range(len(self.params) - 1, -1, -1):
            members.append(i)
            size += self.params[i].numel() * esz
            if size >= bucket_bytes or i == 0:
                self.buckets.append((self.offsets[i], end, members))
                end, members, size = self.offsets[i], [], 0
        for b, (_, _, members) in enumerate(self.buckets):
            for i in members:
                self._bucket_of[i] = b
        self._pending = [len(m) for _, _, m in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._stream = torch.cuda.Stream() if (self.enabled and arena.is_cuda) else None
        self._hooks = []
        if self.enabled:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _make_hook(self, i):
        def hook(_param):
            self.mark_ready(i)
        return hook

    def reset(self):
        self._pending = [len(m) for _, _, m in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []

    def mark_ready(self, i):
        """Gradient of parameter i is final (called from autograd hooks, or directly by fused backward kernels)."""
        if not self.enabled:
            return
        b = self._bucket_of[i]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._launched[b]:
            self._launch(b)

    def _launch(self, b, overlap=True):
        start, end, _ = self.buckets[b]
        view = self.arena[start:end]
        self._launched[b] = True
        if self.pre_launch is not None:
            self.pre_launch(b, start, end)
        if not overlap:
            # called from finish(): nothing left to overlap with, stay on the caller's stream
            dist.all_reduce(view, group=self.group)
        elif self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                dist.all_reduce(view, group=self.group)
        else:
            self._works.append(dist.all_reduce(view, group=self.group, async_op=True))

    def finish(self):
        """Launch whatever has not fired (parameters without gradient this step), join, return the averaging factor."""
        if not self.enabled:
            return 1.0
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b, overlap=False)
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        for w in self._works:
            w.wait()
        self.reset()
        return 1.0 / dist.get_world_size(self.group)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_patches(num_patches, group=None):
    """Indices of the sliding-window patches this rank runs (round-robin, predict.py:111)."""
    return list(range(rank(group), num_patches, world_size(group)))


def reduce_volume(volume, group=None, op="sum"):
    """Merge per-rank aggregator volumes in place.

    Average mode (op="sum"): logit sums and visit counts add.  Crop mode (op="max"): torchio lets a later patch
    overwrite an earlier one where their cropped interiors overlap (the extra window at the far border), so each
    voxel carries an int32 key ((global patch index + 1) << 8 | label); the maximum over ranks is the label written
    by the last patch in sampler order, exactly what a single sequential aggregator produces."""
    if is_parallel(group):
        dist.all_reduce(volume, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=group)
    return volume
