"""Fused Adam over one flat fp32 parameter arena (torch.optim.Adam semantics; train.py:109,214).

All parameters (and their gradients) are re-homed as views into two contiguous buffers so that the optimiser step is
one kernel launch, zero_grad is one memset and the data-parallel gradient all-reduce is one NCCL call over the arena.
"""
import ctypes

import torch
import torch.distributed as dist

from .functional import _call, _ptr, _stream


class _RawDeviceBuffer:
    """A cudaMalloc'ed float buffer exposed through __cuda_array_interface__ (zero-copy view for torch.as_tensor)."""

    def __init__(self, ptr, numel):
        self.ptr, self.numel = ptr, numel
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerGradExchange:
    """Gradient all-reduce over NVLink peer memory (csrc/p2p.cu: b200seg_p2p_grad_allreduce), capturable in the training
    step's CUDA graph -- what NCCL's all-reduce is not on this stack.  Owns the IPC-shareable allocation that FusedAdam
    uses as its zeroed gradient arena, a staging buffer and the flag block, and the peers' mappings of all three."""

    def __init__(self, numel, group=None):
        from ._lib import call
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.numel = numel
        dev = torch.device("cuda", torch.cuda.current_device())
        chunks = (numel + 63) // 64
        sizes = {"buf": chunks * 64 * 4, "red": ((chunks + self.world - 1) // self.world) * 64 * 4, "flags": 256}
        self._own, handles = {}, {}
        for name, nbytes in sizes.items():
            ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            call("b200seg_p2p_alloc_bytes", nbytes, ctypes.byref(ptr), handle)
            self._own[name], handles[name] = ptr.value, handle.raw
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        self._tables = {}
        for name in sizes:
            ptrs = []
            for r, hs in enumerate(gathered):
                if r == self.rank:
                    ptrs.append(self._own[name])
                else:
                    q = ctypes.c_void_p()
                    call("b200seg_p2p_open", ctypes.create_string_buffer(hs[name], 64), ctypes.byref(q))
                    ptrs.append(q.value)
            self._tables[name] = (ctypes.c_void_p * self.world)(*ptrs)
        self.arena = torch.as_tensor(_RawDeviceBuffer(self._own["buf"], chunks * 64), device=dev)
        self.seq = torch.zeros(1, dtype=torch.int32, device=dev)
        self.live, self.live_mask, self.detections = None, None, 0
        torch.cuda.synchronize()
        dist.barrier(group=group)      # every buffer is mapped and zeroed before anybody's first exchange

    def set_live_chunks(self, mask):
        """mask: bool tensor over the 64-float chunks of the arena (True = receives gradients on some rank)."""
        m = mask.to(torch.int32)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)      # identical list on every rank
        self.live = torch.nonzero(m, as_tuple=False).flatten().to(torch.int32).contiguous()
        return int(self.live.numel())

    def all_reduce_(self):
        assert self.live is not None and self.live.numel() > 0
        _call("b200seg_p2p_grad_allreduce", self._tables["buf"], self._tables["red"], self._tables["flags"], _ptr(self.live),
              int(self.live.numel()), self.rank, self.world, _ptr(self.seq), _stream())


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, peer_grads=False, group=None):
        """peer_grads=True (multi-GPU): the gradient arena lives in CUDA-IPC shareable memory and `all_reduce_grads` sums
        it over NVLink peer memory inside the stream (CUDA-graph capturable) instead of calling NCCL."""
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        dev = self.params[0].device
        # 64-element (256 B) alignment per tensor keeps every view 16-byte aligned for vector access
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 63) // 64 * 64
        self.numel = total
        self.param_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        # gradients and the conv kernels' packed weight-gradient accumulators share one allocation: one memset clears both
        padded = (total + 63) // 64 * 64          # the packed accumulators start 256-byte aligned (128-bit loads of dw)
        self._padded = padded
        self.peer = None
        import os
        if peer_grads and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 \
                and os.environ.get("B200SEG_GRADS", "p2p") != "nccl":
            self.peer = PeerGradExchange(padded + total, group)
            self._zeroed = self.peer.arena[:padded + total]
        else:
            self._zeroed = torch.zeros(padded + total, dtype=torch.float32, device=dev)
        self.grad_arena = self._zeroed[:total]
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.dw_arena = self._zeroed[padded:]      # see functional._grad_target
        self._pending = [False]                    # some packed weight gradient has not been transposed yet
        for p, off in zip(self.params, self.offsets):
            view = self.param_arena[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad_arena[off:off + p.numel()].view_as(p)
            p._b200_arena_grad = True               # functional._arena_managed: backward may add into the slot directly
            if p.dim() == 5:      # Conv3d / ConvTranspose3d weights: backward accumulates straight into p.grad
                p._b200_dwp = self.dw_arena[off:off + p.numel()]
                p._b200_pending = self._pending
                p._b200_direct_grad = True
        self._build_pack_table()
        self.refresh_packs()

    # ---- bf16 weight packs for the conv kernels, refreshed by ONE launch after every update -------------------------
    def _build_pack_table(self):
        import struct
        recs, urecs, total = [], [], 0
        self._max_k3, self._tiles = 1, 0
        self._packed = []
        for p, off in zip(self.params, self.offsets):
            if p.dim() != 5 or p.shape[2] != p.shape[3] or p.shape[3] != p.shape[4]:
                continue
            a, b, k3 = p.shape[0], p.shape[1], p.shape[2] ** 3     # Conv3d [cout][cin][k^3]; ConvTranspose3d alike
            n = p.numel()
            self._max_k3 = max(self._max_k3, k3)
            self._tiles += ((a + 31) // 32) * ((b + 7) // 8)
            recs.append(struct.pack("<qqiiii", off, total, a, b, k3, 0))
            recs.append(struct.pack("<qqiiii", off, total + n, a, b, k3, 1))
            urecs.append(struct.pack("<qqiiii", off, off, a, b, k3, 0))     # dw_arena[off] -> grad_arena[off]
            self._packed.append((p, total, n))
            total += 2 * n
        self._build_adam_table()
        if not recs:
            self._pack_desc = None
            return
        dev = self.param_arena.device
        self._pack_arena = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self._pack_desc = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev)
        self._npack = len(recs)
        self._unpack_desc = torch.frombuffer(bytearray(b"".join(urecs)), dtype=torch.uint8).to(dev)
        self._nunpack = len(urecs)
        for p, o, n in self._packed:
            p._b200_pack0 = self._pack_arena[o:o + n]            # fprop layout [k^3][cout][cin]
            p._b200_pack1 = self._pack_arena[o + n:o + 2 * n]    # dgrad layout [k^3 flipped][cin][cout]

    def refresh_packs(self):
        """Re-pack every conv weight (call after changing parameters outside step(), e.g. load_state_dict)."""
        if self._pack_desc is None:
            return
        _call("b200seg_pack_weights_batched", _ptr(self.param_arena), _ptr(self._pack_arena), _ptr(self._pack_desc),
              self._npack, self._max_k3, 2 * self._tiles, _stream())
        for p, _, _ in self._packed:
            p._b200_pack_ver = p._version

    def _build_adam_table(self):
        """Records of the one-launch optimiser step (b200seg_adam_step_fused): a brick list over the conv weights (Adam +
        gradient transpose + both packs) and chunk lists over everything else."""
        import struct
        self._adam_desc = None
        if not self._packed:
            return
        packed = {id(p): o for p, o, _ in self._packed}
        tile_ci = 32 if self._max_k3 <= 27 else 8
        recs, tile0 = [], 0
        for p, off in zip(self.params, self.offsets):
            if id(p) in packed:
                a, b, k3 = p.shape[0], p.shape[1], p.shape[2] ** 3
                recs.append(struct.pack("<qqiiiiii", off, packed[id(p)], a, b, k3, 0, tile0, 0))
                tile0 += ((a + 31) // 32) * ((b + tile_ci - 1) // tile_ci)
            else:
                recs.append(struct.pack("<qqiiiiii", off, 0, p.numel(), 0, 0, 1, tile0, 0))
                tile0 += (p.numel() + 4095) // 4096
        self._adam_desc = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(self.param_arena.device)
        self._nadam, self._adam_tiles = len(recs), tile0

    def finalize_grads(self):
        """The conv kernels leave their weight gradients in the packed [tap][C_in][C_out] accumulators of dw_arena; one
        launch transposes all of them into the torch-layout gradient arena.  Runs before the all-reduce / the update."""
        if self._pending[0] and self._pack_desc is not None:
            _call("b200seg_unpack_wgrads_batched", _ptr(self.dw_arena), _ptr(self.grad_arena), _ptr(self._unpack_desc),
                  self._nunpack, self._max_k3, self._tiles, _stream())
        self._pending[0] = False

    def zero_grad(self, set_to_none=False):
        self._zeroed.zero_()
        self._pending[0] = False
        for p, off in zip(self.params, self.offsets):   # autograd accumulates in place into these views
            if p.grad is None or p.grad.data_ptr() != self.grad_arena.data_ptr() + 4 * off:
                p.grad = self.grad_arena[off:off + p.numel()].view_as(p)

    def attach_reducer(self, group=None, bucket_bytes=32 << 20):
        """Overlap the data-parallel gradient all-reduce with backward (parallel.GradBucketReducer)."""
        import struct
        from .parallel import GradBucketReducer
        self.reducer = GradBucketReducer(self.grad_arena, self.params, self.offsets, bucket_bytes, group,
                                         pre_launch=self._unpack_bucket)
        # per-bucket transposition tables: a bucket's packed conv weight gradients are moved into the gradient arena right
        # before that bucket is all-reduced (from the autograd hook of its last parameter, overlapping the rest of backward)
        self._bucket_unpack = []
        dev = self.param_arena.device
        for start, end, _ in self.reducer.buckets:
            recs, tiles = [], 0
            for p, off in zip(self.params, self.offsets):
                if start <= off < end and any(p is q for q, _, _ in self._packed):
                    a, b, k3 = p.shape[0], p.shape[1], p.shape[2] ** 3
                    recs.append(struct.pack("<qqiiii", off, off, a, b, k3, 0))
                    tiles += ((a + 31) // 32) * ((b + 7) // 8)
            desc = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev) if recs else None
            self._bucket_unpack.append((desc, len(recs), tiles))
        for p in self.params:
            p._b200_hooked = True       # functional: no postponed weight gradients, the hook order must mean "gradient final"
        return self.reducer

    def _unpack_bucket(self, b, start, end):
        desc, n, tiles = self._bucket_unpack[b]
        if desc is not None and self._pending[0]:
            _call("b200seg_unpack_wgrads_batched", _ptr(self.dw_arena), _ptr(self.grad_arena), _ptr(desc), n, self._max_k3,
                  tiles, _stream())

    def all_reduce_grads(self, group=None):
        """Data-parallel gradient averaging (what DDP does inside accelerator.backward, train.py:211).  Returns the
        factor the summed gradients still have to be scaled by (folded into the Adam kernel)."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return 1.0      # single process: the fused step consumes the packed weight gradients where they are
        if self.peer is not None:
            # NVLink peer-memory exchange of both arenas as they are (torch-layout gradients and the packed conv weight
            # gradients): no transpose pass, no NCCL call, capturable; the fused Adam kernel consumes both afterwards
            if torch.cuda.is_current_stream_capturing():
                if self.peer.live is None:
                    raise RuntimeError("FusedAdam: the first gradient exchange must run eagerly (it records which chunks "
                                       "of the arena receive gradients)")
            elif self.peer.detections < 3:
                # the first eager steps record (and accumulate) which chunks receive gradients; fixed afterwards
                mask = self._live_chunk_mask()
                if self.peer.live_mask is not None:
                    mask = mask | self.peer.live_mask
                self.peer.live_mask = mask
                self.peer.set_live_chunks(mask)
                self.peer.detections += 1
            self.peer.all_reduce_()
            return 1.0 / dist.get_world_size(group)
        if getattr(self, "reducer", None) is not None and self.reducer.enabled:
            scale = self.reducer.finish()      # every bucket transposes its own packed gradients before its all-reduce
            self._pending[0] = False
            return scale
        self.finalize_grads()
        dist.all_reduce(self.grad_arena, group=group)
        return 1.0 / dist.get_world_size(group)

    def _live_chunk_mask(self):
        """Which 64-float chunks of [gradient arena | packed weight-gradient arena] hold a gradient after this backward:
        a conv weight's gradient sits in the packed arena (direct accumulation) or the torch-layout one (autograd), never
        both; biases in front of batch statistics get none at all.  Decided from the data once, in an eager step."""
        z = self._zeroed
        n = (z.numel() + 63) // 64 * 64
        buf = self.peer.arena[:n] if self.peer is not None else torch.nn.functional.pad(z, (0, n - z.numel()))
        mask = buf.view(-1, 64).ne(0).any(dim=1)        # what the data shows in this step ...
        # ... plus what the structure says (a gradient that happens to be all-zero today -- dead ReLU channels of a tiny
        # bottleneck -- must still be exchanged tomorrow): every parameter's slot in the arena its gradient is written to
        for p, off in zip(self.params, self.offsets):
            base = self._padded + off if getattr(p, "_b200_dw_used", False) else off
            mask[base // 64:(base + p.numel() + 63) // 64] = True
        return mask

    @property
    def peer_grads(self):
        return self.peer is not None

    def _sync_hyper(self, grad_scale):
        """Hyper-parameters and the step counter live on the device so a CUDA-graph-captured step replays correctly;
        the host copy is only pushed when a value changed (e.g. StepLR, train.py:119-120)."""
        want = (float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                float(grad_scale))
        if getattr(self, "_hyper_host", None) != want:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: hyper-parameters changed inside a CUDA graph capture")
            if getattr(self, "_hyper", None) is None:
                self._hyper = torch.zeros(6, dtype=torch.float32, device=self.param_arena.device)
                self._state = torch.zeros(2, dtype=torch.int32, device=self.param_arena.device)
                self._state[0] = self.step_count
            self._hyper.copy_(torch.tensor(want, dtype=torch.float32))
            self._hyper_host = want

    def step(self, grad_scale=1.0):
        import os
        if getattr(self, "_adam_desc", None) is not None and not os.environ.get("B200SEG_DISABLE_FUSED_ADAM"):
            # one launch: packed weight gradients (if backward left any) + Adam + both bf16 packs of every conv weight
            self._sync_hyper(grad_scale)
            self.step_count += 1
            _call("b200seg_adam_step_fused", _ptr(self.param_arena), _ptr(self.grad_arena), _ptr(self.dw_arena),
                  _ptr(self.exp_avg), _ptr(self.exp_avg_sq), _ptr(self._pack_arena), _ptr(self._adam_desc), self._nadam,
                  self._max_k3, self._adam_tiles, _ptr(self._hyper), _ptr(self._state), 1 if self._pending[0] else 0,
                  _stream())
            self._pending[0] = False     # consumed; p.grad of the conv weights is NOT materialised on this path
            for p, _, _ in self._packed:
                p._b200_pack_ver = p._version
            return
        self.finalize_grads()
        self._sync_hyper(grad_scale)
        self.step_count += 1
        _call("b200seg_adam_step_dev", _ptr(self.param_arena), _ptr(self.grad_arena), _ptr(self.exp_avg),
              _ptr(self.exp_avg_sq), self.numel, _ptr(self._hyper), _ptr(self._state), _stream())
        self.refresh_packs()

    def state_dict(self):
        """torch.optim.Adam's schema ({"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}), so
        that a checkpoint written here resumes under the reference's `optimizer.load_state_dict` (train.py:128) and
        vice versa.  The moments are views into the flat arenas (torch.save writes them out as tensors)."""
        if getattr(self, "_state", None) is not None:
            self.step_count = int(self._state[0].item())   # graph replays advance the device counter only
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off:off + p.numel()].view_as(p),
                        "exp_avg_sq": self.exp_avg_sq[off:off + p.numel()].view_as(p)}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts torch.optim.Adam's state dict (a reference checkpoint; per-parameter state is matched by position in
        parameters() order and checked by shape) and the flat round-1 form {"step", "lr", "exp_avg", "exp_avg_sq"}."""
        if "param_groups" in sd:
            group = sd["param_groups"][0]
            if len(sd["param_groups"]) != 1 or len(group["params"]) != len(self.params):
                raise ValueError("FusedAdam.load_state_dict: expected one parameter group of %d tensors, got %s"
                                 % (len(self.params), [len(g["params"]) for g in sd["param_groups"]]))
            if group.get("amsgrad") or group.get("maximize"):
                raise ValueError("FusedAdam.load_state_dict: amsgrad / maximize are not supported")
            self.lr, self.betas = float(group["lr"]), tuple(float(b) for b in group["betas"])
            self.eps, self.weight_decay = float(group["eps"]), float(group["weight_decay"])
            steps = set()
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            for key, p, off in zip(group["params"], self.params, self.offsets):
                st = sd["state"].get(key)
                if st is None:        # a parameter that never received a gradient has no entry in torch's state
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError("FusedAdam.load_state_dict: state %r has shape %s, parameter has %s"
                                     % (key, tuple(st["exp_avg"].shape), tuple(p.shape)))
                self.exp_avg[off:off + p.numel()].view_as(p).copy_(st["exp_avg"])
                self.exp_avg_sq[off:off + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
            if len(steps) > 1:
                raise ValueError("FusedAdam.load_state_dict: per-parameter step counts differ (%s); the fused kernel keeps "
                                 "one counter" % sorted(steps))
            self.step_count = steps.pop() if steps else 0
        else:
            if sd["exp_avg"].numel() != self.numel:
                raise ValueError("FusedAdam.load_state_dict: arena of %d elements, checkpoint has %d"
                                 % (self.numel, sd["exp_avg"].numel()))
            self.step_count, self.lr = sd["step"], sd["lr"]
            self.betas = tuple(sd.get("betas", self.betas))
            self.eps, self.weight_decay = sd.get("eps", self.eps), sd.get("weight_decay", self.weight_decay)
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self._hyper_host = None
        if getattr(self, "_state", None) is not None:
            self._state[0] = self.step_count
