"""Autograd-facing wrappers around the C ABI (include/b200seg.h).

Activations travel between these functions as channels-last bf16 tensors of shape [N, D, H, W, C]; the innermost
dimension is contiguous and consecutive voxels are `pitch` elements apart, so a tensor may be a channel slice of a
wider buffer (that is how the U-Net skip concatenation is made copy-free).  Parameters, statistics and parameter
gradients are fp32.  PyTorch supplies memory, streams and the autograd tape; all arithmetic is in the kernels.
"""
import contextlib
import ctypes
import os

import torch
import torch.distributed as dist

from ._lib import ConvGeom, call
from .parallel import all_reduce_stats, is_parallel, peer_exchange

ACT = {"none": 0, None: 0, "relu": 1, "leaky_relu": 2, "elu": 3, "prelu": 4}

# kernels launched since the last reset (bench.py reports it as gpu_launches)
_LAUNCHES = [0]


def launches():
    return _LAUNCHES[0]


def reset_launches():
    _LAUNCHES[0] = 0


# optional per-kernel timing (bench.py roofline): list of (name, work, start_event, end_event)
_PROFILE = [None]


def profile_begin():
    _PROFILE[0] = []


def profile_end():
    rec, _PROFILE[0] = _PROFILE[0], None
    torch.cuda.synchronize()
    out = {}
    for name, work, e0, e1 in rec:
        t = out.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
        t["launches"] += 1
        t["ms"] += e0.elapsed_time(e1)
        t["work"] += work
    return out


def _call(name, *args, work=0.0, tag=None):
    _LAUNCHES[0] += 1
    if _PROFILE[0] is None:
        call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    _PROFILE[0].append((tag or name, work, e0, e1))


# Zero-initialised scratch for the many small accumulators of a step (per-layer statistics, backward partial sums):
# engine.TrainStep clears ONE buffer at the start of the step and the helpers below hand out 256-byte aligned slices of
# it, instead of launching a fill kernel per accumulator (~60 per U-Net step).  Outside a step (or when it is full)
# they fall back to torch.zeros.
_SCRATCH = {"buf": None, "pos": 0, "active": False}


def begin_step(device, nbytes=4 << 20):
    if _SCRATCH["buf"] is None or _SCRATCH["buf"].device != torch.device(device) or _SCRATCH["buf"].numel() < nbytes // 4:
        _SCRATCH["buf"] = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
    _SCRATCH["buf"].zero_()
    _SCRATCH["pos"], _SCRATCH["active"] = 0, True
    _COLSUMS.clear()
    _ZERO_TAIL.clear()
    _GRAD_STASH.clear()


def end_step():
    flush_deferred()
    _SCRATCH["active"] = False
    _COLSUMS.clear()
    _ZERO_TAIL.clear()
    _GRAD_STASH.clear()
    _LAST_PADDED_INPUT[0] = None
    flush_counters()


# Weight-gradient launches that were postponed by one layer (multi-GPU SyncBatchNorm only): the backward statistics
# exchange of the NEXT layer in backward order is a send launch and a receive launch (parallel.PeerExchange.
# all_reduce_split_), and the postponed weight gradient -- which nothing in that chain depends on -- runs in between, so
# the NVLink round trip and the wait for the slowest rank cost nothing.  Only inside a training step (begin_step ...
# end_step flushes the last one) and only for weights whose gradient is accumulated directly into the optimiser's arena.
_DEFERRED = []


def flush_deferred():
    while _DEFERRED:
        _DEFERRED.pop(0)()


_COUNTERS = []


def bump_counter(t):
    """num_batches_tracked += 1 (what nn.BatchNorm3d.forward does in training mode).  Inside a training step the
    increments of all norm layers are collected and applied by ONE multi-tensor launch at the end of the step instead of
    one tiny launch per layer."""
    if _SCRATCH["active"] and t.is_cuda:
        _COUNTERS.append(t)
    else:
        t += 1


def flush_counters():
    if _COUNTERS:
        torch._foreach_add_(list(_COUNTERS), 1)
        _COUNTERS.clear()


def _zeros_f32(numel, device):
    if _SCRATCH["active"] and _SCRATCH["buf"].device == torch.device(device):
        pos = _SCRATCH["pos"]
        nxt = pos + (numel + 63) // 64 * 64
        if nxt <= _SCRATCH["buf"].numel():
            _SCRATCH["pos"] = nxt
            return _SCRATCH["buf"][pos:pos + numel]
    return torch.zeros(numel, dtype=torch.float32, device=device)


def umma_launch_count():
    from ._lib import load
    return load().b200seg_umma_launch_count()


def conv_uses_tensor_cores(g):
    from ._lib import load
    return bool(load().b200seg_conv3d_uses_tensor_cores(ctypes.byref(g)))


def _conv_flops(g):
    return 2.0 * g.n * g.od * g.oh * g.ow * g.cout * g.cin * g.k ** 3


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _pitched(t):
    """True when t is a [N,D,H,W,C] bf16 view with unit channel stride and a uniform voxel pitch."""
    if t.dim() != 5 or t.stride(4) != 1:
        return False
    n, d, h, w, c = t.shape
    p = t.stride(3)
    return p >= c and t.stride(2) == w * p and t.stride(1) == h * w * p and t.stride(0) == d * h * w * p


def _as_rows(t):
    """Return (tensor usable by the kernels, pitch)."""
    assert t.is_cuda and t.dtype == torch.bfloat16, "expected a CUDA bf16 NDHWC tensor"
    if not _pitched(t) or (t.data_ptr() % 16 and t.shape[4] % 8 == 0 and t.stride(3) % 8 == 0):
        t = t.contiguous()   # the 128-bit kernels need 16-byte aligned rows
    return t, t.stride(3)


def _adjacent(a, b):
    """a and b are back-to-back channel slices of one wider NDHWC buffer."""
    return (_pitched(a) and _pitched(b) and a.shape[:4] == b.shape[:4] and a.stride() == b.stride()
            and b.data_ptr() == a.data_ptr() + a.shape[4] * a.element_size() and a.stride(3) >= a.shape[4] + b.shape[4])


def merge_channels(a, b):
    """cat((a, b), channel) -- free when the two are adjacent slices of one buffer (see alloc_concat)."""
    if _adjacent(a, b):
        n, d, h, w, c = a.shape
        return a.as_strided((n, d, h, w, c + b.shape[4]), a.stride())
    return torch.cat((a, b), dim=4)


def alloc_concat(n, d, h, w, c_first, c_second, device):
    """One [n,d,h,w,c_first+c_second] buffer and its two channel-slice views (torch.cat replacement, unet3d.py:59)."""
    buf = torch.empty((n, d, h, w, c_first + c_second), dtype=torch.bfloat16, device=device)
    return buf, buf[..., :c_first], buf[..., c_first:]


def conv_out(size, k, stride, pad, dil):
    return (size + 2 * pad - dil * (k - 1) - 1) // stride + 1


def _geom(x_shape, cin, cout, k, stride, pad, dil):
    n, d, h, w = x_shape[:4]
    return ConvGeom(n, d, h, w, cin, conv_out(d, k, stride, pad, dil), conv_out(h, k, stride, pad, dil),
                    conv_out(w, k, stride, pad, dil), cout, k, stride, pad, dil)


def _cached_pack(weight, variant):
    """bf16 pack kept fresh by optim.FusedAdam (one batched launch per step), valid while the parameter has not been
    modified through torch since (its autograd version is unchanged)."""
    if getattr(weight, "_b200_pack_ver", None) == weight._version:
        return weight._b200_pack1 if variant else weight._b200_pack0
    return None


# Inference over many batches with fixed parameters (sliding-window prediction: 10 batches per volume, predict.py:112-139):
# inside `frozen_parameters()` the bf16 weight packs and the eval-mode normalisation constants are computed once and reused,
# instead of once per batch (18 pack + 18 constant launches per U-Net forward).  Keys are object ids; the entries hold the
# objects, so an id cannot be recycled while the cache lives.
_FROZEN = [None]


@contextlib.contextmanager
def frozen_parameters():
    prev = _FROZEN[0]
    if prev is None:
        _FROZEN[0] = {}
    try:
        yield
    finally:
        _FROZEN[0] = prev


def _frozen_get(key, keep, make):
    cache = _FROZEN[0]
    if cache is None or torch.is_grad_enabled():
        return make()
    hit = cache.get(key)
    if hit is None:
        hit = cache[key] = (keep, make())
    return hit[1]


def pack_conv_weight(weight, cin_off=0, cin_cnt=None, dgrad=False):
    cout, cin, k = weight.shape[0], weight.shape[1], weight.shape[2]
    cin_cnt = cin if cin_cnt is None else cin_cnt
    if cin_off == 0 and cin_cnt == cin:
        cached = _cached_pack(weight, 1 if dgrad else 0)
        if cached is not None:
            return cached

    def make():
        w = weight.detach()
        packed = torch.empty(k ** 3 * cout * cin_cnt, dtype=torch.bfloat16, device=w.device)
        _call("b200seg_pack_conv_weight", _ptr(w), _ptr(packed), cout, cin, k, cin_off, cin_cnt, int(dgrad), _stream())
        return packed
    return _frozen_get(("pack", id(weight), cin_off, cin_cnt, bool(dgrad)), weight, make)


# ------------------------------------------------------------------------------------------------ raw op helpers
def _pad16(c):
    return (c + 15) // 16 * 16


def _padded_geom(g):
    return ConvGeom(g.n, g.d, g.h, g.w, _pad16(g.cin), g.od, g.oh, g.ow, _pad16(g.cout), g.k, g.stride, g.pad, g.dil)


def _use_padded(g):
    """Channel counts that are not multiples of 16 (C_in = 1 stems, DenseVoxelNet's 12-channel growth and 16+12i inputs,
    V-Net's 32 -> 2 output conv: densevoxelnet3d.py:22, vnet3d.py:47,111) reach the tensor cores zero-padded: the weights
    get zero rows / columns, activations and gradients a zero-filled copy.  Only where it pays (large volumes)."""
    if g.cin % 16 == 0 and g.cout % 16 == 0:
        return False
    if g.n * g.od * g.oh * g.ow < (1 << 16) and _conv_flops(g) < 2e8:     # tiny: the direct kernels are launch-bound
        return False
    return conv_uses_tensor_cores(_padded_geom(g))


# Tensors that are channel slices [..., :C] of a wider contiguous buffer whose remaining channels are known to be zero
# (written so by the producing kernel, see _Dropout.backward): _pad_channels hands out the wide buffer instead of copying.
# Entries hold the wide tensor, so its memory cannot be recycled under the key (same scheme as _COLSUMS below).
_ZERO_TAIL = {}


def _publish_zero_tail(view, wide):
    if len(_ZERO_TAIL) >= 16:
        _ZERO_TAIL.clear()
    _ZERO_TAIL[(view.data_ptr(), tuple(view.shape), tuple(view.stride()))] = wide


def _pad_channels(x, c_pad):
    """[N,D,H,W,C] -> contiguous [N,D,H,W,c_pad] with zero-filled extra channels."""
    if x.dim() == 5 and _ZERO_TAIL:
        wide = _ZERO_TAIL.pop((x.data_ptr(), tuple(x.shape), tuple(x.stride())), None)
        if wide is not None and wide.shape[4] == c_pad:
            return wide
    x, xp = _as_rows(x) if (x.dim() == 5 and x.stride(4) == 1 and _pitched(x)) else (x.contiguous(), x.shape[4])
    n, d, h, w, c = x.shape
    if c == c_pad and xp == c:
        return x
    out = torch.empty((n, d, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    _call("b200seg_pad_channels", _ptr(x), xp, c, _ptr(out), c_pad, n * d * h * w, _stream())
    return out


def _pack_padded(weight, cin_p, cout_p, dgrad):
    """bf16 fprop / dgrad pack of `weight` widened with zero rows / columns to cout_p x cin_p, one launch."""
    w = weight.detach()
    cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
    packed = torch.empty(k ** 3 * cout_p * cin_p, dtype=torch.bfloat16, device=w.device)
    _call("b200seg_pack_conv_weight_padded", _ptr(w), _ptr(packed), cout, cin, k, cout_p, cin_p, int(dgrad), _stream())
    return packed


def _unpad_stats(stats_p, c, c_pad, spare):
    if c == c_pad:
        return stats_p
    parts = [stats_p[:c], stats_p[c_pad:c_pad + c]]
    if spare:
        parts.append(stats_p[2 * c_pad:2 * c_pad + spare])
    return torch.cat(parts)


_LAST_PADDED_INPUT = [None]


def _conv_scratch(g, dgrad, device):
    """Scratch for the weights-stationary kernel of the K-heavy layers on 8 x 8 planes (b200seg_conv3d_ws_bytes): an fp32
    copy of the output in which the partial sums of the tap rows meet.  (None, 0) for every other geometry."""
    from ._lib import load
    nbytes = int(load().b200seg_conv3d_ws_bytes(ctypes.byref(g), int(dgrad)))
    if not nbytes:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def conv3d_fprop_raw(x, weight, bias, k, stride, pad, dil, want_stats, y_out=None):
    x, xp = _as_rows(x)
    cout, cin = weight.shape[0], weight.shape[1]
    assert x.shape[4] == cin, "input has %d channels, weight expects %d" % (x.shape[4], cin)
    g = _geom(x.shape, cin, cout, k, stride, pad, dil)
    b = bias.detach().float() if bias is not None else None
    if _use_padded(g):
        gp = _padded_geom(g)
        x_p = _pad_channels(x, gp.cin)
        if gp.cout == cout:     # only K is padded (stems): the pack kernel zero-fills the extra input channels itself
            wp = torch.empty(k ** 3 * cout * gp.cin, dtype=torch.bfloat16, device=x.device)
            _call("b200seg_pack_conv_weight", _ptr(weight.detach()), _ptr(wp), cout, cin, k, 0, gp.cin, 0, _stream())
        else:
            wp = _pack_padded(weight, gp.cin, gp.cout, False)
        b_p = b
        if b is not None and gp.cout != cout:
            b_p = b.new_zeros(gp.cout)
            b_p[:cout] = b
        y_p = torch.empty((g.n, g.od, g.oh, g.ow, gp.cout), dtype=torch.bfloat16, device=x.device)
        stats_p = _zeros_f32(2 * gp.cout + 1, x.device) if want_stats else None
        _call("b200seg_conv3d_fprop", ctypes.byref(gp), _ptr(x_p), gp.cin, _ptr(wp), _ptr(b_p), _ptr(y_p), gp.cout,
              _ptr(stats_p), None, 0, _stream(), work=_conv_flops(g), tag="conv_fprop_padded_tc")
        y = y_p[..., :cout] if gp.cout != cout else y_p
        _LAST_PADDED_INPUT[0] = x_p      # _ConvNormAct keeps the widened input for the weight gradient
        return y, (_unpad_stats(stats_p, cout, gp.cout, 1) if want_stats else None), g
    if y_out is not None and _pitched(y_out) and tuple(y_out.shape) == (g.n, g.od, g.oh, g.ow, cout) \
            and y_out.data_ptr() % 16 == 0 and y_out.stride(3) % 8 == 0:
        y = y_out
    else:
        y = torch.empty((g.n, g.od, g.oh, g.ow, cout), dtype=torch.bfloat16, device=x.device)
    # flat {sum[C], sumsq[C], (count)}: the spare float lets the cross-GPU exchange carry the element count
    stats = _zeros_f32(2 * cout + 1, x.device) if want_stats else None
    wp = pack_conv_weight(weight)
    ws, ws_bytes = _conv_scratch(g, 0, x.device)
    _call("b200seg_conv3d_fprop", ctypes.byref(g), _ptr(x), xp, _ptr(wp), _ptr(b), _ptr(y), y.stride(3), _ptr(stats),
          _ptr(ws), ws_bytes, _stream(), work=_conv_flops(g),
          tag="conv_fprop_tc" if conv_uses_tensor_cores(g) else "conv_fprop_direct")
    return y, stats, g


def conv3d_dgrad_raw(g, dy, weight, colsum=False):
    """dx of the conv described by g.  colsum=True also returns {sum[C_in], sumsq[C_in]} of dx over the voxels (fp32,
    from the same epilogue): the first row is the bias gradient of whatever produced the conv's input."""
    if _use_padded(g):
        gp = _padded_geom(g)
        dy_p = _pad_channels(dy, gp.cout)
        wd = _pack_padded(weight, gp.cin, gp.cout, True)
        dx_p = torch.empty((g.n, g.d, g.h, g.w, gp.cin), dtype=torch.bfloat16, device=dy.device)
        stats_p = _zeros_f32(2 * gp.cin, dy.device) if colsum else None
        _call("b200seg_conv3d_dgrad", ctypes.byref(gp), _ptr(dy_p), gp.cout, _ptr(wd), _ptr(dx_p), gp.cin, _ptr(stats_p),
              None, 0, _stream(), work=_conv_flops(g), tag="conv_dgrad_padded")
        dx = dx_p[..., :g.cin] if gp.cin != g.cin else dx_p
        return (dx, _unpad_stats(stats_p, g.cin, gp.cin, 0)) if colsum else dx
    dy, dyp = _as_rows(dy)
    wd = pack_conv_weight(weight, dgrad=True)
    dx = torch.empty((g.n, g.d, g.h, g.w, g.cin), dtype=torch.bfloat16, device=dy.device)
    stats = _zeros_f32(2 * g.cin, dy.device) if colsum else None
    ws, ws_bytes = _conv_scratch(g, 1, dy.device)
    _call("b200seg_conv3d_dgrad", ctypes.byref(g), _ptr(dy), dyp, _ptr(wd), _ptr(dx), g.cin, _ptr(stats), _ptr(ws), ws_bytes,
          _stream(), work=_conv_flops(g), tag="conv_dgrad")
    return (dx, stats) if colsum else dx


# Column sums of gradients that a conv's dgrad epilogue already produced, keyed by the gradient tensor they describe:
# ConvTranspose3d.backward takes its bias gradient from here instead of re-reading dy (unet3d.py:58-59: the up-conv output
# is the first half of the decoder conv's input).  Entries hold the tensor, so its memory cannot be recycled under the key.
_COLSUMS = {}


def _publish_colsum(t, sums):
    if len(_COLSUMS) >= 16:    # nobody took them (the producers were not up-convolutions): do not pin their memory
        _COLSUMS.clear()
    _COLSUMS[(t.data_ptr(), tuple(t.shape), tuple(t.stride()))] = (t, sums)


def _take_colsum(t):
    hit = _COLSUMS.pop((t.data_ptr(), tuple(t.shape), tuple(t.stride())), None)
    return None if hit is None else hit[1]


def _arena_managed(param):
    """The parameter's .grad is a view into optim.FusedAdam's flat gradient arena (zeroed by zero_grad every step)."""
    return param is not None and getattr(param, "_b200_arena_grad", False) and param.grad is not None


def _affine_grad_target(gamma, beta):
    """Where the backward reduction can add d(gamma) | d(beta) directly: the two gradients are adjacent in the arena
    (norm.weight is followed by norm.bias in parameters() order).  None: return them to autograd as usual."""
    if not (_arena_managed(gamma) and _arena_managed(beta)):
        return None
    gg, gb = gamma.grad, beta.grad
    if gg.dtype != torch.float32 or gb.data_ptr() != gg.data_ptr() + 4 * gg.numel() or gb.numel() != gg.numel():
        return None
    return gg


def _grad_target(param):
    """Packed fp32 accumulator ([tap][C_in][C_out]) of a conv weight managed by optim.FusedAdam, else None.  The
    accumulator is a slice of an arena that FusedAdam.zero_grad clears together with the gradients; the optimiser
    transposes ALL accumulators into the gradient arena with one launch (FusedAdam.finalize_grads) before the all-reduce /
    update, so a weight gradient needs no zero-fill, no temporary, no per-layer unpack and no autograd accumulation --
    and a weight used twice in one step (residual_unet3d.py:126-128) simply accumulates twice."""
    if param is None or not getattr(param, "_b200_direct_grad", False) or param.grad is None:
        return None
    param._b200_pending[0] = True
    param._b200_dw_used = True       # optim.PeerGradExchange: this weight's gradient lives in the packed arena
    return param._b200_dwp


def conv3d_wgrad_raw(g, x, dy, weight_shape, weight=None):
    """Returns the weight gradient, or None when it was accumulated straight into weight.grad (FusedAdam arena)."""
    from ._lib import load
    if _use_padded(g) and not (g.cin == 1 and g.k == 3 and g.stride == 1 and load().b200seg_conv3d_workspace_bytes(ctypes.byref(g))):
        # (the C_in = 1 3x3x3 stem keeps its taps-as-channels weight gradient, conv.cu: stem_plan)
        gp = _padded_geom(g)
        x_p, dy_p = _pad_channels(x, gp.cin), _pad_channels(dy, gp.cout)
        dwp_p = torch.zeros(g.k ** 3 * gp.cin * gp.cout, dtype=torch.float32, device=x_p.device)
        ws_bytes = load().b200seg_conv3d_workspace_bytes(ctypes.byref(gp))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x_p.device) if ws_bytes else None
        _call("b200seg_conv3d_wgrad", ctypes.byref(gp), _ptr(x_p), gp.cin, _ptr(dy_p), gp.cout, _ptr(dwp_p), _ptr(ws), ws_bytes,
              _stream(), work=_conv_flops(g), tag="conv_wgrad_padded")
        gw_p = torch.empty((gp.cout, gp.cin) + tuple(weight_shape[2:]), dtype=torch.float32, device=x_p.device)
        _call("b200seg_unpack_conv_wgrad", _ptr(dwp_p), _ptr(gw_p), gp.cout, gp.cin, g.k, 0, gp.cin, 0, _stream())
        return gw_p[:g.cout, :g.cin].contiguous()
    x, xp = _as_rows(x)
    dy, dyp = _as_rows(dy)
    k3 = g.k ** 3
    dwp = _grad_target(weight)
    direct = dwp is not None
    if not direct:
        dwp = torch.zeros(k3 * g.cin * g.cout, dtype=torch.float32, device=x.device)
    ws_bytes = load().b200seg_conv3d_workspace_bytes(ctypes.byref(g))      # split-K partial tiles (0: atomics path)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
    _call("b200seg_conv3d_wgrad", ctypes.byref(g), _ptr(x), xp, _ptr(dy), dyp, _ptr(dwp), _ptr(ws), ws_bytes, _stream(),
          work=_conv_flops(g), tag="conv_wgrad")
    if direct:
        return None
    gw = torch.empty(weight_shape, dtype=torch.float32, device=x.device)
    _call("b200seg_unpack_conv_wgrad", _ptr(dwp), _ptr(gw), g.cout, g.cin, g.k, 0, g.cin, 0, _stream())
    return gw


def channel_stats(x, groups=1, spare=0):
    """[groups][2][C] fp32 sums / sums of squares over the voxels of each group (group = sample for instance norm).
    spare=1 (groups == 1 only) returns the flat {sum[C], sumsq[C], spare} form the cross-GPU exchange uses."""
    x, xp = _as_rows(x)
    n, d, h, w, c = x.shape
    rows = n * d * h * w // groups
    if spare:
        assert groups == 1
        stats = _zeros_f32(2 * c + spare, x.device)
    else:
        stats = _zeros_f32(groups * 2 * c, x.device).view(groups, 2, c)
    _call("b200seg_channel_stats", _ptr(x), xp, rows, groups, c, _ptr(stats), _stream())
    return stats


def _norm_coef(stats, count, groups, c, gamma, beta, running_mean, running_var, momentum, eps, clamp_eps, device):
    coef = torch.empty((groups, 4, c), dtype=torch.float32, device=device)
    _call("b200seg_norm_finalize", _ptr(stats), float(count), groups, c, _ptr(gamma), _ptr(beta),
          _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), int(clamp_eps), _ptr(coef), _stream())
    return coef


def _eval_coef(gamma, beta, running_mean, running_var, eps, c, device, conv_bias=None):
    """Inference-mode normalisation constants [1][4][C] = {mean, inv_std, scale, shift} from the running statistics;
    a convolution bias in front of the norm is folded into the shift (for the fused conv epilogue)."""
    def make():
        coef = torch.empty((1, 4, c), dtype=torch.float32, device=device)
        f32 = lambda t: None if t is None else t.detach().float().contiguous()   # noqa: E731
        g_, b_, rm, rv, cb = f32(gamma), f32(beta), f32(running_mean), f32(running_var), f32(conv_bias)
        _call("b200seg_norm_eval_coef", _ptr(g_), _ptr(b_), _ptr(rm), _ptr(rv), _ptr(cb), float(eps), c, _ptr(coef), _stream())
        return coef
    keep = (gamma, beta, running_mean, running_var, conv_bias)
    return _frozen_get(("coef",) + tuple(id(t) for t in keep) + (float(eps), c), keep, make)


def _copy_into(out, y):
    """out[...] = y by kernel (the identity form of the normalise pass).  `out` is usually a channel slice of a buffer of
    which autograd-visible views exist (concat-free decoders, dense blocks); a torch in-place op on it would be refused."""
    y, yp = _as_rows(y)
    n, d, h, w, c = y.shape
    assert _pitched(out) and tuple(out.shape) == tuple(y.shape)
    if out.data_ptr() % 16 and c % 8 == 0 and out.stride(3) % 8 == 0:      # the 128-bit path would be misaligned
        out.copy_(y)
        return out
    _call("b200seg_norm_act_fwd", _ptr(y), yp, None, n * d * h * w, 1, c, 0, 0.0, None, None, 0, _ptr(out), out.stride(3),
          _stream())
    return out


def conv_fused_eval_supported(g):
    from ._lib import load
    return bool(load().b200seg_conv3d_fprop_act_supported(ctypes.byref(g)))


def conv3d_fprop_eval_fused(x, weight, bias, k, pad, dil, spec, gamma, beta, running_mean, running_var, out=None):
    """Inference: conv -> eval-mode BatchNorm -> {none, ReLU, LeakyReLU} in ONE kernel (scale / shift / activation in the conv
    epilogue).  Returns None when the geometry is not on the fused tensor-core path (the caller then runs the two passes)."""
    x, xp = _as_rows(x)
    cout, cin = weight.shape[0], weight.shape[1]
    g = _geom(x.shape, cin, cout, k, 1, pad, dil)
    wp = None
    if cin % 16 and cout % 16 == 0 and _use_padded(g) and conv_fused_eval_supported(_padded_geom(g)):
        # C_in = 1 stems (unet3d.py:80, vnet3d.py:47): K zero-padded to 16 channels, the same fused epilogue
        g = _padded_geom(g)
        x = _pad_channels(x, g.cin)
        xp = g.cin

        def make():
            packed = torch.empty(k ** 3 * cout * g.cin, dtype=torch.bfloat16, device=x.device)
            _call("b200seg_pack_conv_weight", _ptr(weight.detach()), _ptr(packed), cout, cin, k, 0, g.cin, 0, _stream())
            return packed
        wp = _frozen_get(("packk", id(weight), g.cin), weight, make)
    elif cin % 16 or cout % 16 or not conv_fused_eval_supported(g):
        return None
    coef = _eval_coef(gamma, beta, running_mean, running_var, spec.eps, cout, x.device, conv_bias=bias)
    if out is not None and _pitched(out) and tuple(out.shape) == (g.n, g.od, g.oh, g.ow, cout) and out.data_ptr() % 16 == 0 \
            and out.stride(3) % 8 == 0:
        z = out
    else:
        z = torch.empty((g.n, g.od, g.oh, g.ow, cout), dtype=torch.bfloat16, device=x.device)
    _call("b200seg_conv3d_fprop_act", ctypes.byref(g), _ptr(x), xp, _ptr(wp if wp is not None else pack_conv_weight(weight)), _ptr(coef[0, 2]),
          _ptr(coef[0, 3]), spec.act, spec.act_param, _ptr(z), z.stride(3), _stream(), work=_conv_flops(g),
          tag="conv_fprop_tc" if wp is None else "conv_fprop_padded_tc")
    if out is not None and z is not out:
        z = _copy_into(out, z)
    return z


class NormSpec:
    """What follows a convolution: normalisation kind, activation, and where the statistics come from."""

    def __init__(self, kind=None, act="none", act_param=0.0, eps=1e-5, momentum=0.1, training=True, sync=False,
                 clamp_eps=False, process_group=None):
        assert kind in (None, "batch", "instance")
        self.kind, self.act, self.act_param = kind, ACT[act], float(act_param)
        self.eps, self.momentum, self.training = eps, momentum, training
        self.sync, self.clamp_eps, self.process_group = sync, clamp_eps, process_group


def _norm_forward(y, stats, spec, gamma, beta, running_mean, running_var, prelu_w, residual, out):
    """y (raw conv output or block input) -> z = act(norm(y) [+ residual]).  Returns z, coef, count, groups."""
    y, yp = _as_rows(y)
    n, d, h, w, c = y.shape
    groups, count, coef = 1, float(n * d * h * w), None
    fused_stats = None
    if spec.kind == "instance":
        groups, count = n, float(d * h * w)
        coef = _norm_coef(channel_stats(y, groups), count, groups, c, None, None, None, None, 0.0, spec.eps, False,
                          y.device)
    elif spec.kind == "batch":
        if spec.training:
            if stats is None:
                stats = channel_stats(y, 1, spare=1)
            px = peer_exchange(spec.process_group) if (spec.sync and is_parallel(spec.process_group)) else None
            if px is not None and 2 * c + 1 <= 2112:
                # one kernel: NVLink exchange of {sum, sumsq, count} + _compute_mean_std (batchnorm.py:102-125)
                coef = px.reduce_and_finalize(stats, count, c, gamma, beta, running_mean, running_var, spec.momentum,
                                              spec.eps, spec.clamp_eps)
                count *= px.world      # equal per-rank batches (what DDP + DistributedSampler guarantee)
            else:
                if spec.sync:
                    count *= all_reduce_stats(stats[:2 * c], spec.process_group)
                if c <= 512 and not os.environ.get("B200SEG_SEPARATE_FINALIZE"):
                    fused_stats = stats       # finalised in the prologue of the normalise kernel below (no extra launch)
                else:
                    coef = _norm_coef(stats, count, 1, c, gamma, beta, running_mean, running_var, spec.momentum,
                                      spec.eps, spec.clamp_eps, y.device)
        else:
            coef = _eval_coef(gamma, beta, running_mean, running_var, spec.eps, c, y.device)
    if out is None:
        if c > 16 and c % 16 and n * d * h * w >= 65536:
            # channel counts like DenseVoxelNet's 16 + 12 i (densevoxelnet3d.py:36-42): the consumer is a convolution on
            # the padded tensor-core path, which takes the zero-tailed wide buffer as it is (see _pad_channels)
            wide = torch.zeros((n, d, h, w, _pad16(c)), dtype=torch.bfloat16, device=y.device)
            out = wide[..., :c]
            _publish_zero_tail(out, wide)
        else:
            out = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=y.device)
    assert _pitched(out) and out.shape == y.shape
    res, resp = (None, 0)
    if residual is not None:
        res, resp = _as_rows(residual)
    rows = n * d * h * w // groups
    if fused_stats is not None:
        coef = torch.empty((1, 4, c), dtype=torch.float32, device=y.device)
        _call("b200seg_norm_act_fwd_stats", _ptr(y), yp, _ptr(fused_stats), float(count), _ptr(gamma), _ptr(beta),
              _ptr(running_mean), _ptr(running_var), float(spec.momentum), float(spec.eps), int(spec.clamp_eps), _ptr(coef),
              rows, c, spec.act, spec.act_param, _ptr(prelu_w), _ptr(res), resp, _ptr(out), out.stride(3), _stream(),
              tag="b200seg_norm_act_fwd")
        return out, coef, count, groups
    _call("b200seg_norm_act_fwd", _ptr(y), yp, _ptr(coef), rows, groups, c, spec.act, spec.act_param, _ptr(prelu_w),
          _ptr(res), resp, _ptr(out), out.stride(3), _stream())
    return out, coef, count, groups


def _norm_backward(dz, y, coef, count, groups, spec, prelu_w, residual, want_dres, grad_affine=None, acc_into=None):
    """Returns dy (bf16, contiguous), dres, sums ([groups][2 or 3][C] fp32: sum dpre, sum dpre*xhat[, dPReLU]).
    acc_into: a gradient buffer (view) that already holds what y received through another consumer; the result is added
    to it in place and it is returned as dy."""
    dz, dzp = _as_rows(dz)
    y, yp = _as_rows(y)
    n, d, h, w, c = y.shape
    rows = n * d * h * w // groups
    res, resp = (None, 0)
    if residual is not None:
        res, resp = _as_rows(residual)
    nrow = 3 if spec.act == ACT["prelu"] else 2
    use_batch_stats = spec.kind == "instance" or (spec.kind == "batch" and spec.training)
    sums = None
    if use_batch_stats or nrow == 3 or (spec.kind == "batch" and coef is not None):
        sums = _zeros_f32(groups * nrow * c, y.device).view(groups, nrow, c)
        dprelu = sums[0, 2] if nrow == 3 else None
        _call("b200seg_norm_act_bwd_reduce", _ptr(dz), dzp, _ptr(y), yp, _ptr(coef), rows, groups, c, spec.act,
              spec.act_param, _ptr(prelu_w), _ptr(res), resp, _ptr(sums), _ptr(dprelu), _ptr(grad_affine), _stream())
    red = sums
    if use_batch_stats and spec.kind == "batch" and spec.sync and is_parallel(spec.process_group):
        red = sums.clone()
        all_reduce_stats(red, spec.process_group, between=flush_deferred if _DEFERRED else None)
    dres = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=y.device) if want_dres else None
    if acc_into is not None:
        assert _pitched(acc_into) and tuple(acc_into.shape) == (n, d, h, w, c)
        dy = acc_into
        _call("b200seg_norm_act_bwd_apply_acc", _ptr(dz), dzp, _ptr(y), yp, _ptr(coef),
              _ptr(red if use_batch_stats else None), float(count), rows, groups, c, spec.act, spec.act_param, _ptr(prelu_w),
              _ptr(res), resp, _ptr(dy), dy.stride(3), _ptr(dres), c, _ptr(dy), dy.stride(3), _stream(),
              tag="b200seg_norm_act_bwd_apply")
        return dy, dres, sums
    dy = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=y.device)
    _call("b200seg_norm_act_bwd_apply", _ptr(dz), dzp, _ptr(y), yp, _ptr(coef), _ptr(red if use_batch_stats else None),
          float(count), rows, groups, c, spec.act, spec.act_param, _ptr(prelu_w), _ptr(res), resp, _ptr(dy), c,
          _ptr(dres), c, _stream())
    return dy, dres, sums


# ------------------------------------------------------------------------------------------------ autograd functions
class _ToNDHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        n, c = x.shape[0], x.shape[1]
        spatial = x[0, 0].numel()
        x = x.contiguous().float()
        out = torch.empty((n,) + tuple(x.shape[2:]) + (c,), dtype=torch.bfloat16, device=x.device)
        _call("b200seg_ncdhw_f32_to_ndhwc_bf16", _ptr(x), _ptr(out), n, c, spatial, _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n, c = g.shape[0], g.shape[4]
        spatial = g[0, ..., 0].numel()
        out = torch.empty((n, c) + tuple(g.shape[1:4]), dtype=torch.float32, device=g.device)
        _call("b200seg_ndhwc_bf16_to_ncdhw_f32", _ptr(g), _ptr(out), n, c, spatial, _stream())
        return out


def to_ndhwc(x):
    """fp32 NCDHW (what train.py:195 feeds the model) -> bf16 NDHWC."""
    return _ToNDHWC.apply(x)


class _ConcatInput(torch.autograd.Function):
    """torch.cat((x, maps), dim=1) of two fp32 NCDHW tensors, delivered as one bf16 NDHWC activation (Double_Unet.py:90)."""

    @staticmethod
    def forward(ctx, x, maps):
        n, c1, c2 = x.shape[0], x.shape[1], maps.shape[1]
        spatial = x[0, 0].numel()
        assert maps.shape[0] == n and maps.shape[2:] == x.shape[2:]
        x, maps = x.contiguous().float(), maps.contiguous().float()
        out = torch.empty((n,) + tuple(x.shape[2:]) + (c1 + c2,), dtype=torch.bfloat16, device=x.device)
        _call("b200seg_ncdhw_f32_to_ndhwc_bf16_pitched", _ptr(x), _ptr(out), c1 + c2, n, c1, spatial, _stream())
        _call("b200seg_ncdhw_f32_to_ndhwc_bf16_pitched", _ptr(maps), out.data_ptr() + 2 * c1, c1 + c2, n, c2, spatial,
              _stream())
        ctx.split = (c1, c2)
        return out

    @staticmethod
    def backward(ctx, g):
        c1, c2 = ctx.split
        g = g.contiguous()
        n = g.shape[0]
        spatial = g[0, ..., 0].numel()
        outs = []
        for need, off, c in ((ctx.needs_input_grad[0], 0, c1), (ctx.needs_input_grad[1], c1, c2)):
            if not need:
                outs.append(None)
                continue
            o = torch.empty((n, c) + tuple(g.shape[1:4]), dtype=torch.float32, device=g.device)
            _call("b200seg_ndhwc_bf16_to_ncdhw_f32_pitched", g.data_ptr() + 2 * off, c1 + c2, _ptr(o), n, c, spatial, _stream())
            outs.append(o)
        return tuple(outs)


def concat_input(x, maps):
    return _ConcatInput.apply(x, maps)


class _FromNDHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        n, c = x.shape[0], x.shape[4]
        spatial = x[0, ..., 0].numel()
        out = torch.empty((n, c) + tuple(x.shape[1:4]), dtype=torch.float32, device=x.device)
        _call("b200seg_ndhwc_bf16_to_ncdhw_f32", _ptr(x), _ptr(out), n, c, spatial, _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return _ToNDHWC.forward(None, g)


def from_ndhwc(x):
    return _FromNDHWC.apply(x)


class _ConvNormAct(torch.autograd.Function):
    """conv (+bias) -> [batch / instance norm] -> activation [+ residual before the activation]."""

    @staticmethod
    def forward(ctx, x, x2, weight, bias, gamma, beta, prelu_w, residual, running_mean, running_var, cfg):
        k, stride, pad, dil, spec, out = cfg
        xin = x if x2 is None else merge_channels(x, x2)
        if (spec.kind == "batch" and not spec.training and residual is None and stride == 1 and not torch.is_grad_enabled()
                and spec.act in (ACT["none"], ACT["relu"], ACT["leaky_relu"]) and running_mean is not None):
            # inference: BatchNorm scale / shift + activation in the convolution's epilogue, no separate pass
            z = conv3d_fprop_eval_fused(xin, weight, bias, k, pad, dil, spec, gamma, beta, running_mean, running_var, out)
            if z is not None:
                return z
        fused_stats = spec.kind == "batch" and spec.training
        plain = spec.kind is None and spec.act == 0 and residual is None
        _LAST_PADDED_INPUT[0] = None
        y, stats, g = conv3d_fprop_raw(xin, weight, bias, k, stride, pad, dil, fused_stats, y_out=out if plain else None)
        if _LAST_PADDED_INPUT[0] is not None and not (g.cin == 1 and g.k == 3):   # (that stem has its own weight gradient)
            # the padded tensor-core path widened the input (zero channels): save that copy, so that the weight gradient
            # does not widen it a second time (the extra channels only produce weight-gradient columns that are dropped)
            xin, _LAST_PADDED_INPUT[0] = _LAST_PADDED_INPUT[0], None
        if plain:
            z, coef, count, groups = y, None, 0.0, 1
            if out is not None and y is not out:
                z = _copy_into(out, y)
        else:
            z, coef, count, groups = _norm_forward(y, stats, spec, gamma, beta, running_mean, running_var, prelu_w,
                                                   residual, out)
        # a plain conv's output is not needed by its own backward (and may be modified in place by a residual sum)
        ctx.save_for_backward(xin, None if plain else y, coef, weight, gamma, prelu_w, residual)
        ctx.cfg = (g, spec, count, groups, None if x2 is None else x.shape[4], bias is not None)
        ctx.beta_ref, ctx.bias_ref = beta, bias      # parameters (not saved tensors): only their .grad slots are looked at
        return z

    @staticmethod
    def backward(ctx, dz):
        xin, y, coef, weight, gamma, prelu_w, residual = ctx.saved_tensors
        g, spec, count, groups, split, has_bias = ctx.cfg
        need = ctx.needs_input_grad
        dgamma = dbeta = dprelu = dres = None
        plain = spec.kind is None and spec.act == 0 and residual is None
        if plain:
            dy = dz
        else:
            affine = _affine_grad_target(gamma, ctx.beta_ref) if (need[4] and need[5]) else None
            dy, dres, sums = _norm_backward(dz, y, coef, count, groups, spec, prelu_w, residual,
                                            residual is not None and need[7], grad_affine=affine)
            if gamma is not None and affine is None:
                dgamma, dbeta = (sums[0, 1], sums[0, 0]) if groups == 1 else (sums[:, 1].sum(0), sums[:, 0].sum(0))
            if prelu_w is not None:
                dprelu = sums[0, 2]
        dx = dx2 = dw = db = None
        if _use_padded(g) and dy.shape[4] != _pad16(g.cout):
            dy_full, dy = dy, _pad_channels(dy, _pad16(g.cout))     # once, for both the data and the weight gradient
        else:
            dy_full = dy
        if need[0] or (split is not None and need[1]):
            if split is None:
                dx = conv3d_dgrad_raw(g, dy, weight)
            else:
                dxin, colsum = conv3d_dgrad_raw(g, dy, weight, colsum=True)
                dx, dx2 = dxin[..., :split], dxin[..., split:]
                _publish_colsum(dx, colsum[:split])
        if need[2]:
            if (_SCRATCH["active"] and spec.kind == "batch" and spec.sync and spec.training and is_parallel(spec.process_group)
                    and _arena_managed(weight) and getattr(weight, "_b200_direct_grad", False) and not _use_padded(g)
                    and not getattr(weight, "_b200_hooked", False)):
                # postponed to the middle of the next layer's statistics exchange (see _DEFERRED)
                _DEFERRED.append(lambda g=g, xin=xin, dy=dy, weight=weight: conv3d_wgrad_raw(g, xin, dy, weight.shape, weight))
            else:
                dw = conv3d_wgrad_raw(g, xin, dy, weight.shape, weight)
        if has_bias and need[3]:
            if spec.kind is not None and (spec.training or spec.kind == "instance"):
                # a bias in front of batch/instance statistics has an analytically zero gradient: leave the (zeroed)
                # slot of the gradient arena alone instead of accumulating zeros into it
                db = None if _arena_managed(ctx.bias_ref) else _zeros_f32(g.cout, dz.device)
            else:
                db = channel_stats(dy_full, 1)[0, 0]
        return dx, dx2, dw, db, dgamma, dbeta, dprelu, dres, None, None, None


class _SpaceToDepth(torch.autograd.Function):
    """[n, d, h, w, c] -> [n, od, oh, ow, k^3 * c]: the taps of every s^3 cell side by side (k <= s, no padding)."""

    @staticmethod
    def forward(ctx, x, cfg):
        k, s = cfg
        x, xp = _as_rows(x)
        n, d, h, w, c = x.shape
        od, oh, ow = (d - k) // s + 1, (h - k) // s + 1, (w - k) // s + 1
        y = torch.empty((n, od, oh, ow, k ** 3 * c), dtype=torch.bfloat16, device=x.device)
        _call("b200seg_space_to_depth", _ptr(x), xp, _ptr(y), n, d, h, w, c, k, s, od, oh, ow, _stream())
        ctx.cfg = (k, s, (n, d, h, w, c))
        return y

    @staticmethod
    def backward(ctx, dy):
        k, s, (n, d, h, w, c) = ctx.cfg
        dy = dy.contiguous()
        od, oh, ow = dy.shape[1:4]
        dx = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=dy.device)
        _call("b200seg_depth_to_space", _ptr(dy), _ptr(dx), c, n, d, h, w, c, k, s, od, oh, ow, _stream())
        return dx, None


class _DepthToSpace(torch.autograd.Function):
    """[n, d, h, w, s^3 * c] -> [n, s d, s h, s w, c] (pixel shuffle, the k == s inverse of _SpaceToDepth)."""

    @staticmethod
    def forward(ctx, y, cfg):
        s, out = cfg
        y = y.contiguous()
        n, d, h, w, cc = y.shape
        c = cc // s ** 3
        if out is None:
            out = torch.empty((n, s * d, s * h, s * w, c), dtype=torch.bfloat16, device=y.device)
        assert _pitched(out) and tuple(out.shape) == (n, s * d, s * h, s * w, c)
        _call("b200seg_depth_to_space", _ptr(y), _ptr(out), out.stride(3), n, s * d, s * h, s * w, c, s, s, d, h, w, _stream())
        ctx.cfg = (s, c)
        return out

    @staticmethod
    def backward(ctx, dout):
        s, c = ctx.cfg
        dout, dp = _as_rows(dout)
        n, D, H, W, _ = dout.shape
        dy = torch.empty((n, D // s, H // s, W // s, s ** 3 * c), dtype=torch.bfloat16, device=dout.device)
        _call("b200seg_space_to_depth", _ptr(dout), dp, _ptr(dy), n, D, H, W, c, s, s, D // s, H // s, W // s, _stream())
        return dy, None


def conv_norm_act(x, weight, bias=None, *, x2=None, k=3, stride=1, pad=1, dil=1, spec=None, gamma=None, beta=None,
                  prelu_weight=None, residual=None, running_mean=None, running_var=None, out=None):
    spec = spec or NormSpec()
    if stride > 2 and k <= stride and pad == 0 and dil == 1 and x2 is None:
        # non-overlapping windows (csrnet.py:115-133: Conv3d(k3, s4)): gather the taps into the channel dimension and run a
        # 1x1x1 convolution (a plain GEMM on the tensor cores) with the weight in the same (tap, channel) order
        xs = _SpaceToDepth.apply(x, (int(k), int(stride)))
        w2 = weight.permute(0, 2, 3, 4, 1).reshape(weight.shape[0], -1, 1, 1, 1)
        return _ConvNormAct.apply(xs, None, w2, bias, gamma, beta, prelu_weight, residual, running_mean, running_var,
                                  (1, 1, 0, 1, spec, out))
    return _ConvNormAct.apply(x, x2, weight, bias, gamma, beta, prelu_weight, residual, running_mean, running_var,
                              (k, stride, pad, dil, spec, out))


class _NormAct(torch.autograd.Function):
    """Stand-alone normalisation + activation (pre-activation blocks: convolution.py:48-52, densevoxelnet3d.py:21-24)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, prelu_w, residual, running_mean, running_var, cfg):
        spec, out = cfg[0], cfg[1]
        z, coef, count, groups = _norm_forward(y, None, spec, gamma, beta, running_mean, running_var, prelu_w, residual,
                                               out)
        ctx.save_for_backward(y, coef, gamma, prelu_w, residual)
        ctx.cfg = (spec, count, groups)
        ctx.shared_grad = len(cfg) > 2 and bool(cfg[2])
        return z

    @staticmethod
    def backward(ctx, dz):
        y, coef, gamma, prelu_w, residual = ctx.saved_tensors
        spec, count, groups = ctx.cfg
        acc = None
        if ctx.shared_grad:      # y is also the head of a dense-block concatenation: its gradient buffer already exists
            hit = _GRAD_STASH.pop((y.data_ptr(), tuple(y.shape), tuple(y.stride())), None)
            acc = hit[1] if hit is not None else None
        dy, dres, sums = _norm_backward(dz, y, coef, count, groups, spec, prelu_w, residual,
                                        residual is not None and ctx.needs_input_grad[4], acc_into=acc)
        dgamma = dbeta = dprelu = None
        if gamma is not None:
            dgamma, dbeta = (sums[0, 1], sums[0, 0]) if groups == 1 else (sums[:, 1].sum(0), sums[:, 0].sum(0))
        if prelu_w is not None:
            dprelu = sums[0, 2]
        return dy, dgamma, dbeta, dprelu, dres, None, None, None


def norm_act(y, spec, gamma=None, beta=None, prelu_weight=None, residual=None, running_mean=None, running_var=None,
             out=None, shared_grad=False):
    """shared_grad: y was produced by concat_channels_shared (dense blocks): the backward adds this op's input gradient
    into the buffer that already holds the gradient y received as the head of the next concatenation."""
    return _NormAct.apply(y, gamma, beta, prelu_weight, residual, running_mean, running_var, (spec, out, shared_grad))


class _MaxPool2(torch.autograd.Function):
    """MaxPool3d(2, 2); with `skip=True` it also returns the input as a second output, so that the gradient arriving
    through a skip connection is added inside the pooling backward kernel instead of by a separate autograd add."""

    @staticmethod
    def forward(ctx, x, skip):
        x, xp = _as_rows(x)
        n, d, h, w, c = x.shape
        y = torch.empty((n, d // 2, h // 2, w // 2, c), dtype=torch.bfloat16, device=x.device)
        idx = torch.empty((n, d // 2, h // 2, w // 2, c), dtype=torch.uint8, device=x.device)
        _call("b200seg_maxpool2_fwd", _ptr(x), xp, _ptr(y), c, _ptr(idx), n, d, h, w, c, _stream())
        ctx.save_for_backward(idx)
        ctx.shape = (n, d, h, w, c)
        ctx.mark_non_differentiable(idx)
        if skip:
            return y, idx, x.as_strided(x.shape, x.stride())
        return y, idx

    @staticmethod
    def backward(ctx, dy, _, dskip=None):
        (idx,) = ctx.saved_tensors
        n, d, h, w, c = ctx.shape
        assert d % 2 == 0 and h % 2 == 0 and w % 2 == 0, "MaxPool3d(2,2) backward needs even extents"
        if dy is None:
            return dskip, None
        dy, dyp = _as_rows(dy)
        add, addp = (None, 0)
        if dskip is not None:
            add, addp = _as_rows(dskip)
        dx = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=dy.device)
        _call("b200seg_maxpool2_bwd", _ptr(dy), dyp, _ptr(idx), _ptr(dx), c, _ptr(add), addp, n, d, h, w, c, _stream())
        return dx, None


def max_pool2(x, return_indices=False):
    """nn.MaxPool3d(2, 2).  Indices are uint8 local arg-max codes (see maxpool_indices_to_torch)."""
    y, idx = _MaxPool2.apply(x, False)
    return (y, idx) if return_indices else y


def max_pool2_skip(x):
    """(MaxPool3d(2,2)(x), x): use the second value for the skip connection (unet3d.py:52-68); both gradients of x
    are then combined by one kernel."""
    y, _, xs = _MaxPool2.apply(x, True)
    return y, xs


def maxpool_indices_to_torch(idx, in_shape):
    """uint8 local codes -> torch's int64 flat D*H*W indices, NCDHW layout (for the bit-exact comparison)."""
    n, d, h, w, c = in_shape
    out = torch.empty((n, c, d // 2, h // 2, w // 2), dtype=torch.int64, device=idx.device)
    _call("b200seg_maxpool2_idx_to_torch", _ptr(idx), _ptr(out), n, d, h, w, c, _stream())
    return out


class _ConvT2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, out):
        x, xp = _as_rows(x)
        n, d, h, w, cin = x.shape
        cout = weight.shape[1]
        assert weight.shape[0] == cin and tuple(weight.shape[2:]) == (2, 2, 2)
        if out is None:
            out = torch.empty((n, 2 * d, 2 * h, 2 * w, cout), dtype=torch.bfloat16, device=x.device)
        assert _pitched(out) and tuple(out.shape) == (n, 2 * d, 2 * h, 2 * w, cout)
        wp = _cached_pack(weight, 1)     # ConvT forward = dgrad of the equivalent strided conv
        if wp is None:
            wp = torch.empty(8 * cin * cout, dtype=torch.bfloat16, device=x.device)
            _call("b200seg_pack_convt_weight", _ptr(weight.detach()), _ptr(wp), cin, cout, 0, _stream())
        b = bias.detach().float() if bias is not None else None
        _call("b200seg_convt_k2s2_fwd", _ptr(x), xp, _ptr(wp), _ptr(b), _ptr(out), out.stride(3), n, d, h, w, cin, cout,
              _stream())
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        n, d, h, w, cin = x.shape
        cout = weight.shape[1]
        dy_in = dy
        dy, dyp = _as_rows(dy)
        x, xp = _as_rows(x)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wd = _cached_pack(weight, 0)
            if wd is None:
                wd = torch.empty(8 * cin * cout, dtype=torch.bfloat16, device=x.device)
                _call("b200seg_pack_convt_weight", _ptr(weight.detach()), _ptr(wd), cin, cout, 1, _stream())
            dx = torch.empty((n, d, h, w, cin), dtype=torch.bfloat16, device=x.device)
            _call("b200seg_convt_k2s2_dgrad", _ptr(dy), dyp, _ptr(wd), _ptr(dx), cin, n, d, h, w, cin, cout, _stream())
        if ctx.needs_input_grad[1]:
            dwp = _grad_target(weight)
            direct = dwp is not None
            if not direct:
                dwp = torch.zeros(8 * cin * cout, dtype=torch.float32, device=x.device)
            _call("b200seg_convt_k2s2_wgrad", _ptr(x), xp, _ptr(dy), dyp, _ptr(dwp), n, d, h, w, cin, cout, _stream())
            # packed [8][cout_T][cin_T] is the strided conv's [k^3][cin_S][cout_S]: unpack with cout_S=cin_T, cin_S=cout_T
            if not direct:
                dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
                _call("b200seg_unpack_conv_wgrad", _ptr(dwp), _ptr(dw), cin, cout, 2, 0, cout, 0, _stream())
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _take_colsum(dy_in)
            if db is None:
                db = channel_stats(dy, 1)[0, 0]
        return dx, dw, db, None


def conv_transpose_k2s2(x, weight, bias=None, out=None):
    """nn.ConvTranspose3d(kernel_size=2, stride=2) (unet3d.py:29-43)."""
    return _ConvT2.apply(x, weight, bias, out)


class _ConvTS(torch.autograd.Function):
    """nn.ConvTranspose3d(kernel_size = stride = s, padding 0): non-overlapping up-convolution for any s (csrnet.py:137-149
    uses s = 4).  It is the exact transpose of the strided convolution S: [n, s*d, s*h, s*w, C_out] -> [n, d, h, w, C_in]
    with the same weight memory, so the three passes are S's data-gradient / forward / weight-gradient kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, cfg):
        s, out = cfg
        x, xp = _as_rows(x)
        n, d, h, w, cin = x.shape
        cout = weight.shape[1]
        assert weight.shape[0] == cin and tuple(weight.shape[2:]) == (s, s, s)
        gs = ConvGeom(n, s * d, s * h, s * w, cout, d, h, w, cin, s, s, 0, 1)
        if out is None:
            out = torch.empty((n, s * d, s * h, s * w, cout), dtype=torch.bfloat16, device=x.device)
        assert _pitched(out) and tuple(out.shape) == (n, s * d, s * h, s * w, cout)
        wd = torch.empty(s ** 3 * cin * cout, dtype=torch.bfloat16, device=x.device)
        _call("b200seg_pack_conv_weight", _ptr(weight.detach()), _ptr(wd), cin, cout, s, 0, cout, 1, _stream())
        _call("b200seg_conv3d_dgrad", ctypes.byref(gs), _ptr(x), xp, _ptr(wd), _ptr(out), out.stride(3), None, None, 0, _stream())
        if bias is not None:      # + bias: one scale-shift pass with scale 1
            coef = torch.zeros((1, 4, cout), dtype=torch.float32, device=x.device)
            coef[0, 2] = 1.0
            coef[0, 3] = bias.detach().float()
            _call("b200seg_norm_act_fwd", _ptr(out), out.stride(3), _ptr(coef), n * d * h * w * s ** 3, 1, cout, 0, 0.0, None,
                  None, 0, _ptr(out), out.stride(3), _stream())
        ctx.save_for_backward(x, weight)
        ctx.cfg = (s, gs, bias is not None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        s, gs, has_bias = ctx.cfg
        n, d, h, w, cin = x.shape
        cout = weight.shape[1]
        dy, dyp = _as_rows(dy)
        x, xp = _as_rows(x)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wf = torch.empty(s ** 3 * cin * cout, dtype=torch.bfloat16, device=x.device)
            _call("b200seg_pack_conv_weight", _ptr(weight.detach()), _ptr(wf), cin, cout, s, 0, cout, 0, _stream())
            dx = torch.empty((n, d, h, w, cin), dtype=torch.bfloat16, device=x.device)
            _call("b200seg_conv3d_fprop", ctypes.byref(gs), _ptr(dy), dyp, _ptr(wf), None, _ptr(dx), cin, None, None, 0, _stream())
        if ctx.needs_input_grad[1]:
            dwp = torch.zeros(s ** 3 * cin * cout, dtype=torch.float32, device=x.device)   # [k^3][cin_S = cout][cout_S = cin]
            _call("b200seg_conv3d_wgrad", ctypes.byref(gs), _ptr(dy), dyp, _ptr(x), xp, _ptr(dwp), None, 0, _stream())
            dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
            _call("b200seg_unpack_conv_wgrad", _ptr(dwp), _ptr(dw), cin, cout, s, 0, cout, 0, _stream())
        if has_bias and ctx.needs_input_grad[2]:
            db = channel_stats(dy, 1)[0, 0]
        return dx, dw, db, None


def conv_transpose_kxsx(x, weight, bias=None, stride=2, out=None):
    """nn.ConvTranspose3d(kernel_size = stride).  stride 2 takes the dedicated tensor-core path (conv_transpose_k2s2)."""
    if stride == 2:
        return _ConvT2.apply(x, weight, bias, out)
    s, cin, cout = int(stride), weight.shape[0], weight.shape[1]
    if tuple(weight.shape[2:]) == (s, s, s) and cin % 16 == 0 and (s ** 3 * cout) % 16 == 0 and not os.environ.get("B200SEG_CONVT_DIRECT"):
        # a 1x1x1 convolution to s^3 * C_out channels (tap-major) on the tensor cores, then the pixel shuffle
        w2 = weight.permute(2, 3, 4, 1, 0).reshape(s ** 3 * cout, cin, 1, 1, 1)
        b2 = bias.repeat(s ** 3) if bias is not None else None
        y = _ConvNormAct.apply(x, None, w2, b2, None, None, None, None, None, None, (1, 1, 0, 1, NormSpec(), None))
        return _DepthToSpace.apply(y, (s, out))
    return _ConvTS.apply(x, weight, bias, (s, out))


class _Head(torch.autograd.Function):
    """1x1x1 convolution to class logits, returned as fp32 NCDHW like the reference model's output (unet3d.py:70)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, xp = _as_rows(x)
        n, d, h, w, cin = x.shape
        classes = weight.shape[0]
        w2 = weight.detach().reshape(classes, cin).float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        logits = torch.empty((n, classes, d, h, w), dtype=torch.float32, device=x.device)
        _call("b200seg_head_conv1x1_fwd", _ptr(x), xp, _ptr(w2), _ptr(b), _ptr(logits), n, d * h * w, cin, classes,
              _stream())
        ctx.save_for_backward(x, w2)
        ctx.wshape = weight.shape
        ctx.has_bias = bias is not None
        return logits

    @staticmethod
    def backward(ctx, dl):
        x, w2 = ctx.saved_tensors
        x, xp = _as_rows(x)
        n, d, h, w, cin = x.shape
        classes = w2.shape[0]
        dl = dl.contiguous().float()
        dx = torch.empty((n, d, h, w, cin), dtype=torch.bfloat16, device=x.device)
        gw = _zeros_f32(classes * cin, x.device).view(classes, cin)
        gb = _zeros_f32(classes, x.device)
        _call("b200seg_head_conv1x1_bwd", _ptr(dl), _ptr(x), xp, _ptr(w2), _ptr(dx), cin, _ptr(gw), _ptr(gb), n,
              d * h * w, cin, classes, _stream())
        return dx, gw.reshape(ctx.wshape), (gb if ctx.has_bias else None)


def head_conv1x1(x, weight, bias=None):
    return _Head.apply(x, weight, bias)


class _Upsample2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x, xp = _as_rows(x)
        n, d, h, w, c = x.shape
        y = torch.empty((n, 2 * d, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
        _call("b200seg_upsample2_fwd", _ptr(x), xp, _ptr(y), c, n, d, h, w, c, _stream())
        ctx.shape = (n, d, h, w, c)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, d, h, w, c = ctx.shape
        dy, dyp = _as_rows(dy)
        dx = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=dy.device)
        _call("b200seg_upsample2_bwd", _ptr(dy), dyp, _ptr(dx), c, n, d, h, w, c, _stream())
        return dx


def upsample_nearest2(x):
    return _Upsample2.apply(x)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, out):
        a, ap = _as_rows(a)
        b, bp = _as_rows(b)
        n, d, h, w, c = a.shape
        if out is None:
            out = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=a.device)
        else:
            assert _pitched(out) and out.shape == a.shape
            if out.data_ptr() == a.data_ptr() or out.data_ptr() == b.data_ptr():
                ctx.mark_dirty(out)
        _call("b200seg_add", _ptr(a), ap, _ptr(b), bp, _ptr(out), out.stride(3), n * d * h * w, c, _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g, None


def add(a, b, out=None):
    """a + b (element-wise); `out` may be a channel slice of a concat buffer, or alias `a` (in-place sum)."""
    return _Add.apply(a, b, out)


# ------------------------------------------------------------------------------------------------ losses / metrics
def _labels_u8(target, n, spatial_shape):
    t = target.reshape((n,) + tuple(spatial_shape))
    return t.to(torch.uint8).contiguous()


class _SegLoss(torch.autograd.Function):
    """w_ce*cross_entropy_3D + w_dice*DiceLossss(softmax=True) + w_sdice*DiceLoss(sigmoid) + w_bce*BCEWithLogits,
    from one reduction pass over the logits (loss_function.py:8-16, 102-130, 148-185; train.py:115).  class_weights:
    None or fp32 [2][K] = per-class {cross-entropy, Dice} weights (loss_function.py:13, 176-183)."""

    @staticmethod
    def forward(ctx, logits, labels, weights, class_weights):
        logits = logits.contiguous().float()
        n, classes = logits.shape[0], logits.shape[1]
        spatial = logits[0, 0].numel()
        partial = torch.zeros(1 + 3 * classes + 4, dtype=torch.float64, device=logits.device)
        w_ce, w_dice, w_sdice, w_bce = weights
        terms = (1 if (w_ce or w_dice) else 0) | (2 if (w_sdice or w_bce) else 0)
        _call("b200seg_loss_reduce", _ptr(logits), _ptr(labels), n, spatial, classes, terms or 1, _ptr(class_weights),
              _ptr(partial), _stream())
        ctx.save_for_backward(logits, labels, partial, class_weights)
        ctx.weights = weights
        vox = float(n * spatial)
        smooth = 1e-5
        loss = partial.new_zeros(())
        if w_ce:
            loss = loss + w_ce * partial[0] / vox
        if w_dice:
            pk = partial[1:1 + 3 * classes].view(classes, 3)
            per_class = 1 - (2 * pk[:, 0] + smooth) / (pk[:, 1] + pk[:, 2] + smooth)
            if class_weights is not None:
                per_class = per_class * class_weights[1].double()
            loss = loss + w_dice * per_class.mean()
        tail = partial[1 + 3 * classes:]
        if w_sdice:
            loss = loss + w_sdice * (1 - 2 * (tail[0] + smooth) / (tail[1] + tail[2] + smooth))
        if w_bce:
            loss = loss + w_bce * tail[3] / (vox * classes)
        return loss.float()

    @staticmethod
    def backward(ctx, g):
        logits, labels, partial, class_weights = ctx.saved_tensors
        n, classes = logits.shape[0], logits.shape[1]
        spatial = logits[0, 0].numel()
        w_ce, w_dice, w_sdice, w_bce = ctx.weights
        dl = torch.empty_like(logits)
        gs = g.reshape(1).to(device=logits.device, dtype=torch.float32)
        _call("b200seg_loss_grad", _ptr(logits), _ptr(labels), n, spatial, classes, _ptr(partial), float(w_ce),
              float(w_dice), float(w_sdice), float(w_bce), _ptr(gs), _ptr(class_weights), _ptr(dl), _stream())
        return dl, None, None, None


def seg_loss(logits, target, w_ce=1.0, w_dice=1.0, w_sdice=0.0, w_bce=0.0, ce_class_weight=None, dice_class_weight=None):
    """logits: fp32 [N, K, D, H, W]; target: integer class labels [N, D, H, W] or [N, 1, D, H, W].  The optional
    per-class weights follow nll_loss(weight=...) (sum, then / numel: loss_function.py:13-15) and DiceLossss's `weight`."""
    n, k = logits.shape[0], logits.shape[1]
    labels = _labels_u8(target, n, logits.shape[2:])
    cw = None
    if ce_class_weight is not None or dice_class_weight is not None:
        def vec(w):
            if w is None:
                return torch.ones(k, dtype=torch.float32, device=logits.device)
            w = torch.as_tensor(w, dtype=torch.float32, device=logits.device).reshape(-1)
            assert w.numel() == k, "expected one weight per class"
            return w
        cw = torch.stack((vec(ce_class_weight), vec(dice_class_weight))).contiguous()
    return _SegLoss.apply(logits, labels, (w_ce, w_dice, w_sdice, w_bce), cw)


class _ProbDice(torch.autograd.Function):
    """Dice on probabilities the caller supplies (BinaryDiceLoss, DiceLossss(softmax=False)): per-row losses
    1 - (a*I + smooth) / (X + T + smooth) with I = sum x*t, X = sum x^p, T = sum t^p (b200seg_dice_sums / _dice_grad)."""

    @staticmethod
    def forward(ctx, pred, target_f, target_l, cfg):
        by_class, p_exp, smooth, a = cfg
        pred = pred.contiguous().float()
        n, k = pred.shape[0], pred.shape[1]
        spatial = pred[0, 0].numel()
        rows = k if by_class else n
        partial = torch.zeros((rows, 3), dtype=torch.float64, device=pred.device)
        _call("b200seg_dice_sums", _ptr(pred), _ptr(target_f), _ptr(target_l), n, k, spatial, int(by_class), float(p_exp),
              _ptr(partial), _stream())
        den = partial[:, 1] + partial[:, 2] + smooth
        num = a * partial[:, 0] + smooth
        ctx.save_for_backward(pred, target_f, target_l, num, den)
        ctx.cfg = cfg
        return (1 - num / den).float()

    @staticmethod
    def backward(ctx, g):
        pred, target_f, target_l, num, den = ctx.saved_tensors
        by_class, p_exp, smooth, a = ctx.cfg
        n, k = pred.shape[0], pred.shape[1]
        spatial = pred[0, 0].numel()
        g = g.double()
        coef_t = (-a / den * g).float().contiguous()          # d(1 - num/den)/dx = -a*t/den + num * p x^(p-1) / den^2
        coef_x = (num / (den * den) * g).float().contiguous()
        dp = torch.empty_like(pred)
        _call("b200seg_dice_grad", _ptr(pred), _ptr(target_f), _ptr(target_l), n, k, spatial, int(by_class), float(p_exp),
              _ptr(coef_t), _ptr(coef_x), None, _ptr(dp), _stream())
        return dp, None, None, None


def prob_dice_rows(pred, target=None, labels=None, by_class=False, p=2.0, smooth=1.0, intersect_scale=1.0):
    """Per-row Dice losses on probabilities: pred [N, K, *]; `target` float of the same shape, or integer `labels`
    [N, *] compared with the class index.  Rows are samples (by_class=False) or classes (by_class=True)."""
    n, k = pred.shape[0], pred.shape[1]
    tf = target.contiguous().float().reshape(pred.shape) if target is not None else None
    tl = _labels_u8(labels, n, pred.shape[2:]) if labels is not None else None
    return _ProbDice.apply(pred, tf, tl, (by_class, float(p), float(smooth), float(intersect_scale)))


def argmax_labels(logits):
    """pred.argmax(dim=1, keepdim=True) as uint8 (train.py:204, predict.py:139); ties -> lowest class."""
    logits = logits.contiguous().float()
    n, classes = logits.shape[0], logits.shape[1]
    out = torch.empty((n, 1) + tuple(logits.shape[2:]), dtype=torch.uint8, device=logits.device)
    _call("b200seg_argmax_labels", _ptr(logits), _ptr(out), n, logits[0, 0].numel(), classes, _stream())
    return out


def seg_counts(gt, pred):
    """uint64[4] = {sum gt, sum pred, |gt & pred| != 0, |gt | pred| != 0} (metric.py:36-46), on the device."""
    gt = gt.to(torch.uint8).contiguous()
    pred = pred.to(torch.uint8).contiguous()
    assert gt.numel() == pred.numel()
    counts = torch.zeros(4, dtype=torch.int64, device=gt.device)
    _call("b200seg_seg_counts", _ptr(gt), _ptr(pred), gt.numel(), _ptr(counts), _stream())
    return counts


# ------------------------------------------------------------------------------------------------ more graph ops
def activation(x, act, act_param=0.0, prelu_weight=None, residual=None, out=None):
    """act(x [+ residual]) without normalisation (nn.ELU / nn.PReLU / nn.LeakyReLU / nn.ReLU call sites such as
    vnet3d.py:58,79,103 and residual_unet3d.py:113,120)."""
    return norm_act(x, NormSpec(None, act, act_param), prelu_weight=prelu_weight, residual=residual, out=out)


_SEEDS = {}
_SALT = [0]


def dropout_seed(device):
    """Device-resident seed of the dropout masks (int64).  engine.TrainStep bumps it once per step."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SEEDS:
        _SEEDS[key] = torch.full((1,), 0x5DEECE66D, dtype=torch.int64, device=torch.device("cuda", key))
    return _SEEDS[key]


def advance_dropout_seed(device):
    dropout_seed(device).add_(1)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg):
        p, channel_mode, salt, salt2, out = cfg
        x, xp = _as_rows(x)
        n, d, h, w, c = x.shape
        if out is None:
            out = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=x.device)
        assert _pitched(out) and out.shape == x.shape
        seed = dropout_seed(x.device)
        _call("b200seg_dropout2", _ptr(x), xp, _ptr(out), out.stride(3), n * d * h * w, d * h * w, c, c, float(p),
              _ptr(seed), salt, salt2, int(channel_mode), _stream())
        ctx.cfg = (p, channel_mode, salt, salt2)
        return out

    @staticmethod
    def backward(ctx, g):
        p, channel_mode, salt, salt2 = ctx.cfg
        g, gp = _as_rows(g)
        n, d, h, w, c = g.shape
        # the gradient is written with its channels widened (zeros) to the next multiple of 16: when the producer of x
        # was a convolution on the padded tensor-core path, its backward takes the wide buffer as it is
        c_wide = _pad16(c)
        wide = torch.empty((n, d, h, w, c_wide), dtype=torch.bfloat16, device=g.device)
        seed = dropout_seed(g.device)
        _call("b200seg_dropout2", _ptr(g), gp, _ptr(wide), c_wide, n * d * h * w, d * h * w, c, c_wide, float(p),
              _ptr(seed), salt2 or salt, salt if salt2 else 0, int(channel_mode), _stream())
        if c_wide == c:
            return wide, None
        dx = wide[..., :c]
        _publish_zero_tail(dx, wide)
        return dx, None


def dropout(x, p, training=True, channel=False, out=None, times=1):
    """nn.Dropout (channel=False) / nn.Dropout3d (channel=True).  Identity when not training or p == 0 (how the parity
    tests run: torch's Philox stream cannot be reproduced).  times=2 applies two independent masks in one pass
    (densevoxelnet3d.py:25-32 calls its dropout twice in train mode)."""
    assert times in (1, 2)
    if not training or p == 0.0:
        if out is not None:
            return activation(x, "none", out=out)      # a differentiable copy by kernel
        return x
    _SALT[0] += times
    return _Dropout.apply(x, (p, channel, _SALT[0], _SALT[0] - 1 if times == 2 else 0, out))


class _Pad3d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg):
        pad, mode = cfg
        x, xp = _as_rows(x)
        n, d, h, w, c = x.shape
        y = torch.empty((n, d + 2 * pad, h + 2 * pad, w + 2 * pad, c), dtype=torch.bfloat16, device=x.device)
        _call("b200seg_pad3d_fwd", _ptr(x), xp, _ptr(y), c, n, d, h, w, c, pad, mode, _stream())
        ctx.cfg = (pad, mode, (n, d, h, w, c))
        return y

    @staticmethod
    def backward(ctx, dy):
        pad, mode, (n, d, h, w, c) = ctx.cfg
        dy, dyp = _as_rows(dy)
        dx = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=dy.device)
        _call("b200seg_pad3d_bwd", _ptr(dy), dyp, _ptr(dx), c, n, d, h, w, c, pad, mode, _stream())
        return dx, None


def pad3d(x, pad, mode):
    """F.pad(x, 6*[pad], mode) for mode in {'reflect', 'replicate'} (utils/convolution.py:78-86); 'constant' padding never
    materialises (it is the TMA out-of-bounds fill of the convolution kernels)."""
    if pad == 0:
        return x
    return _Pad3d.apply(x, (int(pad), {"reflect": 1, "replicate": 2}[mode]))


class _AddSlice(torch.autograd.Function):
    """out[..., off:off+c_x] += x, in place (residual.py:74-83: the shortcut zero-padded to the wider channel count)."""

    @staticmethod
    def forward(ctx, out, x, off):
        x, xp = _as_rows(x)
        assert _pitched(out)
        n, d, h, w, c = x.shape
        view = out[..., off:off + c]
        _call("b200seg_add", _ptr(view), out.stride(3), _ptr(x), xp, _ptr(view), out.stride(3), n * d * h * w, c,
              _stream())
        ctx.mark_dirty(out)
        ctx.cfg = (off, c)
        return out

    @staticmethod
    def backward(ctx, g):
        off, c = ctx.cfg
        return g, g[..., off:off + c], None


def add_channel_padded(out, x):
    """x zero-padded symmetrically to out's channel count, plus out (ResidualBlock 'pad' shortcut)."""
    diff = out.shape[4] - x.shape[4]
    if diff == 0:
        return add(out, x)
    return _AddSlice.apply(out, x, diff // 2)


class _ClassmapUp2Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coarse, fine):
        coarse = coarse.contiguous().float()
        n, k, d, h, w = coarse.shape
        out = torch.empty((n, k, 2 * d, 2 * h, 2 * w), dtype=torch.float32, device=coarse.device)
        f = fine.contiguous().float() if fine is not None else None
        _call("b200seg_classmap_up2_add", _ptr(coarse), _ptr(f), _ptr(out), n * k, d, h, w, _stream())
        ctx.shape = (n, k, d, h, w)
        ctx.has_fine = fine is not None
        return out

    @staticmethod
    def backward(ctx, g):
        n, k, d, h, w = ctx.shape
        g = g.contiguous().float()
        dc = torch.empty((n, k, d, h, w), dtype=torch.float32, device=g.device)
        _call("b200seg_classmap_down2_sum", _ptr(g), _ptr(dc), n * k, d, h, w, _stream())
        return dc, (g if ctx.has_fine else None)


def classmap_up2_add(coarse, fine=None):
    """nearest x2 up-sampling of an fp32 NCDHW class-score map, plus `fine` (residual_unet3d.py:196-202)."""
    return _ClassmapUp2Add.apply(coarse, fine)


# ---- attention gates of ER-Net / RE-Net / Double-UNet (csrc/gates.cu) -------------------------------------------------
class _ConvTMap(torch.autograd.Function):
    """ConvTranspose3d(1, 1, kernel 2, stride 2) on an fp32 single-channel NCDHW map (ER_net.py:166-168)."""

    @staticmethod
    def forward(ctx, gmap, weight, bias):
        assert tuple(weight.shape) == (1, 1, 2, 2, 2) and gmap.shape[1] == 1, "single-channel k2s2 transposed convolution"
        gmap = gmap.contiguous().float()
        n, _, d, h, w = gmap.shape
        w8 = weight.detach().reshape(8).float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        out = torch.empty((n, 1, 2 * d, 2 * h, 2 * w), dtype=torch.float32, device=gmap.device)
        _call("b200seg_convt1_k2s2_fwd", _ptr(gmap), _ptr(w8), _ptr(b), _ptr(out), n, d, h, w, _stream())
        ctx.save_for_backward(gmap, w8)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        gmap, w8 = ctx.saved_tensors
        n, _, d, h, w = gmap.shape
        dout = dout.contiguous().float()
        din = torch.empty_like(gmap)
        sums = _zeros_f32(9, gmap.device)
        _call("b200seg_convt1_k2s2_bwd", _ptr(dout), _ptr(gmap), _ptr(w8), _ptr(din), _ptr(sums), n, d, h, w, _stream())
        return din, sums[:8].view(1, 1, 2, 2, 2), (sums[8:9] if ctx.has_bias else None)


def convt_map_k2s2(gmap, weight, bias=None):
    return _ConvTMap.apply(gmap, weight, bias)


class _ReverseGate(torch.autograd.Function):
    """fine * (1 - sigmoid(g)) + fine with a single-channel fp32 gate map g (ER_net.py:184-187, RE_net.py:118-121)."""

    @staticmethod
    def forward(ctx, fine, g, out):
        fine, fp = _as_rows(fine)
        n, d, h, w, c = fine.shape
        g = g.contiguous().float()
        assert g.numel() == n * d * h * w, "gate map and features differ in extent"
        if out is None:
            out = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=fine.device)
        assert _pitched(out) and out.shape == fine.shape
        _call("b200seg_reverse_gate_fwd", _ptr(fine), fp, _ptr(g), _ptr(out), out.stride(3), n * d * h * w, c, _stream())
        ctx.save_for_backward(fine, g)
        return out

    @staticmethod
    def backward(ctx, dout):
        fine, g = ctx.saved_tensors
        fine, fp = _as_rows(fine)
        dout, dp = _as_rows(dout)
        n, d, h, w, c = fine.shape
        dfine = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=fine.device)
        dg = torch.empty_like(g)
        _call("b200seg_reverse_gate_bwd", _ptr(dout), dp, _ptr(fine), fp, _ptr(g), _ptr(dfine), c, _ptr(dg), n * d * h * w, c,
              _stream())
        return dfine, dg, None


def reverse_gate(fine, g, out=None):
    return _ReverseGate.apply(fine, g, out)


class _GatedBlend(torch.autograd.Function):
    """out = x1 * w1[n, c] (+ x2 * w2[n, c]) with (w1, w2) = gate_fn(mean_voxels(x1 (+ x2)), *params).

    The pooled mean (b200seg_channel_stats per sample), the blend and both backward passes are kernels; gate_fn itself --
    two or three Linear layers on an [N, C] tensor (SE.py:32-37, ER_net.py:80-101: <= 2 x 256 x 64 multiply-adds) -- runs
    as torch ops inside, and is differentiated by torch on that tiny graph."""

    @staticmethod
    def forward(ctx, x1, x2, gate_fn, out, *params):
        x1, p1 = _as_rows(x1)
        n, d, h, w, c = x1.shape
        vox = d * h * w
        pooled = channel_stats(x1, n)[:, 0, :]
        if x2 is not None:
            x2, p2 = _as_rows(x2)
            assert x2.shape == x1.shape
            pooled = pooled + channel_stats(x2, n)[:, 0, :]
        with torch.enable_grad():
            s = (pooled / vox).detach().requires_grad_(True)
            ps = [p.detach().requires_grad_(True) for p in params]
            w1, w2 = gate_fn(s, *ps)
            w1 = w1.float().contiguous()
            w2 = w2.float().contiguous() if w2 is not None else None
        assert (w2 is not None) == (x2 is not None)
        if out is None:
            out = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=x1.device)
        assert _pitched(out) and out.shape == x1.shape
        _call("b200seg_channel_blend_fwd", _ptr(x1), p1, _ptr(w1.detach()), _ptr(x2), p2 if x2 is not None else 0,
              _ptr(w2.detach() if w2 is not None else None), _ptr(out), out.stride(3), vox, n, c, _stream())
        ctx.x1, ctx.x2, ctx.graph = x1, x2, (s, ps, w1, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        x1, x2 = ctx.x1, ctx.x2
        s, ps, w1, w2 = ctx.graph
        x1, p1 = _as_rows(x1)
        dout, dp = _as_rows(dout)
        n, d, h, w, c = x1.shape
        vox = d * h * w
        dots = _zeros_f32(2 * n * c, x1.device).view(2, n, c)
        x2p = 0
        if x2 is not None:
            x2, x2p = _as_rows(x2)
        _call("b200seg_channel_blend_bwd_reduce", _ptr(dout), dp, _ptr(x1), p1, _ptr(x2), x2p, _ptr(dots[0]),
              _ptr(dots[1] if x2 is not None else None), vox, n, c, _stream())
        outs, gouts = [w1], [dots[0]]
        if w2 is not None:
            outs.append(w2)
            gouts.append(dots[1])
        grads = torch.autograd.grad(outs, [s] + ps, gouts, allow_unused=True)
        add = (grads[0] / vox).float().contiguous() if grads[0] is not None else None
        dx1 = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=x1.device)
        dx2 = torch.empty_like(dx1) if x2 is not None else None
        _call("b200seg_channel_blend_bwd_apply", _ptr(dout), dp, _ptr(w1.detach()), _ptr(w2.detach() if w2 is not None else None),
              _ptr(add), _ptr(dx1), c, _ptr(dx2), c, vox, n, c, _stream())
        ctx.graph = None
        return (dx1, dx2, None, None) + tuple(grads[1:])


def gated_blend(x1, x2, gate_fn, params, out=None):
    """Squeeze-and-excitation / selective-fusion gates: see _GatedBlend.  gate_fn(pooled_mean [N, C], *params) returns the
    per-(sample, channel) weights (w1, w2); w2 is None when there is no second input."""
    return _GatedBlend.apply(x1, x2, gate_fn, out, *params)


class _SigmoidMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        y = torch.empty_like(x)
        _call("b200seg_f32_sigmoid_fwd", _ptr(x), _ptr(y), x.numel(), _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        _call("b200seg_f32_sigmoid_bwd", _ptr(dy), _ptr(y), _ptr(dx), y.numel(), _stream())
        return dx


def sigmoid_map(x):
    """torch.sigmoid on an fp32 class-score map (RE_net.py:158)."""
    return _SigmoidMap.apply(x)


def alloc_channels(n, d, h, w, c, device):
    """Uninitialised [n, d, h, w, c] activation buffer whose channel slices are filled by `out=` of the producing ops
    (DenseVoxelNet's dense blocks, densevoxelnet3d.py:36-42)."""
    return torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=device)


def zeros_ndhwc(n, d, h, w, c, device):
    return torch.zeros((n, d, h, w, c), dtype=torch.bfloat16, device=device)


def repeat_channels(x, times):
    """x.repeat(1, times, 1, 1, 1) for a channels-last tensor (vnet3d.py:57; the input has 1-2 channels)."""
    return x.repeat(1, 1, 1, 1, times)


def spatial(x):
    """(n, d, h, w) of an activation tensor."""
    return tuple(x.shape[:4])


def channels(x):
    return x.shape[4]


def device_of(x):
    return x.device


def channel_slice(x, lo, hi):
    return x[..., lo:hi]


class _Concat(torch.autograd.Function):
    """torch.cat((a, b), channel) as a graph node: free when the two are the halves of one alloc_concat buffer."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[4]
        return merge_channels(a, b)

    @staticmethod
    def backward(ctx, g):
        return g[..., :ctx.ca], g[..., ctx.ca:]


def concat_channels(a, b):
    return _Concat.apply(a, b)


# Dense blocks (densevoxelnet3d.py:36-42): the tensor `a` of cat((a, new)) is also the input of the layer that produced
# `new`, so its gradient has two parts -- the head slice of the concatenation's gradient and what comes back through that
# layer.  Instead of letting autograd sum a strided slice and a dense tensor (a non-vectorised ATen add per layer), the
# concatenation's backward leaves the head slice here, keyed by `a`, and returns nothing for it; the layer's normalise
# backward (norm_act(..., shared_grad=True) on the same `a`) adds its result INTO that slice and returns it.  Entries hold
# the tensors, so an address cannot be recycled under a key; begin_step / end_step clear the table.
_GRAD_STASH = {}


class _ConcatShared(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[4]
        ctx.key = (a.data_ptr(), tuple(a.shape), tuple(a.stride()))
        ctx.keep = a
        return merge_channels(a, b)

    @staticmethod
    def backward(ctx, g):
        head = g[..., :ctx.ca]
        if _pitched(head) and head.data_ptr() % 16 == 0 and ctx.needs_input_grad[0]:
            _GRAD_STASH[ctx.key] = (ctx.keep, head)
            ctx.keep = None
            return None, g[..., ctx.ca:]
        return head, g[..., ctx.ca:]


def concat_channels_shared(a, b):
    """concat_channels for a tensor `a` that is ALSO consumed by norm_act(a, ..., shared_grad=True): see _GRAD_STASH."""
    return _ConcatShared.apply(a, b)
