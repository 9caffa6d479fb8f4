"""World-size-2 gloo tests (CPU) of the data-parallel host logic in b200seg.parallel: bucketed gradient all-reduce,
SyncBatchNorm statistics exchange, patch sharding / volume merge, metric count reduction.  The collectives move the same
tensors the CUDA path moves; the arithmetic around them is checked against the oracle and the committed golden vectors."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, result_dir):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from b200seg import parallel
    rk, local, ws = parallel.init_from_env("gloo")
    assert (rk, ws) == (rank, world) and parallel.is_parallel()
    try:
        globals()[fn_name](rank, world, parallel)
        open(os.path.join(result_dir, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def _run(fn_name, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), fn_name, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


# ---------------------------------------------------------------------------------------------- worker bodies
def _make_net():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(12, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                               torch.nn.Linear(40, 3))


def _arena_for(net):
    params = [p for p in net.parameters()]
    offsets, total = [], 0
    for p in params:
        offsets.append(total)
        total += (p.numel() + 63) // 64 * 64
    arena = torch.zeros(total)
    for p, off in zip(params, offsets):
        p.grad = arena[off:off + p.numel()].view_as(p)
    return arena, params, offsets


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(5, 12, generator=g), torch.randn(5, 3, generator=g)


def body_grad_buckets(rank, world, parallel):
    net = _make_net()
    arena, params, offsets = _arena_for(net)
    red = parallel.GradBucketReducer(arena, params, offsets, bucket_bytes=256)
    assert len(red.buckets) >= 3                       # several buckets, cut from the end of the arena
    assert red.buckets[0][1] == arena.numel() and red.buckets[-1][0] == 0
    covered = sorted((s, e) for s, e, _ in red.buckets)
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    for step in range(2):                              # twice: the reducer must re-arm itself
        arena.zero_()
        x, y = _data(rank)
        ((net(x) - y) ** 2).mean().backward()
        scale = red.finish()
        assert scale == 1.0 / world
        # expected: mean over ranks of the per-rank gradients, recomputed locally
        ref = _make_net()
        want = [torch.zeros_like(p) for p in ref.parameters()]
        for r in range(world):
            ref.zero_grad()
            xr, yr = _data(r)
            ((ref(xr) - yr) ** 2).mean().backward()
            for w, p in zip(want, ref.parameters()):
                w += p.grad / world
        for p, w in zip(params, want):
            assert torch.allclose(p.grad * scale, w, rtol=1e-5, atol=1e-7)
    red.remove()


def body_syncbn_stats(rank, world, parallel):
    from oracle import syncbn
    g = np.load(os.path.join(GOLDEN, "syncbn.npz"))
    x = torch.from_numpy(g["x%d" % rank])
    s, ss, n = syncbn.local_sums(x)
    stats = torch.stack((s, ss)).unsqueeze(0).contiguous()      # [1][2][C], the layout the kernels exchange
    ranks = parallel.all_reduce_stats(stats)
    assert ranks == world
    # the golden replicas hold 2 and 3 samples: the reference's message carries sum_size (batchnorm.py:58-62)
    total = torch.tensor([float(n)])
    parallel.all_reduce_stats(total)
    mean, inv_std, rm, rv = syncbn.compute_mean_std(stats[0, 0], stats[0, 1], int(total.item()), torch.zeros(6),
                                                    torch.ones(6))
    assert torch.allclose(mean, torch.from_numpy(g["mean"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(inv_std, torch.from_numpy(g["inv_std"]), rtol=1e-6)
    assert torch.allclose(rv, torch.from_numpy(g["running_var"]), rtol=1e-6)
    w, b = torch.from_numpy(g["weight"]), torch.from_numpy(g["bias"])
    shp = (1, -1, 1, 1, 1)
    out = (x - mean.view(shp)) * (inv_std * w).view(shp) + b.view(shp)
    assert torch.allclose(out, torch.from_numpy(g["out%d" % rank]), rtol=1e-5, atol=1e-6)


def body_window_shards(rank, world, parallel):
    """Each rank stitches its round-robin share; keys (patch index, label) merged with MAX reproduce the sequential
    last-writer-wins result of the oracle aggregator, including where two cropped interiors overlap."""
    from oracle import window
    rng = np.random.default_rng(0)
    shape, patch, ov = (40, 36, 50), (16, 16, 16), (4, 4, 6)
    locs = window.grid_locations(shape, patch, ov)
    # patches that DISAGREE where they overlap: label depends on the patch index
    patches = np.stack([rng.integers(0, 3, size=(1,) + patch) for _ in locs]).astype(np.int64)
    seq = window.Aggregator(shape, ov, "crop")
    seq.add_batch(patches, locs)
    want = seq.get_output_tensor()
    mine = parallel.shard_patches(len(locs))
    assert mine == list(range(rank, len(locs), world))
    key = np.zeros((1,) + shape, np.int32)
    for i in mine:                                        # what window_crop does on the device, per patch
        agg = window.Aggregator(shape, ov, "crop")
        c, a, b = agg._crop(patches[i], locs[i])
        region = key[:, a[0]:b[0], a[1]:b[1], a[2]:b[2]]
        np.maximum(region, ((i + 1) << 8) | c.astype(np.int32), out=region)
    merged = parallel.reduce_volume(torch.from_numpy(key), op="max").numpy()
    assert np.array_equal(merged & 255, want)
    acc = torch.full((2, 3), float(rank + 1))
    assert torch.equal(parallel.reduce_volume(acc), torch.full((2, 3), 3.0))


def body_metric_counts(rank, world, parallel):
    from oracle import metric
    g = np.load(os.path.join(GOLDEN, "metric.npz"))
    gt, pred = g["gt%d" % rank], g["pred%d" % rank]
    gi, pi = (gt != 0), (pred != 0)
    counts = torch.tensor([int(gt.astype(np.int64).sum()), int(pred.astype(np.int64).sum()), int((gi & pi).sum()),
                           int((gi | pi).sum())], dtype=torch.int64)
    total = parallel.all_reduce_counts(counts.clone())
    both_gt = np.concatenate([g["gt0"].ravel(), g["gt1"].ravel()])
    both_pred = np.concatenate([g["pred0"].ravel(), g["pred1"].ravel()])
    want = metric.metric(both_gt, both_pred)
    dice = 2 * total[2].item() / (total[0].item() + total[1].item() + 0.001)
    assert abs(dice - want["dice"]) < 1e-12
    assert abs(total[2].item() / (total[3].item() + 0.001) - want["jaccard"]) < 1e-12


# ---------------------------------------------------------------------------------------------- tests
@pytest.mark.parametrize("body", ["body_grad_buckets", "body_syncbn_stats", "body_window_shards", "body_metric_counts"])
def test_world2_gloo(body, tmp_path):
    _run(body, tmp_path)


def test_single_process_paths_are_noops():
    from b200seg import parallel
    assert not parallel.is_parallel() and parallel.world_size() == 1 and parallel.rank() == 0
    t = torch.ones(3)
    assert parallel.all_reduce_stats(t) == 1 and torch.equal(t, torch.ones(3))
    assert parallel.shard_patches(5) == [0, 1, 2, 3, 4]
    net = _make_net()
    arena, params, offsets = _arena_for(net)
    red = parallel.GradBucketReducer(arena, params, offsets)
    assert red.finish() == 1.0


class _SideGrad(torch.autograd.Function):
    """Like the conv weight gradient of the CUDA path: the weight's gradient is written OUTSIDE autograd (into `side`) and
    backward returns None for it -- torch still fires the parameter's post-accumulate-grad hook."""

    @staticmethod
    def forward(ctx, x, w, side):
        ctx.save_for_backward(x, w)
        ctx.side = side
        return x @ w

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        ctx.side.add_(x.t() @ g)          # the "packed accumulator"
        return g @ w.t(), None, None


def body_reducer_pre_launch(rank, world, parallel):
    """GradBucketReducer(pre_launch=...): the callback that moves externally accumulated gradients into a bucket must run
    before that bucket's all-reduce even when the bucket is launched from a hook (round-1 defect: hooks fire for None
    gradients, the bucket went out before its side gradients had been moved in)."""
    torch.manual_seed(3)
    w1 = torch.nn.Parameter(torch.randn(6, 5))
    w2 = torch.nn.Parameter(torch.randn(5, 4))
    b = torch.nn.Parameter(torch.randn(4))
    params, offsets, arena = [w1, w2, b], [0, 64, 128], torch.zeros(192)
    for p, off in zip(params, offsets):
        p.grad = arena[off:off + p.numel()].view_as(p)
    side = [torch.zeros(6, 5), torch.zeros(5, 4)]
    calls = []

    def pre_launch(bucket, start, end):
        calls.append(bucket)
        for p, off, s in zip(params[:2], offsets[:2], side):
            if start <= off < end:
                arena[off:off + p.numel()].view_as(p).copy_(s)
    red = parallel.GradBucketReducer(arena, params, offsets, bucket_bytes=1, pre_launch=pre_launch)   # one bucket per parameter
    g = torch.Generator().manual_seed(10 + rank)
    x = torch.randn(3, 6, generator=g)
    out = _SideGrad.apply(torch.tanh(_SideGrad.apply(x, w1, side[0])), w2, side[1]) + b
    out.pow(2).sum().backward()
    scale = red.finish()
    assert sorted(calls) == [0, 1, 2] and scale == 1.0 / world
    # reference: plain autograd on every rank's data, summed
    tot = [torch.zeros_like(p) for p in params]
    for r in range(world):
        g = torch.Generator().manual_seed(10 + r)
        xr = torch.randn(3, 6, generator=g)
        ps = [p.detach().clone().requires_grad_(True) for p in params]
        ((torch.tanh(xr @ ps[0]) @ ps[1] + ps[2]).pow(2).sum()).backward()
        for t, p in zip(tot, ps):
            t += p.grad
    for p, t in zip(params, tot):
        assert torch.allclose(p.grad, t, rtol=1e-5, atol=1e-6)


def test_reducer_pre_launch_moves_side_gradients_before_the_all_reduce(tmp_path):
    _run("body_reducer_pre_launch", tmp_path)
