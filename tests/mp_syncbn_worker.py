"""Two-rank worker for tests/test_gpu_syncbn.py (launched under torchrun, one process per GPU).

1. SynchronizedBatchNorm3d (train mode) over the NVLink mailbox == single-process batch norm over the concatenated batch
   (batchnorm.py:48-125): outputs, running statistics and the input gradient, which needs the
   backward all-reduce of {sum dy, sum dy*xhat}.
2. engine.TrainStep in multi-GPU CUDA-graph mode (forward + backward + weight-gradient transpose captured, NCCL gradient
   all-reduce + Adam after every replay) follows the eager loop: parameters after 5 steps agree, and both ranks hold
   bit-identical parameters.
Prints MP_SYNCBN_OK on rank 0 when everything passed."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    from b200seg import parallel
    from b200seg.engine import TrainStep
    from b200seg.models.sync_batchnorm.batchnorm import SynchronizedBatchNorm3d, convert_model
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.optim import FusedAdam
    from b200seg.utils.loss_function import DiceCELoss
    from oracle import syncbn as osync
    from oracle import unet3d as ounet
    rank, local, world = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    assert world >= 2 and parallel.peer_exchange() is not None, "NVLink peer exchange unavailable"

    # ---- 1. SyncBN parity (all ranks generate all shards from one seed) ------------------------------------------------
    g = torch.Generator().manual_seed(7)
    c = 24
    shards = [(torch.randn(2, c, 6, 8, 8, generator=g) * (1 + 0.3 * r) + 0.2 * r).bfloat16().float() for r in range(world)]
    gouts = [torch.randn(2, c, 6, 8, 8, generator=g).bfloat16().float() for r in range(world)]
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.2
    bn = SynchronizedBatchNorm3d(c).to(dev)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    bn.train()
    x = shards[rank].to(dev).requires_grad_(True)
    out = bn(x)
    out.backward(gouts[rank].to(dev))
    torch.cuda.synchronize()
    # reference: the oracle's master arithmetic (clamp(eps)) on the concatenated batch, autograd for the gradients
    xs = [s.clone().requires_grad_(True) for s in shards]
    wr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    outs, mean, inv_std, o_rm, o_rv = osync.forward_replicas(xs, wr, br, torch.zeros(c), torch.ones(c))
    sum((o * go).sum() for o, go in zip(outs, gouts)).backward()
    assert rel(out.detach().cpu(), outs[rank].detach()) < 1e-2, ("syncbn out", rel(out.detach().cpu(), outs[rank].detach()))
    assert rel(bn.running_mean.cpu(), o_rm.detach()) < 1e-4 and rel(bn.running_var.cpu(), o_rv.detach()) < 1e-4
    assert rel(x.grad.cpu(), xs[rank].grad) < 2e-2, ("syncbn dx", rel(x.grad.cpu(), xs[rank].grad))
    # local affine gradients sum to the global ones over ranks (what the gradient all-reduce then does)
    gw = bn.weight.grad.clone()
    dist.all_reduce(gw)
    assert rel(gw.cpu(), wr.grad) < 2e-2

    # ---- 2. graph replay == eager in split (multi-GPU) mode -----------------------------------------------------------
    sd = ounet.init_state_dict(1, 2, 16, seed=3)
    torch.manual_seed(100 + rank)
    data = [(torch.randn(2, 1, 32, 32, 32, device=dev), (torch.rand(2, 32, 32, 32, device=dev) > 0.8).to(torch.uint8))
            for _ in range(7)]
    results = []
    # (eager, NCCL bucket reducer) is the yardstick; (graph, NCCL after the replay) and (graph / eager, NVLink peer-memory
    # exchange inside the stream) must follow it
    for use_graph, peer, reducer in ((False, False, False), (True, False, False), (True, True, False), (False, True, False),
                                     (False, False, True)):
        net = UNet3D(1, 2, 16).to(dev)
        net.load_state_dict(sd)
        convert_model(net)
        net.train()
        opt = FusedAdam(net.parameters(), lr=1e-3, peer_grads=peer)
        assert opt.peer_grads == peer
        if reducer:
            opt.attach_reducer()          # bucketed NCCL all-reduce launched from autograd hooks (eager mode only)
        step = TrainStep(net, DiceCELoss(2), opt, use_graph=use_graph, warmup=2)
        losses = [float(step(xb, yb)[0]) for xb, yb in data]
        assert (step.graph is not None) == use_graph
        if use_graph:
            assert step._split == (not peer)
        torch.cuda.synchronize()
        results.append((losses, {k: v.detach().clone() for k, v in net.state_dict().items()}))
        # data-parallel invariant: every rank holds the same parameters, bit for bit
        flat = torch.cat([v.flatten().float() for v in net.parameters()])
        other = flat.clone()
        dist.broadcast(other, 0)
        assert torch.equal(flat, other), "ranks diverged (graph=%s peer=%s reducer=%s): max diff %g" % (
            use_graph, peer, reducer, float((flat - other).abs().max()))
        if hasattr(opt, "reducer"):
            opt.reducer.remove()
    l0, p0 = results[0]
    for (l1, p1), what in zip(results[1:], ("graph+nccl", "graph+peer", "eager+peer", "eager+bucket-reducer")):
        assert max(abs(a - b) for a, b in zip(l0, l1)) < 6e-3, (what, l0, l1)
        for k, tol in (("encoder1.enc1conv1.weight", 2e-2), ("decoder1.dec1conv2.weight", 2e-2), ("conv.weight", 2e-2),
                       ("upconv1.weight", 2e-2), ("encoder2.enc2norm1.running_var", 2e-2)):
            e = rel(p1[k].float().cpu(), p0[k].float().cpu())
            assert e < tol, (what, k, e)
        # the conv weights really moved in the replayed steps (the weight gradients were not dropped after the first replay)
        moved = rel(p1["decoder1.dec1conv2.weight"].cpu(), sd["decoder1.dec1conv2.weight"])
        assert moved > 1e-3, (what, moved)
    dist.barrier()
    if rank == 0:
        print("MP_SYNCBN_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
