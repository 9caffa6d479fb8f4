"""2+ GPU check of the NVLink peer exchange (csrc/p2p.cu): eager, repeated (slot reuse), fused finalize vs the
NCCL + norm_finalize path, and inside a CUDA graph.  torchrun --nproc-per-node N probes/p2p_test.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from b200seg import parallel
import b200seg.functional as F
rank, local, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
px = parallel.peer_exchange()
assert px is not None, "peer exchange unavailable"
torch.manual_seed(rank)
for it in range(200):                      # slot reuse under back-to-back exchanges of varying size
    n = [7, 65, 2049, 1024][it % 4]
    v = torch.randn(n, device="cuda")
    want = v.clone(); dist.all_reduce(want)
    got = px.all_reduce_(v.clone())
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), (it, (got - want).abs().max())
# bit-identical on every rank
chk = got.clone(); dist.broadcast(chk, 0); assert torch.equal(chk, got)
# fused finalize vs all-reduce + norm_finalize
C = 64
stats = torch.zeros(2 * C + 1, device="cuda"); stats[:C] = torch.randn(C, device="cuda") * 50
stats[C:2 * C] = stats[:C] ** 2 / 4096 + torch.rand(C, device="cuda") * 100
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
rm1, rv1, rm2, rv2 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
ref = stats[:2 * C].clone(); dist.all_reduce(ref)
coef_ref = F._norm_coef(ref, 4096.0 * world, 1, C, gamma, beta, rm1, rv1, 0.1, 1e-5, True, stats.device)
coef = px.reduce_and_finalize(stats.clone(), 4096.0, C, gamma, beta, rm2, rv2, 0.1, 1e-5, True)
assert torch.allclose(coef, coef_ref, rtol=1e-5, atol=1e-6), (coef - coef_ref).abs().max()
assert torch.allclose(rm1, rm2) and torch.allclose(rv1, rv2)
# inside a CUDA graph
v = torch.ones(33, device="cuda") * (rank + 1)
out = torch.empty_like(v)
torch.cuda.synchronize(); dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    w = v * 2
    px.all_reduce_(w)
    out.copy_(w)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
assert torch.allclose(out, torch.full_like(out, 2.0 * sum(range(1, world + 1)))), out[:4]
dist.barrier()
if rank == 0:
    # latency
    v = torch.randn(129, device="cuda")
    pass
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
v = torch.randn(129, device="cuda")
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(200):
    px.all_reduce_(v)
e1.record(); torch.cuda.synchronize()
t_p2p = e0.elapsed_time(e1) / 200
e0.record()
for _ in range(200):
    dist.all_reduce(v)
e1.record(); torch.cuda.synchronize()
t_nccl = e0.elapsed_time(e1) / 200
print("rank %d: p2p exchange OK (eager, graph, fused finalize); 129 floats: p2p %.1f us/call, NCCL %.1f us/call (incl. launch)" % (rank, t_p2p * 1e3, t_nccl * 1e3), flush=True)
dist.destroy_process_group()
