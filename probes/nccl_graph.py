"""Does an NCCL all-reduce survive CUDA-graph capture here (main stream and a forked side stream)?"""
import os, sys, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(1024, device="cuda") * (dist.get_rank() + 1)
y = torch.ones(1 << 20, device="cuda")
side = torch.cuda.Stream()
for _ in range(3):
    dist.all_reduce(x)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dist.all_reduce(y)
    torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
print(dist.get_rank(), "eager ok", x[0].item(), flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "main"
g = torch.cuda.CUDAGraph()
x.fill_(dist.get_rank() + 1); y.fill_(1)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    dist.all_reduce(x)
    if mode == "side":
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            dist.all_reduce(y)
        torch.cuda.current_stream().wait_stream(side)
    z = x * 2
torch.cuda.synchronize()
print(dist.get_rank(), "captured", flush=True)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
print(dist.get_rank(), "replayed", mode, x[0].item(), z[0].item(), y[0].item(), flush=True)
dist.destroy_process_group()
