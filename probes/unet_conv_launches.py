"""The convolution launches of ONE UNet3D(1,2,32) step, alone, between cudaProfilerStart/Stop -- the target of the ncu pass
whose DRAM bytes bench.py quotes as roofline.traffic (probes/make_traffic_json.py turns the CSV into profiles/r02_traffic.json).

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none --csv --log-file gpurun_out/r02_conv_train.csv python probes/unet_conv_launches.py train
    ... python probes/unet_conv_launches.py predict        (forward launches of one batch of 16 patches, eval mode)

train:   the 18 forward and 17 data-gradient 3x3x3 launches at batch 2 x 128^3 (the first layer has no data gradient).
predict: the 18 forward launches at batch 16 x 128^3 (one of the 147/16 batches of a 512x512x256 volume)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200seg.functional as F

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
dev = torch.device("cuda")
f, s, n = 32, 128, (2 if mode == "train" else 16)
LAYERS = [(1, f, s), (f, f, s), (f, 2 * f, s // 2), (2 * f, 2 * f, s // 2), (2 * f, 4 * f, s // 4), (4 * f, 4 * f, s // 4),
          (4 * f, 8 * f, s // 8), (8 * f, 8 * f, s // 8), (8 * f, 16 * f, s // 16), (16 * f, 16 * f, s // 16),
          (16 * f, 8 * f, s // 8), (8 * f, 8 * f, s // 8), (8 * f, 4 * f, s // 4), (4 * f, 4 * f, s // 4),
          (4 * f, 2 * f, s // 2), (2 * f, 2 * f, s // 2), (2 * f, f, s), (f, f, s)]
torch.manual_seed(0)
cases = []
for ci, co, e in LAYERS:
    x = torch.randn(n, e, e, e, ci, device=dev).bfloat16()
    w = torch.randn(co, ci, 3, 3, 3, device=dev) * 0.05
    b = torch.zeros(co, device=dev)
    dy = torch.randn(n, e, e, e, co, device=dev).bfloat16() if mode == "train" else None
    cases.append((x, w, b, dy))
# warm-up (tensor maps, attribute calls), then flush the L2 so that every launch starts cold like in a real step
for x, w, b, dy in cases[:2]:
    F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, mode == "train")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i, (x, w, b, dy) in enumerate(cases):
    flush.zero_()
    y, stats, g = F.conv3d_fprop_raw(x, w, b, 3, 1, 1, 1, mode == "train")
    if mode == "train" and i > 0:
        flush.zero_()
        F.conv3d_dgrad_raw(g, dy, w)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", mode, len(cases))
