"""Train-step and sliding-window timings of the other in-scope configurations (BASELINE.json configs[2..4]); diagnostic,
the graded bench line is bench.py (configs[1]).  usage: bench_models.py [vnet res_unet highres densevoxel predict]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200seg.engine import TrainStep
from b200seg.optim import FusedAdam
from b200seg.utils.loss_function import DiceCELoss

dev = torch.device("cuda")
CASES = {
    "vnet": (lambda: __import__("b200seg.models.three_d.vnet3d", fromlist=["VNet"]).VNet(True, 1, 2), 128, 2, 1463.1),
    "res_unet": (lambda: __import__("b200seg.models.three_d.residual_unet3d", fromlist=["UNet"]).UNet(1, 2, 32), 128, 2, 1820.7),
    "highres": (lambda: __import__("b200seg.models.three_d.highresnet", fromlist=["HighRes3DNet"]).HighRes3DNet(1, 2), 96, 2, 1419.7),
    "densevoxel": (lambda: __import__("b200seg.models.three_d.densevoxelnet3d", fromlist=["DenseVoxelNet"]).DenseVoxelNet(1, 2), 96, 2, 144.7),
}


def train_case(name):
    make, size, batch, gflop_fwd = CASES[name]
    torch.manual_seed(0)
    net = make().to(dev).train()
    opt = FusedAdam(net.parameters(), lr=1e-3)
    step = TrainStep(net, DiceCELoss(2), opt, use_graph=True)
    x = torch.randn(batch, 1, size, size, size, device=dev)
    lab = (torch.rand(batch, size, size, size, device=dev) > 0.9).to(torch.uint8)
    for _ in range(6):
        loss, _ = step(x, lab)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        loss, _ = step(x, lab)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"config": name, "patch": size, "batch": batch, "ms_per_step": round(ms, 3),
                      "patches_per_s": round(batch / ms * 1e3, 2), "graph": step.graph is not None,
                      "train_tflops": round(3 * gflop_fwd * batch / ms, 1), "loss": round(float(loss), 4),
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}), flush=True)
    del net, opt, step
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


def predict_case():
    from b200seg.inference import sliding_window_predict
    from b200seg.models.three_d.unet3d import UNet3D
    torch.manual_seed(0)
    net = UNet3D(1, 2, 32).to(dev).eval()
    vol = torch.randn(1, 512, 512, 256, device=dev)
    for mode in ("crop", "average"):
        sliding_window_predict(net, vol[:, :256, :256, :128], (128,) * 3, (64,) * 3, batch_size=4, overlap_mode=mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = sliding_window_predict(net, vol, (128,) * 3, (64,) * 3, batch_size=16, overlap_mode=mode)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"config": "predict 512x512x256, 128^3 patches, overlap 64 (147 patches), batch 16", "mode": mode,
                          "seconds_per_volume": round(dt, 3), "patches_per_s": round(147 / dt, 1),
                          "fwd_tflops": round(147 * 951.3 / dt / 1e3, 1), "labels": list(out.shape)}), flush=True)


for name in (sys.argv[1:] or ["vnet", "res_unet", "highres", "densevoxel", "predict"]):
    try:
        predict_case() if name == "predict" else train_case(name)
    except Exception as e:   # keep going: this is a survey of configurations
        print(json.dumps({"config": name, "error": repr(e)[:300]}), flush=True)
