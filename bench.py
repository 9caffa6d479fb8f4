#!/usr/bin/env python
"""Benchmark of the hot path: 3D U-Net(1,2,32) training step on synthetic 128^3 single-channel patches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = host/HBM batch 2x1x128^3 per GPU -> forward -> Dice+CE -> backward -> (gradient all-reduce) -> Adam.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own log (its version banner when NCCL_DEBUG is set) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

PATCH = 128
FEATURES = 32
BATCH = 2
# DRAM bytes of the fprop + dgrad conv launches of one step (ncu, profiles/r01_step4_final_launch_list.md) and their
# algorithmic in + out + weight bytes (3.68 GB per forward, about the same for the data gradients)
NCU_CONV_TRAFFIC_BYTES_PER_STEP = 5.69e9
CONV_ALGORITHMIC_BYTES_PER_STEP = 7.3e9
WORKLOAD = ("UNet3D(1,2,32) train step, batch 2x1x128^3 per GPU, Dice+CE, BatchNorm (SyncBatchNorm across ranks when "
            "N > 1), Adam (BASELINE.json configs[1])")
TRAIN_GFLOP_PER_PATCH = 2850.4  # fwd 951.3 + wgrad 951.3 + dgrad (951.3 - 3.6 first layer): BASELINE.md section 3


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: keep a private duplicate of fd 1 for it and point fd 1 at stderr
    for the rest of the run, so that whatever a library prints to stdout (NCCL's version banner under NCCL_DEBUG, which
    ignores NCCL_DEBUG_FILE on this stack) cannot end up in front of the result."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_step_factory(size, batch=1):
    """The reference's CPU implementation of the path (oracle port: same torch.nn arithmetic, fp32, all host cores)."""
    import torch
    from oracle import losses as olosses
    from oracle import unet3d as ounet
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    sd = ounet.init_state_dict(1, 2, FEATURES, seed=0)
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3)
    x = torch.randn(batch, 1, size, size, size)
    lab = (torch.rand(batch, size, size, size) > 0.9).long()

    def step():
        opt.zero_grad(set_to_none=True)
        stats = {}
        out = ounet.forward(sd, x, training=True, new_stats=stats)
        loss = olosses.dice_ce(out, lab)
        loss.backward()
        opt.step()
        for k, v in stats.items():
            sd[k] = v
        return float(loss)
    return step


def time_cpu(size, iters, warmup=1):
    step = oracle_step_factory(size)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    size = 64  # bounded sample: one 1x1x64^3 patch per step = 1/8 of a 128^3 patch
    step = oracle_step_factory(size)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = (size / PATCH) ** 3 / dt
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "train_patches_per_s_128cubed", "value": value, "unit": "patches/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * args.gpus, "parallelism": "dp%d" % args.gpus,
                       "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port",
                             "sample": "oracle port of the reference modules, fp32 torch CPU, %d threads; each step = "
                                       "fwd+bwd+Adam on one 1x1x64^3 patch, counted as 1/8 of a 128^3 patch" % cores},
            "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ b200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    import b200seg.functional as F
    from b200seg.models.sync_batchnorm.batchnorm import convert_model
    from b200seg.models.three_d.unet3d import UNet3D
    from b200seg.optim import FusedAdam
    from b200seg.utils.loss_function import DiceCELoss

    from b200seg import parallel
    rank, local, world = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(1234 + rank)

    net = UNet3D(1, 2, FEATURES).to(dev)
    if world > 1:
        parallel.broadcast_parameters(net)
        convert_model(net)       # nn.BatchNorm3d -> SynchronizedBatchNorm3d (statistics all-reduced over NCCL)
    net.train()
    opt = FusedAdam(net.parameters(), lr=1e-3)
    if world > 1:
        opt.attach_reducer()     # bucketed gradient all-reduce on a side stream, overlapped with backward
    crit = DiceCELoss(2)
    vox = PATCH ** 3
    x_host = torch.randn(BATCH, 1, PATCH, PATCH, PATCH).pin_memory()
    lab_host = (torch.rand(BATCH, PATCH, PATCH, PATCH) > 0.9).to(torch.uint8).pin_memory()
    x_dev, lab_dev = x_host.to(dev), lab_host.to(dev)

    from b200seg.engine import TrainStep
    train_step = TrainStep(net, crit, opt, use_graph=not args.no_graph)   # train.py:187-214 as one CUDA graph

    def step(x, lab):
        return train_step(x, lab)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3) + 2):   # TrainStep runs 3 eager steps, captures, then replays
        step(x_dev, lab_dev)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    F.reset_launches()
    umma0 = F.umma_launch_count()
    ms_total = timed(lambda: step(x_dev, lab_dev), args.steps)
    launches = F.launches()
    umma_launches = F.umma_launch_count() - umma0
    if train_step.graph is not None:   # replays do not pass through the Python wrappers: count the graph's kernel nodes
        launches = train_step.kernels_per_step * args.steps
        umma_launches = train_step.umma_per_step * args.steps

    # end to end through the public API (the loop of b200seg/train.py): every step's batch is copied from pinned host memory
    # inside the timed region -- by data.DevicePrefetcher, one batch ahead on a copy stream -- and every step's result
    # (the loss) is read back on the host before the next step starts.
    from b200seg.data import DevicePrefetcher

    def e2e_run(steps):
        def run():
            for x, lab in DevicePrefetcher(((x_host, lab_host) for _ in range(steps)), dev):
                step(x, lab).item()
        return run
    e2e_run(2)()
    ms_e2e = timed(e2e_run(args.steps), 1)
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel timing pass (CUDA events around every launch on the launching stream, eager) -> roofline
    eager = TrainStep(net, crit, opt, use_graph=False)
    eager(x_dev, lab_dev)
    F.profile_begin()
    for _ in range(2):
        eager(x_dev, lab_dev)
    prof = F.profile_end()

    if rank == 0:
        peaks = load_peaks()
        ms_step = ms_total / args.steps
        patches = BATCH * world
        value = patches / (ms_step * 1e-3)
        tc = [v for k, v in prof.items() if k in ("conv_fprop_tc", "conv_fprop_stem_tc", "conv_dgrad", "conv_wgrad")]
        main = prof.get("conv_fprop_tc", {"ms": 0.0, "work": 0.0, "launches": 0})
        dg = prof.get("conv_dgrad", {"ms": 0.0, "work": 0.0, "launches": 0})
        tc_ms = main["ms"] + dg["ms"]
        tc_work = main["work"] + dg["work"]
        achieved = tc_work / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        all_conv_ms = sum(v["ms"] for v in tc) + prof.get("conv_fprop_direct", {"ms": 0})["ms"]
        all_conv_work = sum(v["work"] for v in prof.values())
        cpu_t = time_cpu(64, iters=5, warmup=1)
        cores = os.cpu_count() or 1
        line = {
            "metric": "train_patches_per_s_128cubed", "value": value, "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": patches, "parallelism": "dp%d" % world,
                       "launch": "eager" if args.no_graph else "one CUDA graph per step",
                       "l2": "no flush needed: each step streams >10 GB of activations, far larger than the 126 MB L2"},
            "voxels_per_s": value * vox,
            "e2e": {"value": patches / (ms_e2e / args.steps * 1e-3), "unit": "patches/s",
                    "h2d_bytes_per_step": x_host.numel() * 4 + lab_host.numel(), "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "tcgen05_launches": umma_launches,
            # dominant kernel family: the tcgen05 implicit-GEMM convolutions (conv_umma_roll / conv_umma_plane), all 34
            # forward + data-gradient launches of the step; achieved = their algorithmic FLOPs / their CUDA-event time.
            # traffic = DRAM read + write bytes of the same launches in the committed ncu launch list
            # (profiles/r01_step4_final_launch_list.md), per step like `achieved`.
            "roofline": {"bound": "tensor", "kernel": "conv_umma_roll_kernel + conv_umma_plane_kernel (the fprop + dgrad "
                                                      "launches of the step)",
                         "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops_sustained"], "peak_source": peaks["source"] + " sustained",
                         "traffic": NCU_CONV_TRAFFIC_BYTES_PER_STEP, "algorithmic_bytes": CONV_ALGORITHMIC_BYTES_PER_STEP,
                         "launches_per_step": (main["launches"] + dg["launches"]) // 2,
                         "kernel_ms_per_step": tc_ms / 2,
                         "all_conv_ms_per_step": all_conv_ms / 2,
                         "all_conv_tflops": all_conv_work / (all_conv_ms * 1e-3) / 1e12 if all_conv_ms else 0.0,
                         "per_kernel": {k: {"ms_per_step": v["ms"] / 2, "launches_per_step": v["launches"] // 2,
                                            "tflops": v["work"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else 0.0}
                                        for k, v in prof.items()}},
            "model_tflops": patches * TRAIN_GFLOP_PER_PATCH / ms_step / 1e3,
            "cpu_baseline": {"value": (64 / PATCH) ** 3 / cpu_t, "unit": "patches/s", "cores": cores, "kind": "port",
                             "sample": "oracle port (reference modules' torch CPU arithmetic, fp32, %d threads): median "
                                       "of 5 fwd+bwd+Adam steps on one 1x1x64^3 patch = 1/8 of a 128^3 patch" % cores},
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying "
                                                            "the captured CUDA graph of the step")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__), "--gpus",
               str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()
