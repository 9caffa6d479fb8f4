"""Library bar: the same 3D U-Net graph in stock PyTorch/cuDNN on one B200 (not the product; a yardstick).

Times fwd+bwd+Adam for batch 2x1x128^3 in (a) fp32/TF32 NCDHW as the reference would run and (b) bf16 autocast +
channels_last_3d, the strongest library configuration. Prints one JSON line per configuration.
"""
import json
import sys
import time
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


def block(cin, f, name):
    return nn.Sequential(OrderedDict([
        (name + "conv1", nn.Conv3d(cin, f, 3, padding=1)), (name + "norm1", nn.BatchNorm3d(f)),
        (name + "relu1", nn.ReLU(inplace=True)),
        (name + "conv2", nn.Conv3d(f, f, 3, padding=1)), (name + "norm2", nn.BatchNorm3d(f)),
        (name + "relu2", nn.ReLU(inplace=True))]))


class Net(nn.Module):
    def __init__(self, cin=1, cout=2, f=32):
        super().__init__()
        self.e1, self.e2, self.e3, self.e4 = block(cin, f, "a"), block(f, 2 * f, "b"), block(2 * f, 4 * f, "c"), block(4 * f, 8 * f, "d")
        self.bott = block(8 * f, 16 * f, "e")
        self.u4, self.u3 = nn.ConvTranspose3d(16 * f, 8 * f, 2, 2), nn.ConvTranspose3d(8 * f, 4 * f, 2, 2)
        self.u2, self.u1 = nn.ConvTranspose3d(4 * f, 2 * f, 2, 2), nn.ConvTranspose3d(2 * f, f, 2, 2)
        self.d4, self.d3, self.d2, self.d1 = block(16 * f, 8 * f, "f"), block(8 * f, 4 * f, "g"), block(4 * f, 2 * f, "h"), block(2 * f, f, "i")
        self.head = nn.Conv3d(f, cout, 1)

    def forward(self, x):
        p = lambda t: F.max_pool3d(t, 2, 2)
        e1 = self.e1(x); e2 = self.e2(p(e1)); e3 = self.e3(p(e2)); e4 = self.e4(p(e3))
        b = self.bott(p(e4))
        d4 = self.d4(torch.cat((self.u4(b), e4), 1))
        d3 = self.d3(torch.cat((self.u3(d4), e3), 1))
        d2 = self.d2(torch.cat((self.u2(d3), e2), 1))
        d1 = self.d1(torch.cat((self.u1(d2), e1), 1))
        return self.head(d1)


def dice_ce(logits, lab):
    ce = F.cross_entropy(logits.float(), lab)
    p = torch.softmax(logits.float(), 1)
    loss = 0.0
    for i in range(2):
        t = (lab == i).float()
        loss = loss + 1 - (2 * (p[:, i] * t).sum() + 1e-5) / ((p[:, i] ** 2).sum() + (t * t).sum() + 1e-5)
    return ce + loss / 2


def run(mode, size=128, iters=5):
    torch.manual_seed(0)
    dev = "cuda"
    net = Net().to(dev)
    if mode == "bf16_cl":
        net = net.to(memory_format=torch.channels_last_3d)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    x = torch.randn(2, 1, size, size, size, device=dev)
    lab = (torch.rand(2, size, size, size, device=dev) > 0.9).long()
    def step():
        opt.zero_grad(set_to_none=True)
        if mode == "bf16_cl":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(x.contiguous(memory_format=torch.channels_last_3d))
        else:
            out = net(x)
        loss = dice_ce(out, lab)
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"probe": "torch_unet_bar", "mode": mode, "size": size, "ms_per_step": ms,
                      "patches_per_s": 2 / (ms * 1e-3), "conv_TFLOPs": 2 * 2850e9 * (size / 128) ** 3 / (ms * 1e-3) / 1e12,
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


if __name__ == "__main__":
    torch.backends.cudnn.benchmark = True
    for mode in ("fp32_tf32", "bf16_cl"):
        try:
            run(mode)
        except Exception as e:  # noqa
            print(json.dumps({"probe": "torch_unet_bar", "mode": mode, "error": repr(e)[:300]}), flush=True)
