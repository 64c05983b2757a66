"""Time the tf32 mix / crush kernels over a few channel counts (event-timed, inputs > L2): python tools/unet_shapes.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultrazoom_b200 import unet as N  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(5)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for C in (32, 48, 64, 80, 96, 128, 192):
    H, W = 540, 960 * 96 // C
    x, z = torch.randn(1, H, W, C, generator=g).to(dev), torch.randn(1, H, W, C, generator=g).to(dev)
    w = (torch.randn(C, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5).to(dev)
    a = torch.tensor(0.25)
    ms = timed(lambda: N.adaptive_residual_mix(x, z, w, a))
    print(f"mix   C {C:4d} {W}x{H}: {ms:.4f} ms  {3 * x.numel() * 4 / ms / 1e6:7.0f} GB/s")
    del x, z
for Cin, Cout, f in ((48, 96, 2), (64, 128, 2), (32, 64, 2), (96, 192, 2), (48, 96, 3), (48, 96, 4), (16, 32, 4)):
    H, W = 1080 // f * f, (1920 * 48 // Cin) // f * f
    x = torch.randn(1, H, W, Cin, generator=g).to(dev)
    wc = (torch.randn(Cout, Cin, f, f, generator=g) / (f * f * Cin) ** 0.5).to(dev)
    ms = timed(lambda: N.pixel_crush(x, wc, f))
    nbytes = (x.numel() + (H // f) * (W // f) * Cout) * 4
    print(f"crush {Cin:3d}->{Cout:3d} f{f} {W}x{H}: {ms:.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s")
    del x
