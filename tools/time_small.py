"""Time the non-encoder kernels (stem, FiLM table, head = conv + pixel shuffle + bicubic skip, plain bicubic) at a
workload's shape:  python tools/time_small.py [C r H W B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import _native, ops  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    a = [int(v) for v in sys.argv[1:] if "=" not in v]
    kw = {v.split("=")[0]: int(v.split("=")[1]) for v in sys.argv[1:] if "=" in v}
    C, r, H, W, B = (a + [96, 4, 540, 960, 1][len(a):])[:5]
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    Cp = ops.padded_channels(C)
    x = torch.rand(B, 3, H, W, generator=g).to(dev)
    zb = torch.randn(B, H, W, Cp, generator=g).to(torch.float16).to(dev)
    wh = ops.pack_conv_weight(torch.randn(3 * r * r, C, 3, 3, generator=g) * 0.02, dev)
    ws, bs = torch.randn(C, 3, 1, 1, generator=g), torch.randn(C, generator=g)
    y = torch.empty(B, 3, H * r, W * r, device=dev)
    t = _native.tune(**kw) if kw else None
    out_mb = y.numel() * 4 / 1e6
    us = timeit(lambda: ops.head_shuffle_add(zb, wh, r, x=x, y=y, skip_mode=2, clamp01=True, tune=t))
    print(f"head (skip recomputed)  {us:8.1f} us   {out_mb / us * 1e3:7.1f} GB/s of HR writes")
    us = timeit(lambda: ops.head_shuffle_add(zb, wh, r, x=None, y=y, skip_mode=0, clamp01=True, tune=t))
    print(f"head (no skip)          {us:8.1f} us")
    us = timeit(lambda: ops.head_shuffle_add(zb, wh, r, x=None, y=y, skip_mode=1, clamp01=True, tune=t))
    print(f"head (skip from buffer) {us:8.1f} us")
    us = timeit(lambda: ops.bicubic(x, r))
    print(f"bicubic                 {us:8.1f} us   {(out_mb + x.numel() * 4 / 1e6) / us * 1e3:7.1f} GB/s")
    us = timeit(lambda: ops.stem_pack(x, ws, bs))
    print(f"stem                    {us:8.1f} us   {(x.numel() * 4 + B * H * W * Cp * 6) / 1e6 / us * 1e3:7.1f} GB/s")


if __name__ == "__main__":
    main()
