"""One launch each of the non-encoder kernels at cfg2's shape, for an ncu capture:
    ncu --set full --clock-control none -k regex:"conv_tc_kernel|stem_kernel|bicubic_kernel" -o small python tools/small_once.py
order of launches: head (skip recomputed), head (no skip), head (skip from buffer), bicubic, stem."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import ops  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    C, r, H, W, B = (a + [48, 2, 540, 960, 16][len(a):])[:5]
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    Cp = ops.padded_channels(C)
    x = torch.rand(B, 3, H, W, generator=g).to(dev)
    zb = torch.randn(B, H, W, Cp, generator=g).to(torch.float16).to(dev)
    wh = ops.pack_conv_weight(torch.randn(3 * r * r, C, 3, 3, generator=g) * 0.02, dev)
    ws, bs = torch.randn(C, 3, 1, 1, generator=g), torch.randn(C, generator=g)
    y = torch.empty(B, 3, H * r, W * r, device=dev)
    ops.head_shuffle_add(zb, wh, r, x=x, y=y, skip_mode=2, clamp01=True)
    ops.head_shuffle_add(zb, wh, r, x=None, y=y, skip_mode=0, clamp01=True)
    ops.head_shuffle_add(zb, wh, r, x=None, y=y, skip_mode=1, clamp01=True)
    ops.bicubic(x, r)
    ops.stem_pack(x, ws, bs)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
