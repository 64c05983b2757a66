"""Run a few launches of conv1 and conv2 at a given shape (for ncu): python tools/one_conv.py [C H W B] [k=v ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import _native, ops  # noqa: E402

nums = [int(a) for a in sys.argv[1:] if "=" not in a]
kw = {a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[1:] if "=" in a}
C, H, W, B = (nums + [96, 540, 960, 1][len(nums):])[:4]
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
Cp, hCp = ops.padded_channels(C), ops.padded_channels(2 * C)
zb = torch.randn(B, H, W, Cp, generator=g).to(torch.float16).to(dev)
hid = torch.randn(B, H, W, hCp, generator=g).to(torch.float16).to(dev)
zf = torch.zeros(B, H, W, Cp, device=dev)
w1 = ops.pack_conv_weight(torch.randn(2 * C, C, 3, 3, generator=g) * 0.02, dev)
w2 = ops.pack_conv_weight(torch.randn(C, 2 * C, 3, 3, generator=g) * 0.02, dev)
film = torch.ones(B, 2, hCp, device=dev)
t = _native.tune(**kw) if kw else None
for _ in range(3):
    ops.conv3x3(zb, w1, 0, film, None, tune=t)
    ops.conv3x3(hid, w2, 1, None, zf, tune=t)
torch.cuda.synchronize()
print("ok")
