"""One launch of each U-Net operator kernel at the bench shapes (for ncu captures): python tools/unet_one.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultrazoom_b200 import unet as N  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(5)
C = 96
x, z = torch.randn(1, 540, 960, C, generator=g).to(dev), torch.randn(1, 540, 960, C, generator=g).to(dev)
w = (torch.randn(C, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5).to(dev)
xi = torch.randn(1, 1080, 1920, 48, generator=g).to(dev)
wc = (torch.randn(96, 48, 2, 2, generator=g) / (4 * 48) ** 0.5).to(dev)
xs = torch.randn(1, 540, 960, 192, generator=g).to(dev)
for _ in range(reps):
    N.adaptive_residual_mix(x, z, w, torch.tensor(0.25))
    N.pixel_crush(xi, wc, 2)
    N.pixel_shuffle_nhwc(xs, 2)
    N.crop_feature_maps(x, (536, 956))
torch.cuda.synchronize()
print("ok")
