"""Latency of MewZoom.upscale for one model / shape: device time per call (CUDA events) and host enqueue time per call.
    python tools/time_model.py MewZoom-2X 1 256 256"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import MODEL_CONFIGS, MewZoom  # noqa: E402

name, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = MODEL_CONFIGS[name]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = MewZoom(**cfg).to(dev).eval()
x = torch.rand(B, 3, H, W, device=dev)
c = torch.tensor([[0.5, 0.2, 0.3]], device=dev) if cfg["control_features"] else None
for _ in range(5):
    m.upscale(x, c)
torch.cuda.synchronize()
n = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    m.upscale(x, c)
e1.record()
t_host = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
print(f"{name} {B}x{H}x{W}: device {e0.elapsed_time(e1) / n * 1e3:.1f} us per call, host enqueue {t_host * 1e6:.1f} us per call")

# the same call replayed from a CUDA graph (MewZoom.capture): no host work per kernel
g = m.capture(x, c)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    g.replay()
e1.record()
t_host = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
same = torch.equal(g.replay(), m.upscale(x, c))
print(f"{name} {B}x{H}x{W}: graph replay {e0.elapsed_time(e1) / n * 1e3:.1f} us per call, host enqueue {t_host * 1e6:.1f} us per call, "
      f"bit-identical to the eager call: {same}")
