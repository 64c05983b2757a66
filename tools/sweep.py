"""Time single convolution launches (CUDA events) over a grid of mz_conv_tune settings.

    python tools/sweep.py [C] [H] [W] [B]        (defaults: 96 540 960 1  = MewZoom-4X-Ctrl on a 960x540 frame)
"""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import _native, ops  # noqa: E402


def time_conv(inp, wp, mode, film, zf, tune, reps=10):
    if tune.dbg & 16:
        reps = 1
    for _ in range(2):
        ops.conv3x3(inp, wp, mode, film, zf, tune=tune)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv3x3(inp, wp, mode, film, zf, tune=tune)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 540
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 960
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    Cp, hCp = ops.padded_channels(C), ops.padded_channels(2 * C)
    Cz = int(os.environ.get("ZP", Cp))        # channel pitch of zb (the model keeps 48-channel zb at pitch 64)
    flops = 2.0 * 9 * C * 2 * C * B * H * W
    dt = torch.bfloat16 if os.environ.get("DT") == "bf16" else torch.float16
    zb = torch.zeros(B, H, W, Cz, dtype=dt)
    zb[..., :Cp] = torch.randn(B, H, W, Cp, generator=g).to(dt)
    zb = zb.to(dev)
    hid = torch.randn(B, H, W, hCp, generator=g).to(dt).to(dev)
    zf = torch.zeros(B, H, W, Cp, device=dev)
    w1 = ops.pack_conv_weight(torch.randn(2 * C, C, 3, 3, generator=g) * 0.02, dev, cin_p=Cz, dtype=dt)
    w2 = ops.pack_conv_weight(torch.randn(C, 2 * C, 3, 3, generator=g) * 0.02, dev, dtype=dt)
    film = torch.ones(B, 2, hCp, device=dev)
    print(f"C={C} {W}x{H} B={B}: {flops / 1e9:.1f} GFLOP per conv; 100% of 1644 TF = {flops / 1644e12 * 1e6:.1f} us")
    base = dict(pair=int(os.environ.get("PAIR", "0")))
    if os.environ.get("SWEEP_CFGS"):        # e.g. SWEEP_CFGS="[dict(resident=1), dict(resident=2, rows=2)]"
        cfgs = [dict(base, **c) for c in eval(os.environ["SWEEP_CFGS"], {"dict": dict})]
    elif os.environ.get("SWEEP") == "dbg":
        cfgs = [dict(base, dbg=d) for d in (0, 16, 16, 4, 8, 12, 3, 7)]
    else:
        cfgs = [dict(base, resident=r, rows=rows) for r in (0, 2) for rows in (0, 1, 2, 4)]
    only = os.environ.get("ONLY")
    for which, (inp, wp, mode, fl, z) in (("conv1", (zb, w1, 0, film, None)), ("conv2", (hid, w2, 1, None, zf))):
        if only and only != which:
            continue
        for kw in cfgs:
            cin = inp.shape[-1]
            if kw.get("kc") and cin % kw["kc"]:
                continue
            try:
                us = time_conv(inp, wp, mode, fl, z, _native.tune(**kw))
                print(f"{which} {kw}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{which} {kw}: {str(e)[:100]}", flush=True)


if __name__ == "__main__":
    main()
