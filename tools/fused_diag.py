"""Bring-up diagnostics of the fused encoder block: error maps of one small case against the fp32 reference."""
import os
import sys

import torch
from torch.nn import functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import ops  # noqa: E402


def run(shape, seg_rows=0, max_ctas=0, with_film=True, dt=torch.float16):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(1)
    B, H, W = shape
    zf0 = torch.randn(B, H, W, 48, generator=g)
    zb = zf0.to(dt)
    w1 = torch.randn(96, 48, 3, 3, generator=g) / (3.0 * 48 ** 0.5)
    w2 = torch.randn(48, 96, 3, 3, generator=g) / (3.0 * 96 ** 0.5)
    film = None
    if with_film:
        film = torch.empty(B, 2, 96)
        film[:, 0] = 1 + 0.3 * torch.randn(B, 96, generator=g)
        film[:, 1] = 0.3 * torch.randn(B, 96, generator=g)
    acc1 = F.conv2d(zb.float().permute(0, 3, 1, 2), w1.to(dt).float(), padding=1)
    if film is not None:
        acc1 = acc1 * film[:, 0][:, :, None, None] + film[:, 1][:, :, None, None]
    hid = F.silu(acc1).to(dt).float()
    ref = zf0 + F.conv2d(hid, w2.to(dt).float(), padding=1).permute(0, 2, 3, 1)
    w1p, w2p = ops.pack_conv_weight(w1, dev, dtype=dt), ops.pack_conv_weight(w2, dev, dtype=dt)
    zf = zf0.to(dev).contiguous()
    out = ops.block_fused(zb.to(dev), w1p, w2p, film.to(dev) if film is not None else None, zf, seg_rows=seg_rows, max_ctas=max_ctas)
    torch.cuda.synchronize()
    err = (zf.cpu() - ref).abs()
    print(f"shape {shape} seg_rows {seg_rows} max_ctas {max_ctas} film {with_film}: max err {err.max().item():.4g} "
          f"(delta applied rms {(ref - zf0).pow(2).mean().sqrt().item():.3g}); shadow ok {bool(torch.equal(out, zf.to(dt)))}")
    if err.max().item() > 3e-3:
        e = err[0]                                                   # (H, W, 48)
        rows = e.amax(dim=(1, 2))
        print("  per row   :", " ".join(f"{v:.2g}" for v in rows[:40].tolist()))
        cols = e.amax(dim=(0, 2))
        blocks = [cols[i:i + 16].max().item() for i in range(0, W, 16)]
        print("  per 16 px :", " ".join(f"{v:.2g}" for v in blocks[:40]))
        ch = e.amax(dim=(0, 1))
        print("  per ch    :", " ".join(f"{v:.2g}" for v in ch.tolist()))
        untouched = (zf.cpu() == zf0).float()[0].mean(dim=2)
        print("  fraction of pixels whose zf is unchanged:", untouched.mean().item())
    return err.max().item()


if __name__ == "__main__":
    bad = 0
    for args in [((1, 4, 100),), ((1, 9, 300),), ((1, 9, 300), 0, 0, False), ((1, 23, 260), 3, 2), ((2, 17, 400), 5, 4), ((1, 40, 960),)]:
        try:
            bad += run(*args) > 3e-3
        except Exception as ex:  # noqa: BLE001
            print("FAILED", args, type(ex).__name__, str(ex)[:300])
            bad += 1
            break
    print("fused_diag:", "ok" if not bad else f"{bad} bad")
