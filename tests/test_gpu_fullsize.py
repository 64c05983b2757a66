"""GPU parity at the FULL sizes of BASELINE.json's configs ("no BASELINE config is checked against the
oracle at its real size"): the whole frame runs on the GPU; crops of it are compared with the CPU oracle run on the
same crop plus the receptive-field halo R = 2L+1 (sharding.halo_radius; exact -- tests/test_sharding.py), so that
the oracle finishes in seconds while every kernel configuration the full size selects is the one under test:

    cfg2   MewZoom-2X-Ctrl  960x540  -> 1920x1080  (two frames of the batch of 16; 7.5 column tiles: ragged last tile)
    cfg3   MewZoom-3X-Ctrl  1280x720 -> 3840x2160  (54 -> 64 / 108 -> 128 padded channels)
    cfg4a  MewZoom-4X-Ctrl  960x540  -> 3840x2160  (CTA-pair kernels, ragged last tile)
    cfg4b  MewZoom-4X-Ctrl  1920x1080-> 7680x4320  (= cfg5 un-tiled)

Crops (LR core 40 x 40): the four corners -- the right ones lie inside the half-empty last 128-pixel tile of a 960-pixel
row (x >= 896), the bottom ones in the last patch rows -- one interior crop straddling a 128-pixel tile boundary and an
even/odd patch-row boundary, and one at the right edge in the middle of the frame.

Stated tolerances (un-clamped forward AND clamped upscale, per-image random control vector):
    fp16 operands (default):  2X  max|err| <= 4e-3, PSNR >= 64 dB;  3X  6e-3 / 62 dB;  4X  8e-3 / 60 dB
    bf16 operands:            2X, 3X  max|err| <= 2e-2, PSNR >= 50 dB;  4X  <= 3e-2, PSNR >= 48 dB
BASELINE.json's envelope is max|err| <= 2e-2 and PSNR >= 45 dB: fp16 meets it on every config with a 2.5x margin; bf16
-- the type the north_star names -- does NOT meet the max-abs half of it on the 40-layer model (0.020-0.026 measured,
the extreme-value tail grows with the pixel count), which is why fp16 is the default operand type.
"""
import json
import os

import pytest
import torch

from oracle import make_oracle, max_abs_err, psnr
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu

FULL = {
    "cfg2": ("MewZoom-2X-Ctrl", 2, 540, 960),
    "cfg3": ("MewZoom-3X-Ctrl", 1, 720, 1280),
    "cfg4a": ("MewZoom-4X-Ctrl", 1, 540, 960),
    "cfg4b": ("MewZoom-4X-Ctrl", 1, 1080, 1920),
}
TOL = {  # (max-abs, PSNR) per upscale ratio and operand type
    "float16": {2: (4e-3, 64.0), 3: (6e-3, 62.0), 4: (8e-3, 60.0)},
    "bfloat16": {2: (2e-2, 50.0), 3: (2e-2, 50.0), 4: (3e-2, 48.0)},
}
CORE = 40


def crops_of(H: int, W: int):
    """(name, y0, x0) of the LR cores; every frame here has a tile boundary at x = 128 k and patch rows of 2."""
    xb = (W // 2) // 128 * 128 - CORE // 2          # straddles a 128-pixel tile boundary
    yb = (H // 2) | 1                               # starts on an odd row: straddles patch-row boundaries
    return [("top-left", 0, 0), ("top-right", 0, W - CORE), ("bottom-left", H - CORE, 0),
            ("bottom-right", H - CORE, W - CORE), ("interior", yb, xb), ("right-edge", H // 3, W - CORE)]


_oracle_cache = {}


def oracle_crops(workload: str):
    """Oracle forward (un-clamped) on crop + halo, cropped back to the core: {crop name: (B,3,r*CORE,r*CORE)}."""
    if workload in _oracle_cache:
        return _oracle_cache[workload]
    from ultrazoom_b200.sharding import halo_radius

    name, B, H, W = FULL[workload]
    o = make_oracle(name, seed=0)
    r, R = o.upscale_ratio, halo_radius(o.num_encoder_layers)
    g = torch.Generator().manual_seed(4321)
    x, c = torch.rand(B, 3, H, W, generator=g), torch.rand(B, 3, generator=g)
    refs = {}
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.inference_mode():
        for cname, y0, x0 in crops_of(H, W):
            hy0, hy1, hx0, hx1 = max(0, y0 - R), min(H, y0 + CORE + R), max(0, x0 - R), min(W, x0 + CORE + R)
            full = o.forward(x[:, :, hy0:hy1, hx0:hx1].contiguous(), c)
            oy, ox = (y0 - hy0) * r, (x0 - hx0) * r
            refs[cname] = full[:, :, oy:oy + CORE * r, ox:ox + CORE * r].clone()
    _oracle_cache[workload] = (o, x, c, refs)
    return _oracle_cache[workload]


@pytest.mark.parametrize("operands", ["float16", "bfloat16"])
@pytest.mark.parametrize("workload", list(FULL))
def test_full_size_frames_against_the_oracle_on_crops(workload, operands):
    from ultrazoom_b200 import MODEL_CONFIGS, MewZoom

    dev = torch.device("cuda", 0)
    name, B, H, W = FULL[workload]
    o, x, c, refs = oracle_crops(workload)
    r = o.upscale_ratio
    m = MewZoom(**MODEL_CONFIGS[name], operand_dtype=operands)
    m.load_state_dict(o.state_dict())
    m = m.to(dev).eval()
    fwd = m.forward(x.to(dev), c.to(dev))
    ups = m.upscale(x.to(dev), c.to(dev))
    assert tuple(ups.shape) == (B, 3, H * r, W * r)
    assert float(ups.min()) >= 0.0 and float(ups.max()) <= 1.0          # reference tests/test_model.py:161-169
    assert torch.equal(ups, fwd.clamp(0, 1))                            # the fused clamp is exactly torch.clamp
    if operands == "float16":
        assert not m.saturated()
    tol_abs, tol_psnr = TOL[operands][r]
    report, worst, se, n = {}, 0.0, 0.0, 0
    for cname, y0, x0 in crops_of(H, W):
        ref = refs[cname]
        got = fwd[:, :, y0 * r:(y0 + CORE) * r, x0 * r:(x0 + CORE) * r].cpu()
        e_f = max_abs_err(got, ref)
        e_u = max_abs_err(got.clamp(0, 1), ref.clamp(0, 1))
        assert float(ref.std()) > 0.05, "degenerate crop"               # the residual branch is alive
        report[cname] = {"max_abs_forward": e_f, "max_abs_upscale": e_u, "psnr_upscale": psnr(got.clamp(0, 1), ref.clamp(0, 1))}
        worst = max(worst, e_f, e_u)
        se += float(((got.clamp(0, 1) - ref.clamp(0, 1)).double() ** 2).sum())
        n += ref.numel()
        assert e_f <= tol_abs and e_u <= tol_abs, (workload, operands, cname, e_f, e_u)
    total_psnr = 10.0 * torch.log10(torch.tensor(1.0 / max(se / n, 1e-20))).item()
    assert total_psnr >= tol_psnr, (workload, operands, total_psnr)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):                                          # measured values, for profiles/ and DESIGN.md
        path = os.path.join(out_dir, "fullsize_parity.json")
        try:
            with open(path) as f:
                acc = json.load(f)
        except (OSError, ValueError):
            acc = {}
        acc[f"{workload}/{operands}"] = {"model": name, "lr": [B, H, W], "worst_max_abs": worst, "psnr_over_crops": total_psnr,
                                         "tolerance": [tol_abs, tol_psnr], "crops": report}
        with open(path, "w") as f:
            json.dump(acc, f, indent=1)


def test_cfg5_tiling_is_bit_exact_at_full_size():
    """BASELINE configs[4] at its real size: the 1920x1080 -> 7680x4320 frame cut into the 8-GPU grid of halo-padded
    tiles (all of them on this one GPU, each written straight into the assembled frame by its head kernel) equals the
    un-tiled frame bit for bit."""
    from ultrazoom_b200 import MODEL_CONFIGS, MewZoom
    from ultrazoom_b200.sharding import best_grid, halo_radius, plan_tiles, run_tile_into

    dev = torch.device("cuda", 0)
    o, x, c, _ = oracle_crops("cfg4b")
    m = MewZoom(**MODEL_CONFIGS["MewZoom-4X-Ctrl"])
    m.load_state_dict(o.state_dict())
    m = m.to(dev).eval()
    xd, cd = x.to(dev), c.to(dev)
    full = m.upscale(xd, cd)
    R = halo_radius(40)
    rows, cols = best_grid(1080, 1920, 8, R, align_w=128)
    frame = torch.full_like(full, -1.0)
    for t in plan_tiles(1080, 1920, rows, cols, R, align_w=128):
        run_tile_into(m, xd, cd, t, 4, frame)
    assert torch.equal(frame, full)
