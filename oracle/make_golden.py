"""Generate tests/golden/*.npz from the REFERENCE's own leaf classes.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference's ``MewZoom`` cannot be constructed (model.py:356-358 NameError)
and its 0.2.x control module is absent, so the flat network that
BASELINE.json names is assembled here from the reference classes that DO
import and run -- ``FanOutProjection`` (model.py:212-242), ``InvertedBottleneck``
(:731-778), ``ResidualConnection`` (:781-792), ``SubpixelConv2d`` (:885-930) and
``torch.nn.Upsample(mode="bicubic")`` (:71) -- so that every convolution,
activation, shuffle, skip add and clamp in the fixture was computed by the
reference's code.  Only the FiLM control (ours) is
restated locally.  The fixtures store inputs, weights (oracle state_dict key
names) and outputs; tests/test_oracle_golden.py replays them through
``oracle.OracleMewZoom`` and the GPU tests replay them through the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/src"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_SRC)

from ultrazoom.model import (  # type: ignore  # noqa: E402  (the reference)
    FanOutProjection,
    InvertedBottleneck,
    ResidualConnection,
    SubpixelConv2d,
)
from oracle.mewzoom_oracle import make_oracle  # noqa: E402


class RefFlat(nn.Module):
    """Flat 0.2.x-style network built from reference leaf modules."""

    def __init__(self, r, C, h, L, Fc):
        super().__init__()
        self.r, self.Fc = r, Fc
        self.bicubic = nn.Upsample(scale_factor=r, mode="bicubic")   # model.py:71
        self.stem = FanOutProjection(3, C)
        self.blocks = nn.ModuleList([InvertedBottleneck(C, h) for _ in range(L)])
        self.controls = nn.ModuleList([nn.Linear(Fc, 2 * h * C) for _ in range(L)]) if Fc else None
        self.skip = ResidualConnection()
        self.head = SubpixelConv2d(C, 3, r)

    def forward(self, x, c=None):
        s = self.bicubic.forward(x)
        z = self.stem.forward(x)
        for k, blk in enumerate(self.blocks):
            hdn = blk.conv1.forward(z)
            if self.Fc:
                g = self.controls[k](c)
                hc = hdn.shape[1]
                hdn = (1.0 + g[:, :hc])[:, :, None, None] * hdn + g[:, hc:][:, :, None, None]
            hdn = blk.silu.forward(hdn)
            d = blk.conv2.forward(hdn)
            z = self.skip.forward(z, d)
        z = self.head.forward(z)
        return self.skip.forward(s, z)


def load_from_oracle(ref: RefFlat, orc) -> None:
    sd = orc.state_dict()
    with torch.no_grad():
        ref.stem.conv.weight.copy_(sd["stem.conv.weight"])
        ref.stem.conv.bias.copy_(sd["stem.conv.bias"])
        for k, blk in enumerate(ref.blocks):
            blk.conv1.weight.copy_(sd[f"encoder.{k}.convnet.conv1.weight"])
            blk.conv2.weight.copy_(sd[f"encoder.{k}.convnet.conv2.weight"])
            if ref.Fc:
                ref.controls[k].weight.copy_(sd[f"encoder.{k}.control.linear.weight"])
                ref.controls[k].bias.copy_(sd[f"encoder.{k}.control.linear.bias"])
        ref.head.conv.weight.copy_(sd["head.conv.weight"])


CASES = {
    # name: (cfg, (B,H,W), seed, c)
    "flat_2x_c16_l3": (dict(upscale_ratio=2, num_channels=16, hidden_ratio=2, num_encoder_layers=3, control_features=0),
                       (1, 12, 10), 11, None),
    "ctrl_3x_c16_l2": (dict(upscale_ratio=3, num_channels=16, hidden_ratio=2, num_encoder_layers=2, control_features=3),
                       (2, 9, 11), 12, "per_image"),
    "ctrl_4x_c32_l2": (dict(upscale_ratio=4, num_channels=32, hidden_ratio=2, num_encoder_layers=2, control_features=3),
                       (1, 10, 17), 13, "readme"),
    "ctrl_2x_c48_l2": (dict(upscale_ratio=2, num_channels=48, hidden_ratio=2, num_encoder_layers=2, control_features=3),
                       (1, 9, 19), 14, "readme"),
}


def main() -> None:
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(1)
    for name, (cfg, (B, H, W), seed, cmode) in CASES.items():
        orc = make_oracle(cfg, seed=seed)
        ref = RefFlat(cfg["upscale_ratio"], cfg["num_channels"], cfg["hidden_ratio"],
                      cfg["num_encoder_layers"], cfg["control_features"]).eval()
        load_from_oracle(ref, orc)
        g = torch.Generator().manual_seed(1000 + seed)
        x = torch.rand(B, 3, H, W, generator=g)
        if cmode is None:
            c = None
        elif cmode == "readme":
            c = torch.tensor([[0.5, 0.2, 0.3]]).repeat(B, 1)      # README.md:118-122
        else:
            c = torch.rand(B, 3, generator=g)
        with torch.inference_mode():
            y = ref.forward(x, c)
            y_clamped = torch.clamp(y, 0, 1)                      # model.py:177
            s = ref.bicubic.forward(x)
        arrays = {"x": x.numpy(), "forward": y.numpy(), "upscale": y_clamped.numpy(), "bicubic": s.numpy(),
                  "cfg_keys": np.array(list(cfg.keys())), "cfg_vals": np.array(list(cfg.values()), dtype=np.int64)}
        if c is not None:
            arrays["c"] = c.numpy()
        for k, v in orc.state_dict().items():
            arrays["w:" + k] = v.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        res = (y - s).pow(2).mean().sqrt().item()
        print(f"{name}: out {tuple(y.shape)} residual rms {res:.4f}")

    # leaf known-answer vectors: bicubic (r=2,3,4) and pixel shuffle, straight from torch as the
    # reference calls it (Upsample model.py:71; PixelShuffle model.py:911).
    g = torch.Generator().manual_seed(77)
    x = torch.rand(1, 3, 5, 7, generator=g)
    leaf = {"x": x.numpy()}
    for r in (2, 3, 4):
        leaf[f"bicubic_r{r}"] = nn.Upsample(scale_factor=r, mode="bicubic")(x).numpy()
        z = torch.rand(1, 3 * r * r, 3, 4, generator=g)
        leaf[f"shuffle_in_r{r}"] = z.numpy()
        leaf[f"shuffle_out_r{r}"] = nn.PixelShuffle(r)(z).numpy()
    np.savez_compressed(os.path.join(out_dir, "leaf_ops.npz"), **leaf)
    print("leaf_ops written")


if __name__ == "__main__":
    main()
