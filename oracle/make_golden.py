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


def make_unet_ops() -> None:
    """Known-answer vectors of the 0.3.0 U-Net operators, computed by the reference's OWN leaf
    classes (they import and run although ``MewZoom`` itself cannot be constructed): inputs, the leaf's state_dict and
    its output, per operator.  tests/test_gpu_unet_ops.py replays them through ultrazoom_b200.unet."""
    from ultrazoom.model import (  # type: ignore  (the reference)
        AdaptiveResidualMix, Decoder, EncoderBlock, PixelCrush, QualityAssessor, SR2XBlock, SuperResolver)

    torch.manual_seed(2024)
    g = torch.Generator().manual_seed(99)
    out = {}

    def put(prefix, module, inputs, result):
        for i, t in enumerate(inputs):
            out[f"{prefix}/in{i}"] = t.numpy()
        for k, v in module.state_dict().items():
            out[f"{prefix}/w:{k}"] = v.detach().numpy()
        out[f"{prefix}/out"] = result.detach().numpy()

    with torch.inference_mode():
        C = 16
        mix = AdaptiveResidualMix(C).eval()                        # model.py:795-839
        mix.alpha.fill_(0.7)                                       # (0 is the init value: sigmoid(0) = 0.5)
        x, z = torch.randn(2, C, 9, 13, generator=g), torch.randn(2, C, 9, 13, generator=g)
        put("mix", mix, [x, z], mix.forward(x, z))
        for f, hw in ((2, (10, 14)), (3, (10, 14)), (4, (9, 17))):  # model.py:842-882 (odd sizes: the conv floors)
            crush = PixelCrush(C, 2 * C, f).eval()
            x = torch.randn(2, C, *hw, generator=g)
            put(f"crush{f}", crush, [x], crush.forward(x))
        up = SubpixelConv2d(2 * C, C, 2).eval()                    # mid-network upsampler, Decoder model.py:569-571
        x = torch.randn(1, 2 * C, 7, 9, generator=g)
        put("subpixel", up, [x], up.forward(x))
        for name, size in (("crop_smaller", (6, 8)), ("crop_larger", (11, 13)), ("crop_mixed", (12, 7))):
            x = torch.randn(1, C, 9, 10, generator=g)              # Decoder.crop_feature_maps model.py:650-689
            out[f"{name}/in0"] = x.numpy()
            out[f"{name}/size"] = np.array(size, dtype=np.int64)
            out[f"{name}/out"] = Decoder.crop_feature_maps(x, size).numpy()
        blk = EncoderBlock(C, 2).eval()                            # model.py:487-511 (InvertedBottleneck + gated mix)
        blk.skip.alpha.fill_(-0.3)
        x = torch.randn(1, C, 10, 18, generator=g)
        put("encoder_block", blk, [x], blk.forward(x))
        sr = SR2XBlock(C, 2, C).eval()                             # model.py:975-1001
        x = torch.randn(1, C, 6, 10, generator=g)
        put("sr2x", sr, [x], sr.forward(x))
        head = SuperResolver(C, 2, 4).eval()                       # model.py:933-972: 16 ch -> x2 -> x2 -> 3 ch
        x = torch.randn(1, C, 5, 7, generator=g)
        put("super_resolver", head, [x], head.forward(x))
        qa = QualityAssessor(2 * C, 3).eval()                      # model.py:1004-1032
        x = torch.randn(2, 2 * C, 8, 11, generator=g)
        put("quality", qa, [x], qa.forward(x))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "unet_ops.npz"), **out)
    print(f"unet_ops written ({len(out)} arrays)")


if __name__ == "__main__":
    if "--unet-only" not in sys.argv:
        main()
    make_unet_ops()
