"""CPU restatement (plain torch functional, fp32) of the reference's 0.3.0 U-Net operators -- TEST INFRASTRUCTURE
only: tests/ and bench.py's CPU leg may import this; the product package (ultrazoom_b200) never does.

Pinned: every function here reproduces tests/golden/unet_ops.npz, which oracle/make_golden.py computed with the
reference's own leaf classes (tests/test_oracle_unet.py).  All tensors NCHW, as in the reference."""
from __future__ import annotations

from math import log2
from typing import Dict, Tuple

import torch
from torch import Tensor
from torch.nn import functional as F


def adaptive_residual_mix(x: Tensor, z: Tensor, weight: Tensor, alpha: Tensor) -> Tensor:
    """AdaptiveResidualMix.forward, reference src/ultrazoom/model.py:826-839."""
    beta = torch.sigmoid(F.conv2d(torch.cat([x, z], dim=1), weight))          # :827-830
    w = torch.sigmoid(alpha) * beta                                          # :833-835
    return (1 - w) * x + w * z                                               # :837


def pixel_crush(x: Tensor, weight: Tensor, f: int) -> Tensor:
    """PixelCrush.forward, model.py:859-865,881-882: kernel_size = stride = crush_factor, bias=False."""
    return F.conv2d(x, weight, stride=f)


def inverted_bottleneck(x: Tensor, w1: Tensor, w2: Tensor) -> Tensor:
    """InvertedBottleneck.forward, model.py:773-778."""
    return F.conv2d(F.silu(F.conv2d(x, w1, padding=1)), w2, padding=1)


def subpixel_conv2d(x: Tensor, weight: Tensor, r: int) -> Tensor:
    """SubpixelConv2d.forward, model.py:926-930: 3x3 conv (bias=False) then PixelShuffle(r)."""
    return F.pixel_shuffle(F.conv2d(x, weight, padding=1), r)


def crop_feature_maps(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """Decoder.crop_feature_maps, model.py:650-689: centre crop or zero pad, per axis."""
    _, _, h, w = x.shape
    th, tw = size
    if h > th:
        s = (h - th) // 2
        x = x[:, :, s:s + th, :]
    elif h < th:
        top = (th - h) // 2
        x = F.pad(x, (0, 0, top, th - h - top))
    if w > tw:
        s = (w - tw) // 2
        x = x[:, :, :, s:s + tw]
    elif w < tw:
        left = (tw - w) // 2
        x = F.pad(x, (left, tw - w - left, 0, 0))
    return x


def encoder_block(x: Tensor, sd: Dict[str, Tensor], prefix: str = "") -> Tensor:
    """EncoderBlock.forward (= DecoderBlock), model.py:507-511: skip(x, convnet(x)) with the gated mix."""
    z = inverted_bottleneck(x, sd[prefix + "convnet.conv1.weight"], sd[prefix + "convnet.conv2.weight"])
    return adaptive_residual_mix(x, z, sd[prefix + "skip.conv.weight"], sd[prefix + "skip.alpha"])


def sr2x_block(x: Tensor, sd: Dict[str, Tensor], prefix: str = "") -> Tensor:
    """SR2XBlock.forward, model.py:997-1001."""
    return subpixel_conv2d(encoder_block(x, sd, prefix + "refiner."), sd[prefix + "upscale.conv.weight"], 2)


def super_resolver(x: Tensor, sd: Dict[str, Tensor], upscale_ratio: int) -> Tensor:
    """SuperResolver.forward, model.py:968-972: log2(ratio) SR2XBlocks."""
    for i in range(int(log2(upscale_ratio))):
        x = sr2x_block(x, sd, f"layers.{i}.")
    return x


def quality_assessor(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """QualityAssessor.forward, model.py:1024-1032: conv3x3 + bias -> AdaptiveAvgPool2d(1) -> flatten."""
    return F.conv2d(x, weight, bias, padding=1).mean(dim=(2, 3))
