#!/usr/bin/env python
"""validate.py of the reference (validate.py:23-125) on the B200-native model: PSNR / SSIM / VIF of the bicubic baseline
and of ``MewZoom.upscale`` over paired LR / HR images, the control vector taken from flags.

    python tools/validate.py --checkpoint_path ckpt.pt --lr_images_path dataset/validate/lr --hr_images_path dataset/validate/hr \
        --gaussian_blur 0.1 --gaussian_noise 0.1 --jpeg_compression 0.1 [--device cuda:0] [--operands auto]

The checkpoint is a reference training checkpoint (``model_args`` / ``model`` of 0.2.x as validate.py:53-57 reads it, or
``upscaler_args`` / ``upscaler`` as pretrain.py:335-340 writes it; weight-norm parametrizations and ``_orig_mod.``
prefixes are baked / stripped: MewZoom.from_checkpoint) or a ``save_pretrained`` directory.  ``--synthetic N`` runs N
random pairs through a random-init model instead (no dataset / checkpoint needed: an end-to-end check of the tool).
There is no CPU path: --device must be a B200.
"""
from __future__ import annotations

import argparse
import os
import sys
from os import path, walk

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ultrazoom_b200 import MODEL_CONFIGS, ControlVector, MewZoom  # noqa: E402
from ultrazoom_b200.metrics import PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure, VisualInformationFidelity  # noqa: E402

ALLOWED_EXTENSIONS = frozenset({".png", ".jpg", ".jpeg", ".webp", ".gif"})      # reference data.py:176


def image_pairs(lr_root: str, hr_root: str):
    """ImagePairs (reference data.py:170-221): same walk order in both folders, decode_image(mode=RGB), ToDtype(float32, scale=True)."""
    from torchvision.io import decode_image

    def listing(root):
        return [path.join(folder, f) for folder, _, files in walk(root) for f in files if path.splitext(f)[1] in ALLOWED_EXTENSIONS]

    lr, hr = listing(lr_root), listing(hr_root)
    assert len(lr) == len(hr), f"{len(lr)} LR images but {len(hr)} HR images"
    for a, b in zip(lr, hr):
        yield decode_image(a, mode="RGB").unsqueeze(0).float() / 255.0, decode_image(b, mode="RGB").unsqueeze(0).float() / 255.0


def load_model(args) -> MewZoom:
    if args.synthetic:
        torch.manual_seed(0)
        return MewZoom(**MODEL_CONFIGS[args.model], operand_dtype=args.operands)
    if os.path.isdir(args.checkpoint_path):
        return MewZoom.from_pretrained(args.checkpoint_path)
    ckpt = torch.load(args.checkpoint_path, map_location="cpu", weights_only=True)
    if "model_args" in ckpt:                                    # 0.2.x schema (validate.py:53-57)
        ckpt = {"upscaler_args": ckpt["model_args"], "upscaler": ckpt["model"]}
    return MewZoom.from_checkpoint(ckpt, operand_dtype=args.operands)


def main(argv=None) -> dict:
    ap = argparse.ArgumentParser(description="Single-image super-resolution validation script (B200-native)")
    ap.add_argument("--checkpoint_path", default="./checkpoints/checkpoint.pt", type=str)
    ap.add_argument("--lr_images_path", default="./dataset/validate/lr", type=str)
    ap.add_argument("--hr_images_path", default="./dataset/validate/hr", type=str)
    ap.add_argument("--gaussian_blur", default=0.1, type=float)
    ap.add_argument("--gaussian_noise", default=0.1, type=float)
    ap.add_argument("--jpeg_compression", default=0.1, type=float)
    ap.add_argument("--device", default="cuda:0", type=str)
    ap.add_argument("--operands", default="auto", choices=["auto", "float16", "bfloat16"])
    ap.add_argument("--synthetic", default=0, type=int, help="run N random LR/HR pairs through a random-init model")
    ap.add_argument("--model", default="MewZoom-2X-Ctrl", choices=sorted(MODEL_CONFIGS))
    args = ap.parse_args(argv)
    if "cuda" not in args.device or not torch.cuda.is_available():
        raise RuntimeError("Cuda is not available." if "cuda" in args.device else
                           "ultrazoom_b200 runs on a B200 only: --device must be a CUDA device (no CPU path)")
    dev = torch.device(args.device)
    model = load_model(args).to(dev).eval()
    print("Model checkpoint loaded successfully")
    c_hat = None
    if model.control_features:
        c_hat = ControlVector(gaussian_blur=args.gaussian_blur, gaussian_noise=args.gaussian_noise,
                              jpeg_compression=args.jpeg_compression).to_tensor().to(dev).unsqueeze(0)
    if args.synthetic:
        g = torch.Generator().manual_seed(1)
        r = model.upscale_ratio

        def pairs():
            for _ in range(args.synthetic):
                y = torch.rand(1, 3, 64 * r, 96 * r, generator=g)
                yield torch.nn.functional.interpolate(y, scale_factor=1.0 / r, mode="bicubic", antialias=True).clamp(0, 1), y
        data = pairs()
    else:
        data = image_pairs(args.lr_images_path, args.hr_images_path)
    names = ("psnr", "ssim", "vif")
    mk = lambda: (PeakSignalNoiseRatio(data_range=1.0), StructuralSimilarityIndexMeasure(), VisualInformationFidelity())  # noqa: E731
    bicubic, enhanced = mk(), mk()
    n = 0
    for x, y in data:
        x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
        c = c_hat.repeat(x.size(0), 1) if c_hat is not None else None
        u_pred, u_bicubic = model.test_compare(x, c)            # reference validate.py:97
        for m in bicubic:
            m.update(u_bicubic, y)
        for m in enhanced:
            m.update(u_pred, y)
        n += x.size(0)
    res = {"images": n, "bicubic": {k: m.compute() for k, m in zip(names, bicubic)},
           "enhanced": {k: m.compute() for k, m in zip(names, enhanced)}}
    print(f"Bicubic PSNR: {res['bicubic']['psnr']:.5f}, Bicubic SSIM: {res['bicubic']['ssim']:.5f}, Bicubic VIF: {res['bicubic']['vif']:.5f}")
    print(f"Enhanced PSNR: {res['enhanced']['psnr']:.5f}, Enhanced SSIM: {res['enhanced']['ssim']:.5f}, Enhanced VIF: {res['enhanced']['vif']:.5f}")
    return res


if __name__ == "__main__":
    main()
