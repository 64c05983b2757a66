#!/usr/bin/env python
"""export_model.ipynb of the reference (cells 3-7) for the B200-native model: training checkpoint -> baked plain weights
-> ``save_pretrained`` directory (config.json + model.safetensors, the HuggingFace layout ``MewZoom.from_pretrained``
reads; reference README.md:70,109) -> reload and verify.

    python tools/export_model.py --checkpoint_path checkpoints/checkpoint.pt --out exports/mewzoom-2x-ctrl

The notebook's ONNX half (cells 9-11: torch.onnx.export of ONNXModel + an onnxruntime comparison at rtol 1e-2 / atol 1e-3)
is out of scope here: an ONNX graph is an artefact of the PyTorch modules, and this model's forward is the sm_100a
library behind include/mewzoom_b200.h.  ``ONNXModel`` itself (the wrapper whose forward is ``upscale``) is kept."""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ultrazoom_b200 import MewZoom  # noqa: E402


def main(argv=None) -> str:
    ap = argparse.ArgumentParser(description="checkpoint -> save_pretrained directory")
    ap.add_argument("--checkpoint_path", default="./checkpoints/checkpoint.pt", type=str)
    ap.add_argument("--out", default="./exports/model", type=str)
    args = ap.parse_args(argv)
    ckpt = torch.load(args.checkpoint_path, map_location="cpu", weights_only=True)
    if "model_args" in ckpt:                                    # 0.2.x schema (export_model.ipynb cell 3)
        ckpt = {"upscaler_args": ckpt["model_args"], "upscaler": ckpt["model"]}
    model = MewZoom.from_checkpoint(ckpt).eval()                # add_weight_norms / load / remove_parameterizations, baked
    model.save_pretrained(args.out)                             # cell 5
    again = MewZoom.from_pretrained(args.out)
    a, b = model.state_dict(), again.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a), "reloaded weights differ"
    print(f"exported {model.num_params:,} parameters to {args.out}")
    return args.out


if __name__ == "__main__":
    main()
