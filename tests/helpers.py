"""Shared test helpers (tests may import oracle/; the product package may not)."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import OracleMewZoom

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["flat_2x_c16_l3", "ctrl_3x_c16_l2", "ctrl_4x_c32_l2", "ctrl_2x_c48_l2"]


def load_case(name: str):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = {str(k): int(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    x = torch.from_numpy(z["x"])
    c = torch.from_numpy(z["c"]) if "c" in z.files else None
    out = {k: torch.from_numpy(z[k]) for k in ("forward", "upscale", "bicubic")}
    return cfg, sd, x, c, out


def oracle_from_case(cfg, sd) -> OracleMewZoom:
    m = OracleMewZoom(**cfg)
    m.load_state_dict(sd)
    return m.eval()
