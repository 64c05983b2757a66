"""ControlVector -- drop-in for ``ultrazoom.control.ControlVector`` (0.2.x; the file is absent from the
reference snapshot -- API per reference README.md:94,118-122, feature order per README.md:181-185 and
data.py:162)."""
from __future__ import annotations

import torch
from torch import Tensor


class ControlVector:
    """Assumed degradation levels in [0, 1]: ``[gaussian_blur, gaussian_noise, jpeg_compression]``."""

    NUM_FEATURES = 3

    def __init__(self, gaussian_blur: float = 0.0, gaussian_noise: float = 0.0, jpeg_compression: float = 0.0):
        for name, v in (
            ("gaussian_blur", gaussian_blur),
            ("gaussian_noise", gaussian_noise),
            ("jpeg_compression", jpeg_compression),
        ):
            assert 0.0 <= v <= 1.0, f"{name} must be between 0 and 1, {v} given."
        self.gaussian_blur = float(gaussian_blur)
        self.gaussian_noise = float(gaussian_noise)
        self.jpeg_compression = float(jpeg_compression)

    def to_tensor(self) -> Tensor:
        return torch.tensor([self.gaussian_blur, self.gaussian_noise, self.jpeg_compression], dtype=torch.float32)

    def __repr__(self) -> str:
        return (f"ControlVector(gaussian_blur={self.gaussian_blur}, gaussian_noise={self.gaussian_noise}, "
                f"jpeg_compression={self.jpeg_compression})")
