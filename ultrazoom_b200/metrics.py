"""Image-quality metrics for the evaluation CLI (tools/validate.py) -- the three the reference's validate.py reports
(reference validate.py:84-125 via torchmetrics, which is not installed here): PSNR(data_range=1), SSIM and pixel-domain
VIF, restated in plain torch with torchmetrics' definitions and defaults.  Evaluation tooling, not the hot path: they run
on whatever device the tensors live on.  Each metric accumulates over ``update`` calls like its torchmetrics namesake."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.nn import functional as F


class PeakSignalNoiseRatio:
    """10 log10(data_range^2 / MSE), MSE over every element seen (torchmetrics ``PeakSignalNoiseRatio(data_range=1.0)``,
    reference pretrain.py:209, validate.py:84)."""

    def __init__(self, data_range: float = 1.0):
        self.data_range, self.se, self.n = float(data_range), 0.0, 0

    def update(self, preds: Tensor, target: Tensor) -> None:
        self.se += float(((preds.double() - target.double()) ** 2).sum())
        self.n += target.numel()

    def compute(self) -> float:
        return float(10.0 * torch.log10(torch.tensor(self.data_range ** 2 / max(self.se / max(self.n, 1), 1e-30))))


def _gaussian_kernel(size: int, sigma: float, device, dtype) -> Tensor:
    x = torch.arange(size, device=device, dtype=dtype) - (size - 1) / 2.0
    g = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    g = g / g.sum()
    return g[:, None] * g[None, :]


class StructuralSimilarityIndexMeasure:
    """Mean SSIM per image, averaged over images: 11 x 11 Gaussian window (sigma 1.5), K1 = 0.01, K2 = 0.03, reflect
    padding cropped away again, data_range = max - min over preds and target of the batch when not given (torchmetrics
    ``StructuralSimilarityIndexMeasure()`` defaults; reference validate.py:85)."""

    def __init__(self, data_range: float | None = None, kernel_size: int = 11, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03):
        self.data_range, self.ks, self.sigma, self.k1, self.k2 = data_range, kernel_size, sigma, k1, k2
        self.total, self.count = 0.0, 0

    def update(self, preds: Tensor, target: Tensor) -> None:
        preds, target = preds.float(), target.float()
        dr = self.data_range if self.data_range is not None else float(max(preds.max() - preds.min(), target.max() - target.min()))
        c1, c2 = (self.k1 * dr) ** 2, (self.k2 * dr) ** 2
        C = preds.shape[1]
        pad = (self.ks - 1) // 2
        k = _gaussian_kernel(self.ks, self.sigma, preds.device, preds.dtype).expand(C, 1, self.ks, self.ks)
        p, t = F.pad(preds, (pad, pad, pad, pad), mode="reflect"), F.pad(target, (pad, pad, pad, pad), mode="reflect")
        both = torch.cat([p, t, p * p, t * t, p * t])
        out = F.conv2d(both, k, groups=C)
        mu_p, mu_t, e_pp, e_tt, e_pt = out.split(preds.shape[0])
        s_pp, s_tt, s_pt = e_pp - mu_p ** 2, e_tt - mu_t ** 2, e_pt - mu_p * mu_t
        ssim = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (s_pp + s_tt + c2))
        ssim = ssim[..., pad:-pad, pad:-pad] if pad else ssim
        self.total += float(ssim.reshape(ssim.shape[0], -1).mean(-1).sum())
        self.count += ssim.shape[0]

    def compute(self) -> float:
        return self.total / max(self.count, 1)


class VisualInformationFidelity:
    """Pixel-domain VIF (VIF-P, Sheikh & Bovik) with sigma_n^2 = 2.0, four scales, per channel, averaged
    (torchmetrics ``VisualInformationFidelity()``; reference validate.py:86).  Inputs in [0, 1] are scaled to [0, 255] the
    way the metric is defined; images must be at least 41 x 41."""

    def __init__(self, sigma_n_sq: float = 2.0):
        self.sigma_n_sq, self.total, self.count = sigma_n_sq, 0.0, 0

    def _channel(self, p: Tensor, t: Tensor) -> Tensor:
        eps = 1e-10
        num = torch.zeros(p.shape[0], device=p.device, dtype=p.dtype)
        den = torch.zeros_like(num)
        for scale in range(4):
            n = 2.0 ** (4 - scale) + 1
            k = _gaussian_kernel(int(n), n / 5.0, p.device, p.dtype)[None, None]
            if scale > 0:
                t, p = F.conv2d(t, k)[:, :, ::2, ::2], F.conv2d(p, k)[:, :, ::2, ::2]
            mu_t, mu_p = F.conv2d(t, k), F.conv2d(p, k)
            s_tt = (F.conv2d(t * t, k) - mu_t ** 2).clamp(min=0)
            s_pp = (F.conv2d(p * p, k) - mu_p ** 2).clamp(min=0)
            s_tp = F.conv2d(t * p, k) - mu_t * mu_p
            g = s_tp / (s_tt + eps)
            sv = s_pp - g * s_tp
            mask = s_tt < eps
            g, sv, s_tt = torch.where(mask, torch.zeros_like(g), g), torch.where(mask, s_pp, sv), torch.where(mask, torch.zeros_like(s_tt), s_tt)
            mask = s_pp < eps
            g, sv = torch.where(mask, torch.zeros_like(g), g), torch.where(mask, torch.zeros_like(sv), sv)
            mask = g < 0
            sv, g = torch.where(mask, s_pp, sv), g.clamp(min=0)
            sv = sv.clamp(min=eps)
            num = num + torch.log10(1.0 + g ** 2 * s_tt / (sv + self.sigma_n_sq)).sum(dim=(1, 2, 3))
            den = den + torch.log10(1.0 + s_tt / self.sigma_n_sq).sum(dim=(1, 2, 3))
        return num / den

    def update(self, preds: Tensor, target: Tensor) -> None:
        assert preds.shape[-1] >= 41 and preds.shape[-2] >= 41, "VIF needs images of at least 41 x 41 pixels"
        p, t = preds.double() * 255.0, target.double() * 255.0
        per_channel = [self._channel(p[:, c:c + 1], t[:, c:c + 1]) for c in range(p.shape[1])]
        v = torch.stack(per_channel).mean(0)
        self.total += float(v.sum())
        self.count += v.numel()

    def compute(self) -> float:
        return self.total / max(self.count, 1)
