"""The evaluation CLI (tools/validate.py; reference validate.py:23-125) and its metrics (ultrazoom_b200/metrics.py)."""
import os
import sys

import pytest
import torch

from ultrazoom_b200.metrics import PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure, VisualInformationFidelity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_metrics_definitions():
    g = torch.Generator().manual_seed(0)
    y = torch.rand(2, 3, 64, 80, generator=g)
    p = (y + 0.05 * torch.randn(2, 3, 64, 80, generator=g)).clamp(0, 1)
    q = (y + 0.15 * torch.randn(2, 3, 64, 80, generator=g)).clamp(0, 1)
    m = PeakSignalNoiseRatio(data_range=1.0)
    m.update(p[:1], y[:1])
    m.update(p[1:], y[1:])                                              # accumulates over updates like torchmetrics
    assert m.compute() == pytest.approx(float(10 * torch.log10(1 / ((p - y) ** 2).mean())), abs=1e-4)
    for cls in (StructuralSimilarityIndexMeasure, VisualInformationFidelity):
        same, near, far = cls(), cls(), cls()
        same.update(y, y)
        near.update(p, y)
        far.update(q, y)
        assert same.compute() == pytest.approx(1.0, abs=1e-6)
        assert 0.0 < far.compute() < near.compute() < 1.0
    s = StructuralSimilarityIndexMeasure(data_range=1.0)                # constant images: SSIM = luminance term only
    s.update(torch.full((1, 1, 32, 32), 0.5), torch.full((1, 1, 32, 32), 0.25))
    c1 = 0.01 ** 2
    assert s.compute() == pytest.approx((2 * 0.5 * 0.25 + c1) / (0.25 + 0.0625 + c1), abs=1e-5)


@pytest.mark.gpu
def test_validate_cli_end_to_end(tmp_path):
    """Synthetic pairs through a random-init model; then a reference-style checkpoint (0.2.x schema: model_args / model,
    weight-norm parametrizations, _orig_mod. prefixes -- validate.py:51-65) and an LR / HR folder pair of PNG files."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import validate as V
    from torchvision.io import write_png

    from ultrazoom_b200 import MewZoom

    res = V.main(["--synthetic", "2", "--model", "MewZoom-2X-Ctrl"])
    assert res["images"] == 2 and res["bicubic"]["psnr"] > 5.0 and -1.0 <= res["enhanced"]["ssim"] <= 1.0   # (noise images)
    cfg = dict(upscale_ratio=2, num_channels=16, hidden_ratio=2, num_encoder_layers=2, control_features=3)
    torch.manual_seed(3)
    trained = MewZoom(**cfg)
    with torch.no_grad():
        trained.head.conv.weight.mul_(0.05)                              # a head near zero: output close to bicubic
    trained.add_weight_norms()
    sd = {"_orig_mod." + k: v for k, v in trained.state_dict().items()}
    ckpt = tmp_path / "checkpoint.pt"
    torch.save({"model_args": cfg, "model": sd}, ckpt)
    lr_dir, hr_dir = tmp_path / "lr", tmp_path / "hr"
    lr_dir.mkdir()
    hr_dir.mkdir()
    g = torch.Generator().manual_seed(4)
    for i in range(3):
        hr = torch.nn.functional.interpolate(torch.rand(1, 3, 12, 16, generator=g), size=(96, 128), mode="bicubic").clamp(0, 1)
        lr = torch.nn.functional.interpolate(hr, scale_factor=0.5, mode="bicubic", antialias=True).clamp(0, 1)
        write_png((hr[0] * 255).round().to(torch.uint8), str(hr_dir / f"img{i}.png"))
        write_png((lr[0] * 255).round().to(torch.uint8), str(lr_dir / f"img{i}.png"))
    res = V.main(["--checkpoint_path", str(ckpt), "--lr_images_path", str(lr_dir), "--hr_images_path", str(hr_dir),
                  "--gaussian_blur", "0.5", "--gaussian_noise", "0.2", "--jpeg_compression", "0.3"])
    assert res["images"] == 3
    assert res["bicubic"]["psnr"] > 25.0 and res["bicubic"]["ssim"] > 0.8      # smooth images: bicubic is already close
    assert res["enhanced"]["psnr"] > 12.0 and 0.0 < res["enhanced"]["ssim"] < 1.0   # a random-init residual: worse, but an image
    with pytest.raises(RuntimeError, match="no CPU path"):
        V.main(["--synthetic", "1", "--device", "cpu"])


def test_export_round_trip(tmp_path):
    """tools/export_model.py: reference checkpoint (weight-normed, compiled-model prefixes) -> save_pretrained ->
    from_pretrained with identical weights and constructor arguments (export_model.ipynb cells 3-7)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import export_model as E

    from ultrazoom_b200 import MewZoom

    cfg = dict(upscale_ratio=3, num_channels=16, hidden_ratio=2, num_encoder_layers=2, control_features=3)
    torch.manual_seed(5)
    m = MewZoom(**cfg)
    want = {k: v.clone() for k, v in m.state_dict().items()}
    m.add_weight_norms()
    ckpt = tmp_path / "c.pt"
    torch.save({"upscaler_args": cfg, "upscaler": {"_orig_mod." + k: v for k, v in m.state_dict().items()}}, ckpt)
    out = E.main(["--checkpoint_path", str(ckpt), "--out", str(tmp_path / "export")])
    again = MewZoom.from_pretrained(out)
    assert again.upscale_ratio == 3 and again.control_features == 3 and again.num_encoder_layers == 2
    for k, v in again.state_dict().items():
        assert torch.allclose(v, want[k], atol=1e-6), k
