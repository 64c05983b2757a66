"""GPU: the 0.3.0 U-Net operators (ultrazoom_b200.unet -> csrc/unet_ops.cu + the tcgen05 convolutions) against the fixtures
computed by the reference's OWN leaf classes (tests/golden/unet_ops.npz) and against the pinned
oracle at further shapes.  Operators with fp32 arithmetic end to end (mix, crush, shuffle, crop, assessor) match to
1e-5; blocks that contain 3x3 convolutions run those on fp16 tensor-core operands with fp32 accumulation and are held
to 6e-3 x the output scale (the same budget per convolution as the flat path's stated tolerances)."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as U
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLDEN + "/unet_ops.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _sd(gold, prefix):
    return {k[len(prefix) + 3:]: v for k, v in gold.items() if k.startswith(prefix + "/w:")}


def _close(got_nhwc, want_nchw, tol):
    from ultrazoom_b200.unet import to_nchw

    got = to_nchw(got_nhwc).cpu()
    assert got.shape == want_nchw.shape, (got.shape, want_nchw.shape)
    err = (got - want_nchw).abs().max().item()
    assert err <= tol, err


def test_mix_crush_shuffle_crop_assessor_match_reference_fixtures(dev, gold):
    from ultrazoom_b200 import unet as N

    m = N.AdaptiveResidualMix(16)
    m.load_state_dict(_sd(gold, "mix"))
    _close(m.forward(N.to_nhwc(gold["mix/in0"]).to(dev), N.to_nhwc(gold["mix/in1"]).to(dev)), gold["mix/out"], 1e-5)
    for f in (2, 3, 4):
        c = N.PixelCrush(16, 32, f)
        c.load_state_dict(_sd(gold, f"crush{f}"))
        _close(c.forward(N.to_nhwc(gold[f"crush{f}/in0"]).to(dev)), gold[f"crush{f}/out"], 1e-5)
    for name in ("crop_smaller", "crop_larger", "crop_mixed"):
        size = tuple(int(v) for v in gold[name + "/size"])
        _close(N.crop_feature_maps(N.to_nhwc(gold[name + "/in0"]).to(dev), size), gold[name + "/out"], 0.0)
    q = N.QualityAssessor(32, 3)
    q.load_state_dict(_sd(gold, "quality"))
    got = q.forward(N.to_nhwc(gold["quality/in0"]).to(dev)).cpu()
    assert got.shape == (2, 3) and (got - gold["quality/out"]).abs().max().item() <= 1e-5
    z = np.load(GOLDEN + "/leaf_ops.npz")                                 # PixelShuffle known answers, r = 2, 3, 4
    for r in (2, 3, 4):
        x = torch.from_numpy(z[f"shuffle_in_r{r}"])
        _close(N.pixel_shuffle_nhwc(N.to_nhwc(x).to(dev), r), torch.from_numpy(z[f"shuffle_out_r{r}"]), 0.0)


def test_blocks_with_convolutions_match_reference_fixtures(dev, gold):
    from ultrazoom_b200 import unet as N

    up = N.SubpixelConv2d(32, 16, 2)
    up.load_state_dict(_sd(gold, "subpixel"))
    _close(up.forward(N.to_nhwc(gold["subpixel/in0"]).to(dev)), gold["subpixel/out"], 6e-3 * gold["subpixel/out"].abs().max().item())
    blk = N.EncoderBlock(16, 2)
    blk.load_state_dict(_sd(gold, "encoder_block"))
    _close(blk.forward(N.to_nhwc(gold["encoder_block/in0"]).to(dev)), gold["encoder_block/out"],
           6e-3 * gold["encoder_block/out"].abs().max().item())
    sr = N.SR2XBlock(16, 2, 16)
    sr.load_state_dict(_sd(gold, "sr2x"))
    _close(sr.forward(N.to_nhwc(gold["sr2x/in0"]).to(dev)), gold["sr2x/out"], 6e-3 * gold["sr2x/out"].abs().max().item())
    head = N.SuperResolver(16, 2, 4)
    head.load_state_dict(_sd(gold, "super_resolver"))
    got = head.forward(N.to_nhwc(gold["super_resolver/in0"]).to(dev))
    assert tuple(got.shape) == (1, 20, 28, 3)
    _close(got, gold["super_resolver/out"], 1.2e-2 * gold["super_resolver/out"].abs().max().item())


@pytest.mark.parametrize("C,shape", [(16, (1, 1, 1)), (48, (2, 37, 61)), (96, (1, 20, 130)), (32, (3, 5, 257))])
def test_mix_and_crush_against_the_oracle_at_other_shapes(dev, C, shape):
    from ultrazoom_b200 import unet as N

    g = torch.Generator().manual_seed(C + sum(shape))
    B, H, W = shape
    x, z = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5
    for a in (-2.0, 0.0, 1.5):
        alpha = torch.tensor(a)
        got = N.adaptive_residual_mix(N.to_nhwc(x).to(dev), N.to_nhwc(z).to(dev), w, alpha)
        _close(got, U.adaptive_residual_mix(x, z, w, alpha), 2e-5)
    for f in (2, 3, 4):
        if H < f or W < f:
            with pytest.raises(AssertionError):
                N.pixel_crush(N.to_nhwc(x).to(dev), torch.randn(C, C, f, f), f)
            continue
        wc = torch.randn(2 * C, C, f, f, generator=g) / (f * f * C) ** 0.5
        _close(N.pixel_crush(N.to_nhwc(x).to(dev), wc, f), U.pixel_crush(x, wc, f), 2e-5)
    wq, bq = torch.randn(3, C, 3, 3, generator=g) / (9 * C) ** 0.5, torch.randn(3, generator=g)
    got = N.quality_assessor(N.to_nhwc(x).to(dev), wq, bq).cpu()
    assert (got - U.quality_assessor(x, wq, bq)).abs().max().item() <= 2e-5
    for size in ((H + 3, W - 1 if W > 1 else 1), (max(1, H - 2), W + 4)):
        _close(N.crop_feature_maps(N.to_nhwc(x).to(dev), size), U.crop_feature_maps(x, size), 0.0)


def test_argument_checks_and_no_cpu_fallback(dev):
    from ultrazoom_b200 import unet as N

    with pytest.raises(AssertionError, match="Crush factor"):
        N.PixelCrush(16, 32, 5)                                          # reference model.py:853-857
    with pytest.raises(AssertionError, match="Upscale ratio"):
        N.SuperResolver(16, 2, 3)                                        # model.py:939-943
    with pytest.raises(AssertionError, match="Hidden ratio"):
        N.InvertedBottleneck(16, 3)                                      # model.py:738
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        N.adaptive_residual_mix(torch.rand(1, 2, 2, 16), torch.rand(1, 2, 2, 16), torch.rand(16, 32, 1, 1), torch.tensor(0.0))
    with pytest.raises(AssertionError):
        N.adaptive_residual_mix(torch.rand(1, 2, 2, 16, device=dev), torch.rand(1, 2, 3, 16, device=dev),
                                torch.rand(16, 32, 1, 1), torch.tensor(0.0))
