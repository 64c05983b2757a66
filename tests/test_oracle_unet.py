"""CPU: the restated U-Net operators (oracle/unet_oracle.py) reproduce the fixtures the reference's own leaf classes
produced (tests/golden/unet_ops.npz, oracle/make_golden.py) -- the pin of the oracle for SURVEY.md 8(f) rank 3."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as U
from tests.helpers import GOLDEN


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLDEN + "/unet_ops.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _sd(gold, prefix):
    return {k[len(prefix) + 3:]: v for k, v in gold.items() if k.startswith(prefix + "/w:")}


def test_leaf_operators_match_the_reference_fixtures(gold):
    sd = _sd(gold, "mix")
    got = U.adaptive_residual_mix(gold["mix/in0"], gold["mix/in1"], sd["conv.weight"], sd["alpha"])
    assert torch.allclose(got, gold["mix/out"], atol=1e-6)
    for f in (2, 3, 4):
        got = U.pixel_crush(gold[f"crush{f}/in0"], _sd(gold, f"crush{f}")["conv.weight"], f)
        assert got.shape == gold[f"crush{f}/out"].shape and torch.allclose(got, gold[f"crush{f}/out"], atol=1e-6)
    got = U.subpixel_conv2d(gold["subpixel/in0"], _sd(gold, "subpixel")["conv.weight"], 2)
    assert torch.allclose(got, gold["subpixel/out"], atol=1e-6)
    for name in ("crop_smaller", "crop_larger", "crop_mixed"):
        got = U.crop_feature_maps(gold[name + "/in0"], tuple(int(v) for v in gold[name + "/size"]))
        assert torch.equal(got, gold[name + "/out"])
    got = U.quality_assessor(gold["quality/in0"], _sd(gold, "quality")["conv.weight"], _sd(gold, "quality")["conv.bias"])
    assert torch.allclose(got, gold["quality/out"], atol=1e-6)


def test_composite_blocks_match_the_reference_fixtures(gold):
    assert torch.allclose(U.encoder_block(gold["encoder_block/in0"], _sd(gold, "encoder_block")), gold["encoder_block/out"], atol=1e-5)
    assert torch.allclose(U.sr2x_block(gold["sr2x/in0"], _sd(gold, "sr2x")), gold["sr2x/out"], atol=1e-5)
    got = U.super_resolver(gold["super_resolver/in0"], _sd(gold, "super_resolver"), 4)
    assert got.shape == (1, 3, 20, 28) and torch.allclose(got, gold["super_resolver/out"], atol=1e-5)
