"""``ultrazoom.model`` (reference src/ultrazoom/model.py:43,195): MewZoom and ONNXModel of the B200-native build."""
from ultrazoom_b200.model import (  # noqa: F401
    EncoderBlock,
    FanOutProjection,
    GraphedUpscale,
    InvertedBottleneck,
    MewZoom,
    ONNXModel,
    SubpixelConv2d,
)
