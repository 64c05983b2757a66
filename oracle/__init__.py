"""CPU fp32 oracle for the MewZoom.upscale hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ultrazoom_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the timed CPU baseline.
"""
from .mewzoom_oracle import (  # noqa: F401
    OracleMewZoom,
    OracleControlVector,
    bicubic_upsample_ref,
    pixel_shuffle_ref,
    psnr,
    max_abs_err,
    residual_rms,
    MODEL_CONFIGS,
    make_oracle,
)
