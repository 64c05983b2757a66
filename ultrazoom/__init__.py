"""``ultrazoom`` -- import-path shim so that reference user code runs verbatim on the B200-native forward:

    from ultrazoom.model import MewZoom              # reference README.md:69,102
    from ultrazoom.control import ControlVector      # reference README.md:103

Everything lives in ``ultrazoom_b200``; this package only re-exports it under the reference's module names
(src/ultrazoom/__init__.py, src/ultrazoom/model.py and the 0.2.x src/ultrazoom/control.py).  There is no CPU
fallback behind these names either.
"""
from ultrazoom_b200 import MODEL_CONFIGS, __version__  # noqa: F401
