"""``ultrazoom.control`` (0.2.x src/ultrazoom/control.py, absent from the snapshot; API per README.md:94,118-122)."""
from ultrazoom_b200.control import ControlVector  # noqa: F401
