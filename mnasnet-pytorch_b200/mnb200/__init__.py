"""mnb200: B200-native MNASNet training step (hand-written sm_100a CUDA behind a C ABI)."""
from . import _lib  # noqa: F401
from .engine import Engine, DTYPES  # noqa: F401
