"""lrcn_b200: B200-native LRCN caption-decoder hot path (C-ABI liblrcn_b200.so + host mirror)."""
from . import synth  # noqa: F401

__all__ = ["synth"]
