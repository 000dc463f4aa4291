"""Engine: one shard of trajectories on one B200, driven through the C ABI (include/nqcb200.h)."""
from __future__ import annotations

from . import _abi


class Engine(_abi.CHandle):
    """Opaque handle of ``libnqcb200.so``.  Raises if the CUDA library or a GPU is missing."""

    def __init__(self, cfg: _abi.Config, keepalive=()):
        super().__init__(_abi.load_engine_library(), "nqcb200_", cfg, keepalive)


def device_count() -> int:
    return int(_abi.load_engine_library().nqcb200_device_count())
