"""pltfalcor_b200 — B200-native (sm_100a) implementation of PLTFalcor's PLTPT render-pass hot path.

The product is the C-ABI library `libpltpt.so` (include/pltpt.h, sources in csrc/).  The Python modules are the
procedural scene harness (scene.py), buffer layouts (layouts.py) and a ctypes binding (binding.py) for tests/bench."""
from . import layouts  # noqa: F401

__all__ = ["layouts", "scene", "binding", "build"]
