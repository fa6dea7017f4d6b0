"""better_fastlio2_b200 — B200-native (sm_100a) FAST-LIO2 per-scan hot path.

The product is `libfastlio_b200.so` (hand-written CUDA behind the C ABI of include/fastlio_b200.h) plus the C++
facades in include/fastlio_b200/.  This Python package only holds:
  capi   — ctypes view of the C ABI (used by tests/ and bench.py);
  synth  — seeded synthetic scenes (numpy).
Nothing here computes the path on the CPU; without the CUDA library / a GPU the calls fail loudly.
"""
from . import synth  # noqa: F401

__all__ = ["synth", "capi"]
