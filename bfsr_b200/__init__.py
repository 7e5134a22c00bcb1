"""bfsr_b200 — Blackwell-native flow-SR inference (SRFlow-LP / LINF-LP hot path) behind the reference's interfaces.

Everything numeric runs in `libbfsr_b200.so` (hand-written sm_100a CUDA, C ABI in include/bfsr_b200.h);
this package is the thin host-side mirror of the reference's Python surface.
"""
from . import eval, metrics, models  # noqa: F401
from ._lib import BfsrError, LIB_PATH  # noqa: F401

__all__ = ["models", "metrics", "eval", "BfsrError", "LIB_PATH"]
