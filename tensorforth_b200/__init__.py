"""
tensorforth_b200 — B200-native (sm_100a) implementation of tensorForth's tensor-op hot path:
hand-written CUDA kernels behind the C-ABI of include/t4k.h (libt4k.so) plus the host-side
mirror of the reference's Tensor / Model surface.  See DESIGN.md.
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
