"""Shared runtime of the drop-in modules: one Scanner per process and numpy <-> device staging.

The reference's functions take and return host numpy arrays (SURVEY.md §8b "Data ownership"); the
shims copy the input to the GPU, run the CUDA kernels through the C ABI, synchronise and copy the
result back into a freshly allocated array.  Inputs are never mutated.  No CPU compute path exists:
without a B200 every call raises.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from svb200 import Scanner  # noqa: E402

_scanner = None


def scanner() -> Scanner:
    global _scanner
    if _scanner is None:
        _scanner = Scanner()
    return _scanner


def to_device_u8(a: np.ndarray):
    import torch

    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"expected a uint8 image, got {a.dtype}")
    return torch.from_numpy(a).to(scanner()._dev())


def to_host(t) -> np.ndarray:
    return t.cpu().numpy()
