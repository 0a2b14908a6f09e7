"""svb200 — Python host side of the B200-native sudoku-vision scan path.

The compute lives in libsvb200.so (hand-written sm_100a CUDA behind the C ABI in
include/svb200.h); this package binds it with ctypes, uses PyTorch only for device memory and
streams, and mirrors the reference's Python interface in ../dropin/.
There is no CPU fallback: importing works anywhere (the CPU test tier checks the ABI), but every
compute call needs a B200 and fails loudly otherwise.
"""
from . import _lib  # noqa: F401
from .api import Scanner, load_digitcnn_weights, default_weights_path  # noqa: F401

__all__ = ["Scanner", "load_digitcnn_weights", "default_weights_path"]
