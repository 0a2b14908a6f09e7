"""Drop-in for cv/preprocess.py — same names, arguments, defaults and return types; the arithmetic
runs in libsvb200's CUDA kernels (svb_grayscale / svb_blur / svb_adaptive_threshold /
svb_preprocess_v1), bit-identical to the OpenCV calls the reference makes."""
import os
import sys

import numpy as np
from numpy.typing import NDArray

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


def grayscale(image: NDArray[np.uint8]) -> NDArray[np.uint8]:
    """cv/preprocess.py:15-19."""
    if len(image.shape) == 2:
        return image
    return rt.to_host(rt.scanner().grayscale(rt.to_device_u8(image)[None]))[0]


def blur(image: NDArray[np.uint8], ksize: int = 5) -> NDArray[np.uint8]:
    """cv/preprocess.py:22-29.  ksize other than 5 raises NotImplementedError."""
    return rt.to_host(rt.scanner().blur(rt.to_device_u8(image)[None], ksize))[0]


def threshold(image: NDArray[np.uint8], block_size: int = 11, c: int = 2) -> NDArray[np.uint8]:
    """cv/preprocess.py:32-54 (THRESH_BINARY_INV)."""
    return rt.to_host(rt.scanner().adaptive_threshold(rt.to_device_u8(image)[None], block_size, c, True))[0]


def preprocess_for_grid_detection(image: NDArray[np.uint8]) -> NDArray[np.uint8]:
    """cv/preprocess.py:57-65 — one fused kernel."""
    if len(image.shape) == 2:  # already gray: the reference skips cvtColor
        s = rt.scanner()
        return rt.to_host(s.adaptive_threshold(s.blur(rt.to_device_u8(image)[None], 5), 11, 2, True))[0]
    return rt.to_host(rt.scanner().preprocess(rt.to_device_u8(image)[None]))[0]
