"""Drop-in for cv/preprocess_v2.py — same names, arguments, defaults and return types; every function runs in
libsvb200's CUDA kernels (svb_preprocess_v2 / svb_preprocess_multi_v2 / svb_v2_stage / svb_cell_prep), bit-identical
to the OpenCV + numpy calls the reference makes.  Parameters the reference never varies raise NotImplementedError
when changed (there is no CPU fallback); frames of any size from 32x32 up (sides that do not divide by 8 get OpenCV's REFLECT_101-extended CLAHE tile grid)."""
import os
import sys
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
from numpy.typing import NDArray

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


@dataclass
class PreprocessResult:
    """cv/preprocess_v2.py:21-30."""
    binary: NDArray[np.uint8]
    gray: NDArray[np.uint8]
    enhanced: NDArray[np.uint8]
    illumination_normalized: Optional[NDArray[np.uint8]] = None
    has_glare: bool = False
    has_shadow: bool = False
    method_used: str = "adaptive"


def _stage(op, image, arg=0, want_image=True):
    img, info = rt.scanner().v2_stage(op, rt.to_device_u8(image)[None], arg, want_image)
    return (rt.to_host(img)[0] if img is not None else None), rt.to_host(info)[0]


def _gray2d(image, what):
    if len(image.shape) != 2:
        raise ValueError(f"{what}: expected a 2-D grayscale image")
    return image


def grayscale(image: NDArray[np.uint8]) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:33-37."""
    if len(image.shape) == 2:
        return image
    return rt.to_host(rt.scanner().grayscale(rt.to_device_u8(image)[None]))[0]


def normalize_illumination(gray: NDArray[np.uint8]) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:40-60."""
    return _stage("normalize_illumination", _gray2d(gray, "normalize_illumination"))[0]


def detect_glare(gray: NDArray[np.uint8], threshold: int = 250) -> Tuple[bool, NDArray[np.uint8]]:
    """cv/preprocess_v2.py:63-81."""
    if not 0 < int(threshold) < 255:
        raise NotImplementedError("detect_glare: threshold must be in 1..254")
    mask, info = _stage("detect_glare", _gray2d(gray, "detect_glare"), int(threshold))
    return bool(info[0]), mask


def detect_shadow(gray: NDArray[np.uint8]) -> Tuple[bool, NDArray[np.uint8]]:
    """cv/preprocess_v2.py:84-102."""
    mask, info = _stage("detect_shadow", _gray2d(gray, "detect_shadow"))
    return bool(info[1]), mask


def remove_shadow(gray: NDArray[np.uint8]) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:105-119."""
    return _stage("remove_shadow", _gray2d(gray, "remove_shadow"))[0]


def apply_clahe(gray: NDArray[np.uint8], clip_limit: float = 2.0, tile_size: int = 8) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:122-129."""
    if clip_limit != 2.0 or tile_size != 8:
        raise NotImplementedError("apply_clahe: only clip_limit=2.0, tile_size=8 (the reference's values) are implemented")
    return _stage("clahe8", _gray2d(gray, "apply_clahe"))[0]


def threshold_adaptive(gray: NDArray[np.uint8], block_size: int = 11, c: int = 2) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:132-143."""
    return rt.to_host(rt.scanner().adaptive_threshold(rt.to_device_u8(gray)[None], block_size, c, True))[0]


def threshold_otsu(gray: NDArray[np.uint8]) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:146-149."""
    return _stage("otsu", _gray2d(gray, "threshold_otsu"))[0]


def threshold_sauvola(gray: NDArray[np.uint8], window_size: int = 25, k: float = 0.2) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:152-175."""
    if window_size != 25 or k != 0.2:
        raise NotImplementedError("threshold_sauvola: only window_size=25, k=0.2 (the reference's values) are implemented")
    return _stage("sauvola", _gray2d(gray, "threshold_sauvola"))[0]


def morphological_cleanup(binary: NDArray[np.uint8], close_size: int = 3, open_size: int = 2) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:178-202."""
    if close_size != 3 or open_size != 2:
        raise NotImplementedError("morphological_cleanup: only close_size=3, open_size=2 (the reference's values) are implemented")
    return _stage("cleanup", _gray2d(binary, "morphological_cleanup"))[0]


def _batch1(image):
    t = rt.to_device_u8(image)
    if t.dim() not in (2, 3):
        raise ValueError("expected an (H,W,3) BGR or (H,W) gray image")
    return t[None]


def preprocess_for_grid_detection(image: NDArray[np.uint8], use_illumination_norm: bool = True,
                                  use_shadow_removal: bool = True) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:205-244."""
    mask, _ = rt.scanner().preprocess_v2(_batch1(image), use_illumination_norm, use_shadow_removal)
    return rt.to_host(mask)[0]


def preprocess_multi_strategy(image: NDArray[np.uint8]) -> PreprocessResult:
    """cv/preprocess_v2.py:247-308."""
    s = rt.scanner()
    r = s.preprocess_multi(_batch1(image))
    info = rt.to_host(r["info"])[0]
    return PreprocessResult(
        binary=rt.to_host(r["binary"])[0],
        gray=image if len(image.shape) == 2 else rt.to_host(r["gray"])[0],
        enhanced=rt.to_host(r["enhanced"])[0],
        illumination_normalized=rt.to_host(r["illumination_normalized"])[0],
        has_glare=bool(info[0]),
        has_shadow=bool(info[1]),
        method_used=s.V2_METHODS[int(info[2])],
    )


def preprocess_cell(cell: NDArray[np.uint8], clip_limit: float = 2.0, tile_size: int = 4) -> NDArray[np.uint8]:
    """cv/preprocess_v2.py:311-340: CLAHE(2.0,(4,4)) + adaptive threshold BINARY_INV, then 255 - x (= THRESH_BINARY)."""
    if clip_limit != 2.0 or tile_size != 4:
        raise NotImplementedError("preprocess_cell: only clip_limit=2.0, tile_size=4 (the reference's values) are implemented")
    if len(cell.shape) == 3:
        cell = grayscale(cell)
    if cell.shape != (28, 28):
        raise NotImplementedError("preprocess_cell: only 28x28 cells (what cv/extract.py produces) are implemented")
    thr, _ = rt.scanner().cell_prep(rt.to_device_u8(cell)[None], want_thresh=True, want_pm1=False)
    return rt.to_host(thr)[0]
