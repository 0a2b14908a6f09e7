"""ctypes binding of libsvb200.so (include/svb200.h).  One declaration per exported symbol."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsvb200.so")

SVB_OK, SVB_ERR_INVALID, SVB_ERR_UNSUPPORTED, SVB_ERR_CUDA, SVB_ERR_NOT_LOADED = 0, -1, -2, -3, -4

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_d = C.c_double

# name -> (restype, argtypes); mirrors include/svb200.h line by line
SIGNATURES = {
    "svb_create": (_i, [_i, C.POINTER(_p)]),
    "svb_destroy": (None, [_p]),
    "svb_last_error": (C.c_char_p, []),
    "svb_abi_version": (_i, []),
    "svb_launch_count": (_ll, [_p]),
    "svb_set_option": (_i, [_p, _i, _i]),
    "svb_stage_timing": (_i, [_p, _i]),
    "svb_last_stage_ms": (_i, [_p, _p]),
    "svb_grayscale": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "svb_blur": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "svb_adaptive_threshold": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "svb_preprocess_v1": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "svb_preprocess_v2": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "svb_preprocess_multi_v2": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "svb_v2_stage": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p]),
    "svb_find_grid_contour": (_i, [_p, _p, _i, _i, _i, _d, _d, _p, _p, _p]),
    "svb_detect_grid_contour_v2": (_i, [_p, _p, _i, _i, _i, _d, _p, _p, _p]),
    "svb_find_contours_count": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "svb_find_contours_fetch": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "svb_approx_poly_dp": (_i, [_p, _p, _i, _d, _p, _p, _p]),
    "svb_is_cell_empty": (_i, [_p, _p, _i, _i, _i, _d, _p, _p, _p]),
    "svb_warp_perspective": (_i, [_p, _p, _i, _i, _i, _p, _p, _i, _p, _p]),
    "svb_extract_cells": (_i, [_p, _p, _i, _i, _p, _p]),
    "svb_cell_prep": (_i, [_p, _p, _ll, _p, _p, _p]),
    "svb_cells_from_frames": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "svb_cells_from_frames_bits": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "svb_pack_cells_bits": (_i, [_p, _p, _ll, _p, _p]),
    "svb_digitcnn_forward_bits": (_i, [_p, _p, _ll, _p, _p, _p, _p]),
    "svb_digitcnn_load": (_i, [_p] + [_p] * 8 + [_p]),
    "svb_digitcnn_forward": (_i, [_p, _p, _ll, _p, _p, _p, _p]),
    "svb_set_classifier_mode": (_i, [_p, _i]),
    "svb_digitcnn_v3_load": (_i, [_p, _p, _i, _p]),
    "svb_digitcnn_v3_forward": (_i, [_p, _p, _ll, _p, _p, _p, _p, _p]),
    "svb_scan_batch_v1": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "svb_assess_grid_quality": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "svb_scan_batch_v2": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _p]),
    "svb_solve_batch": (_i, [_p, _p, _i, _p, _p, _p]),
    "svb_scan_batch_v1_host": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "svb_jpeg_decode_host": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "svb_scan_batch_v1_jpeg_host": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
}

_lib = None


class SvbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (built in-tree by sudoku-vision_b200/build.py).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvbError(f"{LIB_PATH} is missing: run `python sudoku-vision_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == SVB_OK:
        return
    msg = load().svb_last_error().decode("utf-8", "replace")
    if rc == SVB_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == SVB_ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise SvbError(f"{what} failed ({rc}): {msg}")
