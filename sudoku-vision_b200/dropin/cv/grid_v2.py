"""Drop-in for the contour method of cv/grid_v2.py (method 1 of detect_grid) and its warp_perspective.
The GPU does contour tracing, approxPolyDP, the quadrilateral validity test and the corner ordering
(svb_detect_grid_contour_v2); the Hough / rotation-retry / Harris+RANSAC fallbacks (methods 2-4, one of them
seeded by an unseeded RNG in the reference) are not built: when method 1 finds nothing and the caller asked for
fallbacks, detect_grid raises NotImplementedError instead of silently answering differently."""
import os
import sys
from dataclasses import dataclass
from typing import Optional

import numpy as np
from numpy.typing import NDArray

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


@dataclass
class GridDetectionResult:
    """cv/grid_v2.py:23-31."""
    corners: Optional[NDArray[np.float32]]
    confidence: float
    method: str
    rotation_angle: float
    is_partial: bool
    debug_info: dict


def find_contours(binary):
    raise NotImplementedError("find_contours: the B200 path does not materialise all contours")


def approximate_polygon(contour, epsilon_ratio: float = 0.02):
    raise NotImplementedError("approximate_polygon: fused into detect_grid_contour on the GPU")


def order_points(pts: NDArray) -> NDArray:
    """cv/grid_v2.py:49-61 (four points: host bookkeeping)."""
    rect = np.zeros((4, 2), dtype=np.float32)
    s = pts.sum(axis=1)
    d = np.diff(pts, axis=1).flatten()
    rect[0], rect[2] = pts[np.argmin(s)], pts[np.argmax(s)]
    rect[1], rect[3] = pts[np.argmin(d)], pts[np.argmax(d)]
    return rect


def is_valid_quadrilateral(corners: NDArray, min_angle: float = 45, max_angle: float = 135) -> bool:
    """cv/grid_v2.py:64-95 for callers that test a quadrilateral themselves (four points, host side); inside
    detect_grid_contour the same test runs on the GPU."""
    if corners.shape != (4, 2):
        return False
    for i in range(4):
        v1, v2 = corners[i] - corners[(i + 1) % 4], corners[(i + 2) % 4] - corners[(i + 1) % 4]
        cos_a = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2) + 1e-6)
        ang = np.degrees(np.arccos(np.clip(cos_a, -1, 1)))
        if ang < min_angle or ang > max_angle:
            return False
    sides = [np.linalg.norm(corners[(i + 1) % 4] - corners[i]) for i in range(4)]
    return not (max(sides) > 2 * min(sides))


def detect_grid_contour(binary: NDArray[np.uint8], min_area_ratio: float = 0.1) -> Optional[NDArray]:
    """cv/grid_v2.py:102-128 -> ordered (4,2) float32, or None."""
    corners, found = rt.scanner().detect_grid_contour_v2(rt.to_device_u8(binary)[None], min_area_ratio)
    f = int(found.cpu()[0])
    if f == 2:
        raise RuntimeError("detect_grid_contour: contour scratch capacity exceeded on the GPU")
    return rt.to_host(corners)[0].astype(np.float32) if f == 1 else None


def detect_grid(binary: NDArray[np.uint8], gray: Optional[NDArray[np.uint8]] = None, try_rotation: bool = True,
                try_multiple_methods: bool = True) -> GridDetectionResult:
    """cv/grid_v2.py:401-508, method 1."""
    corners = detect_grid_contour(binary)
    if corners is not None:
        return GridDetectionResult(corners=corners, confidence=0.9, method="contour", rotation_angle=0,
                                   is_partial=False, debug_info={})
    if not try_multiple_methods:
        return GridDetectionResult(corners=None, confidence=0, method="none", rotation_angle=0, is_partial=False,
                                   debug_info={})
    raise NotImplementedError("detect_grid: the contour method found no grid and methods 2-4 "
                              "(Hough lines, rotation retry, Harris+RANSAC) are not built on the B200 path")


def warp_perspective(image: NDArray[np.uint8], corners: NDArray, output_size: int = 450) -> NDArray[np.uint8]:
    """cv/grid_v2.py:511-529."""
    import torch

    c = np.asarray(corners)
    ci = np.rint(c).astype(np.int32).reshape(1, 4, 2)
    if not np.array_equal(ci.reshape(4, 2).astype(np.float32), c.astype(np.float32).reshape(4, 2)):
        raise NotImplementedError("warp_perspective: non-integer corners are not implemented")
    if image.ndim != 3 or image.shape[2] != 3:
        raise NotImplementedError("warp_perspective: only 3-channel images are implemented")
    board = rt.scanner().warp_perspective(rt.to_device_u8(image)[None], torch.from_numpy(ci), None, output_size)
    return rt.to_host(board)[0]
