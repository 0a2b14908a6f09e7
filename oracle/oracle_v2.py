"""CPU oracle for the v2 contour method (cv/grid_v2.py:49-128).  TEST INFRASTRUCTURE ONLY.
Built on the C oracle's contour primitives (svb_oracle.c) plus numpy float32 for the quadrilateral test."""
from __future__ import annotations

import numpy as np

try:
    from . import oracle as O
except ImportError:  # imported as a top-level module
    import oracle as O


def order_points(pts):
    """cv/grid_v2.py:49-61."""
    rect = np.zeros((4, 2), np.float32)
    s = pts.sum(axis=1)
    d = np.diff(pts, axis=1).flatten()
    rect[0], rect[2] = pts[np.argmin(s)], pts[np.argmax(s)]
    rect[1], rect[3] = pts[np.argmin(d)], pts[np.argmax(d)]
    return rect


def is_valid_quadrilateral(corners, min_angle: float = 45, max_angle: float = 135) -> bool:
    """cv/grid_v2.py:64-95 on float32 corners."""
    if corners.shape != (4, 2):
        return False
    c = corners.astype(np.float32)
    for i in range(4):
        v1, v2 = c[i] - c[(i + 1) % 4], c[(i + 2) % 4] - c[(i + 1) % 4]
        cos_a = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2) + 1e-6)
        ang = np.degrees(np.arccos(np.clip(cos_a, -1, 1)))
        if ang < min_angle or ang > max_angle:
            return False
    sides = [np.linalg.norm(c[(i + 1) % 4] - c[i]) for i in range(4)]
    return not (max(sides) > 2 * min(sides))


def detect_grid_contour(mask, min_area_ratio: float = 0.1):
    """cv/grid_v2.py:102-128 -> ordered (4,2) float32 or None."""
    cs = O.find_contours_external(mask)
    if not cs:
        return None
    floor = min_area_ratio * mask.shape[0] * mask.shape[1]
    areas = [O.contour_area(c) for c in cs]
    for i in sorted(range(len(cs)), key=lambda k: areas[k], reverse=True):  # stable, like sorted(..., reverse=True)
        if areas[i] < floor:
            break
        poly = O.approx_poly_dp_closed(cs[i], 0.02 * O.arc_length_closed(cs[i]))
        if len(poly) == 4:
            q = poly.reshape(4, 2).astype(np.float32)
            if is_valid_quadrilateral(q):
                return order_points(q)
    return None
