"""CPU oracle for cv/grid_quality.py (assess_grid_quality, :228-306).  TEST INFRASTRUCTURE ONLY.
numpy restatement of the OpenCV calls (Laplacian aperture 1 with BORDER_REFLECT_101, calcHist, warpPerspective of the
mask through the C oracle's fixed-point sampler) with numpy's dtypes kept as the reference has them (float64 for
sharpness / contrast / completeness, float32 for the corner arithmetic)."""
from __future__ import annotations

import numpy as np

try:
    from . import oracle as O
except ImportError:  # imported as a top-level module
    import oracle as O

FIELDS = ("overall", "sharpness", "contrast", "completeness", "geometry", "size")


def order_points(pts):
    rect = np.zeros((4, 2), np.float32)
    s = pts.sum(axis=1)
    d = np.diff(pts, axis=1).flatten()
    rect[0], rect[2] = pts[np.argmin(s)], pts[np.argmax(s)]
    rect[1], rect[3] = pts[np.argmin(d)], pts[np.argmax(d)]
    return rect


def sharpness(gray):  # :48-63
    g = np.pad(gray.astype(np.int64), 1, mode="reflect")
    lap = g[:-2, 1:-1] + g[2:, 1:-1] + g[1:-1, :-2] + g[1:-1, 2:] - 4 * g[1:-1, 1:-1]
    return min(100, lap.astype(np.float64).var() / 10)


def contrast(gray):  # :66-88
    hist = np.bincount(gray.ravel(), minlength=256).astype(np.float32)
    cumsum = np.cumsum(hist)
    total = gray.size
    return min(100, (np.searchsorted(cumsum, total * 0.975) - np.searchsorted(cumsum, total * 0.025)) / 2)


def completeness(binary, corners):  # :91-141
    size = 450
    m3 = np.repeat(np.ascontiguousarray(binary)[..., None], 3, axis=2)
    warped = O.warp_perspective(m3, np.asarray(corners).astype(np.int32))[..., 0]
    cell = size // 9
    scores = []
    for i in range(10):
        y = min(i * cell, size - 1)
        scores.append(np.mean(warped[max(0, y - 2):min(size, y + 3), :] > 0))
        scores.append(np.mean(warped[:, max(0, y - 2):min(size, y + 3)] > 0))
    return min(100, np.mean(scores) / 0.5 * 100)


def geometry(corners):  # :144-186
    o = order_points(np.asarray(corners, np.float32))
    sides = [np.linalg.norm(o[(i + 1) % 4] - o[i]) for i in range(4)]
    mean_side = np.mean(sides)
    var = np.std(sides) / mean_side if mean_side > 0 else 1
    angles = []
    for i in range(4):
        v1, v2 = o[i] - o[(i + 1) % 4], o[(i + 2) % 4] - o[(i + 1) % 4]
        c = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2) + 1e-6)
        angles.append(abs(np.degrees(np.arccos(np.clip(c, -1, 1))) - 90))
    return (max(0, 100 - var * 200) + max(0, 100 - np.mean(angles) * 5)) / 2


def size_score(corners):  # :189-211
    o = order_points(np.asarray(corners, np.float32))
    cell = np.mean([np.linalg.norm(o[(i + 1) % 4] - o[i]) for i in range(4)]) / 9
    if cell < 15:
        return cell / 15 * 30
    if cell < 30:
        return 30 + (cell - 15) / 15 * 40
    return min(100, 70 + (cell - 30) / 20 * 30)


def assess(image, binary, corners) -> dict:
    gray = O.gray(image) if image.ndim == 3 else image
    s = dict(sharpness=sharpness(gray), contrast=contrast(gray), completeness=completeness(binary, corners),
             geometry=geometry(corners), size=size_score(corners))
    s["overall"] = (0.25 * s["sharpness"] + 0.15 * s["contrast"] + 0.25 * s["completeness"] + 0.20 * s["geometry"]
                    + 0.15 * s["size"])
    return {k: float(s[k]) for k in FIELDS}
