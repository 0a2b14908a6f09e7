"""Drop-in for cv/grid.py.  find_grid_contour / warp_perspective run on the GPU (svb_find_grid_contour,
svb_warp_perspective); order_points is four-point host bookkeeping exactly as the reference's numpy."""
import os
import sys

import numpy as np
from numpy.typing import NDArray

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


def find_contours(binary: NDArray[np.uint8]) -> list:
    """cv/grid.py:16-21.  The batched path never materialises the full contour list (it traces only
    borders that can reach min_area), so this debug helper is not provided."""
    raise NotImplementedError("find_contours: the B200 path does not materialise all contours; use find_grid_contour")


def approximate_polygon(contour: NDArray, epsilon_ratio: float = 0.02) -> NDArray:
    """cv/grid.py:24-34 — fused inside find_grid_contour on the GPU; not exposed per contour."""
    raise NotImplementedError("approximate_polygon: fused into find_grid_contour on the GPU")


def find_grid_contour(binary: NDArray[np.uint8], min_area_ratio: float = 0.1) -> NDArray | None:
    """cv/grid.py:37-71 -> (4,2) int32 in approxPolyDP order, or None."""
    corners, found = rt.scanner().find_grid_contour(rt.to_device_u8(binary)[None], min_area_ratio, 0.02)
    f = int(found.cpu()[0])
    if f == 2:
        raise RuntimeError("find_grid_contour: contour scratch capacity exceeded on the GPU")
    return rt.to_host(corners)[0] if f == 1 else None


def order_points(pts: NDArray) -> NDArray:
    """cv/grid.py:74-91: TL, TR, BR, BL (first index wins ties, as numpy argmin/argmax)."""
    rect = np.zeros((4, 2), dtype=np.float32)
    s = pts.sum(axis=1)
    d = np.diff(pts, axis=1)
    rect[0], rect[2] = pts[np.argmin(s)], pts[np.argmax(s)]
    rect[1], rect[3] = pts[np.argmin(d)], pts[np.argmax(d)]
    return rect


def warp_perspective(image: NDArray[np.uint8], corners: NDArray, output_size: int = 450,
                     inset_ratio: float = 0.0) -> NDArray[np.uint8]:
    """cv/grid.py:94-133.  inset_ratio other than 0 raises NotImplementedError."""
    if inset_ratio != 0.0:
        raise NotImplementedError("warp_perspective: only inset_ratio=0.0 (the reference's default) is implemented")
    import torch

    c = np.asarray(corners)
    ci = np.rint(c).astype(np.int32).reshape(1, 4, 2)
    if not np.array_equal(ci.reshape(4, 2).astype(np.float32), c.astype(np.float32).reshape(4, 2)):
        raise NotImplementedError("warp_perspective: non-integer corners are not implemented")
    if image.ndim != 3 or image.shape[2] != 3:
        raise NotImplementedError("warp_perspective: only 3-channel images are implemented")
    board = rt.scanner().warp_perspective(rt.to_device_u8(image)[None], torch.from_numpy(ci), None, output_size)
    return rt.to_host(board)[0]
