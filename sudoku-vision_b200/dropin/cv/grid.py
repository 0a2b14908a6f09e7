"""Drop-in for cv/grid.py.  find_contours / approximate_polygon / find_grid_contour / warp_perspective run on the GPU
(svb_find_contours_*, svb_approx_poly_dp, svb_find_grid_contour, svb_warp_perspective); order_points is four-point host
bookkeeping exactly as the reference's numpy."""
import os
import sys

import numpy as np
from numpy.typing import NDArray

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


def find_contours(binary: NDArray[np.uint8]) -> tuple:
    """cv/grid.py:16-21 -> tuple of (n_i,1,2) int32 arrays: every RETR_EXTERNAL / CHAIN_APPROX_SIMPLE contour, in
    cv2's order (svb_find_contours_count / _fetch: flood of the outer background + one border walk per component)."""
    pts, offs = rt.scanner().find_contours(rt.to_device_u8(binary))
    pts, offs = rt.to_host(pts), rt.to_host(offs)
    return tuple(pts[offs[i]:offs[i + 1]].reshape(-1, 1, 2).copy() for i in range(len(offs) - 1))


def approximate_polygon(contour: NDArray, epsilon_ratio: float = 0.02) -> NDArray:
    """cv/grid.py:24-34 -> (m,1,2) int32 (svb_approx_poly_dp: arcLength + closed approxPolyDP in one warp)."""
    import torch

    c = np.ascontiguousarray(np.asarray(contour).reshape(-1, 2))
    if c.dtype != np.int32:
        raise NotImplementedError("approximate_polygon: only int32 contours (what find_contours returns) are implemented")
    return rt.to_host(rt.scanner().approx_poly_dp(torch.from_numpy(c), epsilon_ratio)).reshape(-1, 1, 2)


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
