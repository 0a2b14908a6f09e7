"""Drop-in for cv/extract.py.  extract_cells and is_cell_empty run on the GPU (svb_extract_cells, svb_is_cell_empty)."""
import os
import sys

import numpy as np
from numpy.typing import NDArray

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


def extract_cells(grid_image: NDArray[np.uint8], cell_size: int = 28, margin_ratio: float = 0.1) -> list:
    """cv/extract.py:13-56 -> list of 81 (28,28) uint8 arrays, row-major."""
    if cell_size != 28 or margin_ratio != 0.1:
        raise NotImplementedError("extract_cells: only cell_size=28, margin_ratio=0.1 (the defaults) are implemented")
    if grid_image.ndim != 3 or grid_image.shape[2] != 3 or grid_image.shape[0] != grid_image.shape[1]:
        raise NotImplementedError("extract_cells: only square 3-channel boards are implemented")
    cells = rt.to_host(rt.scanner().extract_cells(rt.to_device_u8(grid_image)[None]))[0]
    return [cells[i].copy() for i in range(81)]


def is_cell_empty(cell: NDArray[np.uint8], threshold: float = 0.02) -> bool:
    """cv/extract.py:59-79: Otsu (THRESH_BINARY_INV) + countNonZero ratio, on the GPU (svb_is_cell_empty)."""
    if cell.ndim != 2 or cell.dtype != np.uint8:
        raise NotImplementedError("is_cell_empty: only 2-D uint8 cells are implemented")
    empty = rt.scanner().is_cell_empty(rt.to_device_u8(cell)[None], threshold)
    return bool(int(empty.cpu()[0]))


def preprocess_cell_for_model(cell: NDArray[np.uint8]) -> NDArray[np.float32]:
    """cv/extract.py:82-99 (dead code in the reference: nothing calls it).  For an already-gray
    28x28 cell it is a plain rescale, done here as host bookkeeping; other inputs are not implemented."""
    if cell.shape != (28, 28):
        raise NotImplementedError("preprocess_cell_for_model: only (28,28) gray cells are implemented")
    return (cell.astype(np.float32) / 255.0).reshape(1, 28, 28)
