"""Drop-in for the reference's `cv` package (cv/__init__.py:8-19 re-exports)."""
from .preprocess import preprocess_for_grid_detection
from .grid import find_grid_contour, warp_perspective, order_points
from .extract import extract_cells, is_cell_empty, preprocess_cell_for_model

__all__ = [
    "preprocess_for_grid_detection",
    "find_grid_contour",
    "warp_perspective",
    "order_points",
    "extract_cells",
    "is_cell_empty",
    "preprocess_cell_for_model",
]
