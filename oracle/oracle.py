"""ctypes front end of the CPU oracle (oracle/svb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.  See svb_oracle.c for what each function
restates and which reference call site (file:line) it follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsvb_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "svb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.svo_contour_area.restype = C.c_double
        _lib.svo_arc_length_closed.restype = C.c_double
        _lib.svo_find_contours_external.restype = C.c_long
        _lib.svo_approx_poly_dp_closed.restype = C.c_long
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8
    return a


def gray(bgr):
    bgr = _u8(bgr)
    H, W = bgr.shape[:2]
    out = np.empty((H, W), np.uint8)
    lib().svo_gray_bgr(_p(bgr), H, W, _p(out))
    return out


def blur5(g):
    g = _u8(g)
    H, W = g.shape
    out = np.empty((H, W), np.uint8)
    lib().svo_blur5(_p(g), H, W, _p(out))
    return out


def gauss11_mean_f32(g):
    g = _u8(g)
    H, W = g.shape
    out = np.empty((H, W), np.float32)
    lib().svo_gauss11_mean_f32(_p(g), H, W, _p(out))
    return out


def adaptive_gauss11(g, inv: bool):
    g = _u8(g)
    H, W = g.shape
    out = np.empty((H, W), np.uint8)
    lib().svo_adaptive_gauss11(_p(g), H, W, int(inv), _p(out))
    return out


def preprocess(bgr):
    """cv/preprocess.py:57-65 preprocess_for_grid_detection."""
    bgr = _u8(bgr)
    H, W = bgr.shape[:2]
    out = np.empty((H, W), np.uint8)
    lib().svo_preprocess(_p(bgr), H, W, _p(out))
    return out


def find_contours_external(mask):
    """cv/grid.py:16-21: list of (n,2) int32 arrays in cv2's order."""
    mask = _u8(mask)
    H, W = mask.shape
    max_pts = int(mask.size) + 16
    max_c = int(mask.size) // 2 + 16
    pts = np.empty((max_pts, 2), np.int32)
    offs = np.empty(max_c + 1, np.int64)
    n = lib().svo_find_contours_external(_p(mask), H, W, _p(pts), C.c_long(max_pts), _p(offs), C.c_long(max_c))
    assert n >= 0
    return [pts[offs[i]:offs[i + 1]].copy() for i in range(n)]


def contour_area(pts):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    return lib().svo_contour_area(_p(pts), C.c_long(len(pts)))


def arc_length_closed(pts):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    return lib().svo_arc_length_closed(_p(pts), C.c_long(len(pts)))


def approx_poly_dp_closed(pts, eps: float):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    out = np.empty_like(pts)
    m = lib().svo_approx_poly_dp_closed(_p(pts), C.c_long(len(pts)), C.c_double(eps), _p(out))
    return out[:m].copy()


def find_grid_contour(mask, min_area_ratio: float = 0.1, eps_ratio: float = 0.02):
    """cv/grid.py:37-71: (4,2) int32 in DP order, or None."""
    mask = _u8(mask)
    H, W = mask.shape
    c = np.zeros((4, 2), np.int32)
    f = lib().svo_find_grid_contour(_p(mask), H, W, C.c_double(min_area_ratio), C.c_double(eps_ratio), _p(c))
    return c if f else None


def order_points(corners):
    c = np.ascontiguousarray(corners, np.int32).reshape(4, 2)
    out = np.empty((4, 2), np.float32)
    lib().svo_order_points(_p(c), _p(out))
    return out


def perspective_matrix(src, dst):
    src = np.ascontiguousarray(src, np.float32).reshape(4, 2)
    dst = np.ascontiguousarray(dst, np.float32).reshape(4, 2)
    M = np.empty(9, np.float64)
    lib().svo_perspective_matrix(_p(src), _p(dst), _p(M))
    return M.reshape(3, 3)


def warp_perspective_matrix(bgr, M, out_size: int = 450):
    bgr = _u8(bgr)
    H, W = bgr.shape[:2]
    M = np.ascontiguousarray(M, np.float64)
    out = np.empty((out_size, out_size, 3), np.uint8)
    lib().svo_warp_perspective(_p(bgr), H, W, _p(M), out_size, _p(out))
    return out


def warp_perspective(bgr, corners, out_size: int = 450):
    """cv/grid.py:94-133 with inset_ratio = 0."""
    o = order_points(corners)
    s = out_size - 1
    dst = np.array([[0, 0], [s, 0], [s, s], [0, s]], np.float32)
    return warp_perspective_matrix(bgr, perspective_matrix(o, dst), out_size)


def resize_linear(src, dh: int, dw: int):
    src = _u8(src)
    out = np.empty((dh, dw), np.uint8)
    lib().svo_resize_linear_u8(_p(src), src.shape[0], src.shape[1], _p(out), dh, dw)
    return out


def extract_cells(board):
    """cv/extract.py:13-56 -> (81,28,28) u8."""
    board = _u8(board)
    assert board.shape[0] == board.shape[1] and board.shape[2] == 3
    out = np.empty((81, 28, 28), np.uint8)
    lib().svo_extract_cells(_p(board), board.shape[0], _p(out))
    return out


def clahe28(cell):
    cell = _u8(cell)
    assert cell.shape == (28, 28)
    out = np.empty((28, 28), np.uint8)
    lib().svo_clahe_28(_p(cell), _p(out))
    return out


def cell_prep(cells):
    """pipeline/run.py:73-95 + :129-135.  (n,28,28) u8 -> (n,28,28) u8 in {0,255}; 255 == +1."""
    cells = _u8(cells).reshape(-1, 28, 28)
    out = np.empty_like(cells)
    for i in range(len(cells)):
        lib().svo_cell_prep(_p(cells[i]), _p(out[i]), None)
    return out


def digitcnn_forward(sd: dict, x):
    """ml/model.py:34-42.  sd: state_dict-like of numpy f32 arrays; x (n,1,28,28) f32 -> (n,10)."""
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 784)
    w = [np.ascontiguousarray(np.asarray(sd[k], dtype=np.float32)) for k in (
        "conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
        "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")]
    out = np.empty((len(x), 10), np.float32)
    lib().svo_digitcnn_forward(*[_p(a) for a in w], _p(x), C.c_long(len(x)), _p(out))
    return out


def scan_frame(bgr):
    """run.py:257-293 + :122-136 for one frame, image stages only."""
    bgr = _u8(bgr)
    H, W = bgr.shape[:2]
    mask = np.empty((H, W), np.uint8)
    corners = np.zeros((4, 2), np.int32)
    ordered = np.zeros((4, 2), np.float32)
    cells_u8 = np.zeros((81, 28, 28), np.uint8)
    cells_in = np.zeros((81, 28, 28), np.uint8)
    f = lib().svo_scan_frame(_p(bgr), H, W, _p(mask), _p(corners), _p(ordered), _p(cells_u8), _p(cells_in))
    return dict(found=bool(f), mask=mask, corners=corners if f else None, ordered=ordered if f else None,
                cells_u8=cells_u8 if f else None, cells_in=cells_in if f else None)


# ---- V1: cv/preprocess_v2.py --------------------------------------------------------------------------
def box_blur(g, k: int):
    g = _u8(g)
    out = np.empty_like(g)
    lib().svo_box_blur_u8(_p(g), g.shape[0], g.shape[1], int(k), _p(out))
    return out


def dilate_ellipse(g, k: int):
    g = _u8(g)
    out = np.empty_like(g)
    lib().svo_dilate_ellipse(_p(g), g.shape[0], g.shape[1], int(k), _p(out))
    return out


def erode_ellipse(g, k: int):
    g = _u8(g)
    out = np.empty_like(g)
    lib().svo_erode_ellipse(_p(g), g.shape[0], g.shape[1], int(k), _p(out))
    return out


def divide_normalize(g, bg):
    g, bg = _u8(g), _u8(bg)
    out = np.empty_like(g)
    lib().svo_divide_normalize(_p(g), _p(bg), C.c_long(g.size), _p(out))
    return out


def gaussian_blur_q8(g, n: int):
    g = _u8(g)
    out = np.empty_like(g)
    lib().svo_gaussian_blur_q8(_p(g), g.shape[0], g.shape[1], int(n), _p(out))
    return out


def clahe_frame(g, tiles: int = 8, clip: float = 2.0):
    g = _u8(g)
    out = np.empty_like(g)
    rc = lib().svo_clahe_frame(_p(g), g.shape[0], g.shape[1], int(tiles), C.c_double(clip), _p(out))
    assert rc == 0, "frame sides must divide by the CLAHE grid"
    return out


def morph_cleanup(m):
    m = _u8(m)
    out = np.empty_like(m)
    lib().svo_morph_cleanup(_p(m), m.shape[0], m.shape[1], _p(out))
    return out


def morph_ellipse_direct(g, k: int, dilate: bool):
    g = _u8(g)
    out = np.empty_like(g)
    lib().svo_morph_ellipse_direct(_p(g), g.shape[0], g.shape[1], int(k), int(dilate), _p(out))
    return out


def otsu_inv(g):
    """cv2.threshold(g, 0, 255, THRESH_BINARY_INV + THRESH_OTSU) -> (level, binary)"""
    g = _u8(g)
    out = np.empty_like(g)
    t = lib().svo_threshold_otsu_inv(_p(g), C.c_long(g.size), _p(out))
    return int(t), out


def sauvola(g, window: int = 25, k: float = 0.2):
    g = _u8(g)
    out = np.empty_like(g)
    lib().svo_threshold_sauvola(_p(g), g.shape[0], g.shape[1], int(window), C.c_float(k), _p(out))
    return out


def preprocess_v2(bgr, use_illumination_norm: bool = True, use_shadow_removal: bool = True):
    """cv/preprocess_v2.py:205-244 -> (mask, has_glare, has_shadow)."""
    bgr = _u8(bgr)
    H, W = bgr.shape[:2]
    out = np.empty((H, W), np.uint8)
    flags = (C.c_int * 2)()
    rc = lib().svo_preprocess_v2_opts(_p(bgr), H, W, int(use_illumination_norm), int(use_shadow_removal), _p(out), flags)
    assert rc == 0, "frame too small (minimum 32x32)"
    return out, bool(flags[0]), bool(flags[1])


METHODS = ("adaptive", "otsu", "sauvola")


def preprocess_multi(bgr):
    """cv/preprocess_v2.py:247-308 -> dict(binary, gray, enhanced, illumination_normalized, has_glare, has_shadow,
    method_used, otsu_level, scores)."""
    bgr = _u8(bgr)
    H, W = bgr.shape[:2]
    o = [np.empty((H, W), np.uint8) for _ in range(4)]
    info = (C.c_int * 4)()
    scores = (C.c_double * 3)()
    rc = lib().svo_preprocess_multi(_p(bgr), H, W, _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]), info, scores)
    assert rc == 0, "frame too small (minimum 32x32)"
    return dict(binary=o[0], gray=o[1], enhanced=o[2], illumination_normalized=o[3], has_glare=bool(info[0]),
                has_shadow=bool(info[1]), method_used=METHODS[info[2]], otsu_level=int(info[3]), scores=list(scores))


# ---- S1: solver/src/sudoku.c ------------------------------------------------------------------------------
def solve_sudoku(grid):
    """solve_sudoku (solver/src/sudoku.c:72) restated: grid (81,) or (9,9) uint8 -> (status, solution (same shape))."""
    g = _u8(grid)
    out = np.empty_like(g)
    lib().svo_solve_sudoku.restype = C.c_int
    st = lib().svo_solve_sudoku(_p(g), _p(out))
    return int(st), out


_ref_solver = None


def ref_solver_available() -> bool:
    return os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_solver.so"))


def ref_solve_sudoku(grid):
    """The REFERENCE's own solve_sudoku, compiled from its sources by oracle/Makefile (kind = "reference")."""
    global _ref_solver
    if _ref_solver is None:
        _ref_solver = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_solver.so"))
        _ref_solver.solve_sudoku.restype = C.c_int
    g = np.ascontiguousarray(np.asarray(grid).reshape(9, 9), dtype=np.int32)
    work = g.copy()
    st = _ref_solver.solve_sudoku(work.ctypes.data_as(C.c_void_p))
    return int(st), (work if st == 1 else g).astype(np.uint8).reshape(np.asarray(grid).shape)
