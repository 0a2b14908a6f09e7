import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sudoku-vision_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


@pytest.fixture(scope="session")
def contour_host():
    """The product's contour core (csrc/contour_core.cuh) compiled for the host: tests/helpers."""
    import ctypes as C

    src = os.path.join(ROOT, "tests", "helpers", "contour_host.cpp")
    out_dir = os.path.join(ROOT, "tests", "helpers", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcontour_host.so")
    deps = [os.path.join(PKG, "csrc", "contour_core.cuh"), os.path.join(PKG, "csrc", "contours_all_core.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max([os.path.getmtime(src)] + [os.path.getmtime(d) for d in deps]):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src])
    lib = C.CDLL(so)

    def find(mask, min_area_ratio=0.1, eps_ratio=0.02, use_bits=False, v2=False):
        mask = np.ascontiguousarray(mask, np.uint8)
        c = np.zeros((4, 2), np.int32)
        f = lib.svbh_find_grid_contour(mask.ctypes.data_as(C.c_void_p), mask.shape[0], mask.shape[1],
                                       C.c_double(min_area_ratio), C.c_double(eps_ratio),
                                       c.ctypes.data_as(C.c_void_p), None, None, int(use_bits) | (2 if v2 else 0))
        return f, c

    def find_all(mask):
        """cv/grid.py:16-21 through the product's contours_all core: list of (n,2) int32 arrays, cv2's order"""
        mask = np.ascontiguousarray(mask, np.uint8)
        max_pts, max_c = int(mask.size) + 16, int(mask.size) // 2 + 16
        pts = np.empty((max_pts, 2), np.int32)
        offs = np.empty(max_c + 1, np.int64)
        rounds = C.c_int(0)
        lib.svbh_find_contours.restype = C.c_longlong
        n = lib.svbh_find_contours(mask.ctypes.data_as(C.c_void_p), mask.shape[0], mask.shape[1], pts.ctypes.data_as(C.c_void_p),
                                   C.c_longlong(max_pts), offs.ctypes.data_as(C.c_void_p), C.c_longlong(max_c), C.byref(rounds))
        assert n >= 0, n
        return [pts[offs[i]:offs[i + 1]].copy() for i in range(n)], rounds.value

    find.find_all = find_all
    return find


@pytest.fixture(scope="session")
def cells_host():
    """The product's per-cell pipeline (csrc/cells_core.cuh) compiled for the host: tests/helpers/cells_host.cpp."""
    import ctypes as C

    src = os.path.join(ROOT, "tests", "helpers", "cells_host.cpp")
    out_dir = os.path.join(ROOT, "tests", "helpers", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcells_host.so")
    dep = os.path.join(PKG, "csrc", "cells_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src])
    lib = C.CDLL(so)
    lib.svbh_map_check.restype = C.c_long

    class H:
        @staticmethod
        def cells_from_frame(bgr, corners):
            bgr = np.ascontiguousarray(bgr, np.uint8)
            c = np.ascontiguousarray(corners, np.int32).reshape(4, 2)
            u8 = np.empty((81, 28, 28), np.uint8)
            bits = np.empty((81, 28), np.uint32)
            pm1 = np.empty((81, 28, 28), np.float32)
            lib.svbh_cells_from_frame(bgr.ctypes.data_as(C.c_void_p), bgr.shape[0], bgr.shape[1], c.ctypes.data_as(C.c_void_p),
                                      u8.ctypes.data_as(C.c_void_p), bits.ctypes.data_as(C.c_void_p), pm1.ctypes.data_as(C.c_void_p))
            return u8, bits, pm1

        @staticmethod
        def cell_prep(cells):
            cells = np.ascontiguousarray(cells, np.uint8).reshape(-1, 28, 28)
            n = len(cells)
            thr = np.empty((n, 28, 28), np.uint8)
            bits = np.empty((n, 28), np.uint32)
            pm1 = np.empty((n, 28, 28), np.float32)
            lib.svbh_cell_prep(cells.ctypes.data_as(C.c_void_p), n, thr.ctypes.data_as(C.c_void_p), pm1.ctypes.data_as(C.c_void_p),
                               bits.ctypes.data_as(C.c_void_p))
            return thr, bits, pm1

        @staticmethod
        def map_check(corners):
            c = np.ascontiguousarray(corners, np.int32).reshape(4, 2)
            ex = C.c_long(0)
            bad = lib.svbh_map_check(c.ctypes.data_as(C.c_void_p), C.byref(ex))
            return int(bad), int(ex.value)

    H.lib = lib
    return H


def bits_to_pm1(bits):
    """bit rows (..., 28) -> +-1 float cells (..., 28, 28): bit x of row y set <=> +1"""
    b = (np.asarray(bits).astype(np.uint32)[..., None] >> np.arange(28, dtype=np.uint32)) & 1
    return np.where(b == 1, 1.0, -1.0).astype(np.float32)


@pytest.fixture(scope="session")
def scanner():
    if not has_cuda():
        pytest.skip("no CUDA device")
    from svb200 import Scanner, load_digitcnn_weights

    return Scanner(weights=load_digitcnn_weights())


@pytest.fixture(scope="session")
def weights():
    from svb200 import load_digitcnn_weights

    return load_digitcnn_weights()
