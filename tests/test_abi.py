"""CPU tier: libsvb200.so loads without a GPU and exports exactly the symbols include/svb200.h declares;
the ctypes table in svb200/_lib.py covers every one of them.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def header_symbols():
    src = open(os.path.join(ROOT, "include", "svb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
    import build

    return build.build_lib()


def test_header_declares_the_path(header_symbols):
    for s in ("svb_preprocess_v1", "svb_find_grid_contour", "svb_warp_perspective", "svb_extract_cells",
              "svb_cell_prep", "svb_cells_from_frames", "svb_digitcnn_load", "svb_digitcnn_forward",
              "svb_scan_batch_v1", "svb_scan_batch_v1_host"):
        assert s in header_symbols


def test_library_exports_every_declared_symbol(header_symbols, lib_path):
    lib = C.CDLL(lib_path)
    for s in header_symbols:
        assert hasattr(lib, s), f"{s} declared in svb200.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = sorted(set(re.findall(r"\bT (svb_[a-z0-9_]+)", out)))
    assert exported == header_symbols, "exported svb_* symbols differ from the header"


def test_ctypes_table_matches_header(header_symbols, lib_path):
    from svb200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols
    lib = _lib.load()
    assert lib.svb_abi_version() == 1


def test_library_is_sm100a_cuda(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from svb200 import Scanner, _lib

    with pytest.raises(_lib.SvbError):
        Scanner()


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (or cv2) anywhere."""
    pkg = os.path.join(ROOT, "sudoku-vision_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            p = os.path.join(d, f)
            txt = open(p).read()
            if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "svb_oracle" in txt and f != "frames.py" and "oracle/svb_oracle.c" not in txt:
                bad.append(p)
            if f != "frames.py" and re.search(r"^\s*import cv2|^\s*from cv2", txt, flags=re.M):
                bad.append(p)
    assert not bad, bad
