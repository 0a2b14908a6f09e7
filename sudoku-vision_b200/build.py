"""Builds libsvb200.so in-tree with nvcc for sm_100a (no torch extension machinery: the boundary is
a plain C ABI, include/svb200.h).  Usage: python sudoku-vision_b200/build.py [--force] [--verbose]"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "svb200", "lib")
OUT = os.path.join(OUT_DIR, "libsvb200.so")
SOURCES = ["c_api.cu", "preprocess.cu", "preprocess_v2.cu", "contour.cu", "cells.cu", "digitcnn.cu", "digitcnn_tc.cu", "digitcnn_v3.cu", "digitcnn_v3_tc.cu", "solver.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=true",  # exact-order arithmetic is spelled with __fmaf_rn/__fmul_rn/__fadd_rn, which never contract
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "svb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libsvb200.so")
    return OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
