"""Builds libsvb200.so in-tree with nvcc for sm_100a (no torch extension machinery: the boundary is
a plain C ABI, include/svb200.h).  Every .cu is compiled to its own object (in parallel, re-used while neither it nor
a header changed) and the objects are linked into the shared library.
Usage: python sudoku-vision_b200/build.py [--force] [--verbose]"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "svb200", "lib")
OBJ_DIR = os.path.join(HERE, "build")
OUT = os.path.join(OUT_DIR, "libsvb200.so")
SOURCES = ["c_api.cu", "preprocess.cu", "preprocess_v2.cu", "contour.cu", "contours_all.cu", "cells.cu", "digitcnn.cu", "digitcnn_tc.cu",
           "digitcnn_v3.cu", "digitcnn_v3_tc.cu", "solver.cu", "jpeg.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=true",  # exact-order arithmetic is spelled with __fmaf_rn/__fmul_rn/__fadd_rn, which never contract
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _headers() -> list:
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(HERE, "..", "include", "svb200.h")]


def _digest(src: str) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in [src] + _headers():
        h.update(open(p, "rb").read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "svb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool, verbose: bool) -> str:
    name = os.path.basename(src)[:-3]
    obj, stamp = os.path.join(OBJ_DIR, name + ".o"), os.path.join(OBJ_DIR, name + ".sha")
    dig = _digest(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed compiling {name}.cu")
    open(stamp, "w").write(dig)
    return obj


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with cf.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    r = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libsvb200.so")
    return OUT


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
