#!/usr/bin/env python3
"""Run the reference's UNCHANGED front ends (pipeline/run.py, pipeline/benchmark.py, cv/test_pipeline.py …)
on the B200 drop-in modules.

    python sudoku-vision_b200/dropin/launch.py --ref baseline/_ref/sudoku-vision pipeline/run.py IMG --no-display

The reference derives every path from its own location and does `sys.path.insert(0, <ref>/cv)` before
`from preprocess import …` (pipeline/run.py:28-35), so shadowing via PYTHONPATH cannot work.  Instead
this launcher pre-seeds `sys.modules` with our modules under the flat names the reference imports
(`preprocess`, `grid`, `extract`, `model`) and the packaged names (`cv`, `cv.preprocess`, …), then
executes the reference script with runpy.  The script's source stays byte-identical.

`--stage SRC` first copies a read-only reference checkout to --ref (git-ignored baseline/_ref/),
writes the DigitCNN fixture weights where run.py:103-108 looks for them (ml/digit_cnn_v2.pt) and
builds solver/sudoku_solver with make.
"""
import argparse
import importlib
import os
import runpy
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)


def stage(src: str, dst: str) -> None:
    import torch

    if not os.path.exists(dst):
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns(".git", "ios", "web", "node_modules"))
    sys.path.insert(0, PKG)
    from svb200 import load_digitcnn_weights

    sd = {k: torch.from_numpy(v) for k, v in load_digitcnn_weights().items()}
    torch.save(sd, os.path.join(dst, "ml", "digit_cnn_v2.pt"))
    subprocess.check_call(["make", "-C", os.path.join(dst, "solver"), "-s"])


def seed_modules() -> None:
    for p in (HERE, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    cv_pkg = importlib.import_module("cv")                  # dropin/cv/__init__.py
    ml_model = importlib.import_module("ml.model") if os.path.exists(os.path.join(HERE, "ml", "__init__.py")) else None
    if ml_model is None:
        sys.path.insert(0, os.path.join(HERE, "ml"))
        ml_model = importlib.import_module("model")
    sys.modules["preprocess"] = importlib.import_module("cv.preprocess")
    sys.modules["grid"] = importlib.import_module("cv.grid")
    sys.modules["extract"] = importlib.import_module("cv.extract")
    sys.modules["model"] = ml_model
    # v2 front ends (pipeline/run_v2.py:37-39, cv/grid_v2.py:534): preprocessing, detection method 1, the quality gate
    # and the classifier run on the GPU
    sys.modules["preprocess_v2"] = importlib.import_module("cv.preprocess_v2")
    sys.modules["grid_v2"] = importlib.import_module("cv.grid_v2")
    sys.modules["grid_quality"] = importlib.import_module("cv.grid_quality")
    try:
        sys.modules["model_v3"] = importlib.import_module("ml.model_v3")
    except ImportError:
        sys.modules["model_v3"] = importlib.import_module("model_v3")
    sys.modules["cv"] = cv_pkg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True, help="writable staged copy of the reference tree")
    ap.add_argument("--stage", default=None, help="copy this reference checkout to --ref first")
    ap.add_argument("--stage-only", action="store_true", help="stop after staging (works without a GPU)")
    ap.add_argument("script", nargs="?", default=None, help="reference script relative to --ref, e.g. pipeline/run.py")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    ref = os.path.abspath(a.ref)
    if a.stage:
        stage(a.stage, ref)
    if a.stage_only or not a.script:
        return
    seed_modules()
    script = os.path.join(ref, a.script)
    sys.argv = [script] + a.args
    os.chdir(os.path.dirname(script))
    sys.path.insert(0, os.path.dirname(script))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
