#!/usr/bin/env python3
"""Runs the UNMODIFIED reference pipeline (pipeline/run.py:run_pipeline, cv2 + torch CPU) from the staged copy
baseline/_ref/sudoku-vision on its own five photos and records what it recognised -> tests/golden/photos_run.npz.
Build container only (needs /root/reference staged by sudoku-vision_b200/dropin/launch.py --stage ... --stage-only)."""
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = HERE.parent.parent / "baseline" / "_ref" / "sudoku-vision"
sys.path.insert(0, str(REF / "pipeline"))
import run as ref_run  # noqa: E402  (unmodified reference; inserts its own cv/ and ml/ into sys.path)

out = {}
for k in range(1, 6):
    res = ref_run.run_pipeline(REF / "data" / "test_images" / f"sample_{k}.jpg")
    found = res.warped_grid is not None
    out[f"s{k}_found"] = np.array(found)
    if found:
        out[f"s{k}_grid"] = np.array(res.recognized_grid, np.uint8)
        out[f"s{k}_conf"] = np.array([p.confidence for p in res.predictions], np.float32)
        out[f"s{k}_warp_rows"] = res.warped_grid[::90].copy()
    print(k, found, res.error)
np.savez_compressed(HERE / "photos_run.npz", **out)
