#!/usr/bin/env python3
"""Generates tests/golden/photos_v2.npz: the CV section of the UNMODIFIED pipeline/run_v2.py (:276-318) on the five
reference photos — preprocess_multi_strategy, detect_grid, assess_grid_quality — imported from /root/reference
(cv2 4.13.0).  Minutes of CPU time (the 365-px elliptical close, twice per photo).  Build container only.

    python tests/golden/make_photos_v2_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SVB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "cv"))
from preprocess_v2 import preprocess_multi_strategy  # noqa: E402
from grid_v2 import detect_grid, warp_perspective  # noqa: E402
from grid_quality import assess_grid_quality  # noqa: E402

assert cv2.__version__ == "4.13.0", cv2.__version__
out = {}
for k in range(1, 6):
    img = cv2.imread(os.path.join(REF, "data", "test_images", f"sample_{k}.jpg"))
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    pre = preprocess_multi_strategy(img)                 # run_v2.py:279
    det = detect_grid(pre.binary, gray)                  # run_v2.py:287
    out[f"s{k}_method_used"] = np.array(pre.method_used)
    out[f"s{k}_flags"] = np.array([pre.has_glare, pre.has_shadow], np.uint8)
    out[f"s{k}_white"] = np.int64(np.count_nonzero(pre.binary))
    out[f"s{k}_found"] = np.uint8(det.corners is not None)
    out[f"s{k}_method"] = np.array(det.method if det.corners is not None else "none")
    if det.corners is not None:
        out[f"s{k}_corners"] = np.asarray(det.corners, np.float32)
        q = assess_grid_quality(img, pre.binary, det.corners)   # run_v2.py:301
        out[f"s{k}_quality"] = np.array([q.overall, q.sharpness, q.contrast, q.completeness, q.geometry, q.size], np.float64)
        out[f"s{k}_warp_rows"] = warp_perspective(img, det.corners)[::90]   # run_v2.py:314
    print(k, out[f"s{k}_method_used"], out[f"s{k}_method"], out[f"s{k}_flags"], flush=True)
np.savez_compressed(os.path.join(HERE, "photos_v2.npz"), **out)
