#!/usr/bin/env python3
"""Generates tests/golden/quality.npz: the UNMODIFIED reference cv/grid_quality.py (imported from /root/reference, cv2
4.13.0) on the v2pre golden frames, with corners from the reference's own detect_grid_contour.
Run in the build container only:  python tests/golden/make_quality_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SVB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "cv"))
import grid_quality as Q  # noqa: E402  (cv/grid_quality.py, unmodified)
import grid_v2 as G  # noqa: E402

FIELDS = ("overall", "sharpness", "contrast", "completeness", "geometry", "size")
v = np.load(os.path.join(HERE, "v2pre.npz"))
out = {}
for case in [f"{t}_{k}" for t in ("a", "b") for k in ("plain", "shadow", "glare", "flat")]:
    img, binary = v[case + "_bgr"], v[case + "_ref_binary"]
    c = G.detect_grid_contour(binary)
    if c is None:
        h, w = binary.shape  # no grid in this mask: a synthetic skewed quadrilateral keeps the case useful
        c = np.array([[w * 0.2, h * 0.15], [w * 0.8, h * 0.2], [w * 0.75, h * 0.9], [w * 0.25, h * 0.8]], np.float32).round()
    q = Q.assess_grid_quality(img, binary, c)                       # cv/grid_quality.py:228
    out[case + "_corners"] = np.asarray(c, np.float32)
    out[case + "_ref_scores"] = np.array([getattr(q, f) for f in FIELDS], np.float64)
    out[case + "_ref_feedback"] = np.array(Q.get_user_feedback(q))
    print(case, np.round(out[case + "_ref_scores"], 3), q.issues)
np.savez_compressed(os.path.join(HERE, "quality.npz"), **out)
