#!/usr/bin/env python3
"""Generates tests/golden/v2pre.npz: the UNMODIFIED reference cv/preprocess_v2.py (imported from /root/reference,
cv2 4.13.0) on small seeded frames — plain, shadowed (has_shadow -> remove_shadow branch) and with a glare patch.
Run in the build container only:  python tests/golden/make_v2pre_golden.py"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SVB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
sys.path.insert(0, os.path.join(REF, "cv"))
import preprocess_v2 as R  # noqa: E402  (cv/preprocess_v2.py, unmodified)

from svb200 import frames as F  # noqa: E402

assert cv2.__version__ == "4.13.0", cv2.__version__


def variants(seed, h, w):
    img = F.add_noise_host(F.make_frame(seed, h, w, 12.0).image, seed)
    sh = img.copy()
    sh[:, : w // 3] = (sh[:, : w // 3] * 0.45).astype(np.uint8)
    gl = img.copy()
    gl[h // 4: h // 2, w // 4: w // 2] = 255
    flat = (img.astype(np.float32) * 0.25 + 120).astype(np.uint8)  # low contrast: has_shadow False at this size
    return {"plain": img, "shadow": sh, "glare": gl, "flat": flat}


def main():
    out = {}
    for tag, (seed, h, w) in {"a": (5151, 272, 480), "b": (5252, 360, 512)}.items():
        for name, img in variants(seed, h, w).items():
            k = f"{tag}_{name}"
            out[k + "_bgr"] = img
            out[k + "_ref_mask"] = R.preprocess_for_grid_detection(img)                   # cv/preprocess_v2.py:205
            out[k + "_ref_mask_noillum"] = R.preprocess_for_grid_detection(img, False, True)
            out[k + "_ref_mask_noshadow"] = R.preprocess_for_grid_detection(img, True, False)
            r = R.preprocess_multi_strategy(img)                                           # cv/preprocess_v2.py:247
            out[k + "_ref_binary"] = r.binary
            out[k + "_ref_enhanced"] = r.enhanced
            out[k + "_ref_illum"] = r.illumination_normalized
            out[k + "_ref_gray"] = r.gray
            out[k + "_ref_flags"] = np.array([r.has_glare, r.has_shadow])
            out[k + "_ref_method"] = np.array(r.method_used)
            bl = cv2.GaussianBlur(r.enhanced, (5, 5), 0)
            out[k + "_ref_otsu"] = R.threshold_otsu(bl)                                    # cv/preprocess_v2.py:146
            out[k + "_ref_sauvola"] = R.threshold_sauvola(bl)                              # cv/preprocess_v2.py:152
            print(k, img.shape, r.method_used, r.has_glare, r.has_shadow)
    # unit vectors for the primitives
    rng = np.random.default_rng(9)
    g = cv2.GaussianBlur(rng.integers(0, 256, (96, 136)).astype(np.uint8), (5, 5), 0)
    out["u_gray"] = g
    out["u_ref_blur13"] = cv2.blur(g, (13, 13))                                            # :89
    out["u_ref_dilate7"] = cv2.dilate(g, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7)))       # :108-109
    out["u_ref_close51"] = cv2.morphologyEx(g, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (51, 51)))  # :51-52
    out["u_ref_gauss21"] = cv2.GaussianBlur(g, (21, 21), 0)                                # :112
    out["u_ref_clahe8"] = R.apply_clahe(g)                                                 # :122
    out["u_ref_illum"] = R.normalize_illumination(g)                                       # :40
    out["u_ref_noshadow"] = R.remove_shadow(g)                                             # :105
    m = ((rng.random((96, 136)) < 0.3) * 255).astype(np.uint8)
    out["u_mask"] = m
    out["u_ref_cleanup"] = R.morphological_cleanup(m)                                      # :178
    for k in (7, 21, 51, 193, 385):
        out[f"u_ellipse{k}"] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    p = os.path.join(HERE, "v2pre.npz")
    np.savez_compressed(p, **out)
    print(p, os.path.getsize(p))


if __name__ == "__main__":
    main()
