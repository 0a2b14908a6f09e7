#!/usr/bin/env python3
"""Generates tests/golden/jpeg.npz: JPEG files written by cv2.imwrite's encoder (libjpeg-turbo 3.1 inside the cv2 4.13
wheel — the decoder pipeline/run.py:250's cv2.imread uses) and the sha256 of what cv2.imdecode returns for each, so the
decoder parity test runs where cv2 is absent.  Also two files outside the supported subset (progressive) and a truncated
one.  Build container only.   python tests/golden/make_jpeg_golden.py"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
from svb200 import frames as F  # noqa: E402

assert cv2.__version__ == "4.13.0", cv2.__version__
rng = np.random.default_rng(3)
S420, S444 = cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444
cases = [("frame540_q90_r8_420", F.add_noise_host(F.make_frame(77, 540, 960).image, 5), 90, 8, S420),
         ("frame540_q95_r60_420", F.add_noise_host(F.make_frame(78, 540, 960).image, 6), 95, 60, S420),
         ("frame540_q75_r0_420", F.make_frame(79, 540, 960).image, 75, 0, S420),
         ("frame270_q90_r4_444", F.make_frame(80, 270, 480).image, 90, 4, S444),
         ("noise37x53_q100_r1_420", rng.integers(0, 256, (37, 53, 3)).astype(np.uint8), 100, 1, S420),
         ("noise200x3_q60_r0_420", rng.integers(0, 256, (200, 3, 3)).astype(np.uint8), 60, 0, S420),
         ("noise1x1_q90_r0_420", rng.integers(0, 256, (1, 1, 3)).astype(np.uint8), 90, 0, S420),
         ("gray64x80_q85_r3", rng.integers(0, 256, (64, 80)).astype(np.uint8), 85, 3, S420)]
out = {}
for name, img, q, rst, samp in cases:
    params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp]
    if rst:
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    ok, buf = cv2.imencode(".jpg", img, params)
    assert ok
    dec = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    out[name + "_file"] = buf.reshape(-1)
    out[name + "_shape"] = np.array(dec.shape, np.int32)
    out[name + "_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(dec).tobytes()).digest(), np.uint8)
    out[name + "_rows"] = dec[:: max(1, dec.shape[0] // 8)].copy()
ok, buf = cv2.imencode(".jpg", cases[3][1], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
out["unsupported_progressive_file"] = buf.reshape(-1)
out["malformed_truncated_file"] = out["frame270_q90_r4_444_file"][:300].copy()
np.savez_compressed(os.path.join(HERE, "jpeg.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith("_file")})
