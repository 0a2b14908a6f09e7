#!/usr/bin/env python3
"""Generates tests/golden/helpers.npz and sudoku-vision_b200/svb200/weights/digitcnn_coreml.npz from the UNMODIFIED reference
(/root/reference, cv2 4.13.0).  Build container only; the fixtures are committed.

    python tests/golden/make_helpers_golden.py

helpers.npz — the stand-alone helpers the reference's debug callers use (cv/test_pipeline.py:21-23):
  * cv/grid.py:16-21 find_contours on a seeded synthetic frame's mask and on sample_4.jpg decimated 4x
    (contour count, total points, sha256 of the concatenated points in order, first 40 contours in full);
  * cv/grid.py:24-34 approximate_polygon of the 12 largest contours of each mask at three ratios;
  * cv/extract.py:59-79 is_cell_empty of the 81 cells of a synthetic frame and of seeded random cells.
digitcnn_coreml.npz — the only trained weights the reference ships: the fp16 DigitCNN tensors inside
ios/SudokuVision/Resources/DigitClassifier.mlpackage/Data/com.apple.CoreML/weights/weight.bin, decoded at the offsets the
survey lists, stored as float16 under the state_dict keys of ml/model.py.  BASELINE.md §4(ii) names them as the
brittle-margin stress set.
"""
import hashlib
import os
import struct
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SVB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
sys.path.insert(0, os.path.join(REF, "cv"))
from preprocess import preprocess_for_grid_detection  # noqa: E402
from grid import find_contours, approximate_polygon, find_grid_contour, warp_perspective  # noqa: E402
from extract import extract_cells, is_cell_empty  # noqa: E402

from svb200 import frames as F  # noqa: E402

assert cv2.__version__ == "4.13.0", cv2.__version__
out = {}


def contour_block(tag, mask):
    cs = find_contours(mask)
    allp = np.concatenate([c.reshape(-1, 2) for c in cs]).astype(np.int32) if cs else np.zeros((0, 2), np.int32)
    out[f"{tag}_mask"] = np.packbits(mask > 0, axis=1)
    out[f"{tag}_shape"] = np.array(mask.shape, np.int32)
    out[f"{tag}_n"] = np.int64(len(cs))
    out[f"{tag}_npts"] = np.int64(len(allp))
    out[f"{tag}_sha"] = np.frombuffer(hashlib.sha256(allp.tobytes()).digest(), np.uint8)
    out[f"{tag}_lens"] = np.array([len(c) for c in cs], np.int32)
    for i, c in enumerate(cs[:40]):
        out[f"{tag}_c{i}"] = c.reshape(-1, 2).astype(np.int32)
    big = sorted(range(len(cs)), key=lambda i: -cv2.contourArea(cs[i]))[:12]
    out[f"{tag}_big"] = np.array(big, np.int32)
    for j, i in enumerate(big):
        out[f"{tag}_bigc{j}"] = cs[i].reshape(-1, 2).astype(np.int32)
        for r in (0.005, 0.02, 0.05):
            out[f"{tag}_poly{j}_{r}"] = approximate_polygon(cs[i], r).reshape(-1, 2).astype(np.int32)


fr = F.make_frame(4242, 540, 960)
m = preprocess_for_grid_detection(fr.image)
contour_block("syn", m)
img = cv2.imread(os.path.join(REF, "data", "test_images", "sample_4.jpg"))
small = np.ascontiguousarray(img[::4, ::4])
ms = preprocess_for_grid_detection(small)
contour_block("photo", ms)

corners = find_grid_contour(m)
cells = extract_cells(warp_perspective(fr.image, corners))
out["cells"] = np.stack(cells)
out["cells_empty"] = np.array([is_cell_empty(c) for c in cells], np.uint8)
rng = np.random.default_rng(5)
rc = []
for k in range(64):
    c = np.full((28, 28), int(rng.integers(120, 250)), np.uint8)
    if k % 2:
        cv2.putText(c, str(k % 10), (6, 22), cv2.FONT_HERSHEY_SIMPLEX, 0.8, int(rng.integers(0, 90)), 2)
    c = np.clip(c.astype(np.int32) + rng.integers(-6, 7, c.shape), 0, 255).astype(np.uint8)
    rc.append(c)
out["rcells"] = np.stack(rc)
out["rcells_empty"] = np.array([is_cell_empty(c) for c in rc], np.uint8)
out["rcells_empty_t10"] = np.array([is_cell_empty(c, 0.10) for c in rc], np.uint8)
np.savez_compressed(os.path.join(HERE, "helpers.npz"), **out)
print("helpers.npz:", int(out["syn_n"]), "+", int(out["photo_n"]), "contours;", int(out["cells_empty"].sum()), "empty cells")

# ---- CoreML-recovered DigitCNN weights ---------------------------------------------------------------------------
wb = os.path.join(REF, "ios", "SudokuVision", "Resources", "DigitClassifier.mlpackage", "Data", "com.apple.CoreML", "weights",
                  "weight.bin")
raw = open(wb, "rb").read()
hdrs = (0x40, 0x2c0, 0x340, 0x9380, 0x9440, 0xcd480, 0xcd5c0, 0xce000)
keys = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")
shapes = ((32, 1, 3, 3), (32,), (64, 32, 3, 3), (64,), (128, 3136), (128,), (10, 128), (10,))
sd = {}
for h, k, shp in zip(hdrs, keys, shapes):
    magic, dtype, nbytes, off = struct.unpack_from("<IIQQ", raw, h)
    assert magic == 0xDEADBEEF and dtype == 1 and nbytes == 2 * int(np.prod(shp)), (hex(magic), dtype, nbytes, shp)
    sd[k] = np.frombuffer(raw, np.float16, int(np.prod(shp)), off).reshape(shp).copy()
np.savez_compressed(os.path.join(ROOT, "sudoku-vision_b200", "svb200", "weights", "digitcnn_coreml.npz"), **sd)
print("digitcnn_coreml.npz:", sum(v.size for v in sd.values()), "values")
