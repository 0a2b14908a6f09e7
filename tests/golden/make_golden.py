#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
cv2 4.13.0 + torch CPU) on small seeded inputs.  Run in the build container only; the fixtures are
committed because /root/reference does not travel to the GPU box.

    python tests/golden/make_golden.py

Every array named ref_* is an output of a reference function (file:line in the comment beside it).
"""
import hashlib
import os
import sys

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SVB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
sys.path.insert(0, os.path.join(REF, "pipeline"))  # run.py inserts cv/ and ml/ itself
import run as ref_run  # noqa: E402  (pipeline/run.py, unmodified)
from preprocess import grayscale, blur, threshold, preprocess_for_grid_detection  # noqa: E402
from grid import find_grid_contour, order_points, warp_perspective  # noqa: E402
from extract import extract_cells  # noqa: E402
from model import DigitCNN  # noqa: E402

from svb200 import frames as F  # noqa: E402
from svb200.api import load_digitcnn_weights  # noqa: E402

assert cv2.__version__ == "4.13.0", cv2.__version__


def ref_model():
    m = DigitCNN()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in load_digitcnn_weights().items()})
    return m.eval()


def full_path(img, model):
    out = {"bgr": img}
    out["ref_gray"] = grayscale(img)                              # cv/preprocess.py:15
    out["ref_blur"] = blur(out["ref_gray"], 5)                    # cv/preprocess.py:22
    out["ref_mask"] = threshold(out["ref_blur"], 11, 2)           # cv/preprocess.py:32
    assert np.array_equal(out["ref_mask"], preprocess_for_grid_detection(img))
    c = find_grid_contour(out["ref_mask"])                        # cv/grid.py:37
    out["ref_found"] = np.array(c is not None)
    if c is None:
        return out
    out["ref_corners"] = c.astype(np.int32)
    out["ref_ordered"] = order_points(c.astype(np.float32))       # cv/grid.py:74
    board = warp_perspective(img, c)                              # cv/grid.py:94
    out["ref_board_sha256"] = np.frombuffer(hashlib.sha256(board.tobytes()).digest(), np.uint8)
    out["ref_board_rows"] = board[::50].copy()                    # every 50th row, for debugging
    cells = extract_cells(board)                                  # cv/extract.py:13
    out["ref_cells_u8"] = np.stack(cells)
    prepped = [ref_run.preprocess_cell(x) for x in cells]         # pipeline/run.py:73
    out["ref_cells_thresh"] = np.stack(prepped)
    ts = [((torch.from_numpy(255 - p).float().unsqueeze(0).unsqueeze(0) / 255.0) - 0.5) / 0.5 for p in prepped]  # run.py:129-135
    with torch.no_grad():
        logits = torch.cat([model(t) for t in ts], 0)             # run.py:139-140 (batch 1, 81 times)
    out["ref_logits"] = logits.numpy()
    out["ref_digits"] = logits.argmax(1).numpy().astype(np.uint8)
    out["ref_conf"] = torch.softmax(logits, 1).max(1).values.numpy()
    return out


def main():
    model = ref_model()
    # 1. small synthetic frames (three seeds incl. one strongly rotated)
    for name, seed, rot, hw in (("frame_a", 4242, 15.0, (270, 480)), ("frame_b", 4343, 35.0, (360, 480)),
                                ("frame_c", 4444, 15.0, (272, 496))):
        f = F.make_frame(seed, hw[0], hw[1], rot)
        img = F.add_noise_host(f.image, seed)
        out = full_path(img, model)
        out["gt_digits"] = f.digits
        out["gt_corners"] = f.corners
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, img.shape, "found", bool(out["ref_found"]),
              "acc", None if not out["ref_found"] else float((out["ref_digits"].reshape(9, 9) == f.digits).mean()))
    # 2. a frame with no grid (reference returns None)
    rng = np.random.default_rng(5)
    img = rng.integers(60, 200, (200, 320, 3)).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "frame_none.npz"), **full_path(img, model))
    # 3. one of the reference's own photos, decimated 8x (keeps the fixture small)
    p = os.path.join(REF, "data", "test_images", "sample_4.jpg")
    img = cv2.imread(p)[::8, ::8].copy()
    out = full_path(img, model)
    np.savez_compressed(os.path.join(HERE, "photo4_dec8.npz"), **out)
    print("photo4_dec8", img.shape, "found", bool(out["ref_found"]))
    # 4. unit vectors: random images of awkward sizes through the three preprocess stages, and cells
    unit = {}
    for i, (h, w) in enumerate(((28, 28), (37, 53), (64, 80), (45, 66))):
        g = rng.integers(0, 256, (h, w)).astype(np.uint8)
        unit[f"g{i}"] = g
        unit[f"ref_blur{i}"] = blur(g, 5)
        unit[f"ref_thr_inv{i}"] = threshold(g, 11, 2)
        unit[f"ref_thr_bin{i}"] = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)  # run.py:87
    cells = np.stack([cv2.GaussianBlur(rng.integers(0, 256, (28, 28)).astype(np.uint8), (3, 3), 0) for _ in range(12)])
    unit["cells"] = cells
    unit["ref_cells_thresh"] = np.stack([ref_run.preprocess_cell(c) for c in cells])
    unit["ref_clahe"] = np.stack([cv2.createCLAHE(clipLimit=2.0, tileGridSize=(4, 4)).apply(c) for c in cells])  # run.py:80-81
    # the five photos' corners as cv2 4.13.0 + the reference produce them (recorded, inputs not stored)
    for k in range(1, 6):
        im = cv2.imread(os.path.join(REF, "data", "test_images", f"sample_{k}.jpg"))
        c = find_grid_contour(preprocess_for_grid_detection(im))
        unit[f"photo{k}_corners"] = np.zeros((0, 2), np.int32) if c is None else c.astype(np.int32)
    # v2 contour method (cv/grid_v2.py:102-128) on the golden masks and on quadrilaterals that fail its validity test
    sys.path.insert(0, os.path.join(REF, "cv"))
    import grid_v2 as ref_v2  # noqa: E402
    v2 = {}
    for name in ("frame_a", "frame_b", "frame_c", "photo4_dec8", "frame_none"):
        m = np.load(os.path.join(HERE, name + ".npz"))["ref_mask"]
        c = ref_v2.detect_grid_contour(m)
        v2[f"{name}_corners"] = np.zeros((0, 2), np.float32) if c is None else c
    shapes = []
    for k, pts in enumerate(([[20, 30], [200, 25], [210, 60], [15, 70]],          # flat: sides ratio > 2
                             [[30, 20], [220, 30], [120, 180], [20, 170]],         # sheared: an angle outside [45,135]
                             [[40, 40], [210, 35], [215, 200], [35, 205]])):       # fine
        m = np.zeros((240, 256), np.uint8)
        cv2.polylines(m, [np.array(pts, np.int32)], True, 255, 2)
        c = ref_v2.detect_grid_contour(m, 0.05)
        v2[f"shape{k}_mask"] = m
        v2[f"shape{k}_corners"] = np.zeros((0, 2), np.float32) if c is None else c
    np.savez_compressed(os.path.join(HERE, "v2.npz"), **v2)
    # DigitCNNv3 (ml/model_v3.py), eval mode, deterministic weights/inputs from tests/helpers/v3_weights.py
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    from v3_weights import make_v3_inputs, make_v3_state  # noqa: E402
    from model_v3 import DigitCNNv3  # noqa: E402  (reference, via run.py's sys.path entry for ml/)
    net = DigitCNNv3()
    missing = net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in make_v3_state().items()}, strict=True)
    net.eval()
    xin = torch.from_numpy(make_v3_inputs())
    with torch.no_grad():
        np.savez_compressed(os.path.join(HERE, "v3.npz"), ref_logits=net(xin).numpy(),
                            ref_features=net(xin, return_features=True).numpy(),
                            n_state_entries=np.array(len(net.state_dict())))
    np.savez_compressed(os.path.join(HERE, "unit.npz"), **unit)
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
