"""Library-level port of the reference scan path (cv2 + torch CPU).  TEST INFRASTRUCTURE ONLY.

The reference's hot path is a few lines of Python per stage that call OpenCV and PyTorch
(cv/preprocess.py:57-65, cv/grid.py:37-133, cv/extract.py:13-56, pipeline/run.py:73-152,
ml/model.py:19-42).  /root/reference does not exist on the GPU box, so bench.py's CPU legs
(`cpu_baseline`, `--impl reference`) and the cv2 cross-checks in tests/ use this port: the same
library calls, in the same order, with the same arguments, written as one class so that the model
is loaded once (run.py:301 reloads it per image; the hoisted variant is the fair baseline,
BASELINE.md §4).  It is the closest thing to "the reference's own CPU implementation" that can
travel; svb_oracle.c is the closed-form restatement of the arithmetic inside these calls.

Never imported by the product package.
"""
from __future__ import annotations

import numpy as np


def _cv2():
    import cv2

    return cv2


class TorchDigitCNN:
    """ml/model.py:19-42 rebuilt from torch.nn primitives (eval mode)."""

    def __init__(self, state_dict: dict):
        import torch
        import torch.nn as nn

        class Net(nn.Module):
            def __init__(self):
                super().__init__()
                self.conv1 = nn.Conv2d(1, 32, 3, padding=1)
                self.conv2 = nn.Conv2d(32, 64, 3, padding=1)
                self.fc1 = nn.Linear(64 * 7 * 7, 128)
                self.fc2 = nn.Linear(128, 10)

            def forward(self, x):
                F = torch.nn.functional
                x = F.max_pool2d(F.relu(self.conv1(x)), 2, 2)
                x = F.max_pool2d(F.relu(self.conv2(x)), 2, 2)
                x = x.view(x.size(0), -1)
                return self.fc2(F.relu(self.fc1(x)))

        self.torch = torch
        self.net = Net().eval()
        self.net.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()})

    def __call__(self, x):
        with self.torch.no_grad():
            return self.net(x)


def preprocess_for_grid_detection(image):
    cv2 = _cv2()
    g = image if image.ndim == 2 else cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    g = cv2.GaussianBlur(g, (5, 5), 0)
    return cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 11, 2)


def find_grid_contour(binary, min_area_ratio: float = 0.1):
    cv2 = _cv2()
    cs, _ = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not cs:
        return None
    floor = min_area_ratio * binary.shape[0] * binary.shape[1]
    for c in sorted(cs, key=cv2.contourArea, reverse=True):
        if cv2.contourArea(c) < floor:
            break
        poly = cv2.approxPolyDP(c, 0.02 * cv2.arcLength(c, True), True)
        if len(poly) == 4:
            return poly.reshape(4, 2)
    return None


def order_points(pts):
    out = np.zeros((4, 2), np.float32)
    s = pts.sum(axis=1)
    d = np.diff(pts, axis=1)
    out[0], out[2] = pts[np.argmin(s)], pts[np.argmax(s)]
    out[1], out[3] = pts[np.argmin(d)], pts[np.argmax(d)]
    return out


def warp_perspective(image, corners, output_size: int = 450):
    cv2 = _cv2()
    src = order_points(corners.astype(np.float32))
    e = output_size - 1
    dst = np.array([[0, 0], [e, 0], [e, e], [0, e]], np.float32)
    return cv2.warpPerspective(image, cv2.getPerspectiveTransform(src, dst), (output_size, output_size))


def extract_cells(board, cell_size: int = 28, margin_ratio: float = 0.1):
    cv2 = _cv2()
    ch, cw = board.shape[0] // 9, board.shape[1] // 9
    mh, mw = int(ch * margin_ratio), int(cw * margin_ratio)
    out = []
    for r in range(9):
        for c in range(9):
            cell = board[r * ch + mh:(r + 1) * ch - mh, c * cw + mw:(c + 1) * cw - mw]
            if cell.ndim == 3:
                cell = cv2.cvtColor(cell, cv2.COLOR_BGR2GRAY)
            out.append(cv2.resize(cell, (cell_size, cell_size)))
    return out


def preprocess_cell(cell):
    cv2 = _cv2()
    cell = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(4, 4)).apply(cell)
    return cv2.adaptiveThreshold(cell, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)


class RefScanner:
    """run.py:257-318 (CV + ML sections) with the model hoisted out of the per-image loop."""

    def __init__(self, state_dict: dict, batch_cells: bool = False):
        self.model = TorchDigitCNN(state_dict)
        self.batch_cells = batch_cells  # False = 81 batch-1 forwards, as run.py:122-150 does

    def scan(self, image):
        import torch

        binary = preprocess_for_grid_detection(image)
        corners = find_grid_contour(binary)
        if corners is None:
            return dict(found=False, mask=binary)
        board = warp_perspective(image, corners)
        cells = extract_cells(board)
        prepped = [255 - preprocess_cell(c) for c in cells]
        ts = [(torch.from_numpy(p).float().unsqueeze(0).unsqueeze(0) / 255.0 - 0.5) / 0.5 for p in prepped]
        if self.batch_cells:
            logits = self.model(torch.cat(ts, 0))
        else:
            logits = torch.cat([self.model(t) for t in ts], 0)
        probs = torch.softmax(logits, dim=1)
        digits = logits.argmax(dim=1)
        conf = probs[torch.arange(81), digits]
        return dict(found=True, mask=binary, corners=corners, board=board,
                    cells_u8=np.stack(cells), cells_in=np.stack(prepped),
                    logits=logits.numpy(), digits=digits.numpy().astype(np.uint8), conf=conf.numpy())
