#!/usr/bin/env python3
"""Train DigitCNN weights for the benchmark/parity fixtures (host tooling, run once, output committed).

No PyTorch weights ship with the reference (.gitignore:14-16), so the measurement needs its own.
Recipe = the reference's (ml/train.py:305-312: Adam lr 1e-3, weight_decay 1e-4, CrossEntropy,
batch 64) on cells that went through the reference's inference-time cell prep
(ml/datasets.py:18-46 == pipeline/run.py:73-95 + invert), here harvested from our synthetic frames
through the cv2 port (oracle/ref_port.py) so that the classifier is accurate on the bench frames.
Output: sudoku-vision_b200/svb200/weights/digitcnn_synth.npz (state_dict keys of ml/model.py).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))

from oracle import ref_port  # noqa: E402
from svb200 import frames as F  # noqa: E402


def harvest(n_frames, seed0, h, w):
    xs, ys = [], []
    for i in range(n_frames):
        f = F.make_frame(seed0 + i, h, w)
        img = F.add_noise_host(f.image, seed0 + i)
        m = ref_port.preprocess_for_grid_detection(img)
        c = ref_port.find_grid_contour(m)
        if c is None:
            continue
        o = ref_port.order_points(c.astype(np.float32))
        if np.abs(o - f.corners).max() > 4:
            continue
        cells = ref_port.extract_cells(ref_port.warp_perspective(img, c))
        xs.append(np.stack([255 - ref_port.preprocess_cell(x) for x in cells]))
        ys.append(f.digits.reshape(-1))
    x = np.concatenate(xs).astype(np.float32) / 255.0
    return (x - 0.5) / 0.5, np.concatenate(ys).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--out", default=os.path.join(ROOT, "sudoku-vision_b200", "svb200", "weights", "digitcnn_synth.npz"))
    a = ap.parse_args()
    torch.manual_seed(a.seed)
    np.random.seed(a.seed)
    x, y = harvest(a.frames, 500_000, 1080, 1920)
    xv, yv = harvest(24, 900_000, 1080, 1920)
    print("train cells", x.shape, "val cells", xv.shape, "class hist", np.bincount(y))
    init = ref_port.TorchDigitCNN.__new__(ref_port.TorchDigitCNN)
    ref_port.TorchDigitCNN.__init__(init, {k: v.numpy() for k, v in _fresh_state().items()})
    net = init.net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
    lossf = nn.CrossEntropyLoss()
    X, Y = torch.from_numpy(x).unsqueeze(1), torch.from_numpy(y)
    XV, YV = torch.from_numpy(xv).unsqueeze(1), torch.from_numpy(yv)
    drop = nn.Dropout(0.5)
    for ep in range(a.epochs):
        perm = torch.randperm(len(X))
        tot = 0.0
        for s in range(0, len(X), 64):
            idx = perm[s:s + 64]
            xb = X[idx]
            Fn = torch.nn.functional
            h = Fn.max_pool2d(Fn.relu(net.conv1(xb)), 2, 2)
            h = Fn.max_pool2d(Fn.relu(net.conv2(h)), 2, 2)
            h = drop(Fn.relu(net.fc1(h.view(h.size(0), -1))))
            loss = lossf(net.fc2(h), Y[idx])
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += loss.item() * len(idx)
        net.eval()
        with torch.no_grad():
            acc = (net(XV).argmax(1) == YV).float().mean().item()
        net.train()
        print(f"epoch {ep + 1}: loss {tot / len(X):.4f}  val acc {acc:.4f}")
    sd = {k: v.detach().numpy().astype(np.float32) for k, v in net.state_dict().items()}
    np.savez_compressed(a.out, **sd)
    print("saved", a.out, os.path.getsize(a.out))


def _fresh_state():
    m = nn.ModuleDict(dict(conv1=nn.Conv2d(1, 32, 3, padding=1), conv2=nn.Conv2d(32, 64, 3, padding=1),
                           fc1=nn.Linear(3136, 128), fc2=nn.Linear(128, 10)))
    return {k: v.detach() for k, v in m.state_dict().items()}


if __name__ == "__main__":
    main()
