"""K1 alone on a device-resident batch of synthetic 1080p frames: ms per launch and GB/s of algorithmic bytes
(parity is the tests' job: tests/test_gpu_parity.py, tests/soak.py).
    python tools/k1_ab.py [frames] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sudoku-vision_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from svb200 import Scanner, load_digitcnn_weights
from svb200 import frames as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H, W = 1080, 1920
dev = torch.device("cuda", 0)
sc = Scanner(device=0, weights=load_digitcnn_weights())
clean = torch.from_numpy(np.stack([F.make_frame(31000 + i, H, W).image for i in range(8)])).to(dev)
batch = F.noisy_batch_device(clean, n, seed=7)
out = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
for _ in range(3):
    sc.preprocess(batch, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    sc.preprocess(batch, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
gbs = 4 * H * W * n / (ms * 1e-3) / 1e9
print(f"K1: {n} frames {ms:.3f} ms/launch  {gbs:.0f} GB/s  frac {gbs / 6454:.3f}")
