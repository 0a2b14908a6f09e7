"""One warm launch + a few launches of a single cv/preprocess_v2.py stage, for ncu.  python tools/prof_v2_one.py OP [frames] [arg]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
import numpy as np
import torch
from svb200 import Scanner
from svb200 import frames as F

op = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
arg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sc = Scanner()
clean = torch.from_numpy(np.stack([F.make_frame(31000 + i, 1080, 1920).image for i in range(4)])).cuda()
batch = F.noisy_batch_device(clean, n, seed=7)
gray = sc.grayscale(batch)
for _ in range(3):
    if op == "multi":
        sc.preprocess_multi(batch)
    elif op == "grid":
        sc.preprocess_v2(batch)
    else:
        sc.v2_stage(op, gray, arg)
torch.cuda.synchronize()
print("done")
