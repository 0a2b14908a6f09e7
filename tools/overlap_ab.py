"""A/B of SVB_OPT_OVERLAP (sub-batches of svb_scan_batch_v1 on two internal streams) on 1024 device-resident 1080p frames.
    python tools/overlap_ab.py [frames]"""
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
sc = Scanner(device=0, weights=load_digitcnn_weights())
clean = torch.from_numpy(np.stack([F.make_frame(31000 + i, 1080, 1920).image for i in range(8)])).cuda()
batch = F.noisy_batch_device(clean, n, seed=7)
out = sc.alloc_outputs(n)
for parts in (0, 2, 3, 4, 8, 0):
    sc.set_option("overlap", parts)
    for _ in range(2):
        sc.scan_batch(batch, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sc.scan_batch(batch, out)
    e1.record()
    torch.cuda.synchronize()
    print(f"overlap parts={parts}: {e0.elapsed_time(e1) / 5:.3f} ms per {n} frames")
