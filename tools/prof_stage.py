"""Profiling driver: runs one stage (or the whole path) a few times on a device-resident synthetic batch.
    python tools/prof_stage.py [all|k1|k2|k4|k5|k5f] [frames] [iters]
k4 = the batched cells kernel (bit-row output), k5 = the classifier on bit rows, k5f = the classifier on float cells."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
import numpy as np
import torch
import svb200._lib as _L

if os.environ.get("SVB_LIB"):  # A/B of two builds (tools/build_variant.py); a tool-side override, the product loader has none
    _L.LIB_PATH = os.path.abspath(os.environ["SVB_LIB"])
from svb200 import Scanner, load_digitcnn_weights
from svb200 import frames as F

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc = Scanner(weights=load_digitcnn_weights())
clean = torch.from_numpy(np.stack([F.make_frame(31000 + i, 1080, 1920).image for i in range(4)])).cuda()
batch = F.noisy_batch_device(clean, n, seed=7)
mask = sc.preprocess(batch)
corners, found = sc.find_grid_contour(mask)
bits = sc.cells_from_frames_bits(batch, corners, found)
xb = bits.view(-1, 28)
sc.digitcnn_forward_bits(xb)  # warm-up: arenas, function attributes
x5 = None
if what == "k5f":
    _, pm1 = sc.cells_from_frames(batch, corners, found, want_u8=False)
    x5 = pm1.view(-1, 1, 28, 28)
    sc.digitcnn_forward(x5)
out = sc.scan_batch(batch)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(iters):
    if what == "k1":
        sc.preprocess(batch, out=mask)
    elif what == "k2":
        sc.find_grid_contour(mask)
    elif what == "k4":
        sc.cells_from_frames_bits(batch, corners, found)
    elif what == "k5":
        sc.digitcnn_forward_bits(xb)
    elif what == "k5f":
        sc.digitcnn_forward(x5)
    else:
        sc.scan_batch(batch, out)
ev[1].record()
torch.cuda.synchronize()
print(what, "frames", n, "ms/iter", ev[0].elapsed_time(ev[1]) / iters, "found", int((found == 1).sum()))
