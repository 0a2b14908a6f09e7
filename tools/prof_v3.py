"""DigitCNNv3 forward (K6) timing on device-resident +-1 cells: python tools/prof_v3.py [cells] [iters]
(SVB_LIB=ab/<name>/libsvb200.so selects another build, see tools/build_variant.py)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sudoku-vision_b200")]
import torch
import svb200._lib as _L

if os.environ.get("SVB_LIB"):
    _L.LIB_PATH = os.path.abspath(os.environ["SVB_LIB"])
from svb200 import Scanner
from svb200.v3_init import random_v3_state

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = Scanner(device=0)
sc.load_weights_v3(random_v3_state())
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.where(torch.rand((n, 1, 28, 28), device="cuda", generator=g) < 0.25, 1.0, -1.0)
lg = sc.digitcnn_v3_forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    lg = sc.digitcnn_v3_forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"v3 forward: {n} cells {ms:.3f} ms  {n / ms / 1e3:.3f} M cells/s  logits checksum {float(lg.double().sum()):.6f} absmax {float(lg.abs().max()):.4f}")
