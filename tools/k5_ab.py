"""DigitCNN classifier alone (K5: conv + fc): ms per launch and cells/s on +-1 cells given as bit rows (the batched path) and
as floats (the drop-in forward), and the logits of one against the other.   python tools/k5_ab.py [cells] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sudoku-vision_b200")):
    sys.path.insert(0, p)
import torch
import svb200._lib as _L

if os.environ.get("SVB_LIB"):  # A/B of two builds (tools/build_variant.py)
    _L.LIB_PATH = os.path.abspath(os.environ["SVB_LIB"])
from svb200 import Scanner, load_digitcnn_weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 82944
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
sc = Scanner(device=0, weights=load_digitcnn_weights())
g = torch.Generator(device=dev).manual_seed(5)
x = torch.where(torch.rand((n, 1, 28, 28), device=dev, generator=g) < 0.25, 1.0, -1.0)
xb = sc.pack_cells_bits(x.view(n, 28, 28))
ref = None
for name, fn in (("bit rows", lambda: sc.digitcnn_forward_bits(xb)), ("floats", lambda: sc.digitcnn_forward(x))):
    for _ in range(2):
        lg = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lg = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    d = "" if ref is None else f"  max |dlogit| vs bit rows: {float((lg - ref).abs().max()):.2e}"
    ref = lg if ref is None else ref
    print(f"K5 {name}: {n} cells {ms:.3f} ms/launch  {n / ms / 1e3:.2f} M cells/s{d}")
