"""DigitCNNv3: tensor-core path vs the fp32 CUDA-core path vs the CPU oracle on seeded weights/inputs, plus timing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
import numpy as np
import torch
from svb200 import Scanner
from svb200.v3_init import random_v3_state

n = int(sys.argv[1]) if len(sys.argv) > 1 else 37
sc = Scanner()
sd = random_v3_state()
sc.load_weights_v3(sd)
rng = np.random.default_rng(3)
x = torch.from_numpy(np.where(rng.random((n, 1, 28, 28)) < 0.25, 1.0, -1.0).astype(np.float32)).cuda()
sc.set_classifier_mode("fp32")
ref = sc.digitcnn_v3_forward(x).cpu().numpy()
sc.set_classifier_mode("tc")
got = sc.digitcnn_v3_forward(x).cpu().numpy()
torch.cuda.synchronize()
d = np.abs(got - ref)
print("tc vs fp32: max |d| =", d.max(), "argmax equal:", (got.argmax(1) == ref.argmax(1)).mean(), "finite:", np.isfinite(got).all())
if n <= 64:
    from oracle import model_v3_oracle as M
    want = M.forward(sd, x.cpu().numpy())
    print("tc vs oracle: max |d| =", np.abs(got - want).max(), " fp32 vs oracle:", np.abs(ref - want).max())
if n >= 1024:
    for mode in ("fp32", "tc"):
        sc.set_classifier_mode(mode)
        sc.digitcnn_v3_forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sc.digitcnn_v3_forward(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{mode}: {ms:.2f} ms for {n} cells = {n / ms * 1e3:.0f} cells/s")
