"""Profiling driver for row V1 (cv/preprocess_v2.py kernels): times the whole front end and each stage on a
device-resident synthetic batch.   python tools/prof_v2.py [frames] [iters] [h] [w] [multi|grid|stages]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
import numpy as np
import torch
from svb200 import Scanner
from svb200 import frames as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
w = int(sys.argv[4]) if len(sys.argv) > 4 else 1920
what = sys.argv[5] if len(sys.argv) > 5 else "all"
sc = Scanner()
clean = torch.from_numpy(np.stack([F.make_frame(31000 + i, h, w).image for i in range(4)])).cuda()
clean[1, :, : w // 3] = (clean[1, :, : w // 3].float() * 0.45).to(torch.uint8)  # one shadowed frame in four
batch = F.noisy_batch_device(clean, n, seed=7)
gray = sc.grayscale(batch)


def timed(name, fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms:9.3f} ms / {n} frames = {1000 * ms / n:8.2f} us/frame", flush=True)


if what in ("all", "grid"):
    timed("preprocess_v2 (grid_det)", lambda: sc.preprocess_v2(batch))
if what in ("all", "multi"):
    timed("preprocess_multi", lambda: sc.preprocess_multi(batch))
    timed("preprocess_multi (no aux)", lambda: sc.preprocess_multi(batch, want_aux=False))
if what in ("all", "stages"):
    k = max(h, w) // 10
    k += (k % 2 == 0)
    timed(f"dilate_ellipse k={k}", lambda: sc.v2_stage("dilate_ellipse", gray, k))
    timed("normalize_illumination", lambda: sc.v2_stage("normalize_illumination", gray))
    timed("detect_shadow (box blur)", lambda: sc.v2_stage("detect_shadow", gray, want_image=False))
    timed("remove_shadow", lambda: sc.v2_stage("remove_shadow", gray))
    timed("clahe8", lambda: sc.v2_stage("clahe8", gray))
    timed("otsu", lambda: sc.v2_stage("otsu", gray))
    timed("sauvola", lambda: sc.v2_stage("sauvola", gray))
    timed("cleanup", lambda: sc.v2_stage("cleanup", gray))
    timed("blur5", lambda: sc.blur(gray))
    timed("adaptive", lambda: sc.adaptive_threshold(gray))
r = sc.preprocess_multi(batch[:8])
print("info", r["info"].cpu().numpy().tolist())
