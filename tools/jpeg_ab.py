"""Frame ingest alone: svb_jpeg_decode_host on a batch of synthetic 1080p JPEGs (cv2 encoder): ms per stage by CUDA events
around whole calls, frames/s, and equality with cv2.imdecode.   python tools/jpeg_ab.py [frames] [quality] [rst_interval]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sudoku-vision_b200")):
    sys.path.insert(0, p)
import cv2
import numpy as np
import torch
import svb200._lib as _L

if os.environ.get("SVB_LIB"):  # A/B of two builds (tools/build_variant.py)
    _L.LIB_PATH = os.path.abspath(os.environ["SVB_LIB"])
from svb200 import Scanner, load_digitcnn_weights
from svb200 import frames as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
q = int(sys.argv[2]) if len(sys.argv) > 2 else 95
rst = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sc = Scanner(device=0, weights=load_digitcnn_weights())
files, ref = [], []
for i in range(8):
    im = F.add_noise_host(F.make_frame(31000 + i, 1080, 1920).image, 100 + i)
    params = [cv2.IMWRITE_JPEG_QUALITY, q] + ([cv2.IMWRITE_JPEG_RST_INTERVAL, rst] if rst else [])
    ok, buf = cv2.imencode(".jpg", im, params)
    files.append(buf.tobytes())
    ref.append(cv2.imdecode(buf, cv2.IMREAD_COLOR))
blob, offs = Scanner.pack_jpegs([files[i % 8] for i in range(n)])
pb = torch.from_numpy(blob).pin_memory()
out = sc.jpeg_decode(pb, offs, 1080, 1920)
torch.cuda.synchronize()
bad = sum(int((out[i].cpu().numpy() != ref[i % 8]).sum()) for i in (0, 1, 7, n - 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = sc.jpeg_decode(pb, offs, 1080, 1920)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"jpeg decode: {n} frames q{q} rst{rst} ({len(blob) / n / 1e3:.0f} KB/frame): {ms:.3f} ms per call = {ms * 1e3 / n:.1f} us/frame "
      f"= {n / ms * 1e3:.0f} frames/s; mismatching bytes vs cv2.imdecode in 4 frames: {bad}")
