"""GPU-side bisect of a whole-path parity failure on a bench-style batch: every frame against the C oracle, stage by stage
(corners -> u8 cells -> +-1 bits -> logits); frames that differ are dumped to gpurun_out/ for CPU reproduction with the
host-compiled cores.  Usage (GPU box): python tests/bisect_1024.py [n_frames]"""
import concurrent.futures as cf
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sudoku-vision_b200")]
from oracle import oracle as O  # noqa: E402
from svb200 import Scanner, load_digitcnn_weights  # noqa: E402
from svb200 import frames as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w = load_digitcnn_weights()
sc = Scanner(device=0, weights=w)
clean = np.stack([F.make_frame(31000 + i, 1080, 1920).image for i in range(16)])
batch = F.noisy_batch_device(torch.from_numpy(clean).cuda(), n, seed=7)
out = sc.scan_batch(batch, want_logits=True)
torch.cuda.synchronize()
got = {k: v.cpu().numpy() for k, v in out.items()}
frames = batch.cpu().numpy()


def one(i):
    r = O.scan_frame(frames[i])
    if not r["found"]:
        return i, None
    return i, r


with cf.ThreadPoolExecutor(max_workers=max(1, len(os.sched_getaffinity(0)))) as ex:
    res = list(ex.map(one, range(n)))
bad = []
for i, r in res:
    if r is None:
        if got["found"][i] == 1:
            print(f"frame {i}: GPU found a grid, oracle did not")
            bad.append(i)
        continue
    if got["found"][i] != 1 or not np.array_equal(got["corners"][i], r["corners"]):
        print(f"frame {i}: found/corners differ: {got['found'][i]} {got['corners'][i].tolist()} vs {r['corners'].tolist()}")
        bad.append(i)
        continue
    lg = O.digitcnn_forward(w, (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5)
    d = np.abs(got["logits"][i] - lg).max(1)
    if d.max() >= 1e-3:
        bad.append(i)
        cells = np.nonzero(d >= 1e-3)[0]
        print(f"frame {i}: logits differ in cells {cells.tolist()} (max {d.max():.4f})")
print(f"{len(bad)} of {n} frames differ: {bad}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for i in bad[:6]:
    fr = batch[i:i + 1]
    r = dict(res)[i]
    c = torch.from_numpy(r["corners"][None].astype(np.int32)).cuda()
    u8, pm1 = sc.cells_from_frames(fr, c, want_u8=True)
    bits = sc.cells_from_frames_bits(fr, c)
    torch.cuda.synchronize()
    u8, pm1, bits = u8.cpu().numpy()[0], pm1.cpu().numpy()[0], bits.cpu().numpy()[0]
    want_pm1 = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
    bits_as_pm1 = np.where((bits[:, :, None] >> np.arange(28)[None, None, :]) & 1, 1.0, -1.0).astype(np.float32)
    print(f"frame {i}: u8 cells differ in {np.nonzero((u8 != r['cells_u8']).any((1, 2)))[0].tolist()}, "
          f"float +-1 in {np.nonzero((pm1 != want_pm1).any((1, 2)))[0].tolist()}, "
          f"bit rows in {np.nonzero((bits_as_pm1 != want_pm1).any((1, 2)))[0].tolist()}")
    # the staged drop-in path for the same frame
    board = sc.warp_perspective(fr, c)
    cells2 = sc.extract_cells(board)
    torch.cuda.synchronize()
    print(f"   staged extract_cells differ in {np.nonzero((cells2.cpu().numpy()[0] != r['cells_u8']).any((1, 2)))[0].tolist()}")
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"dbg1024_frame{i}.npz"), frame=frames[i], corners=r["corners"],
                        gpu_u8=u8, gpu_pm1=pm1, gpu_bits=bits, gpu_logits=got["logits"][i], want_u8=r["cells_u8"],
                        want_in=r["cells_in"])
