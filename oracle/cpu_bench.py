#!/usr/bin/env python3
"""CPU baseline runner (bench.py's `cpu_baseline` leg and `--impl reference` arm).  TEST/BENCH
INFRASTRUCTURE: times the reference's CPU implementation of the path (pipeline/run.py:257-318 with the model load
hoisted) on the host cores.  With --ref-root pointing at a staged copy of the reference tree
(baseline/_ref/sudoku-vision: git-ignored, shipped to the GPU box) the UNMODIFIED reference modules are imported and
called — pipeline/run.py's own preprocess_for_grid_detection / find_grid_contour / warp_perspective / extract_cells /
predict_cells (kind = "reference"); otherwise oracle/ref_port.py, the same cv2 + torch call sequence restated
(kind = "port").  CUDA is hidden from this process, so the reference's load_model picks the CPU (run.py:98).

Mode "pool": one worker process per core, each with cv2/torch pinned to 1 thread, frames pre-decoded
in memory (BASELINE.md §4 mode B: the most throughput the host can give the reference).
Mode "single": one process, library-default threads (mode A, what `python pipeline/run.py` does).
Prints one JSON object.  Runs in its own process so that forking is safe (no CUDA context here).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)  # ref_port.py lives beside this file

_G = {}


def ref_root_usable(root) -> bool:
    return bool(root) and os.path.exists(os.path.join(root, "pipeline", "run.py")) and \
        os.path.exists(os.path.join(root, "ml", "digit_cnn_v2.pt"))


class _StagedReference:
    """The unmodified reference: run.py's own functions, model loaded once (run.py:301 reloads it per image)."""

    def __init__(self, root):
        sys.path.insert(0, os.path.join(root, "pipeline"))
        import run as R  # the reference's pipeline/run.py; it puts its cv/ and ml/ on sys.path itself

        assert os.path.realpath(R.__file__).startswith(os.path.realpath(root)), R.__file__
        self.R = R
        self.model, self.device = R.load_model()
        assert self.device.type == "cpu"

    def scan(self, image):
        R = self.R
        binary = R.preprocess_for_grid_detection(image)       # run.py:261
        corners = R.find_grid_contour(binary)                 # run.py:268
        if corners is None:
            return dict(found=False)
        warped = R.warp_perspective(image, corners)           # run.py:278
        cells = R.extract_cells(warped)                       # run.py:286
        preds = R.predict_cells(cells, self.model, self.device)  # run.py:302
        return dict(found=True, grid=R.build_grid(preds), corners=corners)     # run.py:306


def _init(frames_path, weights_path, threads, ref_root=None):
    import cv2
    import torch

    if threads:
        cv2.setNumThreads(threads)
        torch.set_num_threads(threads)
    if ref_root_usable(ref_root):
        _G["scanner"] = _StagedReference(ref_root)
    else:
        import ref_port

        z = np.load(weights_path)
        _G["scanner"] = ref_port.RefScanner({k: z[k] for k in z.files})
    _G["frames"] = np.load(frames_path, mmap_mode="r")


def _work(args):
    i0, count = args
    fr = _G["frames"]
    sc = _G["scanner"]
    found = 0
    for k in range(count):
        r = sc.scan(np.ascontiguousarray(fr[(i0 + k) % len(fr)]))
        found += int(r["found"])
    return count, found


def _result_of(i):
    """what the scanner reads off sample frame i: (found, corners (4,2) int32 in approxPolyDP order, 9x9 grid)"""
    r = _G["scanner"].scan(np.ascontiguousarray(_G["frames"][i]))
    if not r["found"]:
        return 0, np.zeros((4, 2), np.int32), np.zeros((9, 9), np.uint8)
    grid = r["grid"] if "grid" in r else r["digits"]  # the staged reference builds a 9x9 list, the port keeps 81 digits
    return 1, np.asarray(r["corners"], np.int32).reshape(4, 2), np.asarray(grid, np.uint8).reshape(9, 9)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", required=True, help=".npy of (n,H,W,3) uint8 frames")
    ap.add_argument("--weights", required=True)
    ap.add_argument("--seconds", type=float, default=12.0)
    ap.add_argument("--workers", type=int, default=0, help="0 = os.cpu_count()")
    ap.add_argument("--mode", choices=["pool", "single"], default="pool")
    ap.add_argument("--ref-root", default=None, help="staged copy of the reference tree (baseline/_ref/sudoku-vision)")
    ap.add_argument("--results", default=None, help="write found / corners / grid of every sample frame to this .npz (untimed)")
    a = ap.parse_args()
    os.environ["CUDA_VISIBLE_DEVICES"] = ""  # the CPU baseline: the reference's load_model must not find a GPU
    kind = "reference" if ref_root_usable(a.ref_root) else "port"
    ncpu = os.cpu_count() or 1
    workers = a.workers or ncpu
    t_setup = time.perf_counter()
    if a.mode == "single":
        _init(a.frames, a.weights, 0, a.ref_root)
        _work((0, 2))  # warm-up
        if a.results:
            res = [_result_of(i) for i in range(len(_G["frames"]))]
            np.savez(a.results, found=np.array([r[0] for r in res], np.uint8), corners=np.stack([r[1] for r in res]),
                     grid=np.stack([r[2] for r in res]))
        done = found = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < a.seconds:
            c, f = _work((done, 1))
            done += c
            found += f
        dt = time.perf_counter() - t0
        import cv2
        import torch

        cores = max(cv2.getNumThreads(), torch.get_num_threads())
    else:
        import multiprocessing as mp

        ctx = mp.get_context("fork")
        with ctx.Pool(workers, initializer=_init, initargs=(a.frames, a.weights, 1, a.ref_root)) as pool:
            pool.map(_work, [(i, 1) for i in range(workers)])  # warm-up: import + first frame per worker
            if a.results:  # untimed: the boards of the sample frames, for bench.py's parity block
                nfr = len(np.load(a.frames, mmap_mode="r"))
                res = pool.map(_result_of, range(nfr))
                np.savez(a.results, found=np.array([r[0] for r in res], np.uint8), corners=np.stack([r[1] for r in res]),
                         grid=np.stack([r[2] for r in res]))
            done = found = 0
            chunk = 2
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < a.seconds:
                res = pool.map(_work, [(done + i * chunk, chunk) for i in range(workers)])
                done += sum(c for c, _ in res)
                found += sum(f for _, f in res)
            dt = time.perf_counter() - t0
        cores = workers
    print(json.dumps({"frames_per_s": done / dt, "frames": done, "seconds": dt, "cores": cores, "mode": a.mode,
                      "found": found, "host_cpus": ncpu, "setup_s": t0 - t_setup, "kind": kind}))


if __name__ == "__main__":
    main()
