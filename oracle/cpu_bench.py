#!/usr/bin/env python3
"""CPU baseline runner (bench.py's `cpu_baseline` leg and `--impl reference` arm).  TEST/BENCH
INFRASTRUCTURE: times oracle/ref_port.py — the reference's own OpenCV + PyTorch-CPU call sequence
(pipeline/run.py:257-318 with the model load hoisted) — on the host cores.

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


def _init(frames_path, weights_path, threads):
    import cv2
    import torch

    if threads:
        cv2.setNumThreads(threads)
        torch.set_num_threads(threads)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", required=True, help=".npy of (n,H,W,3) uint8 frames")
    ap.add_argument("--weights", required=True)
    ap.add_argument("--seconds", type=float, default=12.0)
    ap.add_argument("--workers", type=int, default=0, help="0 = os.cpu_count()")
    ap.add_argument("--mode", choices=["pool", "single"], default="pool")
    a = ap.parse_args()
    ncpu = os.cpu_count() or 1
    workers = a.workers or ncpu
    t_setup = time.perf_counter()
    if a.mode == "single":
        _init(a.frames, a.weights, 0)
        _work((0, 2))  # warm-up
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
        with ctx.Pool(workers, initializer=_init, initargs=(a.frames, a.weights, 1)) as pool:
            pool.map(_work, [(i, 1) for i in range(workers)])  # warm-up: import + first frame per worker
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
                      "found": found, "host_cpus": ncpu, "setup_s": t0 - t_setup}))


if __name__ == "__main__":
    main()
