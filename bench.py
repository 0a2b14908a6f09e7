#!/usr/bin/env python3
"""bench.py — frames/sec of the sudoku-vision scan path (preprocess -> corners -> warp -> 81 cells ->
DigitCNN) on synthetic 1080p frames, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames-per-gpu F] [--impl ours|reference]
    N > 1 is launched by torchrun (one rank per GPU); ranks shard frames by image, no data-path
    collective, weak scaling: every GPU scans its own F-frame batch per step.

A step = one pass of the whole hot path (svb_scan_batch_v1: K1..K5) over a batch of F device-resident
frames (BASELINE.json configs[1]: 1024 synthetic 1080p frames).  Prints ONE JSON line (rank 0).
`value`  device-timed (CUDA events, max over ranks) frames/s with inputs resident in HBM.
`e2e`    the same metric through svb_scan_batch_v1_host with pinned HOST buffers: H2D of the
         frames and D2H of the boards inside the timed region.
`roofline` the longest kernel of a step: K1 (HBM-bound: 3HW read + HW written per frame) or the classifier's
         convolution kernel (tensor-bound); the other one is reported as `roofline_other`.
`cpu_baseline` oracle/ref_port.py (the reference's cv2 + torch-CPU call sequence) on the host cores.
--impl reference runs only that CPU leg and prints it in the same schema.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sudoku-vision_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W = 1080, 1920
METRIC = "frames/sec end-to-end (preprocess->warp->81-cell CNN) at 1080p"
K1_BYTES_PER_FRAME = 3 * H * W + H * W  # 6,220,800 read + 2,073,600 written
# dram__bytes_read.sum + dram__bytes_write.sum of k1w::fused_preprocess_warp_kernel per 1080p frame, from the ncu --set full
# capture profiles/r1d_k1w_raw.csv (256 frames: 1.7692 GB read + 0.5105 GB written; the 32 staged halo/pad columns of every
# 240-column strip are the excess over the algorithmic 6.22 + 2.07 MB)
K1_TRAFFIC_PER_FRAME = int((1.769238e9 + 0.510463e9) / 256)
# k5tc::tc_conv_kernel: conv1 + conv2 of ml/model.py:36-37, true MACs only: 225,792 + 3,612,672 per cell
K5_CONV_FLOP_PER_CELL = 2 * (225792 + 3612672)


def measured_peaks():
    """(HBM GB/s, sustained dense bf16 TFLOP/s, where they come from)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1500.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        busy = [r for r in self.rows if r and r[0].isdigit() and int(r[0]) > 500]  # samples taken under load
        sm = sorted(int(r[0]) for r in (busy or self.rows) if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa_node(local: int) -> str:
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the NUMA node its GPU hangs off:
    with one rank per GPU all copying 55 GB/s from host memory at once, cross-socket traffic is what breaks e2e scaling."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local).pci_bus_id  # e.g. 0000:1B:00.0 (older torch: absent)
    except Exception:
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            bus = out[-12:] if len(out) >= 12 else out  # nvidia-smi prints an 8-digit domain
        except Exception:
            return "unbound"
    try:
        bus = bus.lower()
        if len(bus.split(":")[0]) > 4:
            bus = bus[-12:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "unbound (no NUMA information)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unbound"
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception:
        return "unbound"


def cpu_leg(frames_np, seconds: float, mode: str) -> dict:
    """Runs oracle/cpu_bench.py in a clean subprocess on a bounded sample of the same workload."""
    import numpy as np
    from svb200.api import default_weights_path

    with tempfile.TemporaryDirectory() as td:
        fp = os.path.join(td, "frames.npy")
        np.save(fp, frames_np)
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_bench.py"), "--frames", fp, "--weights",
               default_weights_path(), "--seconds", str(seconds), "--mode", mode,
               "--ref-root", os.path.join(ROOT, "baseline", "_ref", "sudoku-vision")]
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    if out.returncode != 0:
        raise RuntimeError("cpu_bench failed: " + out.stderr[-2000:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    what = ("the UNMODIFIED reference (baseline/_ref/sudoku-vision): pipeline/run.py's own preprocess / find_grid_contour / "
            "warp / extract_cells / predict_cells on cv2 + torch CPU" if r.get("kind") == "reference"
            else "oracle/ref_port.py = the reference's cv2+torch-CPU call sequence")
    return {"value": r["frames_per_s"], "unit": "frames/s", "cores": r["cores"], "kind": r.get("kind", "port"),
            "sample": f"{r['frames']} synthetic 1080p frames in {r['seconds']:.1f} s, {r['mode']} mode "
                      f"({r['cores']} worker(s) on {r['host_cpus']} host CPUs), {what}, model load hoisted; "
                      f"grids found {r['found']}/{r['frames']}"}


def host_frames(n_unique: int, seed0: int = 31000):
    import numpy as np
    from svb200 import frames as F

    return np.stack([F.make_frame(seed0 + i, H, W).image for i in range(n_unique)])


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    import numpy as np
    from svb200 import frames as F

    clean = host_frames(8)
    frames = np.stack([F.add_noise_host(clean[i % 8], 100 + i) for i in range(16)])
    per_step = max(4.0, min(20.0, 90.0 / max(args.steps + args.warmup, 1)))
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_leg(frames, per_step, "pool")
        if s >= args.warmup:
            vals.append(r)
    v = sum(x["value"] for x in vals) / len(vals)
    base = vals[-1]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
            "config": {"workload": "synthetic 1080p sudoku frames (svb200/frames.py), bounded CPU sample per step",
                       "frame": [H, W, 3]},
            "cpu_baseline": dict(base, value=v),
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def other_configs(sc, dev) -> dict:
    """Bounded device-timed runs of the other BASELINE configs (parity is covered in tests/): classifier-only cells/s
    (configs[2]) and the v2 path preprocess_multi -> contour+validity -> cells -> DigitCNNv3 -> top-3 (configs[3]) at
    1080p and 4K.  DigitCNNv3 has no shipped weights: seeded random init, as run_v2.py itself falls back to."""
    import numpy as np
    import torch
    from svb200 import frames as F
    from svb200.v3_init import random_v3_state

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {}
    sc.load_weights_v3(random_v3_state())
    g = torch.Generator(device=dev).manual_seed(5)
    for name, n_cells, fwd in (("digitcnn_tc", 262144, sc.digitcnn_forward), ("digitcnn_v3_tc", 32768, sc.digitcnn_v3_forward)):
        x = torch.where(torch.rand((n_cells, 1, 28, 28), device=dev, generator=g) < 0.25, 1.0, -1.0)
        ms = timed(lambda: fwd(x), 3)
        out["classifier_only_" + name] = {"cells": n_cells, "ms": round(ms, 3), "cells_per_s": n_cells / (ms * 1e-3)}
        del x
    for tag, (hh, ww, n) in {"1080p": (1080, 1920, 64), "4k": (2160, 3840, 16)}.items():
        clean = torch.from_numpy(np.stack([F.make_frame(41000 + i, hh, ww).image for i in range(2)])).to(dev)
        batch = F.noisy_batch_device(clean, n, seed=11)
        ms_pre = timed(lambda: sc.preprocess_multi(batch, want_aux=False), 2)
        ms_all = timed(lambda: sc.scan_batch_v2(batch), 2)
        r = sc.scan_batch_v2(batch)
        out["v2_path_" + tag] = {"frames": n, "frame": [hh, ww, 3], "ms": round(ms_all, 3), "frames_per_s": n / (ms_all * 1e-3),
                                 "preprocess_multi_ms": round(ms_pre, 3), "grids_found": int((r["found"] == 1).sum().item())}
        del batch, clean, r
    # the step after the path: batched solve of recognised boards, puzzles of mixed difficulty
    from svb200 import frames as _F  # noqa: F401
    rng = np.random.default_rng(9)
    base = np.array([[(3 * (r % 3) + r // 3 + c) % 9 + 1 for c in range(9)] for r in range(9)], np.uint8)
    grids = np.stack([(rng.permutation(9) + 1).astype(np.uint8)[base - 1].reshape(-1) for _ in range(16384)])
    for g_ in grids:
        g_[rng.permutation(81)[: int(rng.integers(40, 58))]] = 0
    gd = torch.from_numpy(grids).to(dev)
    ms = timed(lambda: sc.solve_batch(gd), 2)
    _, stt = sc.solve_batch(gd)
    out["solve_batch"] = {"puzzles": len(grids), "ms": round(ms, 3), "puzzles_per_s": len(grids) / (ms * 1e-3),
                          "solved": int((stt == 1).sum().item())}
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames-per-gpu", type=int, default=1024)
    ap.add_argument("--e2e-frames", type=int, default=1024, help="frames per end-to-end step (pinned host memory: 6.2 MB each)")
    ap.add_argument("--unique", type=int, default=16, help="distinct clean frames rendered on the host")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the bounded BASELINE configs[2..3] side measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from svb200 import Scanner, load_digitcnn_weights
    from svb200 import frames as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "unbound (single rank)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    Fn = args.frames_per_gpu
    sc = Scanner(device=local, weights=load_digitcnn_weights())
    clean = torch.from_numpy(host_frames(args.unique, 31000 + 1000 * rank)).to(dev)
    batch = F.noisy_batch_device(clean, Fn, seed=7 + rank)  # Fn x 6.2 MB: far larger than the 126 MB L2
    out = sc.alloc_outputs(Fn)
    sc.stage_timing(True)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.3 s to start: begin before the warm-up, read the median under load
    for _ in range(args.warmup):
        sc.scan_batch(batch, out)
    barrier()
    launches0 = sc.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {k: 0.0 for k in Scanner.STAGES}
    ev0.record()
    for _ in range(args.steps):
        sc.scan_batch(batch, out)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = sc.launches - launches0
    last = sc.last_stage_ms()  # stages of the last timed step (events recorded inside the timed region)
    # per-stage averages over K more steps, identical launches (kept out of `value`)
    for _ in range(args.steps):
        sc.scan_batch(batch, out)
        s = sc.last_stage_ms()
        for k in stage_ms:
            stage_ms[k] += s[k] / args.steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    found = int((out["found"] == 1).sum().item())

    # ---- e2e: pinned host frames -> svb_scan_batch_v1_host -> host boards ---------------------------
    En = min(args.e2e_frames, Fn)
    host_in = torch.empty((En, H, W, 3), dtype=torch.uint8).pin_memory()
    host_in.copy_(batch[:En])
    host_out = dict(digits=torch.empty((En, 81), dtype=torch.uint8).pin_memory(),
                    conf=torch.empty((En, 81), dtype=torch.float32).pin_memory(),
                    corners=torch.empty((En, 4, 2), dtype=torch.int32).pin_memory(),
                    found=torch.empty((En,), dtype=torch.uint8).pin_memory())
    for _ in range(2):
        sc.scan_batch_host(host_in, host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sc.scan_batch_host(host_in, host_out)  # synchronous: returns after the D2H copies
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    clocks = sampler.stop() if rank == 0 else None
    same = bool(torch.equal(host_out["digits"].to(dev), out["digits"][:En]))

    # final gather of the 81-digit boards (outside the timed region; the only collective)
    if world > 1:
        from svb200.shard import gather_boards

        boards = gather_boards(out["digits"], Fn * world)
        assert boards.shape == (Fn * world, 81)

    # ---- side measurements of BASELINE configs[2] and [3] (bounded; N = 1 only; not the headline) --------------
    other_cfg = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        other_cfg = other_configs(sc, dev)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core, whatever this rank was bound to
        cpu = cpu_leg(batch[:16].cpu().numpy(), args.cpu_seconds, "pool")
    if rank == 0:
        peak, tpeak, how = measured_peaks()
        k1_ms = stage_ms["k1_preprocess"]
        achieved = K1_BYTES_PER_FRAME * Fn / (k1_ms * 1e-3) / 1e9
        value = Fn * world * args.steps / (ms_total * 1e-3)
        roof_k1 = {"bound": "hbm", "kernel": "k1w::fused_preprocess_warp_kernel", "achieved": achieved, "peak": peak,
                   "unit": "GB/s", "frac": achieved / peak, "traffic": K1_TRAFFIC_PER_FRAME * Fn,
                   "traffic_source": "ncu --set full capture of the same kernel through svb_preprocess_v1, profiles/r1d_k1w_raw.csv, scaled per "
                                     "frame; in the whole-path call it also writes the 0.27 MB/frame bit mask K2 traces (not in `traffic`, "
                                     "nor in the algorithmic bytes)",
                   "peak_source": how, "algorithmic_bytes_per_launch": K1_BYTES_PER_FRAME * Fn, "launch_ms": k1_ms}
        # the classifier's convolution kernel is the other large launch of a step: tensor-pipe bound, reported against the
        # sustained dense bf16 peak (it runs inside a long step); fp16 hi/lo split = 3 hardware MACs per algorithmic MAC
        k5_ms = stage_ms["k5_conv"]
        tf = K5_CONV_FLOP_PER_CELL * 81 * Fn / (k5_ms * 1e-3) / 1e12
        roof_k5 = {"bound": "tensor", "kernel": "k5tc::tc_conv_kernel", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                   "frac": tf / tpeak, "traffic": None, "peak_source": how + ", bf16 dense sustained",
                   "algorithmic_flops_per_launch": K5_CONV_FLOP_PER_CELL * 81 * Fn, "launch_ms": k5_ms,
                   "note": "operands are split fp16 hi+lo (three tcgen05 products per algorithmic product) to keep logits within 1e-3 of fp32"}
        dominant, other = (roof_k1, roof_k5) if k1_ms >= k5_ms else (roof_k5, roof_k1)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
            "config": {"workload": f"batch of {Fn} synthetic 1080p sudoku frames per GPU (BASELINE configs[1]), "
                                   f"device-resident, whole path K1..K5 per step",
                       "frame": [H, W, 3], "frames_per_gpu": Fn, "classifier": "DigitCNN (ml/model.py) on tcgen05: fp16 hi+lo split operands, fp32 TMEM accumulators (logits within 1e-3 of fp32)",
                       "l2": f"inputs {Fn * H * W * 3 / 1e9:.1f} GB per step >> 126 MB L2 (no flush needed)",
                       "sharding": f"image-sharded x{world}, no data-path collective", "grids_found": f"{found}/{Fn}",
                       "host_binding_rank0": numa},
            "e2e": {"value": En * world * args.steps / e2e_s, "unit": "frames/s",
                    "h2d_bytes_per_step": En * H * W * 3 * world, "d2h_bytes_per_step": En * (81 + 81 * 4 + 32 + 1) * world,
                    "frames_per_step": En, "matches_device_path": same},
            "gpu_launches": int(launches),
            "stage_ms_per_step": {k: round(v, 4) for k, v in stage_ms.items()},
            "stage_ms_last_timed_step": {k: round(v, 4) for k, v in last.items()},
            "roofline": dominant, "roofline_other": other,
            "cpu_baseline": cpu, "clocks": clocks, "other_configs": other_cfg,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
