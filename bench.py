#!/usr/bin/env python3
"""bench.py — frames/sec of the sudoku-vision scan path (preprocess -> corners -> warp -> 81 cells ->
DigitCNN) on synthetic 1080p frames, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames-per-gpu F] [--impl ours|reference]
    N > 1 is launched by torchrun (one rank per GPU); ranks shard frames by image, no data-path
    collective, weak scaling: every GPU scans its own F-frame batch per step.

A step = one pass of the whole hot path (svb_scan_batch_v1: K1..K5) over a batch of F device-resident
frames (BASELINE.json configs[1]: 1024 synthetic 1080p frames).  Prints ONE JSON line (rank 0).
`value`     device-timed (CUDA events, max over ranks) frames/s with inputs resident in HBM.
`e2e`       the same metric through svb_scan_batch_v1_host with pinned HOST buffers: H2D of the frames and D2H of the
            boards inside the timed region; `h2d_ceiling_gbs` = plain pinned cudaMemcpyAsync on all ranks at once.
`roofline`  the longest kernel of a step; `rooflines` = every stage (K1, K2, K4, K5 conv, K5 fc) with its bound.
`parity`    the benched frames themselves checked against the CPU oracle (and, for the frames the CPU baseline ran,
            against the UNMODIFIED reference's boards): found / corners / masks / +-1 cells / logits / digits.
`stream`    BASELINE configs[4]: >= --stream-seconds of sustained scanning, from a ring of pinned host batches (H2D in)
            and from a ring of device-resident batches.
`cpu_baseline` the reference's cv2 + torch-CPU path on the host cores, run twice BEFORE any GPU work of this process.
`other_configs` (N = 1) BASELINE configs[2] (1 M cells, DigitCNNv3 and DigitCNN) and configs[3] (v2 path, 4096 4K frames).
--impl reference runs only the CPU leg and prints it in the same schema.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
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
# dram__bytes_read.sum + dram__bytes_write.sum per 1080p frame from the ncu --set full captures under profiles/ (see
# profiles/README.md for the capture each constant comes from); None = no capture of the current kernel yet
NCU_TRAFFIC_PER_FRAME = {"k1": int((1.769168e9 + 0.510258176e9) / 256),         # profiles/r2w_k1w_summary.csv
                         "k4": int((435.430656e6 + 6.659072e6) / 256)}            # profiles/r2w_k4_summary.csv
# true MACs only, per cell: conv1 225,792 + conv2 3,612,672; fc1 401,408 + fc2 1,280
K5_CONV_FLOP_PER_CELL = 2 * (225792 + 3612672)
K5_FC_FLOP_PER_CELL = 2 * (401408 + 1280)
V3_FLOP_PER_CELL = 2 * 66078976


def measured_peaks():
    """(HBM GB/s, sustained dense bf16 TFLOP/s, where they come from)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


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
    """Runs oracle/cpu_bench.py in a clean subprocess on a bounded sample of the same workload.  The boards the reference
    read off the sample frames come back with it (`results`: found / corners / grid per frame) for the parity block."""
    import numpy as np
    from svb200.api import default_weights_path

    with tempfile.TemporaryDirectory() as td:
        fp, rp = os.path.join(td, "frames.npy"), os.path.join(td, "results.npz")
        np.save(fp, frames_np)
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "cpu_bench.py"), "--frames", fp, "--weights",
               default_weights_path(), "--seconds", str(seconds), "--mode", mode, "--results", rp,
               "--ref-root", os.path.join(ROOT, "baseline", "_ref", "sudoku-vision")]
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        if out.returncode != 0:
            raise RuntimeError("cpu_bench failed: " + out.stderr[-2000:])
        res = dict(np.load(rp)) if os.path.exists(rp) else None
    r = json.loads(out.stdout.strip().splitlines()[-1])
    what = ("the UNMODIFIED reference (baseline/_ref/sudoku-vision): pipeline/run.py's own preprocess / find_grid_contour / "
            "warp / extract_cells / predict_cells on cv2 + torch CPU" if r.get("kind") == "reference"
            else "oracle/ref_port.py = the reference's cv2+torch-CPU call sequence")
    return {"value": r["frames_per_s"], "unit": "frames/s", "cores": r["cores"], "kind": r.get("kind", "port"),
            "sample": f"{r['frames']} synthetic 1080p frames in {r['seconds']:.1f} s, {r['mode']} mode "
                      f"({r['cores']} worker(s) on {r['host_cpus']} host CPUs), {what}, model load hoisted; "
                      f"grids found {r['found']}/{r['frames']}",
            "results": res}


def host_frames(n_unique: int, seed0: int = 31000):
    import numpy as np
    from svb200 import frames as F

    return np.stack([F.make_frame(seed0 + i, H, W).image for i in range(n_unique)])


def cpu_sample_frames(clean, n: int = 16):
    """the n noisy host frames the CPU legs run (and the GPU arm scans as frames 0..n-1 of its batch)"""
    import numpy as np
    from svb200 import frames as F

    return np.stack([F.add_noise_host(clean[i % len(clean)], 100 + i) for i in range(n)])


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    frames = cpu_sample_frames(host_frames(8))
    per_step = max(4.0, min(20.0, 90.0 / max(args.steps + args.warmup, 1)))
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_leg(frames, per_step, "pool")
        r.pop("results", None)
        if s >= args.warmup:
            vals.append(r)
    v = sum(x["value"] for x in vals) / len(vals)
    base = vals[-1]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
            "config": {"workload": "synthetic 1080p sudoku frames (svb200/frames.py), bounded CPU sample per step",
                       "frame": [H, W, 3]},
            "cpu_baseline": dict(base, value=v, runs=[x["value"] for x in vals]),
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- parity of the benched frames -----------------------------------------------------------------------------------
def parity_block(sc, batch, idx, weights, ref_results, n_ref: int) -> dict:
    """The frames of the timed batch at `idx` against the C oracle (oracle/svb_oracle.c, test infrastructure: the checker,
    never the thing measured), stage by stage, and frames 0..n_ref-1 against the boards the unmodified reference read."""
    import numpy as np
    import torch
    from oracle import oracle as O

    sel = batch[torch.as_tensor(idx, device=batch.device)].contiguous()
    got = sc.scan_batch(sel, want_logits=True)
    masks = sc.preprocess(sel)
    _, pm1 = sc.cells_from_frames(sel, got["corners"], got["found"], want_u8=False)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in got.items() if v is not None}
    masks, pm1, frames = masks.cpu().numpy(), pm1.cpu().numpy(), sel.cpu().numpy()

    def one(i):
        r = O.scan_frame(frames[i])
        lg = None
        if r["found"]:
            x = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
            lg = O.digitcnn_forward(weights, x)
        return r, lg

    with cf.ThreadPoolExecutor(max_workers=max(1, len(os.sched_getaffinity(0)))) as ex:  # ctypes releases the GIL
        ora = list(ex.map(one, range(len(idx))))
    n = len(idx)
    c = dict(frames=n, found_agree=0, corners_exact=0, corners_le1px=0, corners_le2px=0, mask_px_agree=0.0, pm1_cells_exact=0,
             pm1_cells=0, max_abs_dlogit=0.0, digits_cells=0, digits_identical=0, digits_cells_same_corners=0,
             digits_identical_same_corners=0)
    for i, (r, lg) in enumerate(ora):
        f = bool(g["found"][i] == 1)
        c["found_agree"] += int(f == r["found"])
        c["mask_px_agree"] += float((masks[i] == r["mask"]).mean()) / n
        if not (f and r["found"]):
            continue
        err = int(np.abs(g["corners"][i].astype(np.int64) - r["corners"]).max())
        c["corners_exact"] += int(err == 0)
        c["corners_le1px"] += int(err <= 1)
        c["corners_le2px"] += int(err <= 2)
        want_pm1 = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        same = (pm1[i] == want_pm1).reshape(81, -1).all(1)
        c["pm1_cells"] += 81
        c["pm1_cells_exact"] += int(same.sum())
        dig = lg.argmax(1).astype(np.uint8)
        c["digits_cells"] += 81
        c["digits_identical"] += int((g["digits"][i] == dig).sum())
        if err == 0:
            c["digits_cells_same_corners"] += 81
            c["digits_identical_same_corners"] += int((g["digits"][i] == dig).sum())
            c["max_abs_dlogit"] = max(c["max_abs_dlogit"], float(np.abs(g["logits"][i] - lg).max()))
    c["digits_identical_frac"] = c["digits_identical"] / max(c["digits_cells"], 1)
    c["digits_identical_frac_same_corners"] = c["digits_identical_same_corners"] / max(c["digits_cells_same_corners"], 1)
    c["checker"] = "oracle/svb_oracle.c (C restatement pinned to the reference by tests/golden)"
    if ref_results is not None and n_ref:
        # frames 0..n_ref-1 of the batch are the CPU baseline's sample: the unmodified reference's own outputs
        rr = dict(frames=int(n_ref), found_agree=0, corners_exact=0, digits_cells=0, digits_identical=0)
        first = sc.scan_batch(batch[:n_ref].contiguous())
        torch.cuda.synchronize()
        fd, co, dg = first["found"].cpu().numpy(), first["corners"].cpu().numpy(), first["digits"].cpu().numpy()
        for i in range(n_ref):
            rf = bool(ref_results["found"][i])
            rr["found_agree"] += int(rf == bool(fd[i] == 1))
            if rf and fd[i] == 1:
                rr["corners_exact"] += int(np.array_equal(co[i], ref_results["corners"][i]))
                rr["digits_cells"] += 81
                rr["digits_identical"] += int((dg[i] == ref_results["grid"][i].reshape(81)).sum())
        rr["digits_identical_frac"] = rr["digits_identical"] / max(rr["digits_cells"], 1)
        rr["checker"] = "the unmodified reference (cv2 + torch CPU) on the CPU baseline's sample frames"
        c["vs_reference"] = rr
    return c


def parity_coreml(sc_factory, batch, idx) -> dict:
    """The same frames through a second context loaded with the only trained weights the reference ships (fp16 DigitCNN
    recovered from its CoreML package: MNIST-era, brittle top-2 margins — BASELINE.md §4(ii)): argmax agreement with the
    oracle's fp32 forward on the oracle's cells, conditioned on identical corners."""
    import numpy as np
    import torch
    from oracle import oracle as O
    from svb200.api import coreml_weights_path, load_digitcnn_weights

    if not os.path.exists(coreml_weights_path()):
        return {"unavailable": "digitcnn_coreml.npz missing"}
    w = load_digitcnn_weights(coreml_weights_path())
    sc2 = sc_factory(w)
    sel = batch[torch.as_tensor(idx, device=batch.device)].contiguous()
    got = sc2.scan_batch(sel, want_logits=True)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in got.items() if v is not None}
    frames = sel.cpu().numpy()

    def one(i):
        r = O.scan_frame(frames[i])
        if not r["found"]:
            return None
        x = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        return r, O.digitcnn_forward(w, x)

    with cf.ThreadPoolExecutor(max_workers=max(1, len(os.sched_getaffinity(0)))) as ex:
        ora = list(ex.map(one, range(len(idx))))
    cells = same = 0
    mx = 0.0
    margins = []
    for i, o in enumerate(ora):
        if o is None or g["found"][i] != 1 or not np.array_equal(g["corners"][i], o[0]["corners"]):
            continue
        lg = o[1]
        cells += 81
        same += int((g["digits"][i] == lg.argmax(1)).sum())
        mx = max(mx, float(np.abs(g["logits"][i] - lg).max()))
        s = np.sort(lg, 1)
        margins.append(s[:, -1] - s[:, -2])
    sc2.close()
    m = np.concatenate(margins) if margins else np.zeros(1)
    return {"weights": "CoreML-recovered fp16 DigitCNN (ios/.../DigitClassifier.mlpackage weight.bin)", "frames": len(idx),
            "cells_same_corners": cells, "argmax_identical": same, "argmax_identical_frac": same / max(cells, 1),
            "max_abs_dlogit": mx, "top2_margin_p1": float(np.percentile(m, 1)), "top2_margin_min": float(m.min())}


# ---- BASELINE configs[2], [3] at their stated sizes (N = 1) ---------------------------------------------------------------
def other_configs(sc, dev, tpeak: float) -> dict:
    """configs[2]: 1,000,000 cells through DigitCNNv3 and DigitCNN (classifier only, device-resident +-1 inputs).
    configs[3]: the v2 path (preprocess_multi_strategy -> contour + validity -> cells -> DigitCNNv3 -> top-3) on 4096
    synthetic 4K frames, scanned as resident chunks.  DigitCNNv3 has no shipped weights: seeded random init, as
    run_v2.py itself falls back to.  Parity of these paths is covered in tests/ and, for 4 of the 4K frames, here."""
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
    n_cells = 1_000_000
    x = torch.empty((n_cells, 1, 28, 28), dtype=torch.float32, device=dev)
    for s in range(0, n_cells, 125_000):  # about a quarter of the pixels are ink, as in real cells
        x[s:s + 125_000] = torch.where(torch.rand((125_000, 1, 28, 28), device=dev, generator=g) < 0.25, 1.0, -1.0)
    for name, fwd, flop in (("digitcnn_v3", sc.digitcnn_v3_forward, V3_FLOP_PER_CELL),
                            ("digitcnn", sc.digitcnn_forward, K5_CONV_FLOP_PER_CELL + K5_FC_FLOP_PER_CELL)):
        ms = timed(lambda: fwd(x), 2)
        tf = flop * n_cells / (ms * 1e-3) / 1e12
        out["config3_classifier_only_" + name] = {
            "workload": "BASELINE configs[2]: 1,000,000 synthetic 28x28 +-1 cells, classifier only, one call", "cells": n_cells,
            "ms": round(ms, 3), "cells_per_s": n_cells / (ms * 1e-3), "algorithmic_tflops": tf, "tensor_frac": tf / tpeak,
            "peak_tflops": tpeak}
    # the batched path's own form of the DigitCNN input: 28 bit rows per cell (conv1 by pattern table)
    xb = sc.pack_cells_bits(x.view(n_cells, 28, 28))
    ms = timed(lambda: sc.digitcnn_forward_bits(xb), 2)
    tf = (K5_CONV_FLOP_PER_CELL + K5_FC_FLOP_PER_CELL) * n_cells / (ms * 1e-3) / 1e12
    out["config3_classifier_only_digitcnn_bit_rows"] = {
        "workload": "the same 1,000,000 cells as 28 bit rows each (svb_digitcnn_forward_bits, what svb_scan_batch_v1 runs)", "cells": n_cells,
        "ms": round(ms, 3), "cells_per_s": n_cells / (ms * 1e-3), "algorithmic_tflops": tf, "tensor_frac": tf / tpeak, "peak_tflops": tpeak}
    del x, xb
    torch.cuda.empty_cache()
    # configs[3]: 4096 4K frames = 101.9 GB of BGR: 16 resident chunks of 256 frames (6.4 GB each), regenerated per chunk
    hh, ww, total, chunk = 2160, 3840, 4096, 256
    clean = torch.from_numpy(np.stack([F.make_frame(41000 + i, hh, ww).image for i in range(4)])).to(dev)
    # oracle parity of 4 of those frames: ~40 s of CPU each (the 385-px elliptical close), so the checker threads start
    # now and run beside the GPU work below (ctypes releases the GIL); collected after the timed chunks
    from oracle import oracle as O
    from oracle import oracle_v2 as O2

    b4 = F.noisy_batch_device(clean, chunk, seed=11)
    frames4 = b4[:4].cpu().numpy()

    def oracle_v2_frame(i):
        o = O.preprocess_multi(frames4[i])
        return o["binary"], O2.detect_grid_contour(o["binary"])

    pool = cf.ThreadPoolExecutor(max_workers=4)
    futs = [pool.submit(oracle_v2_frame, i) for i in range(4)]
    ms_all = ms_pre = 0.0
    found = 0
    sample = None
    for ci in range(total // chunk):
        if ci:
            b4 = F.noisy_batch_device(clean, chunk, seed=11 + ci)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        pm = sc.preprocess_multi(b4, want_aux=False)
        e[1].record()
        r = sc.scan_batch_v2(b4)
        e[2].record()
        torch.cuda.synchronize()
        ms_pre += e[0].elapsed_time(e[1])
        ms_all += e[1].elapsed_time(e[2])
        found += int((r["found"] == 1).sum().item())
        if ci == 0:
            sample = (pm["binary"][:4].cpu().numpy(), {k: v[:4].cpu().numpy() for k, v in r.items() if v is not None})
        del b4, r, pm
    cfg4 = {"workload": "BASELINE configs[3]: v2 pipeline (preprocess_v2 multi-strategy + grid_v2 contour method + DigitCNNv3 + "
                        "top-3) on 4096 synthetic 4K frames, as 16 device-resident chunks of 256 (6.4 GB each)",
            "frames": total, "frame": [hh, ww, 3], "ms": round(ms_all, 1), "frames_per_s": total / (ms_all * 1e-3),
            "preprocess_multi_only_ms": round(ms_pre, 1), "grids_found": found}
    try:
        bins, r4 = sample
        agree = dict(frames=4, binary_px_agree=0.0, found_agree=0, corners_exact=0,
                     checker="oracle/svb_oracle.c svo_preprocess_multi + oracle_v2.detect_grid_contour (cv/preprocess_v2.py:247-308, "
                             "cv/grid_v2.py:102-128 restated)")
        for i, fu in enumerate(futs):
            binary, c = fu.result(timeout=600)
            agree["binary_px_agree"] += float((bins[i] == binary).mean()) / 4
            agree["found_agree"] += int((c is not None) == bool(r4["found"][i] == 1))
            if c is not None and r4["found"][i] == 1:
                agree["corners_exact"] += int(np.array_equal(np.asarray(c, np.float32).reshape(4, 2), r4["corners"][i].astype(np.float32)))
        cfg4["oracle_parity"] = agree
    except Exception as ex:  # the bench must not die on the side check
        cfg4["oracle_parity"] = {"error": repr(ex)[:200]}
    pool.shutdown(wait=False)
    out["config4_v2_path_4k"] = cfg4
    del clean
    # the step after the path: batched solve of recognised boards, puzzles of mixed difficulty
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
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs[2..3] (N = 1 only anyway)")
    ap.add_argument("--parity-frames", type=int, default=128, help="benched frames checked against the oracle (0 = skip)")
    ap.add_argument("--stream-seconds", type=float, default=30.0, help="BASELINE configs[4]: sustained streaming (0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np

    # ---- CPU baseline FIRST: before this process (or, under torchrun, any rank that matters) touches the GPU or joins the
    # NCCL group.  The other ranks sit in init_process_group's TCP rendezvous (a blocking socket wait, no spinning).
    clean_host = host_frames(args.unique, 31000 + 1000 * rank)
    n_ref = 16
    cpu = None
    ref_results = None
    if rank == 0 and not args.no_cpu_baseline:
        sample = cpu_sample_frames(clean_host, n_ref)
        runs = [cpu_leg(sample, args.cpu_seconds, "pool") for _ in range(2)]
        ref_results = runs[0].pop("results", None)
        runs[1].pop("results", None)
        cpu = dict(runs[1], value=(runs[0]["value"] + runs[1]["value"]) / 2, runs=[runs[0]["value"], runs[1]["value"]],
                   when="before any GPU work or process-group init of this run")

    import torch
    import torch.distributed as dist
    from svb200 import Scanner, load_digitcnn_weights
    from svb200 import frames as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "unbound (single rank)"
    if world > 1:
        import datetime

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=30))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    Fn = args.frames_per_gpu
    weights = load_digitcnn_weights()
    sc = Scanner(device=local, weights=weights)
    clean = torch.from_numpy(clean_host).to(dev)
    batch = F.noisy_batch_device(clean, Fn, seed=7 + rank)  # Fn x 6.2 MB: far larger than the 126 MB L2
    if rank == 0 and cpu is not None:
        batch[:n_ref] = torch.from_numpy(cpu_sample_frames(clean_host, n_ref)).to(dev)  # the CPU baseline's own frames
    out = sc.alloc_outputs(Fn)

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
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = sc.launches - launches0
    # per-stage times: K more steps with the library's stage events on: the same launches in the same order (kept out of `value`)
    sc.stage_timing(True)
    sc.scan_batch(batch, out)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for _ in range(args.steps):
        sc.scan_batch(batch, out)
        s = sc.last_stage_ms()
        for k in stage_ms:
            stage_ms[k] += s[k] / args.steps
    es1.record()
    torch.cuda.synchronize()
    ms_serial = es0.elapsed_time(es1) / args.steps
    sc.stage_timing(False)
    found_mask = (out["found"] == 1)
    found = int(found_mask.sum().item())
    # measured quad area (shoelace of the detected corners): K4's algorithmic source bytes are 3 x A_quad
    cf32 = out["corners"].to(torch.float64)
    x, y = cf32[:, :, 0], cf32[:, :, 1]
    area = 0.5 * (x * torch.roll(y, -1, 1) - torch.roll(x, -1, 1) * y).sum(1).abs()
    a_quad = float(area[found_mask].mean().item()) if found else 0.0

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
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    same = bool(torch.equal(host_out["digits"].to(dev), out["digits"][:En]))
    # what the link can do: the same bytes as plain pinned cudaMemcpyAsync copies, all ranks at once
    dst = torch.empty((min(En, 256), H, W, 3), dtype=torch.uint8, device=dev)
    nb = dst.shape[0]
    dst.copy_(host_in[:nb], non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for r_ in range(4):
        dst.copy_(host_in[(r_ % max(En // nb, 1)) * nb:(r_ % max(En // nb, 1)) * nb + nb], non_blocking=True)
    barrier()
    h2d_s = max_over_ranks(time.perf_counter() - t0)
    h2d_ceiling = 4 * nb * H * W * 3 * world / h2d_s / 1e9
    del dst
    # ---- e2e_jpeg: the same frames arrive COMPRESSED (cv2.imwrite's encoder, quality 95, one restart marker per 8 MCUs): H2D of
    # the JPEG bytes, GPU decode (svb_scan_batch_v1_jpeg_host), whole path, D2H of the boards.  Reported beside `e2e`, with its
    # own parity counters: the decoder is lossless against cv2.imdecode, but JPEG itself changes the pixels of the raw frames
    e2e_jpeg = None
    try:
        import cv2

        uniq = min(args.unique, 16)
        raw_u = batch[:uniq].cpu().numpy()
        files = []
        for i in range(uniq):
            ok, buf = cv2.imencode(".jpg", raw_u[i], [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_RST_INTERVAL, 8])
            files.append(buf.tobytes())
        blob_u, offs_u = Scanner.pack_jpegs([files[i % uniq] for i in range(En)])
        jblob = torch.from_numpy(blob_u).pin_memory()
        joffs = torch.from_numpy(offs_u)
        jout = dict(digits=torch.empty((En, 81), dtype=torch.uint8).pin_memory(), conf=torch.empty((En, 81), dtype=torch.float32).pin_memory(),
                    corners=torch.empty((En, 4, 2), dtype=torch.int32).pin_memory(), found=torch.empty((En,), dtype=torch.uint8).pin_memory())
        for _ in range(2):
            sc.scan_batch_jpeg_host(jblob, joffs, H, W, jout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sc.scan_batch_jpeg_host(jblob, joffs, H, W, jout)
        barrier()
        jpeg_s = max_over_ranks(time.perf_counter() - t0)
        e2e_jpeg = {"value": En * world * args.steps / jpeg_s, "unit": "frames/s", "h2d_bytes_per_step": int(offs_u[-1]) * world,
                    "d2h_bytes_per_step": En * (81 + 81 * 4 + 32 + 1) * world, "frames_per_step": En,
                    "compressed_bytes_per_frame": int(offs_u[-1]) // En,
                    "what": "svb_scan_batch_v1_jpeg_host: pinned JPEG bytes (cv2 encoder, quality 95, DRI 8 MCUs) -> GPU Huffman + IDCT + "
                            "colour -> K1..K5 -> boards in host memory"}
        if rank == 0:
            # parity of the ingest path: GPU-decoded frames vs cv2.imdecode of the same files, and the boards vs the device path
            dec = sc.jpeg_decode(blob_u[:int(offs_u[uniq]) + 64], offs_u[:uniq + 1], H, W)
            ref = np.stack([cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR) for f in files])
            dref = torch.from_numpy(ref).to(dev)
            a, b = sc.scan_batch(dec), sc.scan_batch(dref)
            raw = sc.scan_batch(batch[:uniq].contiguous())
            torch.cuda.synchronize()
            ok_f = (raw["found"] == 1) & (a["found"] == 1)
            e2e_jpeg["parity"] = {
                "frames": uniq, "decoded_px_equal_cv2_imdecode": float((dec == dref).float().mean().item()),
                "digits_identical_vs_cv2_decoded": float((a["digits"] == b["digits"]).float().mean().item()),
                "corners_exact_vs_cv2_decoded": int((a["corners"] == b["corners"]).all(-1).all(-1).sum().item()),
                "found_vs_raw_frames": int((a["found"] == raw["found"]).sum().item()),
                "corners_le1px_vs_raw_frames": int(((a["corners"] - raw["corners"]).abs().amax((1, 2)) <= 1)[ok_f].sum().item()),
                "digits_identical_vs_raw_frames": float((a["digits"] == raw["digits"])[ok_f].float().mean().item()) if bool(ok_f.any()) else None,
                "note": "vs cv2.imdecode: the decoder's own parity (same files); vs raw frames: what JPEG quality 95 itself changes"}
            del dec, dref
        del jblob
    except ImportError:
        e2e_jpeg = {"unavailable": "cv2 (the encoder for the synthetic JPEG frames) is not importable here"}
    clocks = sampler.stop() if rank == 0 else None

    # ---- BASELINE configs[4]: sustained streaming ------------------------------------------------------------------------
    stream = None
    if args.stream_seconds > 0:
        half = max(En // 2, 1)
        ring = [(host_in[:half], {k: v[:half] for k, v in host_out.items()}),
                (host_in[half:2 * half], {k: v[half:2 * half] for k, v in host_out.items()})]
        barrier()
        t0 = time.perf_counter()
        n_host = 0
        while time.perf_counter() - t0 < args.stream_seconds:
            a, b = ring[(n_host // half) & 1]
            sc.scan_batch_host(a, b)
            n_host += half
        host_s = time.perf_counter() - t0
        barrier()
        host_fps = n_host / host_s
        # device-resident ring: two batches already in HBM, scanned alternately for the same duration
        ring_d = [batch, F.noisy_batch_device(clean, Fn, seed=1007 + rank)]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        n_dev = 0
        while time.perf_counter() - t0 < args.stream_seconds:
            for _ in range(8):
                sc.scan_batch(ring_d[(n_dev // Fn) & 1], out)
                n_dev += Fn
            torch.cuda.current_stream().synchronize()  # bound the queue: the clock above is the host's
        e1.record()
        torch.cuda.synchronize()
        dev_fps = n_dev / (e0.elapsed_time(e1) * 1e-3)
        barrier()
        del ring_d
        stream = {"workload": f"BASELINE configs[4]: 1080p frames image-sharded over {world} B200, sustained >= {args.stream_seconds:.0f} s per leg",
                  "seconds_per_leg": args.stream_seconds,
                  "host_ring": {"what": f"two pinned host batches of {half} frames, svb_scan_batch_v1_host in a loop: H2D of every frame "
                                        "and D2H of the boards included", "frames_per_s_aggregate": sum_over_ranks(host_fps),
                                "frames_per_s_per_gpu_min": -max_over_ranks(-host_fps), "frames_per_s_per_gpu_max": max_over_ranks(host_fps)},
                  "device_ring": {"what": f"two device-resident batches of {Fn} frames scanned alternately (no H2D)",
                                  "frames_per_s_aggregate": sum_over_ranks(dev_fps),
                                  "frames_per_s_per_gpu_min": -max_over_ranks(-dev_fps), "frames_per_s_per_gpu_max": max_over_ranks(dev_fps)}}

    # final gather of the 81-digit boards (outside the timed region; the only collective)
    if world > 1:
        from svb200.shard import gather_boards

        boards = gather_boards(out["digits"], Fn * world)
        assert boards.shape == (Fn * world, 81)

    peak, tpeak, how = measured_peaks()
    parity = coreml = other_cfg = None
    if rank == 0 and args.parity_frames > 0:
        npar = min(args.parity_frames, Fn)
        idx = sorted(set(range(min(n_ref, npar))) | set(int(i) for i in np.linspace(0, Fn - 1, npar).astype(int)))[:max(npar, 1)]
        parity = parity_block(sc, batch, idx, weights, ref_results if cpu is not None else None, n_ref if cpu is not None else 0)
        coreml = parity_coreml(lambda w: Scanner(device=local, weights=w), batch, idx[:32])
    if rank == 0 and world == 1 and not args.no_other_configs:
        del host_in
        other_cfg = other_configs(sc, dev, tpeak)

    if rank == 0:
        def hbm(kernel, ms, bytes_per_frame, traffic_per_frame, note=None):
            ach = bytes_per_frame * Fn / (ms * 1e-3) / 1e9
            d = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "traffic": None if traffic_per_frame is None else traffic_per_frame * Fn, "peak_source": how,
                 "algorithmic_bytes_per_launch": bytes_per_frame * Fn, "launch_ms": ms}
            if note:
                d["note"] = note
            return d

        def tensor(kernel, ms, flop_per_cell, note):
            tf = flop_per_cell * 81 * Fn / (ms * 1e-3) / 1e12
            return {"bound": "tensor", "kernel": kernel, "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                    "traffic": None, "peak_source": how + ", bf16 dense sustained", "algorithmic_flops_per_launch": flop_per_cell * 81 * Fn,
                    "launch_ms": ms, "note": note}

        value = Fn * world * args.steps / (ms_total * 1e-3)
        k4_bytes = 3.0 * a_quad + 63504.0
        roofs = [
            hbm("k1w::fused_preprocess_warp_kernel", stage_ms["k1_preprocess"], K1_BYTES_PER_FRAME, NCU_TRAFFIC_PER_FRAME["k1"],
                "3HW read + HW written per frame (SURVEY 8d); in the whole-path call it also writes the 0.27 MB/frame bit mask K2 "
                "traces (neither in `traffic` nor in the algorithmic bytes)"),
            {"bound": "latency", "kernel": "k2::find_crossings + trace_segments + link_loops + select_quad", "launch_ms": stage_ms["k2_contour"],
             "us_per_frame": 1000.0 * stage_ms["k2_contour"] / Fn, "achieved": None, "peak": None, "unit": "us/frame", "frac": None, "traffic": None,
             "note": "border walks are chains of dependent steps (SURVEY 8d: latency / serial-bound, reported as us per frame)"},
            hbm("k4::cells_from_frames_kernel (+ k4::homography_kernel)", stage_ms["k34_cells"], k4_bytes, NCU_TRAFFIC_PER_FRAME["k4"],
                f"3 x A_quad source bytes + 81 x 784 cell bytes per frame (SURVEY 8d); A_quad = {a_quad:.0f} px, the mean shoelace "
                "area of the detected quads of this batch"),
            tensor("k5tc::tc_conv_kernel", stage_ms["k5_conv"], K5_CONV_FLOP_PER_CELL,
                   "conv1 + conv2 true MACs; operands are split fp16 hi+lo (three tcgen05 products per algorithmic product) to keep "
                   "logits within 1e-3 of fp32"),
            tensor("k5tc::tc_fc_tma_kernel", stage_ms["k5_fc"], K5_FC_FLOP_PER_CELL,
                   "fc1 + fc2 true MACs, same split; the kernel streams the conv features from HBM (12.5 KB per cell)"),
        ]
        timed_roofs = [r for r in roofs if r["bound"] != "latency"]
        dominant = max(timed_roofs, key=lambda r: r["launch_ms"])
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
                    "frames_per_step": En, "matches_device_path": same,
                    "h2d_gbs": En * H * W * 3 * world * args.steps / e2e_s / 1e9, "h2d_ceiling_gbs": h2d_ceiling,
                    "frac_of_ceiling": (En * H * W * 3 * world * args.steps / e2e_s / 1e9) / h2d_ceiling,
                    "ceiling_how": f"plain pinned cudaMemcpyAsync host->device of {nb} frames x4, all {world} rank(s) at once, wall clock max over ranks"},
            "e2e_jpeg": e2e_jpeg,
            "gpu_launches": int(launches),
            "stage_ms_per_step": {k: round(v, 4) for k, v in stage_ms.items()},
            "stage_ms_how": "CUDA events the library records between the stages (svb_stage_timing) of K extra steps, same launches as the timed steps",
            "ms_per_step_with_stage_events": ms_serial,
            "roofline": dominant, "rooflines": roofs, "parity": parity, "parity_coreml_weights": coreml,
            "stream": stream, "cpu_baseline": cpu, "clocks": clocks, "other_configs": other_cfg,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
