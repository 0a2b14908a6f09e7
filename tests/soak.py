"""Parity soak on the GPU box: many frames of several sizes / rotations through the product path, every stage against the
C oracle (test infrastructure) on all host cores.  Rare-event hunting — the 1024-frame sweep of round 2 found a one-pixel
difference in 4 of 82,944 cells that no smaller test saw.  Prints one summary line per section; mismatching inputs are
dumped to gpurun_out/soak_*.npz.  Usage: python tests/soak.py [--frames-per-size 192] [--cells 300000] [--sections v1,cells,k1,jpeg]"""
import argparse
import concurrent.futures as cf
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sudoku-vision_b200")]
from oracle import oracle as O  # noqa: E402
from svb200 import Scanner, load_digitcnn_weights  # noqa: E402
from svb200 import frames as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames-per-size", type=int, default=192)
ap.add_argument("--unique", type=int, default=24)
ap.add_argument("--cells", type=int, default=300000)
ap.add_argument("--v2-frames", type=int, default=48)
ap.add_argument("--seed-offset", type=int, default=0, help="shifts every seed: another draw of frames, cells and noise")
ap.add_argument("--sections", default="v1,cells,k1,jpeg,v2")
args = ap.parse_args()
sections = args.sections.split(",")
NCPU = max(1, len(os.sched_getaffinity(0)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
w = load_digitcnn_weights()
sc = Scanner(device=0, weights=w)
pool = cf.ThreadPoolExecutor(max_workers=NCPU)


def bits_to_pm1(bits):
    return np.where((bits[..., None] >> np.arange(28)) & 1, 1.0, -1.0).astype(np.float32)


def v1_size(h, wd, rot, seed):
    n, k = args.frames_per_size, args.unique
    t0 = time.time()
    clean = np.stack(list(pool.map(lambda i: F.make_frame(seed + i, h, wd, max_rot_deg=rot).image, range(k))))
    batch = F.noisy_batch_device(torch.from_numpy(clean).cuda(), n, seed=seed, sigma=2.0 + (seed % 5))
    out = sc.scan_batch(batch, want_logits=True)
    mask = sc.preprocess(batch)
    torch.cuda.synchronize()
    got = {kk: v.cpu().numpy() for kk, v in out.items()}
    frames, masks = batch.cpu().numpy(), mask.cpu().numpy()

    def one(i):
        r = O.scan_frame(frames[i])
        lg = O.digitcnn_forward(w, (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5) if r["found"] else None
        return r, lg

    res = list(pool.map(one, range(n)))
    st = dict(frames=n, found=0, mask_bad=0, found_bad=0, corners_bad=0, logits_bad=0, digits_bad=0, bits_bad=0, maxd=0.0)
    oc = np.zeros((n, 4, 2), np.int32)
    of = np.zeros(n, np.uint8)
    bad = []
    for i, (r, lg) in enumerate(res):
        if not np.array_equal(masks[i], r["mask"]):
            st["mask_bad"] += 1
            bad.append(i)
        if bool(got["found"][i] == 1) != r["found"]:
            st["found_bad"] += 1
            bad.append(i)
            continue
        if not r["found"]:
            continue
        st["found"] += 1
        oc[i], of[i] = r["corners"], 1
        if not np.array_equal(got["corners"][i], r["corners"]):
            st["corners_bad"] += 1
            bad.append(i)
            continue
        d = float(np.abs(got["logits"][i] - lg).max())
        st["maxd"] = max(st["maxd"], d)
        if d >= 1e-3:
            st["logits_bad"] += 1
            bad.append(i)
        if not np.array_equal(got["digits"][i], lg.argmax(1).astype(np.uint8)):
            st["digits_bad"] += 1
    # K4 alone on the oracle's corners: bit rows and the float tensor
    bits = sc.cells_from_frames_bits(batch, torch.from_numpy(oc).cuda(), torch.from_numpy(of).cuda())
    u8, pm1 = sc.cells_from_frames(batch, torch.from_numpy(oc).cuda(), torch.from_numpy(of).cuda())
    torch.cuda.synchronize()
    bits, u8, pm1 = bits.cpu().numpy(), u8.cpu().numpy(), pm1.cpu().numpy()
    for i, (r, lg) in enumerate(res):
        if not r["found"]:
            continue
        want = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        if not (np.array_equal(bits_to_pm1(bits[i]), want) and np.array_equal(pm1[i], want) and np.array_equal(u8[i], r["cells_u8"])):
            st["bits_bad"] += 1
            bad.append(i)
    bad = sorted(set(bad))
    for i in bad[:3]:
        np.savez_compressed(os.path.join(OUT, f"soak_v1_{h}x{wd}_{seed}_{i}.npz"), frame=frames[i], gpu_corners=got["corners"][i],
                            gpu_found=got["found"][i], gpu_mask=masks[i])
    print(f"v1 {h}x{wd} rot<={rot} seed {seed}: {st} bad frames {bad[:10]} ({time.time() - t0:.0f} s)", flush=True)
    return len(bad)


def cells_section():
    t0 = time.time()
    rng = np.random.default_rng(77 + args.seed_offset)
    n = args.cells
    nbad = 0
    for s in range(0, n, 50000):
        m = min(50000, n - s)
        tex = (rng.normal(150, 30, (m, 1, 1)) + rng.normal(0, 1, (m, 28, 28)) * rng.uniform(1, 40, (m, 1, 1))
               + np.linspace(-1, 1, 28)[None, None, :] * rng.normal(0, 25, (m, 1, 1))
               + np.linspace(-1, 1, 28)[None, :, None] * rng.normal(0, 25, (m, 1, 1)))
        # a dark stroke through half of the cells, as digits have
        stroke = (np.abs(np.arange(28)[None, None, :] - rng.integers(4, 24, (m, 1, 1))) < rng.integers(1, 4, (m, 1, 1))) & (rng.random((m, 1, 1)) < 0.5)
        cells = np.clip(tex - 90 * stroke, 0, 255).astype(np.uint8)
        thr, pm1 = sc.cell_prep(torch.from_numpy(cells).cuda())
        parts = list(pool.map(O.cell_prep, np.array_split(cells, NCPU)))
        ink = np.concatenate(parts)
        torch.cuda.synchronize()
        b = np.nonzero(((thr.cpu().numpy() != 255 - ink) | (pm1.cpu().numpy() != np.where(ink == 255, 1.0, -1.0))).any((1, 2)))[0]
        if len(b):
            np.savez_compressed(os.path.join(OUT, f"soak_cells_{s}.npz"), cells=cells[b[:64]])
        nbad += len(b)
    print(f"cells: {n} textured cells through cell_prep, {nbad} differ from the oracle ({time.time() - t0:.0f} s)", flush=True)
    return nbad


def k1_section():
    """smooth images (many rounding ties) and raw noise at widths that take the fused kernel and the stage kernels"""
    t0 = time.time()
    nbad = tot = 0
    rng = np.random.default_rng(3)
    for (h, wd) in [(1080, 1920), (720, 1280), (1088, 1936), (487, 1000), (2160, 3840), (600, 808)]:
        yy, xx = np.mgrid[0:h, 0:wd]
        imgs = []
        for j in range(6):
            a, b, c = rng.uniform(15, 80), rng.uniform(7, 40), rng.uniform(10, 60)
            base = np.stack([120 + 60 * np.sin(xx / a) + 20 * np.cos(yy / b), 140 + 50 * np.sin((xx + yy) / c), 100 + 70 * np.cos(xx / b) * np.sin(yy / a)], -1)
            imgs.append(np.clip(base + rng.normal(0, j, base.shape), 0, 255).astype(np.uint8))
        imgs = np.stack(imgs)
        m = sc.preprocess(torch.from_numpy(imgs).cuda()).cpu().numpy()
        want = list(pool.map(O.preprocess, imgs))
        for i in range(len(imgs)):
            tot += 1
            if not np.array_equal(m[i], want[i]):
                nbad += 1
                np.savez_compressed(os.path.join(OUT, f"soak_k1_{h}x{wd}_{i}.npz"), img=imgs[i])
    print(f"k1: {tot} smooth / noisy images at six sizes, {nbad} masks differ from the oracle ({time.time() - t0:.0f} s)", flush=True)
    return nbad


def jpeg_section():
    import cv2

    t0 = time.time()
    nbad = tot = 0
    for (h, wd) in [(1080, 1920), (720, 1280), (540, 960)]:
        for q in (60, 85, 95, 100):
            imgs = list(pool.map(lambda i: F.make_frame(5000 + 13 * q + i, h, wd, max_rot_deg=30).image, range(16)))
            rng = np.random.default_rng(q)
            imgs = [np.clip(im.astype(np.int16) + rng.integers(-6, 7, im.shape), 0, 255).astype(np.uint8) for im in imgs]
            for dri in (8, 1):
                files = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, dri])[1].tobytes() for im in imgs]
                blob, offs = sc.pack_jpegs(files)
                dec = sc.jpeg_decode(blob, offs, h, wd).cpu().numpy()
                for i, fb in enumerate(files):
                    tot += 1
                    if not np.array_equal(dec[i], cv2.imdecode(np.frombuffer(fb, np.uint8), cv2.IMREAD_COLOR)):
                        nbad += 1
    print(f"jpeg: {tot} files (three sizes, four qualities, DRI 8 / 1), {nbad} decode differently from cv2.imdecode ({time.time() - t0:.0f} s)", flush=True)
    return nbad


def v2_section():
    """svb_scan_batch_v2's CV stages (preprocess_multi_strategy with per-frame glare / shadow flags, contour + validity) against
    the oracle's composition on frames with synthetic shadows and glare spots"""
    from oracle import oracle_v2 as O2

    t0 = time.time()
    nbad = tot = 0
    rng = np.random.default_rng(21 + args.seed_offset)
    for (h, wd) in [(544, 960), (720, 1280)]:
        n = args.v2_frames
        base = list(pool.map(lambda i: F.make_frame(7000 + h + i + 100000 * args.seed_offset, h, wd, max_rot_deg=25).image, range(n)))
        yy, xx = np.mgrid[0:h, 0:wd].astype(np.float32)
        imgs = []
        for i, im in enumerate(base):
            f = im.astype(np.float32)
            kind = i % 4
            if kind in (1, 3):  # shadow: a smooth darkening ramp across the frame
                ang = rng.uniform(0, 6.28)
                ramp = ((xx - wd / 2) * np.cos(ang) + (yy - h / 2) * np.sin(ang)) / max(h, wd)
                f *= np.clip(0.75 + rng.uniform(0.6, 1.0) * ramp, 0.35, 1.0)[..., None]
            if kind in (2, 3):  # glare: a saturating blob
                cx, cy, rad = rng.uniform(0.2, 0.8) * wd, rng.uniform(0.2, 0.8) * h, rng.uniform(0.05, 0.15) * wd
                f += 200.0 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * rad * rad))[..., None]
            f += rng.normal(0, 2.0, f.shape)
            imgs.append(np.clip(f, 0, 255).astype(np.uint8))
        imgs = np.stack(imgs)
        out = sc.scan_batch_v2(torch.from_numpy(imgs).cuda(), want_logits=False)
        pm = sc.preprocess_multi(torch.from_numpy(imgs).cuda(), want_aux=False)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in out.items() if hasattr(v, "cpu")}
        gbin = pm["binary"].cpu().numpy()

        def one(i):
            r = O.preprocess_multi(imgs[i])
            return r, O2.detect_grid_contour(r["binary"])

        res = list(pool.map(one, range(n)))
        st = dict(frames=n, found=0, glare=0, shadow=0, method_bad=0, binary_bad=0, found_bad=0, corners_bad=0)
        for i, (r, c) in enumerate(res):
            tot += 1
            st["glare"] += int(bool(r.get("has_glare", False)))
            st["shadow"] += int(bool(r.get("has_shadow", False)))
            bad = False
            if sc.V2_METHODS[int(got["info"][i][2])] != r["method_used"]:
                st["method_bad"] += 1
                bad = True
            if not np.array_equal(gbin[i], r["binary"]):
                st["binary_bad"] += 1
                bad = True
            if bool(got["found"][i] == 1) != (c is not None):
                st["found_bad"] += 1
                bad = True
            elif c is not None:
                st["found"] += 1
                if not np.array_equal(got["corners"][i].astype(np.float32), c):
                    st["corners_bad"] += 1
                    bad = True
            if bad:
                nbad += 1
                if nbad <= 3:
                    np.savez_compressed(os.path.join(OUT, f"soak_v2_{h}x{wd}_{i}.npz"), frame=imgs[i])
        print(f"v2 {h}x{wd}: {st}", flush=True)
    print(f"v2: {tot} frames with shadows / glare, {nbad} differ from the oracle ({time.time() - t0:.0f} s)", flush=True)
    return nbad


total = 0
if "v1" in sections:
    for k, (h, wd, rot) in enumerate([(1080, 1920, 15.0), (1080, 1920, 40.0), (720, 1280, 30.0), (540, 960, 25.0), (480, 640, 35.0),
                                      (750, 1000, 20.0), (1200, 1600, 30.0), (2160, 3840, 20.0)]):
        total += v1_size(h, wd, rot, 41000 + 1000 * k + 100000 * args.seed_offset)
if "cells" in sections:
    total += cells_section()
if "k1" in sections:
    total += k1_section()
if "jpeg" in sections:
    total += jpeg_section()
if "v2" in sections:
    sc.load_weights_v3(__import__("svb200.v3_init", fromlist=["random_v3_state"]).random_v3_state())
    total += v2_section()
print(f"soak: {total} mismatches in total")
