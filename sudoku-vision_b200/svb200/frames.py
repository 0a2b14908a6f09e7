"""Synthetic full-frame sudoku generator (host tooling for tests and bench.py; not on the hot path).

The reference ships only a 28x28 *cell* generator (ml/generate_synthetic.py:82-189) and five phone
photos; BASELINE.json's configs 2/4/5 ask for synthetic 1080p / 4K *frames*, so this compositor is
ours.  It follows the reference's style where there is one: black PIL-rendered digits on near-white
paper (generate_synthetic.py:82-123), per-frame seeding `random.seed(s); np.random.seed(s)` with
s = base_seed + index (generate_synthetic.py:291-292), 0 = empty cell.

A frame is a flat 9x9 board (thin/thick dark lines, ~40 % filled cells) mapped by a random
homography (side 0.55-0.85 of the frame height, rotation +-15 deg, corner jitter) onto a mid-gray
textured background, then 3x3-blurred; sensor noise N(0, sigma) and a lighting gradient are added
either on the host (`add_noise_host`) or, for large batches, on the device with torch
(`noisy_batch_device`) so that a 1024-frame 1080p batch does not cost minutes of host RNG.
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np

_FONT_CACHE: dict = {}


def _font(size: int):
    from PIL import ImageFont

    if size not in _FONT_CACHE:
        try:
            _FONT_CACHE[size] = ImageFont.load_default(size=size)
        except TypeError:  # very old Pillow: bitmap font only
            _FONT_CACHE[size] = ImageFont.load_default()
    return _FONT_CACHE[size]


@dataclass
class Frame:
    image: np.ndarray      # (H, W, 3) uint8 BGR, noise-free
    digits: np.ndarray     # (9, 9) uint8 ground truth, 0 = empty
    corners: np.ndarray    # (4, 2) float32 TL, TR, BR, BL of the board's outer border (ground truth)


def render_board(digits: np.ndarray, size: int, rng: random.Random) -> np.ndarray:
    """Flat board, (size, size) uint8 gray: paper, grid lines, digits."""
    from PIL import Image, ImageDraw

    paper = rng.randint(225, 250)
    img = Image.new("L", (size, size), color=paper)
    draw = ImageDraw.Draw(img)
    cell = size / 9.0
    thin = max(2, int(round(size / 450.0)))
    thick = max(4, int(round(size / 180.0)))
    ink = rng.randint(10, 50)
    for i in range(10):
        w = thick if i % 3 == 0 else thin
        p = int(round(i * cell))
        p0 = min(max(p - w // 2, 0), size - w)
        draw.rectangle([p0, 0, p0 + w - 1, size - 1], fill=ink)
        draw.rectangle([0, p0, size - 1, p0 + w - 1], fill=ink)
    fsize = int(cell * rng.uniform(0.62, 0.78))
    font = _font(fsize)
    for r in range(9):
        for c in range(9):
            d = int(digits[r, c])
            if d == 0:
                continue
            text = str(d)
            bbox = draw.textbbox((0, 0), text, font=font)
            tw, th = bbox[2] - bbox[0], bbox[3] - bbox[1]
            jx = rng.uniform(-0.05, 0.05) * cell
            jy = rng.uniform(-0.05, 0.05) * cell
            x = c * cell + (cell - tw) / 2 - bbox[0] + jx
            y = r * cell + (cell - th) / 2 - bbox[1] + jy
            draw.text((x, y), text, font=font, fill=rng.randint(0, 60))
    return np.array(img)


def random_digits(rng: random.Random, fill: float = 0.4) -> np.ndarray:
    d = np.zeros((9, 9), np.uint8)
    for r in range(9):
        for c in range(9):
            if rng.random() < fill:
                d[r, c] = rng.randint(1, 9)
    return d


def make_frame(seed: int, height: int = 1080, width: int = 1920, max_rot_deg: float = 15.0) -> Frame:
    """One noise-free synthetic frame; deterministic in `seed`."""
    import cv2

    rng = random.Random(seed)
    nrng = np.random.RandomState(seed & 0x7FFFFFFF)
    digits = random_digits(rng)
    bsize = int(round(min(height, width) * 0.85))
    board = render_board(digits, bsize, rng)

    # target quad: square of side s, rotated, jittered
    s = rng.uniform(0.55, 0.85) * min(height, width)
    ang = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg))
    cx = width / 2 + rng.uniform(-0.5, 0.5) * max(width - 1.25 * s, 0)
    cy = height / 2 + rng.uniform(-0.5, 0.5) * max(height - 1.25 * s, 0)
    base = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64) * (s / 2)
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    quad = base @ rot.T + np.array([cx, cy])
    quad += nrng.uniform(-0.06, 0.06, (4, 2)) * s
    quad[:, 0] = np.clip(quad[:, 0], 8, width - 9)
    quad[:, 1] = np.clip(quad[:, 1], 8, height - 9)
    quad = quad.astype(np.float32)

    # background: mid-gray with a smooth low-frequency texture
    bg_level = rng.randint(90, 180)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    tex = 6.0 * np.sin(xx / rng.uniform(90, 260) + rng.uniform(0, 6)) * np.cos(yy / rng.uniform(90, 260))
    bg = np.clip(bg_level + tex, 0, 255).astype(np.uint8)

    src = np.array([[0, 0], [bsize - 1, 0], [bsize - 1, bsize - 1], [0, bsize - 1]], np.float32)
    M = cv2.getPerspectiveTransform(src, quad)
    warped = cv2.warpPerspective(board, M, (width, height), flags=cv2.INTER_LINEAR, borderValue=0)
    cover = cv2.warpPerspective(np.full_like(board, 255), M, (width, height), flags=cv2.INTER_LINEAR, borderValue=0)
    a = cover.astype(np.float32) / 255.0
    gray = (warped.astype(np.float32) * a + bg.astype(np.float32) * (1 - a))
    gray = cv2.blur(np.clip(gray, 0, 255).astype(np.uint8), (3, 3))
    # slight colour cast so that BGR2GRAY weights matter
    tint = np.array([rng.uniform(0.94, 1.0), rng.uniform(0.96, 1.0), rng.uniform(0.94, 1.0)], np.float32)
    img = np.clip(gray[..., None].astype(np.float32) * tint, 0, 255).astype(np.uint8)
    return Frame(image=np.ascontiguousarray(img), digits=digits, corners=quad)


def add_noise_host(image: np.ndarray, seed: int, sigma: float = 3.0) -> np.ndarray:
    nrng = np.random.RandomState((seed * 7919 + 13) & 0x7FFFFFFF)
    noise = nrng.normal(0.0, sigma, image.shape).astype(np.float32)
    return np.clip(image.astype(np.float32) + noise, 0, 255).astype(np.uint8)


def make_frames(n: int, height: int = 1080, width: int = 1920, base_seed: int = 1000, sigma: float = 3.0,
                max_rot_deg: float = 15.0):
    """n noisy frames as (n,H,W,3) uint8 + digits (n,9,9) + corners (n,4,2); host only."""
    imgs = np.empty((n, height, width, 3), np.uint8)
    dig = np.empty((n, 9, 9), np.uint8)
    cor = np.empty((n, 4, 2), np.float32)
    for i in range(n):
        f = make_frame(base_seed + i, height, width, max_rot_deg)
        imgs[i] = add_noise_host(f.image, base_seed + i, sigma)
        dig[i] = f.digits
        cor[i] = f.corners
    return imgs, dig, cor


def noisy_batch_device(clean, n: int, seed: int = 0, sigma: float = 3.0):
    """Expand `clean` (k,H,W,3) uint8 CUDA tensor to n frames on the device: frame i is
    clean[i % k] + N(0, sigma) + a per-frame brightness offset, so every frame is distinct."""
    import torch

    k = clean.shape[0]
    g = torch.Generator(device=clean.device)
    g.manual_seed(seed)
    out = torch.empty((n,) + tuple(clean.shape[1:]), dtype=torch.uint8, device=clean.device)
    chunk = 16
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        idx = torch.arange(s, e, device=clean.device) % k
        x = clean[idx].to(torch.float16)
        x += torch.randn(x.shape, generator=g, device=clean.device, dtype=torch.float16) * sigma
        x += (torch.rand((e - s, 1, 1, 1), generator=g, device=clean.device, dtype=torch.float16) - 0.5) * 12.0
        out[s:e] = x.clamp_(0, 255).round_().to(torch.uint8)
    return out
