"""Batched device-resident API over the C ABI.  Tensors in, tensors out, everything on the caller's
current CUDA stream; nothing here computes — it only passes pointers."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_WEIGHT_KEYS = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
                "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")
_WEIGHT_SHAPES = ((32, 1, 3, 3), (32,), (64, 32, 3, 3), (64,), (128, 3136), (128,), (10, 128), (10,))


def default_weights_path() -> str:
    return os.path.join(_HERE, "weights", "digitcnn_synth.npz")


def coreml_weights_path() -> str:
    """the only trained DigitCNN weights the reference ships (fp16, recovered from its iOS CoreML package: SURVEY.md 8c;
    tests/golden/make_helpers_golden.py): MNIST-era, brittle margins — the stress set for argmax parity"""
    return os.path.join(_HERE, "weights", "digitcnn_coreml.npz")


def load_digitcnn_weights(path: str | None = None) -> dict:
    """state_dict-shaped dict of float32 numpy arrays (keys of ml/model.py:22-32)."""
    z = np.load(path or default_weights_path())
    return {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in _WEIGHT_KEYS}


def _torch():
    import torch

    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Scanner:
    """One context per (process, device): owns scratch and packed weights inside the library."""

    def __init__(self, device: int | None = None, weights: dict | None = None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.SvbError("svb200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        _lib.check(self.lib.svb_create(self.device, C.byref(h)), "svb_create")
        self._h = h
        self._weights_dev = None
        if weights is not None:
            self.load_weights(weights)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.svb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        return _torch().device("cuda", self.device)

    def _chk_u8(self, t, ndim, what):
        torch = _torch()
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()
                and t.dim() == ndim and t.device.index == self.device):
            raise ValueError(f"{what}: expected a contiguous uint8 CUDA tensor with {ndim} dims on cuda:{self.device}")
        return t

    @property
    def launches(self) -> int:
        return int(self.lib.svb_launch_count(self._h))

    OPTIONS = dict(overlap=1)

    def set_option(self, name: str, value: int):
        """'overlap' (default 0): p >= 2 makes scan_batch run p sub-batches on two internal streams (see include/svb200.h)."""
        _lib.check(self.lib.svb_set_option(self._h, self.OPTIONS[name], int(value)), "svb_set_option")

    STAGES = ("k1_preprocess", "k2_contour", "k34_cells", "k5_conv", "k5_fc")

    def stage_timing(self, enable: bool = True):
        _lib.check(self.lib.svb_stage_timing(self._h, int(enable)), "svb_stage_timing")

    def last_stage_ms(self) -> dict:
        ms = (C.c_float * len(self.STAGES))()
        _lib.check(self.lib.svb_last_stage_ms(self._h, ms), "svb_last_stage_ms")
        return dict(zip(self.STAGES, [float(x) for x in ms]))

    # -- classifier ------------------------------------------------------------------------------
    def load_weights(self, sd: dict):
        """sd: mapping with the parameter names of ml/model.py (numpy arrays or torch tensors)."""
        torch = _torch()
        ts = []
        for k, shp in zip(_WEIGHT_KEYS, _WEIGHT_SHAPES):
            v = sd[k]
            t = v.detach() if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            t = t.to(device=self._dev(), dtype=torch.float32).contiguous()
            if tuple(t.shape) != shp:
                raise ValueError(f"{k}: expected shape {shp}, got {tuple(t.shape)}")
            ts.append(t)
        _lib.check(self.lib.svb_digitcnn_load(self._h, *[_ptr(t) for t in ts], self._stream()), "svb_digitcnn_load")
        self._weights_dev = ts  # keep alive until the pack kernel has run
        return self

    def set_classifier_mode(self, mode: str):
        """'tc' = tcgen05/TMEM kernels (default), 'fp32' = CUDA-core fp32 kernels."""
        _lib.check(self.lib.svb_set_classifier_mode(self._h, {"tc": 0, "fp32": 1}[mode]), "svb_set_classifier_mode")

    def digitcnn_forward(self, x, want_digits: bool = False):
        """x: (B,1,28,28) float32 CUDA -> logits (B,10) [, digits (B,) u8, conf (B,) f32]."""
        torch = _torch()
        x = x.to(device=self._dev(), dtype=torch.float32).contiguous()
        n = x.shape[0]
        logits = torch.empty((n, 10), dtype=torch.float32, device=x.device)
        digits = torch.empty((n,), dtype=torch.uint8, device=x.device) if want_digits else None
        conf = torch.empty((n,), dtype=torch.float32, device=x.device) if want_digits else None
        _lib.check(self.lib.svb_digitcnn_forward(self._h, _ptr(x), n, _ptr(logits), _ptr(digits), _ptr(conf),
                                                 self._stream()), "svb_digitcnn_forward")
        return (logits, digits, conf) if want_digits else logits

    def digitcnn_forward_bits(self, bits, want_digits: bool = False):
        """bits: (B,28) or (...,28) int32/uint32 CUDA bit rows of +-1 cells -> logits (B,10) [, digits, conf]."""
        torch = _torch()
        if not (bits.is_cuda and bits.dtype in (torch.int32, torch.uint32) and bits.is_contiguous() and bits.shape[-1] == 28):
            raise ValueError("digitcnn_forward_bits: expected contiguous int32 CUDA bit rows (..., 28)")
        n = bits.numel() // 28
        logits = torch.empty((n, 10), dtype=torch.float32, device=bits.device)
        digits = torch.empty((n,), dtype=torch.uint8, device=bits.device) if want_digits else None
        conf = torch.empty((n,), dtype=torch.float32, device=bits.device) if want_digits else None
        _lib.check(self.lib.svb_digitcnn_forward_bits(self._h, _ptr(bits), n, _ptr(logits), _ptr(digits), _ptr(conf),
                                                      self._stream()), "svb_digitcnn_forward_bits")
        return (logits, digits, conf) if want_digits else logits

    def pack_cells_bits(self, pm1):
        """+-1 float cells (..., 28, 28) CUDA -> bit rows (..., 28) int32 (bit x of row y set <=> pixel > 0)."""
        torch = _torch()
        pm1 = pm1.to(device=self._dev(), dtype=torch.float32).contiguous()
        n = pm1.numel() // 784
        bits = torch.empty(tuple(pm1.shape[:-1]), dtype=torch.int32, device=pm1.device)
        _lib.check(self.lib.svb_pack_cells_bits(self._h, _ptr(pm1), n, _ptr(bits), self._stream()), "svb_pack_cells_bits")
        return bits

    # -- DigitCNNv3 ------------------------------------------------------------------------------------
    def load_weights_v3(self, sd: dict):
        """sd: state_dict of ml/model_v3.DigitCNNv3 (torch tensors or numpy).  Folds BatchNorm (running stats,
        eps 1e-5) into the convolutions — a one-off load-time step — and hands the 38 folded tensors to the library."""
        torch = _torch()

        def t(k):
            v = sd[k]
            v = v.detach() if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            return v.to(device=self._dev(), dtype=torch.float32)

        def fold(conv, bn):
            g = t(bn + ".weight") / torch.sqrt(t(bn + ".running_var") + 1e-5)
            w = t(conv + ".weight") * g.view(-1, 1, 1, 1)
            b = t(bn + ".bias") - t(bn + ".running_mean") * g
            return [w.contiguous(), b.contiguous()]

        ts = fold("stem.0", "stem.1")
        for L in range(1, 6):
            p = f"layer{L}"
            ts += fold(p + ".conv1", p + ".bn1") + fold(p + ".conv2", p + ".bn2")
            ts += [t(p + ".se.excite.0.weight").contiguous(), t(p + ".se.excite.2.weight").contiguous()]
            if L in (2, 4):
                ts += fold(p + ".shortcut.0", p + ".shortcut.1")
        ts += [t("fc.weight").contiguous(), t("fc.bias").contiguous()]
        arr = (C.c_void_p * len(ts))(*[x.data_ptr() for x in ts])
        _lib.check(self.lib.svb_digitcnn_v3_load(self._h, arr, len(ts), self._stream()), "svb_digitcnn_v3_load")
        self._weights_v3_dev = ts
        return self

    def digitcnn_v3_forward(self, x, want_digits: bool = False, want_features: bool = False):
        torch = _torch()
        x = x.to(device=self._dev(), dtype=torch.float32).contiguous()
        n = x.shape[0]
        logits = torch.empty((n, 10), dtype=torch.float32, device=x.device)
        digits = torch.empty((n,), dtype=torch.uint8, device=x.device) if want_digits else None
        conf = torch.empty((n,), dtype=torch.float32, device=x.device) if want_digits else None
        feats = torch.empty((n, 128), dtype=torch.float32, device=x.device) if want_features else None
        _lib.check(self.lib.svb_digitcnn_v3_forward(self._h, _ptr(x), n, _ptr(logits), _ptr(digits), _ptr(conf), _ptr(feats),
                                                    self._stream()), "svb_digitcnn_v3_forward")
        if want_features:
            return feats
        return (logits, digits, conf) if want_digits else logits

    # -- image stages (batched, device tensors) --------------------------------------------------
    def grayscale(self, bgr):
        self._chk_u8(bgr, 4, "grayscale")
        n, h, w, _ = bgr.shape
        out = _torch().empty((n, h, w), dtype=_torch().uint8, device=bgr.device)
        _lib.check(self.lib.svb_grayscale(self._h, _ptr(bgr), n, h, w, _ptr(out), self._stream()), "svb_grayscale")
        return out

    def blur(self, gray, ksize: int = 5):
        self._chk_u8(gray, 3, "blur")
        n, h, w = gray.shape
        out = _torch().empty_like(gray)
        _lib.check(self.lib.svb_blur(self._h, _ptr(gray), n, h, w, int(ksize), _ptr(out), self._stream()), "svb_blur")
        return out

    def adaptive_threshold(self, gray, block_size: int = 11, c: int = 2, inverted: bool = True):
        self._chk_u8(gray, 3, "adaptive_threshold")
        n, h, w = gray.shape
        out = _torch().empty_like(gray)
        _lib.check(self.lib.svb_adaptive_threshold(self._h, _ptr(gray), n, h, w, int(block_size), int(c),
                                                   int(bool(inverted)), _ptr(out), self._stream()),
                   "svb_adaptive_threshold")
        return out

    def preprocess(self, bgr, out=None):
        self._chk_u8(bgr, 4, "preprocess")
        n, h, w, _ = bgr.shape
        if out is None:
            out = _torch().empty((n, h, w), dtype=_torch().uint8, device=bgr.device)
        _lib.check(self.lib.svb_preprocess_v1(self._h, _ptr(bgr), n, h, w, _ptr(out), self._stream()),
                   "svb_preprocess_v1")
        return out

    # -- cv/preprocess_v2.py (row V1) ----------------------------------------------------------------
    V2_OPS = dict(normalize_illumination=1, detect_glare=2, detect_shadow=3, remove_shadow=4, clahe8=5, otsu=6,
                  sauvola=7, cleanup=8, dilate_ellipse=9, erode_ellipse=10, box_blur=11, gauss21=12)
    V2_METHODS = ("adaptive", "otsu", "sauvola")

    def _frames_nhw(self, frames, what):
        """(n,H,W,3) BGR or (n,H,W) gray uint8 CUDA -> n, h, w, channels"""
        if frames.dim() == 4:
            self._chk_u8(frames, 4, what)
            if frames.shape[3] != 3:
                raise ValueError(f"{what}: colour frames must be (n,H,W,3)")
            return frames.shape[0], frames.shape[1], frames.shape[2], 3
        self._chk_u8(frames, 3, what)
        return frames.shape[0], frames.shape[1], frames.shape[2], 1

    def preprocess_v2(self, frames, use_illumination_norm: bool = True, use_shadow_removal: bool = True):
        """cv/preprocess_v2.py:205-244 for a batch -> (mask (n,H,W) u8, info (n,4) u8 = glare, shadow, -, -)."""
        torch = _torch()
        n, h, w, ch = self._frames_nhw(frames, "preprocess_v2")
        mask = torch.empty((n, h, w), dtype=torch.uint8, device=frames.device)
        info = torch.empty((n, 4), dtype=torch.uint8, device=frames.device)
        _lib.check(self.lib.svb_preprocess_v2(self._h, _ptr(frames), n, h, w, ch, int(bool(use_illumination_norm)),
                                              int(bool(use_shadow_removal)), _ptr(mask), _ptr(info), self._stream()),
                   "svb_preprocess_v2")
        return mask, info

    def preprocess_multi(self, frames, want_aux: bool = True) -> dict:
        """cv/preprocess_v2.py:247-308 for a batch: dict(binary, gray, enhanced, illumination_normalized, info)."""
        torch = _torch()
        n, h, w, ch = self._frames_nhw(frames, "preprocess_multi")

        def plane():
            return torch.empty((n, h, w), dtype=torch.uint8, device=frames.device)

        out = dict(binary=plane(), gray=plane() if want_aux else None, enhanced=plane() if want_aux else None,
                   illumination_normalized=plane() if want_aux else None,
                   info=torch.empty((n, 4), dtype=torch.uint8, device=frames.device))
        _lib.check(self.lib.svb_preprocess_multi_v2(self._h, _ptr(frames), n, h, w, ch, _ptr(out["binary"]), _ptr(out["gray"]),
                                                    _ptr(out["enhanced"]), _ptr(out["illumination_normalized"]),
                                                    _ptr(out["info"]), self._stream()), "svb_preprocess_multi_v2")
        return out

    def v2_stage(self, op: str, src, arg: int = 0, want_image: bool = True):
        """One function of cv/preprocess_v2.py on (n,H,W) u8 images -> (image or None, info (n,4) u8)."""
        torch = _torch()
        self._chk_u8(src, 3, "v2_stage")
        n, h, w = src.shape
        dst = torch.empty_like(src) if want_image else None
        info = torch.zeros((n, 4), dtype=torch.uint8, device=src.device)
        _lib.check(self.lib.svb_v2_stage(self._h, self.V2_OPS[op], _ptr(src), n, h, w, int(arg), _ptr(dst), _ptr(info),
                                         self._stream()), f"svb_v2_stage({op})")
        return dst, info

    def find_grid_contour(self, mask, min_area_ratio: float = 0.1, eps_ratio: float = 0.02):
        self._chk_u8(mask, 3, "find_grid_contour")
        torch = _torch()
        n, h, w = mask.shape
        corners = torch.empty((n, 4, 2), dtype=torch.int32, device=mask.device)
        found = torch.empty((n,), dtype=torch.uint8, device=mask.device)
        _lib.check(self.lib.svb_find_grid_contour(self._h, _ptr(mask), n, h, w, float(min_area_ratio),
                                                  float(eps_ratio), _ptr(corners), _ptr(found), self._stream()),
                   "svb_find_grid_contour")
        return corners, found

    def find_contours(self, mask):
        """cv/grid.py:16-21 for ONE mask (H,W) u8 CUDA: (points (P,2) int32, offsets (C+1,) int64), cv2's order."""
        self._chk_u8(mask, 2, "find_contours")
        torch = _torch()
        h, w = mask.shape
        nc, npnt = C.c_longlong(0), C.c_longlong(0)
        _lib.check(self.lib.svb_find_contours_count(self._h, _ptr(mask), h, w, C.byref(nc), C.byref(npnt), self._stream()),
                   "svb_find_contours_count")
        pts = torch.empty((max(npnt.value, 1), 2), dtype=torch.int32, device=mask.device)
        offs = torch.empty((nc.value + 1,), dtype=torch.int64, device=mask.device)
        _lib.check(self.lib.svb_find_contours_fetch(self._h, _ptr(mask), h, w, _ptr(pts), _ptr(offs), self._stream()),
                   "svb_find_contours_fetch")
        return pts[:npnt.value], offs

    def approx_poly_dp(self, contour, epsilon_ratio: float = 0.02):
        """cv/grid.py:24-34 for one contour (n,2) int32 CUDA -> polygon (m,2) int32."""
        torch = _torch()
        contour = contour.to(device=self._dev(), dtype=torch.int32).contiguous().view(-1, 2)
        n = contour.shape[0]
        out = torch.empty((n, 2), dtype=torch.int32, device=contour.device)
        m = torch.zeros((1,), dtype=torch.int32, device=contour.device)
        _lib.check(self.lib.svb_approx_poly_dp(self._h, _ptr(contour), n, float(epsilon_ratio), _ptr(out), _ptr(m),
                                               self._stream()), "svb_approx_poly_dp")
        mv = int(m.item())
        if mv == -1:
            raise NotImplementedError("approx_poly_dp: coordinates outside [0, 65535] are not implemented")
        if mv < 0:
            raise _lib.SvbError("approx_poly_dp: scratch exhausted on the GPU")
        return out[:mv]

    def is_cell_empty(self, cells, threshold: float = 0.02, want_info: bool = False):
        """cv/extract.py:59-79 for a batch of gray cells (n,h,w) u8 CUDA -> empty (n,) u8 [, info (n,2) int32]."""
        self._chk_u8(cells, 3, "is_cell_empty")
        torch = _torch()
        n, ch, cw = cells.shape
        empty = torch.empty((n,), dtype=torch.uint8, device=cells.device)
        info = torch.empty((n, 2), dtype=torch.int32, device=cells.device) if want_info else None
        _lib.check(self.lib.svb_is_cell_empty(self._h, _ptr(cells), n, ch, cw, float(threshold), _ptr(empty), _ptr(info),
                                              self._stream()), "svb_is_cell_empty")
        return (empty, info) if want_info else empty

    def detect_grid_contour_v2(self, mask, min_area_ratio: float = 0.1):
        """cv/grid_v2.py:102-128 (method 1 of detect_grid): ordered corners (n,4,2) int32 + found (n,)."""
        self._chk_u8(mask, 3, "detect_grid_contour_v2")
        torch = _torch()
        n, h, w = mask.shape
        corners = torch.empty((n, 4, 2), dtype=torch.int32, device=mask.device)
        found = torch.empty((n,), dtype=torch.uint8, device=mask.device)
        _lib.check(self.lib.svb_detect_grid_contour_v2(self._h, _ptr(mask), n, h, w, float(min_area_ratio), _ptr(corners),
                                                       _ptr(found), self._stream()), "svb_detect_grid_contour_v2")
        return corners, found

    def warp_perspective(self, bgr, corners, found=None, out_size: int = 450):
        self._chk_u8(bgr, 4, "warp_perspective")
        torch = _torch()
        n, h, w, _ = bgr.shape
        corners = corners.to(device=bgr.device, dtype=torch.int32).contiguous()
        board = torch.empty((n, out_size, out_size, 3), dtype=torch.uint8, device=bgr.device)
        _lib.check(self.lib.svb_warp_perspective(self._h, _ptr(bgr), n, h, w, _ptr(corners), _ptr(found),
                                                 int(out_size), _ptr(board), self._stream()), "svb_warp_perspective")
        return board

    def extract_cells(self, board):
        self._chk_u8(board, 4, "extract_cells")
        torch = _torch()
        n, s, s2, ch = board.shape
        if s != s2 or ch != 3:
            raise ValueError("extract_cells: board must be (n, S, S, 3)")
        cells = torch.empty((n, 81, 28, 28), dtype=torch.uint8, device=board.device)
        _lib.check(self.lib.svb_extract_cells(self._h, _ptr(board), n, s, _ptr(cells), self._stream()),
                   "svb_extract_cells")
        return cells

    def cell_prep(self, cells, want_thresh: bool = True, want_pm1: bool = True):
        torch = _torch()
        if not (cells.is_cuda and cells.dtype == torch.uint8 and cells.is_contiguous() and cells.shape[-2:] == (28, 28)):
            raise ValueError("cell_prep: expected contiguous uint8 CUDA cells (..., 28, 28)")
        n = cells.numel() // 784
        thr = torch.empty_like(cells) if want_thresh else None
        pm1 = torch.empty(cells.shape, dtype=torch.float32, device=cells.device) if want_pm1 else None
        _lib.check(self.lib.svb_cell_prep(self._h, _ptr(cells), n, _ptr(thr), _ptr(pm1), self._stream()), "svb_cell_prep")
        return thr, pm1

    def cells_from_frames(self, bgr, corners, found=None, want_u8: bool = True):
        self._chk_u8(bgr, 4, "cells_from_frames")
        torch = _torch()
        n, h, w, _ = bgr.shape
        corners = corners.to(device=bgr.device, dtype=torch.int32).contiguous()
        u8 = torch.empty((n, 81, 28, 28), dtype=torch.uint8, device=bgr.device) if want_u8 else None
        pm1 = torch.empty((n, 81, 28, 28), dtype=torch.float32, device=bgr.device)
        _lib.check(self.lib.svb_cells_from_frames(self._h, _ptr(bgr), n, h, w, _ptr(corners), _ptr(found), _ptr(u8),
                                                  _ptr(pm1), self._stream()), "svb_cells_from_frames")
        return u8, pm1

    def cells_from_frames_bits(self, bgr, corners, found=None):
        """the batched path's K4: (n,H,W,3) frames + corners -> classifier input as bit rows (n,81,28) int32."""
        self._chk_u8(bgr, 4, "cells_from_frames_bits")
        torch = _torch()
        n, h, w, _ = bgr.shape
        corners = corners.to(device=bgr.device, dtype=torch.int32).contiguous()
        bits = torch.empty((n, 81, 28), dtype=torch.int32, device=bgr.device)
        _lib.check(self.lib.svb_cells_from_frames_bits(self._h, _ptr(bgr), n, h, w, _ptr(corners), _ptr(found), _ptr(bits),
                                                       self._stream()), "svb_cells_from_frames_bits")
        return bits

    # -- whole path ------------------------------------------------------------------------------
    def alloc_outputs(self, n: int, want_logits: bool = False):
        torch = _torch()
        dev = self._dev()
        return dict(
            digits=torch.empty((n, 81), dtype=torch.uint8, device=dev),
            conf=torch.empty((n, 81), dtype=torch.float32, device=dev),
            logits=torch.empty((n, 81, 10), dtype=torch.float32, device=dev) if want_logits else None,
            corners=torch.empty((n, 4, 2), dtype=torch.int32, device=dev),
            found=torch.empty((n,), dtype=torch.uint8, device=dev),
        )

    def scan_batch(self, bgr, out: dict | None = None, want_logits: bool = False) -> dict:
        """pipeline/run.py:257-318 for a device-resident batch (n,H,W,3) u8.  Asynchronous."""
        self._chk_u8(bgr, 4, "scan_batch")
        n, h, w, _ = bgr.shape
        if out is None:
            out = self.alloc_outputs(n, want_logits)
        _lib.check(self.lib.svb_scan_batch_v1(self._h, _ptr(bgr), n, h, w, _ptr(out["digits"]), _ptr(out["conf"]),
                                              _ptr(out.get("logits")), _ptr(out["corners"]), _ptr(out["found"]),
                                              self._stream()), "svb_scan_batch_v1")
        return out

    QUALITY_FIELDS = ("overall", "sharpness", "contrast", "completeness", "geometry", "size")

    def assess_grid_quality(self, frames, binary, corners, found=None):
        """cv/grid_quality.py:228-306 for a batch: frames (n,H,W,3) BGR or (n,H,W) gray, binary (n,H,W), corners
        (n,4,2) int32 -> scores (n,6) float64 in QUALITY_FIELDS order."""
        torch = _torch()
        n, h, w, ch = self._frames_nhw(frames, "assess_grid_quality")
        self._chk_u8(binary, 3, "assess_grid_quality")
        corners = corners.to(device=frames.device, dtype=torch.int32).contiguous()
        scores = torch.empty((n, 6), dtype=torch.float64, device=frames.device)
        _lib.check(self.lib.svb_assess_grid_quality(self._h, _ptr(frames), n, h, w, ch, _ptr(binary), _ptr(corners), _ptr(found),
                                                    _ptr(scores), self._stream()), "svb_assess_grid_quality")
        return scores

    def scan_batch_v2(self, bgr, want_logits: bool = False, min_quality_score: float = -1.0) -> dict:
        """pipeline/run_v2.py:276-330 (detection method 1) for a device-resident batch (n,H,W,3) u8:
        preprocess_multi_strategy -> contour + validity -> assess_grid_quality [gate when min_quality_score >= 0; the
        reference's default is 40, -1 = --no-quality-check] -> cells -> DigitCNNv3 -> top-3.  Needs load_weights_v3."""
        torch = _torch()
        self._chk_u8(bgr, 4, "scan_batch_v2")
        n, h, w, _ = bgr.shape
        dev = bgr.device
        out = dict(digits=torch.empty((n, 81), dtype=torch.uint8, device=dev),
                   conf=torch.empty((n, 81), dtype=torch.float32, device=dev),
                   alt_digits=torch.empty((n, 81, 2), dtype=torch.uint8, device=dev),
                   alt_conf=torch.empty((n, 81, 2), dtype=torch.float32, device=dev),
                   logits=torch.empty((n, 81, 10), dtype=torch.float32, device=dev) if want_logits else None,
                   corners=torch.empty((n, 4, 2), dtype=torch.int32, device=dev),
                   found=torch.empty((n,), dtype=torch.uint8, device=dev),
                   info=torch.empty((n, 4), dtype=torch.uint8, device=dev),
                   quality=torch.empty((n, 6), dtype=torch.float64, device=dev))
        _lib.check(self.lib.svb_scan_batch_v2(self._h, _ptr(bgr), n, h, w, _ptr(out["digits"]), _ptr(out["conf"]),
                                              _ptr(out["alt_digits"]), _ptr(out["alt_conf"]), _ptr(out["logits"]),
                                              _ptr(out["corners"]), _ptr(out["found"]), _ptr(out["info"]),
                                              _ptr(out["quality"]), float(min_quality_score), self._stream()),
                   "svb_scan_batch_v2")
        return out

    def solve_batch(self, grids):
        """solve_sudoku (solver/src/sudoku.c:72) for a batch: grids (n,81) or (n,9,9) uint8 CUDA, 0 = empty ->
        (solutions same shape, status (n,) int8: 1 solved, 0 no solution, -1 invalid)."""
        torch = _torch()
        if not (grids.is_cuda and grids.dtype == torch.uint8 and grids.is_contiguous() and grids.numel() % 81 == 0):
            raise ValueError("solve_batch: expected contiguous uint8 CUDA grids with 81 cells each")
        n = grids.numel() // 81
        sol = torch.empty_like(grids)
        status = torch.empty((n,), dtype=torch.int8, device=grids.device)
        _lib.check(self.lib.svb_solve_batch(self._h, _ptr(grids), n, _ptr(sol), _ptr(status), self._stream()), "svb_solve_batch")
        return sol, status

    def scan_batch_host(self, frames: np.ndarray, out: dict | None = None) -> dict:
        """Same through HOST buffers (numpy or pinned torch CPU tensors): H2D + path + D2H, synchronous."""
        torch = _torch()
        if isinstance(frames, torch.Tensor):
            assert frames.device.type == "cpu" and frames.dtype == torch.uint8 and frames.is_contiguous()
            n, h, w, _ = frames.shape
            src = C.c_void_p(frames.data_ptr())
        else:
            frames = np.ascontiguousarray(frames, dtype=np.uint8)
            n, h, w, _ = frames.shape
            src = frames.ctypes.data_as(C.c_void_p)
        if out is None:
            out = dict(digits=np.empty((n, 81), np.uint8), conf=np.empty((n, 81), np.float32),
                       corners=np.empty((n, 4, 2), np.int32), found=np.empty((n,), np.uint8))

        def hp(a):
            return C.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a.ctypes.data_as(C.c_void_p)

        _lib.check(self.lib.svb_scan_batch_v1_host(self._h, src, n, h, w, hp(out["digits"]), hp(out["conf"]),
                                                   hp(out["corners"]), hp(out["found"])), "svb_scan_batch_v1_host")
        return out

    # -- frame ingest (cv2.imread's decode step on the GPU) -------------------------------------------------------------
    @staticmethod
    def pack_jpegs(files) -> tuple:
        """list of bytes-like JPEG files -> (blob uint8 ndarray, offsets int64 ndarray of len n + 1)"""
        offs = np.zeros(len(files) + 1, np.int64)
        for i, f in enumerate(files):
            offs[i + 1] = offs[i] + len(f)
        blob = np.empty(int(offs[-1]) + 64, np.uint8)
        for i, f in enumerate(files):
            blob[offs[i]:offs[i + 1]] = np.frombuffer(f, np.uint8)
        blob[offs[-1]:] = 0
        return blob, offs

    def jpeg_decode(self, blob, offsets, h: int, w: int, want_status: bool = False):
        """blob / offsets: host arrays from pack_jpegs (or pinned torch CPU tensors) -> BGR frames (n,h,w,3) uint8 CUDA."""
        torch = _torch()
        n = len(offsets) - 1
        out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self._dev())
        status = torch.empty((n,), dtype=torch.uint8, device=self._dev()) if want_status else None
        _lib.check(self.lib.svb_jpeg_decode_host(self._h, _hptr(blob), _hptr(offsets), n, h, w, _ptr(out), _ptr(status), self._stream()),
                   "svb_jpeg_decode_host")
        return (out, status) if want_status else out

    def scan_batch_jpeg_host(self, blob, offsets, h: int, w: int, out: dict | None = None) -> dict:
        """scan_batch_host for compressed frames: H2D of the JPEG bytes + decode + path + D2H, synchronous."""
        n = len(offsets) - 1
        if out is None:
            out = dict(digits=np.empty((n, 81), np.uint8), conf=np.empty((n, 81), np.float32),
                       corners=np.empty((n, 4, 2), np.int32), found=np.empty((n,), np.uint8))
        _lib.check(self.lib.svb_scan_batch_v1_jpeg_host(self._h, _hptr(blob), _hptr(offsets), n, h, w, _hptr(out["digits"]),
                                                        _hptr(out["conf"]), _hptr(out["corners"]), _hptr(out["found"])),
                   "svb_scan_batch_v1_jpeg_host")
        return out


def _hptr(a):
    """host pointer of a numpy array or a CPU torch tensor"""
    if hasattr(a, "data_ptr"):
        assert a.device.type == "cpu" and a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    a = np.ascontiguousarray(a)
    return a.ctypes.data_as(C.c_void_p)
