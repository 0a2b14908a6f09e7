"""CPU tier: the product's JPEG decoder core (csrc/jpeg_core.cuh — header parser, Huffman decode, islow IDCT, fancy
upsampling, colour conversion: the functions jpeg.cu's kernels call, run serially by tests/helpers/jpeg_host.cpp) against
cv2.imdecode bit for bit: the committed golden (tests/golden/jpeg.npz) and, where cv2 is importable, fresh encodes."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT


@pytest.fixture(scope="module")
def jpeg_host():
    src = os.path.join(ROOT, "tests", "helpers", "jpeg_host.cpp")
    out_dir = os.path.join(ROOT, "tests", "helpers", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libjpeg_host.so")
    dep = os.path.join(PKG, "csrc", "jpeg_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)

    def decode(buf):
        b = np.ascontiguousarray(np.frombuffer(bytes(buf), np.uint8))
        info = (C.c_int * 6)()
        rc = lib.svbh_jpeg_info(b.ctypes.data_as(C.c_void_p), C.c_longlong(len(b)), info)
        if rc:
            return rc, None
        out = np.empty((info[1], info[0], 3), np.uint8)
        rc = lib.svbh_jpeg_decode(b.ctypes.data_as(C.c_void_p), C.c_longlong(len(b)), out.ctypes.data_as(C.c_void_p))
        return rc, out

    return decode


def test_jpeg_core_golden(jpeg_host, golden):
    g = golden("jpeg")
    names = [k[:-5] for k in g if k.endswith("_file") and not k.startswith(("unsupported", "malformed"))]
    assert len(names) >= 8
    for nm in names:
        rc, got = jpeg_host(g[nm + "_file"])
        assert rc == 0, nm
        assert tuple(got.shape) == tuple(g[nm + "_shape"]), nm
        assert np.array_equal(got[:: max(1, got.shape[0] // 8)], g[nm + "_rows"]), nm
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest() == g[nm + "_sha"].tobytes(), nm
    assert jpeg_host(g["unsupported_progressive_file"])[0] == -2   # valid JPEG, outside the supported subset
    assert jpeg_host(g["malformed_truncated_file"])[0] == -1


def test_jpeg_core_vs_cv2_fresh_encodes(jpeg_host):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:96, 0:130]
    imgs = [rng.integers(0, 256, (50, 70, 3)).astype(np.uint8), rng.integers(0, 256, (8, 8, 3)).astype(np.uint8),
            rng.integers(0, 256, (17, 2, 3)).astype(np.uint8), rng.integers(0, 256, (3, 200, 3)).astype(np.uint8),
            np.stack([xx * 2 % 256, (yy * 3 + xx) % 256, (yy * yy // 40) % 256], -1).astype(np.uint8)]
    n = 0
    for im in imgs:
        for q in (10, 50, 92, 100):
            for rst in (0, 1, 3):
                for samp in (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444):
                    params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp]
                    if rst:
                        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
                    ok, buf = cv2.imencode(".jpg", im, params)
                    rc, got = jpeg_host(buf.tobytes())
                    assert rc == 0 and np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR)), (im.shape, q, rst, samp)
                    n += 1
    assert n == 120
