"""CPU tier: the oracle (oracle/svb_oracle.c) against the golden vectors minted from the unmodified
reference (tests/golden/make_golden.py).  Bit-exact for every integer/byte/index stage; logits
within 1e-3 (fp32, north-star tolerance)."""
import hashlib

import numpy as np
import pytest

FRAMES = ["frame_a", "frame_b", "frame_c", "photo4_dec8", "frame_none"]


@pytest.mark.parametrize("name", FRAMES)
def test_image_stages_bit_exact(oracle, golden, name):
    g = golden(name)
    img = g["bgr"]
    assert np.array_equal(oracle.gray(img), g["ref_gray"])
    assert np.array_equal(oracle.blur5(g["ref_gray"]), g["ref_blur"])
    assert np.array_equal(oracle.adaptive_gauss11(g["ref_blur"], True), g["ref_mask"])
    assert np.array_equal(oracle.preprocess(img), g["ref_mask"])
    c = oracle.find_grid_contour(g["ref_mask"])
    assert (c is not None) == bool(g["ref_found"])
    if c is None:
        return
    assert np.array_equal(c, g["ref_corners"])
    assert np.array_equal(oracle.order_points(c), g["ref_ordered"])
    board = oracle.warp_perspective(img, c)
    assert hashlib.sha256(board.tobytes()).digest() == g["ref_board_sha256"].tobytes()
    cells = oracle.extract_cells(board)
    assert np.array_equal(cells, g["ref_cells_u8"])
    assert np.array_equal(255 - oracle.cell_prep(cells), g["ref_cells_thresh"])


@pytest.mark.parametrize("name", FRAMES[:4])
def test_scan_frame_and_logits(oracle, golden, weights, name):
    g = golden(name)
    r = oracle.scan_frame(g["bgr"])
    assert r["found"]
    assert np.array_equal(r["corners"], g["ref_corners"])
    assert np.array_equal(r["cells_u8"], g["ref_cells_u8"])
    assert np.array_equal(255 - r["cells_in"], g["ref_cells_thresh"])
    x = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
    assert set(np.unique(x)) <= {-1.0, 1.0}
    logits = oracle.digitcnn_forward(weights, x)
    assert np.abs(logits - g["ref_logits"]).max() < 1e-3
    assert np.array_equal(logits.argmax(1).astype(np.uint8), g["ref_digits"])


def test_unit_vectors_awkward_sizes(oracle, golden):
    """Widths that are not multiples of 8 exercise OpenCV's scalar tail arithmetic (svb_oracle.c)."""
    u = golden("unit")
    for i in range(4):
        g = u[f"g{i}"]
        assert np.array_equal(oracle.blur5(g), u[f"ref_blur{i}"])
        assert np.array_equal(oracle.adaptive_gauss11(g, True), u[f"ref_thr_inv{i}"])
        assert np.array_equal(oracle.adaptive_gauss11(g, False), u[f"ref_thr_bin{i}"])
    cells = u["cells"]
    assert np.array_equal(np.stack([oracle.clahe28(c) for c in cells]), u["ref_clahe"])
    assert np.array_equal(255 - oracle.cell_prep(cells), u["ref_cells_thresh"])


def test_gauss_kernel_bits(oracle):
    import ctypes as C

    k = np.zeros(11, np.float32)
    oracle.lib().svo_gauss11_kernel(k.ctypes.data_as(C.c_void_p))
    want = np.array([0x3c10612b, 0x3cde5c35, 0x3d855a85, 0x3df92326, 0x3e353f0f, 0x3e4d6105], np.uint32)
    assert np.array_equal(k[:6].view(np.uint32), want)
    assert np.array_equal(k[:5], k[:5:-1])
    # the constants compiled into the CUDA kernels (csrc/common.cuh)
    lits = np.array([0.00881222915, 0.0271435771, 0.0651140586, 0.121649072, 0.176998362, 0.200565413], np.float32)
    assert np.array_equal(lits.view(np.uint32), want)


def test_product_contour_core_on_frames(oracle, golden, contour_host):
    """The product's contour core (host build of csrc/contour_core.cuh), byte view and tiled-bit view with the
    register-window walker, on the golden masks and on a 1080p synthetic frame."""
    from svb200 import frames as F

    masks = [(golden(n)["ref_mask"], golden(n)) for n in ("frame_a", "frame_b", "frame_c", "frame_none")]
    for m, g in masks:
        for bits in ((False, True) if m.shape[1] % 32 == 0 else (False,)):
            f, c = contour_host(m, 0.1, 0.02, use_bits=bits)
            assert (f == 1) == bool(g["ref_found"])
            if f == 1:
                assert np.array_equal(c, g["ref_corners"])
    fr = F.make_frame(777, 1080, 1920, 40.0)
    m = oracle.preprocess(F.add_noise_host(fr.image, 777))
    want = oracle.find_grid_contour(m)
    for bits in (False, True):
        f, c = contour_host(m, 0.1, 0.02, use_bits=bits)
        assert (f == 1) == (want is not None)
        if want is not None:
            assert np.array_equal(c, want)


def test_v2_contour_method_golden(oracle, golden, contour_host):
    """cv/grid_v2.py:102-128 (detect_grid_contour): oracle_v2 and the product's contour core vs the reference."""
    from oracle import oracle_v2

    v = golden("v2")
    cases = [(golden(n)["ref_mask"], v[f"{n}_corners"], 0.1) for n in ("frame_a", "frame_b", "frame_c", "photo4_dec8", "frame_none")]
    cases += [(v[f"shape{k}_mask"], v[f"shape{k}_corners"], 0.05) for k in range(3)]
    n_none = 0
    for m, want, ratio in cases:
        got = oracle_v2.detect_grid_contour(m, ratio)
        f, c = contour_host(m, ratio, 0.02, use_bits=(m.shape[1] % 32 == 0), v2=True)
        if len(want) == 0:
            n_none += 1
            assert got is None and f == 0
        else:
            assert np.array_equal(got, want)
            assert f == 1 and np.array_equal(c.astype(np.float32), want)
    assert n_none >= 2


def test_digitcnn_v3_oracle_golden(golden):
    """oracle/model_v3_oracle.py (torch functional restatement) vs the reference DigitCNNv3 (golden), 1e-4."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from v3_weights import make_v3_inputs, make_v3_state

    from oracle import model_v3_oracle as M

    g = golden("v3")
    sd = make_v3_state()
    assert len(sd) == int(g["n_state_entries"]) == 91
    x = make_v3_inputs()
    assert np.abs(M.forward(sd, x) - g["ref_logits"]).max() < 1e-4
    assert np.abs(M.forward(sd, x, return_features=True) - g["ref_features"]).max() < 1e-4
