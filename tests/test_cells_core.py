"""CPU tier: the product's per-cell pipeline (csrc/cells_core.cuh — the phase functions cells.cu's kernels call, run here
thread id by thread id) against the oracle, bit for bit: u8 cells, +-1 tensors and the bit rows the batched classifier
reads; the fast map evaluation against cv2's own operation order on every board pixel."""
import numpy as np
import pytest

from conftest import bits_to_pm1


def _want(oracle, img, corners):
    board = oracle.warp_perspective(img, corners)
    cells = oracle.extract_cells(board)
    ink = oracle.cell_prep(cells)  # 255 == +1 (ink): the thresholded cell after run.py's inversion
    return cells, np.where(ink == 255, 1.0, -1.0).astype(np.float32)


@pytest.mark.parametrize("hw,rot", [((540, 960), 15.0), ((1080, 1920), 25.0), ((300, 500), 10.0), ((96, 48), 5.0)])
def test_cells_from_frame_vs_oracle(oracle, cells_host, hw, rot):
    from svb200 import frames as F

    imgs, _, _ = F.make_frames(2, hw[0], hw[1], base_seed=hw[0] + 7, max_rot_deg=rot)
    n = 0
    for im in imgs:
        r = oracle.scan_frame(im)
        if not r["found"]:
            continue
        n += 1
        u8, bits, pm1 = cells_host.cells_from_frame(im, r["corners"])
        want = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        assert np.array_equal(u8, r["cells_u8"])
        assert np.array_equal(pm1, want)
        assert np.array_equal(bits_to_pm1(bits), want)
        assert (bits >> 28).max() == 0
        bad, exact = cells_host.map_check(r["corners"])
        assert bad == 0 and exact < 450 * 450 // 100  # the fall-back to cv2's order stays rare
    assert n >= 1


def test_cells_with_corners_outside_the_frame(oracle, cells_host):
    """quads that leave the frame (BORDER_CONSTANT taps), degenerate and far-away quads (saturating coordinates)"""
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (200, 320, 3)).astype(np.uint8)
    quads = [[[-40, -30], [350, -10], [360, 230], [-20, 215]], [[10, 10], [300, 20], [310, 190], [5, 180]],
             [[100, 100], [100, 100], [100, 100], [100, 100]], [[0, 0], [319, 0], [319, 199], [0, 199]],
             [[5000, 5000], [9000, 5100], [9100, 9000], [5100, 9100]], [[50, 20], [250, 20], [50, 20], [250, 180]]]
    for q in quads:
        c = np.array(q, np.int32)
        cells, want = _want(oracle, img, c)
        u8, bits, pm1 = cells_host.cells_from_frame(img, c)
        assert np.array_equal(u8, cells), q
        assert np.array_equal(pm1, want), q
        assert np.array_equal(bits_to_pm1(bits), want), q


def test_cell_prep_vs_oracle_and_golden(oracle, cells_host, golden):
    rng = np.random.default_rng(8)
    cells = [rng.integers(0, 256, (28, 28)), rng.integers(100, 140, (28, 28)), np.full((28, 28), 7), np.zeros((28, 28)),
             np.full((28, 28), 255), np.tile(np.arange(28) * 9, (28, 1)), np.tile(np.arange(28) * 9, (28, 1)).T,
             rng.integers(0, 2, (28, 28)) * 255, rng.integers(0, 40, (28, 28)), rng.integers(250, 256, (28, 28))]
    cells += [np.clip(rng.normal(128, s, (28, 28)), 0, 255) for s in (1, 5, 20, 60)]
    cells = np.stack(cells).astype(np.uint8)
    g = golden("unit")
    thr, bits, pm1 = cells_host.cell_prep(cells)
    ink = oracle.cell_prep(cells)  # 255 == +1
    assert np.array_equal(thr, 255 - ink)  # preprocess_cell's own return value: white = 255
    assert np.array_equal(pm1, np.where(ink == 255, 1.0, -1.0).astype(np.float32))
    assert np.array_equal(bits_to_pm1(bits), pm1)
    # the reference-minted unit vectors: pipeline/run.py:73-95 preprocess_cell on 12 cells
    t2, _, _ = cells_host.cell_prep(g["cells"])
    assert np.array_equal(t2, g["ref_cells_thresh"])


def test_conv1_from_bits_vs_float_conv(cells_host, weights):
    """csrc/digitcnn_bits_core.cuh: the 512-pattern table + border classes + work-item order of tc_conv_kernel<true>
    against a plain float conv1 + bias + ReLU + 2x2 max-pool (ml/model.py:36) on random and structured +-1 cells."""
    import ctypes as C

    rng = np.random.default_rng(12)
    cells = [rng.random((28, 28)) < p for p in (0.0, 1.0, 0.5, 0.25, 0.75, 0.1)]
    edge = np.zeros((28, 28), bool)
    edge[0, :] = edge[-1, :] = edge[:, 0] = edge[:, -1] = True
    cells += [edge, ~edge, np.eye(28, dtype=bool)]
    x = np.stack(cells)
    bits = (x.astype(np.uint32) << np.arange(28, dtype=np.uint32)).sum(-1).astype(np.uint32)
    assert np.array_equal(bits_to_pm1(bits) > 0, x)
    w, b = weights["conv1.weight"].astype(np.float32), weights["conv1.bias"].astype(np.float32)
    out = np.empty((len(x), 32, 14, 14), np.float32)
    cover = np.empty(784, np.int32)
    lib = cells_host.lib
    lib.svbh_conv1_bits(np.ascontiguousarray(bits).ctypes.data_as(C.c_void_p), len(x), np.ascontiguousarray(w).ctypes.data_as(C.c_void_p),
                        b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), cover.ctypes.data_as(C.c_void_p))
    assert (cover == 1).all()  # every (channel group, pooled pixel) exactly once
    pm1 = np.where(x, 1.0, -1.0).astype(np.float64)
    xp = np.pad(pm1, ((0, 0), (1, 1), (1, 1)))
    conv = np.zeros((len(x), 32, 28, 28))
    for ky in range(3):
        for kx in range(3):
            conv += w[None, :, 0, ky, kx, None, None] * xp[:, None, ky:ky + 28, kx:kx + 28]
    conv += b[None, :, None, None]
    want = np.maximum(conv, 0).reshape(len(x), 32, 14, 2, 14, 2).max((3, 5))
    assert np.abs(out - want).max() < 2e-6
