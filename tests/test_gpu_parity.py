"""GPU tier (-m gpu): every CUDA kernel, called through the C ABI, against the CPU oracle on the same
seeded inputs and against the golden fixtures.  Bit-exact for all integer/byte/index stages;
logits within 1e-3 absolute (fp32, the north star's tolerance)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3


def _t(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _frames(n, h, w, seed, rot=15.0):
    from svb200 import frames as F

    return F.make_frames(n, h, w, base_seed=seed, max_rot_deg=rot)


# ---- P1..P4 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(28, 28), (37, 53), (64, 80), (45, 66), (120, 168)])
def test_stage_kernels_any_size(scanner, oracle, hw):
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    img = rng.integers(0, 256, (2,) + hw + (3,)).astype(np.uint8)
    g = scanner.grayscale(_t(img))
    for i in range(2):
        assert np.array_equal(g[i].cpu().numpy(), oracle.gray(img[i]))
    b = scanner.blur(g)
    for i in range(2):
        assert np.array_equal(b[i].cpu().numpy(), oracle.blur5(g[i].cpu().numpy()))
    for inv in (True, False):
        t = scanner.adaptive_threshold(b, inverted=inv)
        for i in range(2):
            assert np.array_equal(t[i].cpu().numpy(), oracle.adaptive_gauss11(b[i].cpu().numpy(), inv))
    # odd sizes take the three-kernel path inside svb_preprocess_v1
    m = scanner.preprocess(_t(img))
    for i in range(2):
        assert np.array_equal(m[i].cpu().numpy(), oracle.preprocess(img[i]))


@pytest.mark.parametrize("hw", [(64, 64), (72, 480), (100, 496), (270, 480), (136, 976), (300, 1920), (1080, 1920)])
def test_fused_preprocess_bit_exact(scanner, oracle, hw):
    """K1 (TMA-staged, row-streaming): edge strips, partial strips, short segments, full 1080p."""
    rng = np.random.default_rng(hw[0] + hw[1])
    n = 3 if hw[0] * hw[1] < 500_000 else 2
    img = rng.integers(0, 256, (n,) + hw + (3,)).astype(np.uint8)
    # second image: smooth content with lines, closer to real frames (more rounding ties)
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    img[1] = np.stack([(120 + 60 * np.sin(xx / 37.0) + 20 * np.cos(yy / 11.0)), (140 + 50 * np.sin((xx + yy) / 23.0)),
                       (100 + 80 * np.cos(yy / 51.0))], -1).clip(0, 255).astype(np.uint8)
    m = scanner.preprocess(_t(img)).cpu().numpy()
    for i in range(n):
        want = oracle.preprocess(img[i])
        bad = np.argwhere(m[i] != want)
        assert len(bad) == 0, f"{len(bad)} mask pixels differ, first at {bad[:5].tolist()}"


def test_fused_preprocess_golden(scanner, golden):
    for name in ("frame_a", "frame_c"):  # widths 480 / 496: fused path
        g = golden(name)
        m = scanner.preprocess(_t(g["bgr"][None])).cpu().numpy()[0]
        assert np.array_equal(m, g["ref_mask"])
    for name in ("photo4_dec8", "frame_none"):  # 342 / 320 wide: stage-kernel path
        g = golden(name)
        m = scanner.preprocess(_t(g["bgr"][None])).cpu().numpy()[0]
        assert np.array_equal(m, g["ref_mask"])


# ---- G1..G2 ---------------------------------------------------------------------------------------
def test_find_grid_contour_golden(scanner, golden):
    for name in ("frame_a", "frame_b", "frame_c", "photo4_dec8", "frame_none"):
        g = golden(name)
        c, f = scanner.find_grid_contour(_t(g["ref_mask"][None]))
        assert int(f[0]) == int(bool(g["ref_found"]))
        if g["ref_found"]:
            assert np.array_equal(c[0].cpu().numpy(), g["ref_corners"])


def test_find_grid_contour_shapes(scanner, oracle):
    """Adversarial masks (nesting, frame rings, thin strokes, noise) incl. the None outcome."""
    rng = np.random.default_rng(23)
    H, W = 160, 224
    masks = []
    for s in range(96):
        m = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(1, 8))):
            x0, x1 = sorted(rng.integers(0, W, 2).tolist())
            y0, y1 = sorted(rng.integers(0, H, 2).tolist())
            t = int(rng.integers(1, 4))
            k = int(rng.integers(0, 3))
            if k == 0:
                m[y0:y1 + 1, x0:x1 + 1] = 255
            elif k == 1:
                m[y0:y1 + 1, x0:x0 + t] = 255
                m[y0:y1 + 1, max(x1 - t + 1, 0):x1 + 1] = 255
                m[y0:y0 + t, x0:x1 + 1] = 255
                m[max(y1 - t + 1, 0):y1 + 1, x0:x1 + 1] = 255
            else:
                for i in range(max(x1 - x0, 1)):  # diagonal stroke, 8-connected only
                    yy = y0 + (i * (y1 - y0)) // max(x1 - x0, 1)
                    m[min(yy, H - 1), min(x0 + i, W - 1)] = 255
        if rng.random() < 0.3:
            m |= ((rng.random((H, W)) < 0.04) * 255).astype(np.uint8)
        if rng.random() < 0.25:
            m[0:2, :] = 255
            m[-2:, :] = 255
            m[:, 0:2] = 255
            m[:, -2:] = 255
        masks.append(m)
    masks = np.stack(masks)
    n_found = 0
    for ratio in (0.03, 0.1, 0.3):
        c, f = scanner.find_grid_contour(_t(masks), ratio)
        c, f = c.cpu().numpy(), f.cpu().numpy()
        for i in range(len(masks)):
            want = oracle.find_grid_contour(masks[i], ratio, 0.02)
            assert f[i] in (0, 1), f"mask {i}: status {f[i]}"
            assert (want is not None) == bool(f[i]), f"mask {i} ratio {ratio}"
            if want is not None:
                n_found += 1
                assert np.array_equal(c[i], want), f"mask {i} ratio {ratio}: {c[i].tolist()} vs {want.tolist()}"
    assert n_found > 30


def test_find_grid_contour_synthetic_frames(scanner, oracle):
    imgs, _, _ = _frames(6, 1080, 1920, 9100, rot=40.0)
    m = scanner.preprocess(_t(imgs))
    c, f = scanner.find_grid_contour(m)
    c, f, m = c.cpu().numpy(), f.cpu().numpy(), m.cpu().numpy()
    for i in range(len(imgs)):
        want = oracle.find_grid_contour(m[i])
        assert (want is not None) == (f[i] == 1)
        if want is not None:
            assert np.array_equal(c[i], want)


@pytest.mark.parametrize("hw", [(270, 480), (200, 736), (300, 1920), (540, 960), (1080, 1920)])
def test_scan_fused_bit_mask_matches_staged_path(scanner, hw):
    """svb_scan_batch_v1 lets K1 write K2's tiled bit mask (no packing pass); the staged calls pack the byte mask inside
    find_grid_contour.  Both must see the same contours: identical found flags and corners, including partial strips,
    frames whose height is not a multiple of the 32-row tiles and repeated calls with a smaller batch."""
    imgs, _, _ = _frames(5, hw[0], hw[1], 9300 + hw[0], rot=25.0)
    imgs[3] = 127  # no grid
    for n in (5, 2):
        batch = _t(imgs[:n])
        out = scanner.scan_batch(batch)
        m = scanner.preprocess(batch)
        c, f = scanner.find_grid_contour(m)
        assert np.array_equal(out["found"].cpu().numpy(), f.cpu().numpy())
        assert np.array_equal(out["corners"].cpu().numpy(), c.cpu().numpy())
    assert int(f.sum()) >= 1


def test_v2_contour_method(scanner, golden):
    """svb_detect_grid_contour_v2 vs the reference's cv/grid_v2.detect_grid_contour (golden)."""
    v = golden("v2")
    cases = [(golden(n)["ref_mask"], v[f"{n}_corners"], 0.1) for n in ("frame_a", "frame_b", "frame_c", "photo4_dec8", "frame_none")]
    cases += [(v[f"shape{k}_mask"], v[f"shape{k}_corners"], 0.05) for k in range(3)]
    for m, want, ratio in cases:
        c, f = scanner.detect_grid_contour_v2(_t(m[None]), ratio)
        assert int(f[0]) == (1 if len(want) else 0)
        if len(want):
            assert np.array_equal(c[0].cpu().numpy().astype(np.float32), want)


# ---- G3/G4, E1, C1/C2 -----------------------------------------------------------------------------
def test_warp_extract_cellprep_golden(scanner, golden):
    import torch

    for name in ("frame_a", "frame_b", "frame_c", "photo4_dec8"):
        g = golden(name)
        bgr = _t(g["bgr"][None])
        corners = torch.from_numpy(g["ref_corners"][None]).cuda()
        board = scanner.warp_perspective(bgr, corners)
        b = board[0].cpu().numpy()
        assert hashlib.sha256(b.tobytes()).digest() == g["ref_board_sha256"].tobytes(), name
        cells = scanner.extract_cells(board)
        assert np.array_equal(cells[0].cpu().numpy(), g["ref_cells_u8"])
        thr, pm1 = scanner.cell_prep(cells)
        assert np.array_equal(thr[0].cpu().numpy(), g["ref_cells_thresh"])
        want_pm1 = ((255 - g["ref_cells_thresh"]).astype(np.float32) / 255.0 - 0.5) / 0.5
        assert np.array_equal(pm1[0].cpu().numpy(), want_pm1)
        # fused kernel: frames + corners -> cells, no board
        u8, pm1f = scanner.cells_from_frames(bgr, corners)
        assert np.array_equal(u8[0].cpu().numpy(), g["ref_cells_u8"])
        assert np.array_equal(pm1f[0].cpu().numpy(), want_pm1)


def test_cell_prep_unit_vectors(scanner, golden):
    u = golden("unit")
    thr, _ = scanner.cell_prep(_t(u["cells"]))
    assert np.array_equal(thr.cpu().numpy(), u["ref_cells_thresh"])


def test_cell_prep_tail_columns_regression(scanner, oracle):
    """Columns 24..27 of a cell take OpenCV's scalar tail (multiply, round, add).  ptxas contracted the packed form of that
    into FFMA2 and 4 cells of a 1024-frame batch lost one pixel each (round 2): those four cells, and 60,000 textured cells
    (a rounding tie in the tail is a 1-in-20,000-cells event)."""
    import os

    reg = np.load(os.path.join(os.path.dirname(__file__), "golden", "tail_cells.npz"))["cells"]
    rng = np.random.default_rng(5)
    n = 60000
    tex = (rng.normal(150, 25, (n, 1, 1)) + rng.normal(0, 1, (n, 28, 28)) * rng.uniform(2, 30, (n, 1, 1))
           + np.linspace(-20, 20, 28)[None, None, :] * rng.normal(0, 1, (n, 1, 1)))
    cells = np.concatenate([reg, np.clip(tex, 0, 255).astype(np.uint8)])
    thr, pm1 = scanner.cell_prep(_t(cells))
    ink = oracle.cell_prep(cells)
    assert np.array_equal(thr.cpu().numpy(), 255 - ink)
    assert np.array_equal(pm1.cpu().numpy(), np.where(ink == 255, 1.0, -1.0).astype(np.float32))


def test_cells_from_frames_vs_oracle_1080p(scanner, oracle):
    imgs, _, _ = _frames(4, 1080, 1920, 7700)
    res = [oracle.scan_frame(im) for im in imgs]
    import torch

    corners = torch.from_numpy(np.stack([r["corners"] for r in res])).cuda()
    u8, pm1 = scanner.cells_from_frames(_t(imgs), corners)
    for i, r in enumerate(res):
        assert r["found"]
        assert np.array_equal(u8[i].cpu().numpy(), r["cells_u8"])
        want = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        assert np.array_equal(pm1[i].cpu().numpy(), want)


def test_cells_bits_vs_oracle_and_float_path(scanner, oracle):
    """the batched path's K4 output (28 bit rows per cell) carries exactly the +-1 tensor of the float path / the oracle"""
    from conftest import bits_to_pm1

    imgs, _, _ = _frames(3, 1080, 1920, 9100, rot=25.0)
    imgs = np.concatenate([imgs, np.zeros((1, 1080, 1920, 3), np.uint8)])  # a frame without a grid
    d = _t(imgs)
    corners, found = scanner.find_grid_contour(scanner.preprocess(d))
    bits = scanner.cells_from_frames_bits(d, corners, found).cpu().numpy().astype(np.uint32)
    _, pm1 = scanner.cells_from_frames(d, corners, found, want_u8=False)
    pm1 = pm1.cpu().numpy()
    assert int(found[-1]) == 0 and (bits[-1] == 0).all()
    for i in range(3):
        r = oracle.scan_frame(imgs[i])
        assert r["found"] and int(found[i]) == 1
        want = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        assert np.array_equal(pm1[i], want)
        assert np.array_equal(bits_to_pm1(bits[i]), want)
        assert (bits[i] >> 28).max() == 0


@pytest.mark.parametrize("n", [1, 81, 500, 20000])
def test_digitcnn_bits_path_equals_float_path(scanner, oracle, weights, n):
    """svb_digitcnn_forward_bits (conv1 by 512-pattern table) against svb_digitcnn_forward on the same +-1 cells and the oracle"""
    import torch

    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.where(torch.rand((n, 1, 28, 28), device="cuda", generator=g) < 0.3, 1.0, -1.0)
    if n >= 81:  # structured cells: borders, full / empty, diagonals exercise every pixel class
        x[0] = 1.0
        x[1] = -1.0
        x[2] = -1.0
        x[2, 0, 0, :] = x[2, 0, -1, :] = x[2, 0, :, 0] = x[2, 0, :, -1] = 1.0
        x[3] = torch.where(torch.eye(28, device="cuda") > 0, 1.0, -1.0)
    bits = scanner.pack_cells_bits(x.view(n, 28, 28))
    lb, db, cb = scanner.digitcnn_forward_bits(bits, want_digits=True)
    lf, df, _ = scanner.digitcnn_forward(x, want_digits=True)
    assert float((lb - lf).abs().max()) < 2e-4
    k = min(n, 300)
    want = oracle.digitcnn_forward(weights, x[:k].cpu().numpy())
    assert np.abs(lb[:k].cpu().numpy() - want).max() < LOGIT_TOL
    assert np.array_equal(db[:k].cpu().numpy(), want.argmax(1).astype(np.uint8))
    assert torch.equal(db, df)


def test_not_found_frames_are_zero_filled(scanner):
    import torch

    bgr = torch.zeros((2, 64, 64, 3), dtype=torch.uint8, device="cuda")
    corners = torch.zeros((2, 4, 2), dtype=torch.int32, device="cuda")
    found = torch.zeros((2,), dtype=torch.uint8, device="cuda")
    u8, pm1 = scanner.cells_from_frames(bgr, corners, found)
    assert int(u8.abs().sum()) == 0 and float(pm1.abs().sum()) == 0.0


# ---- M1/M2 ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["tc", "fp32"])
def test_digitcnn_logits(scanner, oracle, weights, golden, mode):
    """Both classifier implementations (tcgen05/TMEM with fp16 hi/lo split, and plain fp32) against
    the reference's logits (golden) and the oracle, 1e-3 absolute."""
    import torch

    scanner.set_classifier_mode(mode)
    g = golden("frame_b")
    x = ((255 - g["ref_cells_thresh"]).astype(np.float32) / 255.0 - 0.5) / 0.5
    logits, digits, conf = scanner.digitcnn_forward(torch.from_numpy(x).cuda().unsqueeze(1), want_digits=True)
    assert np.abs(logits.cpu().numpy() - g["ref_logits"]).max() < LOGIT_TOL
    assert np.array_equal(digits.cpu().numpy(), g["ref_digits"])
    assert np.abs(conf.cpu().numpy() - g["ref_conf"]).max() < 1e-4
    # arbitrary float inputs (the drop-in forward must accept any tensor, not only +-1)
    rng = np.random.default_rng(5)
    xr = rng.normal(0, 1, (70, 1, 28, 28)).astype(np.float32)
    got = scanner.digitcnn_forward(torch.from_numpy(xr).cuda()).cpu().numpy()
    want = oracle.digitcnn_forward(weights, xr)
    assert np.abs(got - want).max() < LOGIT_TOL
    # a batch that is not a multiple of the 128-cell tile, +-1 inputs as the path produces them
    xb = np.where(rng.random((300, 1, 28, 28)) < 0.2, 1.0, -1.0).astype(np.float32)
    got = scanner.digitcnn_forward(torch.from_numpy(xb).cuda()).cpu().numpy()
    want = oracle.digitcnn_forward(weights, xb)
    assert np.abs(got - want).max() < LOGIT_TOL
    scanner.set_classifier_mode("tc")


def test_digitcnn_v3_logits(scanner, golden):
    """svb_digitcnn_v3_forward (BN folded at load) vs the reference DigitCNNv3 (golden) and through the drop-in module."""
    import os
    import sys

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "helpers"))
    from v3_weights import make_v3_inputs, make_v3_state

    g = golden("v3")
    sd = make_v3_state()
    x = torch.from_numpy(make_v3_inputs()).cuda()
    scanner.load_weights_v3(sd)
    logits, digits, conf = scanner.digitcnn_v3_forward(x, want_digits=True)
    assert np.abs(logits.cpu().numpy() - g["ref_logits"]).max() < LOGIT_TOL
    assert np.array_equal(digits.cpu().numpy(), g["ref_logits"].argmax(1).astype(np.uint8))
    feats = scanner.digitcnn_v3_forward(x, want_features=True)
    assert np.abs(feats.cpu().numpy() - g["ref_features"]).max() < LOGIT_TOL
    sys.path.insert(0, os.path.join(root, "sudoku-vision_b200", "dropin", "ml"))
    sys.modules.pop("model_v3", None)
    import model_v3

    net = model_v3.DigitCNNv3().to("cuda")
    assert len(net.state_dict()) == 91
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    net.eval()
    with torch.no_grad():
        out = net(x)
        pred, cf = net.get_confidence(x)
    assert np.abs(out.cpu().numpy() - g["ref_logits"]).max() < LOGIT_TOL
    assert np.array_equal(pred.cpu().numpy(), g["ref_logits"].argmax(1))


@pytest.mark.parametrize("n", [1, 3, 301, 2051, 5003])
def test_digitcnn_v3_tensor_core_vs_fp32_and_oracle(scanner, n):
    """K6 on tcgen05 (fp16 hi/lo split, digitcnn_v3_tc.cu) against the fp32 CUDA-core kernels (svb_set_classifier_mode)
    and the CPU oracle: ragged cell counts (partial passes of 1 / 2 / 4 cells; 5003 cells = three chunks of 2,368, the weight
    ring running across many passes)."""
    import os
    import sys

    import torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from v3_weights import make_v3_state

    from oracle import model_v3_oracle as M

    sd = make_v3_state(seed=4321)
    scanner.load_weights_v3(sd)
    rng = np.random.default_rng(n)
    x = np.where(rng.random((n, 1, 28, 28)) < 0.3, 1.0, -1.0).astype(np.float32)
    x[0] = rng.standard_normal((1, 28, 28))  # the drop-in accepts any float tensor, not only +-1
    xd = torch.from_numpy(x).cuda()
    try:
        scanner.set_classifier_mode("fp32")
        ref = scanner.digitcnn_v3_forward(xd).cpu().numpy()
    finally:
        scanner.set_classifier_mode("tc")
    got, digits, conf = scanner.digitcnn_v3_forward(xd, want_digits=True)
    got = got.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < LOGIT_TOL
    m = min(n, 48)
    want = M.forward(sd, x[:m])
    assert np.abs(got[:m] - want).max() < LOGIT_TOL
    top2 = np.sort(want, 1)
    clear = (top2[:, -1] - top2[:, -2]) > 10 * LOGIT_TOL
    assert np.array_equal(digits.cpu().numpy()[:m][clear], want.argmax(1).astype(np.uint8)[clear])
    feats = scanner.digitcnn_v3_forward(xd[:m], want_features=True).cpu().numpy()
    assert np.abs(feats - M.forward(sd, x[:m], return_features=True)).max() < LOGIT_TOL
    if n > 2000:  # a wider oracle sample across both chunks of the batch: 384 cells, checked in parallel (numpy oracle)
        import concurrent.futures as cf

        idx = np.concatenate([np.arange(48, 176), np.arange(n // 2 - 64, n // 2 + 64), np.arange(n - 128, n)])
        with cf.ThreadPoolExecutor(max_workers=8) as ex:
            parts = list(ex.map(lambda ii: M.forward(sd, x[ii]), np.array_split(idx, 12)))
        want = np.concatenate(parts)
        assert np.abs(got[idx] - want).max() < LOGIT_TOL
        top2 = np.sort(want, 1)
        clear = (top2[:, -1] - top2[:, -2]) > 10 * LOGIT_TOL
        assert np.array_equal(digits.cpu().numpy()[idx][clear], want.argmax(1).astype(np.uint8)[clear])


# ---- whole path -----------------------------------------------------------------------------------
def test_scan_batch_vs_oracle(scanner, oracle, weights):
    imgs, digits_gt, _ = _frames(5, 1080, 1920, 8800)
    imgs = np.concatenate([imgs, np.random.default_rng(1).integers(80, 180, (1, 1080, 1920, 3)).astype(np.uint8)])
    out = scanner.scan_batch(_t(imgs), want_logits=True)
    found = out["found"].cpu().numpy()
    for i, im in enumerate(imgs):
        r = oracle.scan_frame(im)
        assert bool(found[i] == 1) == r["found"], i
        if not r["found"]:
            assert int(out["digits"][i].sum()) == 0
            continue
        assert np.array_equal(out["corners"][i].cpu().numpy(), r["corners"])
        x = (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        want = oracle.digitcnn_forward(weights, x)
        assert np.abs(out["logits"][i].cpu().numpy() - want).max() < LOGIT_TOL
        assert np.array_equal(out["digits"][i].cpu().numpy(), want.argmax(1).astype(np.uint8))
        assert (out["digits"][i].cpu().numpy().reshape(9, 9) == digits_gt[i]).mean() > 0.95  # and it reads the board
    assert found[-1] == 0


def test_scan_batch_1024_frames_vs_oracle(scanner, oracle, weights):
    """BASELINE configs[1] at its stated size: all 1024 frames of a bench-style batch (16 rendered frames x device noise, three
    frames without a grid) against the C oracle: found, corners exact, digits identical (the benched batch itself
    was never compared with anything).  The oracle runs on all host cores (ctypes releases the GIL)."""
    import concurrent.futures as cf
    import os

    import torch
    from svb200 import frames as F

    n = 1024
    clean = np.stack([F.make_frame(31000 + i, 1080, 1920).image for i in range(16)])
    batch = F.noisy_batch_device(torch.from_numpy(clean).cuda(), n, seed=7)
    batch[5] = 0
    batch[500] = 255
    batch[1023] = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (1080, 1920, 3)).astype(np.uint8)).cuda()
    out = scanner.scan_batch(batch, want_logits=True)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    frames = batch.cpu().numpy()

    def one(i):
        r = oracle.scan_frame(frames[i])
        if not r["found"]:
            return i, False, None, None
        lg = oracle.digitcnn_forward(weights, (r["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5)
        return i, True, r["corners"], lg

    with cf.ThreadPoolExecutor(max_workers=max(1, len(os.sched_getaffinity(0)))) as ex:
        res = list(ex.map(one, range(n)))
    nfound = 0
    for i, f, corners, lg in res:
        assert bool(got["found"][i] == 1) == f, i
        if not f:
            assert (got["digits"][i] == 0).all()
            continue
        nfound += 1
        assert np.array_equal(got["corners"][i], corners), i
        assert np.abs(got["logits"][i] - lg).max() < LOGIT_TOL, i
        assert np.array_equal(got["digits"][i], lg.argmax(1).astype(np.uint8)), i
    assert nfound == n - 3


@pytest.mark.parametrize("hw", [(1080, 1920), (750, 1000), (300, 500), (96, 48)])
def test_scan_batch_host_equals_device(scanner, hw):
    """fused K1 geometry, two sizes that take the stage-kernel fallback inside the host path (w % 16 != 0: ADVICE r1, the
    staged frames must not alias the fallback's scratch) and w < 64"""
    imgs, _, _ = _frames(3, hw[0], hw[1], 6600)
    dev = scanner.scan_batch(_t(imgs))
    host = scanner.scan_batch_host(imgs)
    assert np.array_equal(host["digits"], dev["digits"].cpu().numpy())
    assert np.array_equal(host["corners"], dev["corners"].cpu().numpy())
    assert np.array_equal(host["found"], dev["found"].cpu().numpy())
    assert np.array_equal(host["conf"], dev["conf"].cpu().numpy())
    if hw[0] >= 300:
        assert int(host["found"].sum()) == 3


def test_scan_batch_host_multi_chunk_ragged_tail(scanner):
    """more than two ~200 MB chunks with a short last one: both workers are reused and the tail chunk is exercised"""
    import torch
    from svb200 import frames as F

    clean = np.stack([F.make_frame(6700 + i, 540, 960).image for i in range(3)])
    n = 2 * ((200 << 20) // (540 * 960 * 3)) + 37
    big = F.noisy_batch_device(torch.from_numpy(clean).cuda(), n, seed=5)
    dev = scanner.scan_batch(big)
    host = scanner.scan_batch_host(big.cpu().numpy())
    for k in ("digits", "corners", "found", "conf"):
        assert np.array_equal(host[k], dev[k].cpu().numpy()), k
    assert int(host["found"].sum()) == n


def test_scan_batch_overlapped_equals_in_order(scanner):
    """with SVB_OPT_OVERLAP = 4 svb_scan_batch_v1 cuts batches of >= 64 frames into four parts on two internal streams: every
    output must be bit-identical to the in-order single-stream form, including a ragged part and a no-grid frame."""
    import torch
    from svb200 import frames as F

    clean = np.stack([F.make_frame(5200 + i, 1080, 1920).image for i in range(3)])
    big = F.noisy_batch_device(torch.from_numpy(clean).cuda(), 131, seed=4)
    big[77] = 0
    try:
        scanner.set_option("overlap", 4)
        a = scanner.scan_batch(big, want_logits=True)
        torch.cuda.synchronize()
        scanner.set_option("overlap", 0)
        b = scanner.scan_batch(big, want_logits=True)
        torch.cuda.synchronize()
    finally:
        scanner.set_option("overlap", 0)
    for k in ("digits", "conf", "corners", "found", "logits"):
        assert torch.equal(a[k], b[k]), k
    assert int(a["found"][77]) == 0 and int((a["found"] == 1).sum()) == 130


def test_scan_batch_properties_full_size(scanner):
    """BASELINE config 2 shape (1080p, a large batch): size-independent properties — determinism,
    batch-composition independence (frame i's result does not depend on its neighbours), and
    permutation equivariance."""
    import torch
    from svb200 import frames as F

    clean = np.stack([F.make_frame(5000 + i, 1080, 1920).image for i in range(4)])
    big = F.noisy_batch_device(torch.from_numpy(clean).cuda(), 96, seed=3)
    a = scanner.scan_batch(big, want_logits=True)
    b = scanner.scan_batch(big, want_logits=True)
    for k in ("digits", "corners", "found", "logits"):
        assert torch.equal(a[k], b[k]), f"{k} not deterministic"
    perm = torch.randperm(96, device="cuda")
    c = scanner.scan_batch(big[perm].contiguous(), want_logits=True)
    for k in ("digits", "corners", "found", "logits"):
        assert torch.equal(a[k][perm], c[k]), f"{k} depends on batch order"
    sub = scanner.scan_batch(big[40:47].contiguous())
    assert torch.equal(sub["digits"], a["digits"][40:47]) and torch.equal(sub["corners"], a["corners"][40:47])
    assert int((a["found"] == 1).sum()) >= 90


# ---- drop-in modules ------------------------------------------------------------------------------
def test_dropin_modules_match_golden(golden, weights):
    import os
    import sys

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "sudoku-vision_b200", "dropin", "cv"))
    sys.path.insert(0, os.path.join(root, "sudoku-vision_b200", "dropin", "ml"))
    for m in ("preprocess", "grid", "extract", "model"):
        sys.modules.pop(m, None)
    import extract
    import grid
    import model
    import preprocess

    g = golden("frame_b")
    img = g["bgr"]
    assert np.array_equal(preprocess.grayscale(img), g["ref_gray"])
    assert np.array_equal(preprocess.blur(g["ref_gray"]), g["ref_blur"])
    assert np.array_equal(preprocess.threshold(g["ref_blur"]), g["ref_mask"])
    binary = preprocess.preprocess_for_grid_detection(img)
    assert np.array_equal(binary, g["ref_mask"])
    corners = grid.find_grid_contour(binary)
    assert corners.dtype == np.int32 and np.array_equal(corners, g["ref_corners"])
    assert np.array_equal(grid.order_points(corners.astype(np.float32)), g["ref_ordered"])
    warped = grid.warp_perspective(img, corners)
    assert warped.shape == (450, 450, 3)
    cells = extract.extract_cells(warped)
    assert isinstance(cells, list) and len(cells) == 81 and np.array_equal(np.stack(cells), g["ref_cells_u8"])
    assert grid.find_grid_contour(golden("frame_none")["ref_mask"]) is None
    with pytest.raises(NotImplementedError):
        preprocess.blur(g["ref_gray"], ksize=7)
    with pytest.raises(NotImplementedError):
        grid.warp_perspective(img, corners, inset_ratio=0.05)
    net = model.DigitCNN().to("cuda")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    net.eval()
    assert model.count_parameters(net) == 421642
    x = ((255 - g["ref_cells_thresh"]).astype(np.float32) / 255.0 - 0.5) / 0.5
    with torch.no_grad():
        outs = torch.cat([net(torch.from_numpy(x[i:i + 1]).unsqueeze(0).cuda()) for i in range(0, 81, 9)], 0)
    assert np.abs(outs.cpu().numpy() - g["ref_logits"][::9]).max() < LOGIT_TOL
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 28, 28))


# ---- edge cases --------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(3, 3), (5, 9), (8, 64), (9, 80), (63, 64), (64, 1024), (17, 2000), (130, 4000)])
def test_preprocess_edge_sizes(scanner, oracle, hw):
    """Smallest legal images, heights below one fused segment, widths that force partial strips / the stage path."""
    rng = np.random.default_rng(hw[0] * 31 + hw[1])
    img = rng.integers(0, 256, (2,) + hw + (3,)).astype(np.uint8)
    m = scanner.preprocess(_t(img)).cpu().numpy()
    for i in range(2):
        assert np.array_equal(m[i], oracle.preprocess(img[i])), hw


def test_contour_degenerate_masks(scanner, oracle):
    H, W = 96, 128
    masks = np.zeros((6, H, W), np.uint8)
    masks[1] = 255                                   # everything foreground: one component touching all borders
    masks[2, 40, 60] = 255                           # a single pixel
    masks[3, 0, :] = 255                             # a line on the top border
    masks[4, 10:80, 10:110] = 255                    # a filled rectangle (a valid quadrilateral)
    masks[5, 10:80, 10:110] = 255
    masks[5, 20:70, 20:100] = 0                      # a ring with a nested filled box inside its hole
    masks[5, 30:60, 30:90] = 255
    c, f = scanner.find_grid_contour(_t(masks))
    c, f = c.cpu().numpy(), f.cpu().numpy()
    for i in range(len(masks)):
        want = oracle.find_grid_contour(masks[i])
        assert (want is not None) == (f[i] == 1), i
        if want is not None:
            assert np.array_equal(c[i], want), i
    assert f[0] == 0 and f[2] == 0 and f[4] == 1


def test_cells_with_corners_outside_the_frame(scanner, oracle):
    """Corners beyond the image: taps outside contribute 0 (BORDER_CONSTANT) — exercises the bounds-checked sampler."""
    import torch

    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (2, 300, 400, 3)).astype(np.uint8)
    corners = np.array([[[-20, 30], [40, 340], [430, 250], [300, -15]], [[5, 5], [395, 2], [398, 297], [2, 295]]], np.int32)
    u8, pm1 = scanner.cells_from_frames(_t(img), torch.from_numpy(corners).cuda())
    board = scanner.warp_perspective(_t(img), torch.from_numpy(corners).cuda())
    for i in range(2):
        want_board = oracle.warp_perspective(img[i], corners[i])
        assert np.array_equal(board[i].cpu().numpy(), want_board)
        want_cells = oracle.extract_cells(want_board)
        assert np.array_equal(u8[i].cpu().numpy(), want_cells)
        want_in = (oracle.cell_prep(want_cells).astype(np.float32) / 255.0 - 0.5) / 0.5
        assert np.array_equal(pm1[i].cpu().numpy(), want_in)


@pytest.mark.parametrize("n", [1, 2, 127, 129, 300, 2500])
def test_digitcnn_batch_sizes(scanner, oracle, weights, n):
    import torch

    rng = np.random.default_rng(n)
    x = np.where(rng.random((n, 1, 28, 28)) < 0.3, 1.0, -1.0).astype(np.float32)
    got = scanner.digitcnn_forward(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(got - oracle.digitcnn_forward(weights, x)).max() < LOGIT_TOL


def test_classifier_only_many_cells(scanner, oracle, weights):
    """BASELINE config 3 in miniature: a large cell batch through the tcgen05 classifier vs its own fp32 CUDA-core
    twin on all cells (argmax must agree everywhere) and vs the oracle on a subsample."""
    import torch

    rng = np.random.default_rng(11)
    n = 50_000
    x = torch.from_numpy(np.where(rng.random((n, 1, 28, 28)) < 0.22, 1.0, -1.0).astype(np.float32)).cuda()
    scanner.set_classifier_mode("tc")
    lt, dt, _ = scanner.digitcnn_forward(x, want_digits=True)
    scanner.set_classifier_mode("fp32")
    lf, df, _ = scanner.digitcnn_forward(x, want_digits=True)
    scanner.set_classifier_mode("tc")
    assert float((lt - lf).abs().max()) < LOGIT_TOL
    top2 = lf.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 2 * LOGIT_TOL  # argmax is only defined up to the tolerance
    assert torch.equal(dt[decided], df[decided])
    idx = rng.choice(n, 64, replace=False)
    want = oracle.digitcnn_forward(weights, x[idx].cpu().numpy())
    assert np.abs(lt[idx].cpu().numpy() - want).max() < LOGIT_TOL
