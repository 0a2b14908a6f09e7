"""CPU tier: the oracle against the LIVE OpenCV wheel (skipped if cv2 is absent).  This is the check
that detects a different OpenCV build on the box: the closed-form restatement in svb_oracle.c was
derived for cv2 4.13.0 (AVX2 dispatch)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _rng(seed):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("hw", [(28, 28), (33, 47), (64, 96), (120, 168), (270, 480)])
def test_preprocess_stages(oracle, hw):
    rng = _rng(hw[0] * 1000 + hw[1])
    img = rng.integers(0, 256, hw + (3,)).astype(np.uint8)
    g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(oracle.gray(img), g)
    b = cv2.GaussianBlur(g, (5, 5), 0)
    assert np.array_equal(oracle.blur5(g), b)
    for inv, flag in ((True, cv2.THRESH_BINARY_INV), (False, cv2.THRESH_BINARY)):
        t = cv2.adaptiveThreshold(b, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, flag, 11, 2)
        assert np.array_equal(oracle.adaptive_gauss11(b, inv), t)
    mf = cv2.GaussianBlur(b.astype(np.float32), (11, 11), 0, borderType=cv2.BORDER_REPLICATE)
    assert np.array_equal(oracle.gauss11_mean_f32(b), mf), "float Gaussian differs: another OpenCV SIMD dispatch?"


def _shapes_mask(rng, h, w):
    m = np.zeros((h, w), np.uint8)
    for _ in range(int(rng.integers(1, 9))):
        p1 = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        p2 = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        k = int(rng.integers(0, 4))
        if k == 0:
            cv2.rectangle(m, p1, p2, 255, int(rng.integers(1, 4)))
        elif k == 1:
            cv2.rectangle(m, p1, p2, 255, -1)
        elif k == 2:
            cv2.circle(m, p1, int(rng.integers(3, 70)), 255, int(rng.integers(1, 4)))
        else:
            pts = rng.integers(0, [w, h], (4, 2)).astype(np.int32)
            cv2.polylines(m, [pts], True, 255, int(rng.integers(1, 3)))
    if rng.random() < 0.3:
        m |= ((rng.random((h, w)) < 0.04) * 255).astype(np.uint8)
    if rng.random() < 0.25:  # a ring touching the frame: everything else becomes nested
        m[0:2, :] = 255
        m[-2:, :] = 255
        m[:, 0:2] = 255
        m[:, -2:] = 255
    return m


def test_contours_and_dp(oracle):
    rng = _rng(11)
    n_dp = 0
    for s in range(60):
        m = _shapes_mask(rng, int(rng.integers(30, 180)), int(rng.integers(30, 220)))
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        oc = oracle.find_contours_external(m)
        assert len(cs) == len(oc)
        for a, b in zip(cs, oc):
            a2 = a.reshape(-1, 2)
            assert np.array_equal(a2, b)
            assert cv2.contourArea(a) == oracle.contour_area(b)
            per = cv2.arcLength(a, True)
            assert per == oracle.arc_length_closed(b)
            for r in (0.005, 0.02, 0.08):
                ap = cv2.approxPolyDP(a, r * per, True).reshape(-1, 2)
                assert np.array_equal(ap, oracle.approx_poly_dp_closed(b, r * per))
                n_dp += 1
    assert n_dp > 300


def _ref_find(binary, ratio):
    cs, _ = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    floor = ratio * binary.shape[0] * binary.shape[1]
    for c in sorted(cs, key=cv2.contourArea, reverse=True):
        if cv2.contourArea(c) < floor:
            break
        poly = cv2.approxPolyDP(c, 0.02 * cv2.arcLength(c, True), True)
        if len(poly) == 4:
            return poly.reshape(4, 2)
    return None


def test_find_grid_contour_oracle_and_product_core(oracle, contour_host):
    """cv2 call sequence of cv/grid.py:37-71 vs the oracle vs the PRODUCT's contour core (the code
    the CUDA kernels run, compiled for the host)."""
    rng = _rng(13)
    hits = 0
    for s in range(150):
        bits = s % 2 == 1  # odd cases: width a multiple of 32, traced through the bit-packed view
        m = _shapes_mask(rng, int(rng.integers(40, 200)), 32 * int(rng.integers(2, 9)) if bits else int(rng.integers(40, 260)))
        for ratio in (0.03, 0.1, 0.3):
            want = _ref_find(m, ratio)
            got = oracle.find_grid_contour(m, ratio, 0.02)
            f, c = contour_host(m, ratio, 0.02, use_bits=bits)
            assert (want is None) == (got is None)
            assert f in (0, 1) and (f == 1) == (want is not None)
            if want is not None:
                hits += 1
                assert np.array_equal(want, got)
                assert np.array_equal(want, c)
    assert hits > 40


def test_product_core_on_noisy_masks_and_dense_probe_grids(oracle, contour_host):
    """Salt noise, full-width / full-height lines and tiny area floors: thousands of small loops, single-pixel loops and line
    ends ON the probe lines, i.e. every alias case of the four crossing kinds (contour_core.cuh).  The product core must
    either agree with the oracle or report the candidate-capacity status (2), never a different quad."""
    rng = _rng(2024)
    n = hits = cap = 0
    for s in range(240):
        bits = s % 2 == 1
        h = int(rng.integers(24, 160))
        w = 32 * int(rng.integers(1, 7)) if bits else int(rng.integers(24, 200))
        m = _shapes_mask(rng, h, w)
        if s % 3 == 0:
            m[rng.random((h, w)) < 0.03] = 255
            m[int(rng.integers(0, h)), :] = 255
            m[:, int(rng.integers(0, w))] = 255
        if s % 5 == 0:
            m[rng.random((h, w)) < 0.5] = 0
        for ratio in (0.0005, 0.004, 0.02, 0.1):
            want = oracle.find_grid_contour(m, ratio, 0.02)
            f, c = contour_host(m, ratio, 0.02, use_bits=bits)
            n += 1
            if f == 2:
                cap += 1
                continue
            assert f in (0, 1) and (f == 1) == (want is not None), (s, ratio)
            if want is not None:
                hits += 1
                assert np.array_equal(want, c), (s, ratio)
    assert hits > 150 and cap < n // 10


def test_warp_extract_cellprep(oracle):
    rng = _rng(17)
    img = cv2.GaussianBlur(rng.integers(0, 256, (300, 400, 3)).astype(np.uint8), (5, 5), 0)
    for corners in ([[60, 40], [50, 250], [330, 270], [350, 30]], [[10, 10], [380, 5], [390, 290], [5, 280]],
                    [[-20, 30], [40, 340], [430, 250], [300, -15]]):
        c = np.array(corners, np.int32)
        src = oracle.order_points(c)
        dst = np.array([[0, 0], [449, 0], [449, 449], [0, 449]], np.float32)
        M = cv2.getPerspectiveTransform(src, dst)
        want = cv2.warpPerspective(img, M, (450, 450))
        got = oracle.warp_perspective(img, c)
        assert np.array_equal(want, got)
        cells = oracle.extract_cells(got)
        for i in range(81):
            r, cc = divmod(i, 9)
            crop = cv2.cvtColor(want[r * 50 + 5:r * 50 + 45, cc * 50 + 5:cc * 50 + 45], cv2.COLOR_BGR2GRAY)
            assert np.array_equal(cv2.resize(crop, (28, 28)), cells[i])
        clahe = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(4, 4))
        prep = oracle.cell_prep(cells)
        for i in range(0, 81, 7):
            e = clahe.apply(cells[i])
            assert np.array_equal(e, oracle.clahe28(cells[i]))
            t = cv2.adaptiveThreshold(e, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
            assert np.array_equal(255 - t, prep[i])


def test_ref_port_matches_oracle_on_synthetic_frame(oracle, weights):
    """oracle/ref_port.py (the cv2+torch call sequence used as CPU baseline) == C oracle."""
    from oracle import ref_port
    from svb200 import frames as F

    f = F.make_frame(31337, 270, 480)
    img = F.add_noise_host(f.image, 31337)
    r = ref_port.RefScanner(weights).scan(img)
    o = oracle.scan_frame(img)
    assert r["found"] == o["found"]
    assert np.array_equal(r["mask"], o["mask"])
    if r["found"]:
        assert np.array_equal(r["corners"], o["corners"])
        assert np.array_equal(r["cells_u8"], o["cells_u8"])
        assert np.array_equal(r["cells_in"], o["cells_in"])
        x = (o["cells_in"].astype(np.float32) / 255.0 - 0.5) / 0.5
        assert np.abs(oracle.digitcnn_forward(weights, x) - r["logits"]).max() < 1e-3
