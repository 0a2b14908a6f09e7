"""CPU tier, row V1 (cv/preprocess_v2.py): the oracle's restatement against golden vectors minted from the
unmodified reference module (tests/golden/make_v2pre_golden.py) and, when importable, the live cv2 wheel.
Everything on this row is integer / byte work or elementwise fp32 with a fixed operation order: bit-exact."""
import numpy as np
import pytest

CASES = [f"{t}_{v}" for t in ("a", "b") for v in ("plain", "shadow", "glare", "flat")]


@pytest.fixture(scope="module")
def v2pre(golden):
    return golden("v2pre")


@pytest.mark.parametrize("case", CASES)
def test_preprocess_for_grid_detection_v2(oracle, v2pre, case):
    img = v2pre[case + "_bgr"]
    mask, glare, shadow = oracle.preprocess_v2(img)
    assert np.array_equal(mask, v2pre[case + "_ref_mask"])
    assert [glare, shadow] == v2pre[case + "_ref_flags"].tolist()
    assert np.array_equal(oracle.preprocess_v2(img, False, True)[0], v2pre[case + "_ref_mask_noillum"])
    assert np.array_equal(oracle.preprocess_v2(img, True, False)[0], v2pre[case + "_ref_mask_noshadow"])


@pytest.mark.parametrize("case", CASES)
def test_preprocess_multi_strategy(oracle, v2pre, case):
    img = v2pre[case + "_bgr"]
    r = oracle.preprocess_multi(img)
    for k, ref in (("binary", "_ref_binary"), ("gray", "_ref_gray"), ("enhanced", "_ref_enhanced"),
                   ("illumination_normalized", "_ref_illum")):
        assert np.array_equal(r[k], v2pre[case + ref]), k
    assert [r["has_glare"], r["has_shadow"]] == v2pre[case + "_ref_flags"].tolist()
    assert r["method_used"] == str(v2pre[case + "_ref_method"])
    bl = oracle.blur5(r["enhanced"])
    assert np.array_equal(oracle.otsu_inv(bl)[1], v2pre[case + "_ref_otsu"])
    assert np.array_equal(oracle.sauvola(bl), v2pre[case + "_ref_sauvola"])


def test_flags_cover_both_branches(v2pre):
    fl = np.array([v2pre[c + "_ref_flags"] for c in CASES])
    assert fl[:, 0].any() and not fl[:, 0].all() and fl[:, 1].any() and not fl[:, 1].all()
    assert len({str(v2pre[c + "_ref_method"]) for c in CASES}) >= 2


def test_primitives(oracle, v2pre):
    import ctypes as C

    g = v2pre["u_gray"]
    assert np.array_equal(oracle.box_blur(g, 13), v2pre["u_ref_blur13"])
    assert np.array_equal(oracle.dilate_ellipse(g, 7), v2pre["u_ref_dilate7"])
    assert np.array_equal(oracle.erode_ellipse(oracle.dilate_ellipse(g, 51), 51), v2pre["u_ref_close51"])
    assert np.array_equal(oracle.gaussian_blur_q8(g, 21), v2pre["u_ref_gauss21"])
    assert np.array_equal(oracle.clahe_frame(g), v2pre["u_ref_clahe8"])
    assert np.array_equal(oracle.divide_normalize(g, v2pre["u_ref_close51"]), v2pre["u_ref_illum"])
    assert np.array_equal(oracle.divide_normalize(g, oracle.gaussian_blur_q8(oracle.dilate_ellipse(g, 7), 21)),
                          v2pre["u_ref_noshadow"])
    assert np.array_equal(oracle.morph_cleanup(v2pre["u_mask"]), v2pre["u_ref_cleanup"])
    for k in (7, 21, 51, 193, 385):  # getStructuringElement(MORPH_ELLIPSE): chord half-widths
        hw = (C.c_int * k)()
        oracle.lib().svo_ellipse_rows(k, hw)
        e = v2pre[f"u_ellipse{k}"]
        want = [(int(r.sum()) - 1) // 2 for r in e]
        assert list(hw) == want
        for i, r in enumerate(e):  # chords are centred
            assert r[k // 2 - want[i]: k // 2 + want[i] + 1].all()
    # the chord-wise evaluation equals the direct definition
    for k in (7, 21):
        for d in (True, False):
            a = oracle.dilate_ellipse(g, k) if d else oracle.erode_ellipse(g, k)
            assert np.array_equal(oracle.morph_ellipse_direct(g, k, d), a)


def test_against_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for (h, w) in ((64, 88), (130, 97)):
        g = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (3, 3), 0)
        assert np.array_equal(oracle.box_blur(g, 9), cv2.blur(g, (9, 9)))
        for k in (7, 33):
            se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
            assert np.array_equal(oracle.dilate_ellipse(g, k), cv2.dilate(g, se))
            assert np.array_equal(oracle.erode_ellipse(g, k), cv2.erode(g, se))
        assert np.array_equal(oracle.gaussian_blur_q8(g, 21), cv2.GaussianBlur(g, (21, 21), 0))
        t, b = cv2.threshold(g, 0, 255, cv2.THRESH_BINARY_INV + cv2.THRESH_OTSU)
        lvl, ob = oracle.otsu_inv(g)
        assert lvl == int(t) and np.array_equal(ob, b)
        m = ((rng.random((h, w)) < 0.4) * 255).astype(np.uint8)
        c = cv2.morphologyEx(m, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3)))
        c = cv2.morphologyEx(c, cv2.MORPH_OPEN, cv2.getStructuringElement(cv2.MORPH_RECT, (2, 2)))
        assert np.array_equal(oracle.morph_cleanup(m), c)
    # CLAHE incl. sides that do not divide by the 8x8 tile grid (OpenCV extends BOTH sides by REFLECT_101 then)
    for hw in ((96, 160), (97, 160), (96, 163), (33, 47), (129, 131), (270, 484)):
        g = cv2.GaussianBlur(rng.integers(0, 256, hw).astype(np.uint8), (3, 3), 0)
        assert np.array_equal(oracle.clahe_frame(g), cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(g)), hw


def test_grid_quality_oracle_vs_reference_golden(golden, v2pre):
    """oracle/oracle_quality.py vs the unmodified cv/grid_quality.py (tests/golden/make_quality_golden.py)."""
    from oracle import oracle_quality as Q

    g = golden("quality")
    for case in CASES:
        r = Q.assess(v2pre[case + "_bgr"], v2pre[case + "_ref_binary"], g[case + "_corners"])
        got = np.array([r[f] for f in Q.FIELDS])
        assert np.abs(got - g[case + "_ref_scores"]).max() < 1e-9, case
    ov = np.array([g[c + "_ref_scores"][0] for c in CASES])
    assert ov.min() < 55 < ov.max()
