"""GPU tier (-m gpu), row V1: the cv/preprocess_v2.py kernels (csrc/preprocess_v2.cu) through the C ABI against the CPU
oracle on seeded inputs, against the golden vectors minted from the reference module, and at BASELINE's full frame
sizes (1080p, 4K).  Everything on this row is byte/integer work or fixed-order float32: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [f"{t}_{v}" for t in ("a", "b") for v in ("plain", "shadow", "glare", "flat")]


def _t(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _np(t):
    return t.cpu().numpy()


def _smooth(rng, h, w):
    """noise + structure: smooth gradients, dark lines, a bright patch (ties and saturation both occur)"""
    yy, xx = np.mgrid[0:h, 0:w]
    g = 130 + 60 * np.sin(xx / 41.0) + 40 * np.cos(yy / 29.0) + rng.normal(0, 6, (h, w))
    g[:, ::37] -= 90
    g[::23, :] -= 70
    g[h // 5: h // 3, w // 2: w // 2 + w // 6] = 255
    return g.clip(0, 255).astype(np.uint8)


@pytest.fixture(scope="module")
def v2pre(golden):
    return golden("v2pre")


@pytest.mark.parametrize("hw", [(40, 56), (96, 136), (130, 97), (264, 520)])
def test_primitives_vs_oracle(scanner, oracle, hw):
    rng = np.random.default_rng(hw[0] * 31 + hw[1])
    imgs = np.stack([_smooth(rng, *hw), rng.integers(0, 256, hw).astype(np.uint8)])
    d = _t(imgs)
    for k in (3, 9, 13, 25):
        got = _np(scanner.v2_stage("box_blur", d, k)[0])
        for i in range(2):
            assert np.array_equal(got[i], oracle.box_blur(imgs[i], k)), f"box blur k={k}"
    for k in (1, 3, 7, 21, 33):
        if k // 2 >= min(hw):
            continue
        for op, fn in (("dilate_ellipse", oracle.dilate_ellipse), ("erode_ellipse", oracle.erode_ellipse)):
            got = _np(scanner.v2_stage(op, d, k)[0])
            for i in range(2):
                assert np.array_equal(got[i], fn(imgs[i], k)), f"{op} k={k}"
    got = _np(scanner.v2_stage("gauss21", d)[0])
    rs = _np(scanner.v2_stage("remove_shadow", d)[0])
    ni = _np(scanner.v2_stage("normalize_illumination", d)[0])
    ot, oinfo = scanner.v2_stage("otsu", d)
    sv = _np(scanner.v2_stage("sauvola", d)[0])
    ki = max(max(hw) // 10 + (max(hw) // 10 % 2 == 0), 51)  # cv/preprocess_v2.py:46-49
    for i in range(2):
        assert np.array_equal(got[i], oracle.gaussian_blur_q8(imgs[i], 21))
        assert np.array_equal(rs[i], oracle.divide_normalize(imgs[i], oracle.gaussian_blur_q8(oracle.dilate_ellipse(imgs[i], 7), 21)))
        assert np.array_equal(ni[i], oracle.divide_normalize(imgs[i], oracle.erode_ellipse(oracle.dilate_ellipse(imgs[i], ki), ki)))
        lvl, want = oracle.otsu_inv(imgs[i])
        assert int(_np(oinfo)[i, 3]) == lvl and np.array_equal(_np(ot)[i], want)
        assert np.array_equal(sv[i], oracle.sauvola(imgs[i]))
    cl = _np(scanner.v2_stage("clahe8", d)[0])  # sides that do not divide by 8: OpenCV's REFLECT_101-extended tile grid
    for i in range(2):
        assert np.array_equal(cl[i], oracle.clahe_frame(imgs[i]))
    masks = np.stack([((rng.random(hw) < p) * 255).astype(np.uint8) for p in (0.3, 0.7)])
    cu = _np(scanner.v2_stage("cleanup", _t(masks))[0])
    for i in range(2):
        assert np.array_equal(cu[i], oracle.morph_cleanup(masks[i]))


def test_ellipse_large_elements_and_wide_frames(scanner, oracle):
    """k = 193 / 385 (the 1080p / 4K illumination kernels) and frames wider than one 2048-column strip."""
    rng = np.random.default_rng(77)
    for (h, w, k) in ((260, 400, 193), (120, 2304, 51), (96, 4100, 33), (400, 420, 385)):
        img = np.stack([_smooth(rng, h, w)])
        for op, fn in (("dilate_ellipse", oracle.dilate_ellipse), ("erode_ellipse", oracle.erode_ellipse)):
            got = _np(scanner.v2_stage(op, _t(img), k)[0])[0]
            assert np.array_equal(got, fn(img[0], k)), f"{op} k={k} {h}x{w}"


@pytest.mark.parametrize("hw", [(270, 484), (275, 480), (301, 517), (100, 203)])
def test_whole_path_on_sides_that_do_not_divide_by_8(scanner, oracle, hw):
    """preprocess_for_grid_detection / preprocess_multi_strategy on ragged frame sizes (CLAHE's padded tile grid, the
    scalar column tails of the float Gaussian, unaligned rows everywhere): bit-exact against the oracle, which matches the
    reference module on these sizes (checked in the build container)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sudoku-vision_b200"))
    from svb200 import frames as F

    h, w = hw
    imgs = []
    for k in range(2):
        img = F.make_frame(7100 + h + k, h, w).image
        if k:  # a shadow gradient and a glare patch, so both flags and another strategy come into play
            xx = np.mgrid[0:h, 0:w][1]
            img = (img.astype(np.float32) * (0.45 + 0.55 * xx[..., None] / w)).clip(0, 255).astype(np.uint8)
            img[10:40, 10:60] = 255
        imgs.append(img)
    batch = _t(np.stack(imgs))
    mask, info = scanner.preprocess_v2(batch)
    r = scanner.preprocess_multi(batch)
    for i, img in enumerate(imgs):
        want, glare, shadow = oracle.preprocess_v2(img)
        assert np.array_equal(_np(mask)[i], want)
        assert _np(info)[i, :2].astype(bool).tolist() == [glare, shadow]
        om = oracle.preprocess_multi(img)
        for key in ("binary", "gray", "enhanced", "illumination_normalized"):
            assert np.array_equal(_np(r[key])[i], om[key]), key
        assert scanner.V2_METHODS[int(_np(r["info"])[i, 2])] == om["method_used"]


def test_detect_glare_and_shadow(scanner, oracle, v2pre):
    for case in CASES:
        g = v2pre[case + "_ref_gray"]
        flags = v2pre[case + "_ref_flags"].tolist()
        gm, gi = scanner.v2_stage("detect_glare", _t(g[None]))
        sm, si = scanner.v2_stage("detect_shadow", _t(g[None]))
        assert [bool(_np(gi)[0, 0]), bool(_np(si)[0, 1])] == flags
        assert np.array_equal(_np(gm)[0], (g > 250).astype(np.uint8) * 255)
        k = max(g.shape) // 20
        k += (k % 2 == 0)
        want = ((g.astype(np.int32) - oracle.box_blur(g, k).astype(np.int32)) < -30).astype(np.uint8) * 255
        assert np.array_equal(_np(sm)[0], want)


@pytest.mark.parametrize("case", CASES)
def test_golden_preprocess_v2(scanner, v2pre, case):
    img = _t(v2pre[case + "_bgr"][None])
    mask, info = scanner.preprocess_v2(img)
    assert np.array_equal(_np(mask)[0], v2pre[case + "_ref_mask"])
    assert _np(info)[0, :2].astype(bool).tolist() == v2pre[case + "_ref_flags"].tolist()
    assert np.array_equal(_np(scanner.preprocess_v2(img, False, True)[0])[0], v2pre[case + "_ref_mask_noillum"])
    assert np.array_equal(_np(scanner.preprocess_v2(img, True, False)[0])[0], v2pre[case + "_ref_mask_noshadow"])
    assert np.array_equal(_np(scanner.preprocess_v2(img, False, False)[0])[0].shape, v2pre[case + "_ref_mask"].shape)


@pytest.mark.parametrize("case", CASES)
def test_golden_preprocess_multi(scanner, v2pre, case):
    r = scanner.preprocess_multi(_t(v2pre[case + "_bgr"][None]))
    for k, ref in (("binary", "_ref_binary"), ("gray", "_ref_gray"), ("enhanced", "_ref_enhanced"),
                   ("illumination_normalized", "_ref_illum")):
        assert np.array_equal(_np(r[k])[0], v2pre[case + ref]), k
    info = _np(r["info"])[0]
    assert info[:2].astype(bool).tolist() == v2pre[case + "_ref_flags"].tolist()
    assert scanner.V2_METHODS[int(info[2])] == str(v2pre[case + "_ref_method"])


def test_mixed_batch_flags_are_per_frame(scanner, v2pre):
    """Frames with and without shadow / glare in ONE batch: the conditional remove_shadow and the strategy choice are
    decided per frame on the device."""
    names = [c for c in CASES if c.startswith("a_")]
    batch = np.stack([v2pre[c + "_bgr"] for c in names])
    r = scanner.preprocess_multi(_t(batch))
    m, _ = scanner.preprocess_v2(_t(batch))
    for i, c in enumerate(names):
        assert np.array_equal(_np(r["binary"])[i], v2pre[c + "_ref_binary"]), c
        assert np.array_equal(_np(r["illumination_normalized"])[i], v2pre[c + "_ref_illum"]), c
        assert np.array_equal(_np(m)[i], v2pre[c + "_ref_mask"]), c
    flags = _np(r["info"])[:, :2].astype(bool)
    assert flags[:, 1].any() and not flags[:, 1].all()


def test_gray_input_passes_through(scanner, oracle, v2pre):
    g = v2pre["a_plain_ref_gray"]
    bgr = np.repeat(g[..., None], 3, axis=2)  # gray(B=G=R=g) == g for this fixed-point formula
    assert np.array_equal(oracle.gray(bgr), g)
    a, _ = scanner.preprocess_v2(_t(g[None]))
    b, _ = scanner.preprocess_v2(_t(bgr[None]))
    assert np.array_equal(_np(a), _np(b))


def _full_size_frames(h, w, seed):
    from svb200 import frames as F

    img = F.add_noise_host(F.make_frame(seed, h, w, 12.0).image, seed)
    sh = img.copy()  # dark bands a 32nd of the width wide, every fourth one: ~24 % "shadow" pixels -> has_shadow
    band = (np.arange(w) // (w // 32)) % 4 == 0
    sh[:, band] = (sh[:, band] * 0.45).astype(np.uint8)
    return np.stack([img, sh])


def test_full_size_1080p(scanner, oracle):
    batch = _full_size_frames(1080, 1920, 606)
    r = scanner.preprocess_multi(_t(batch))
    m, info = scanner.preprocess_v2(_t(batch))
    shadows = []
    for i in range(2):
        o = oracle.preprocess_multi(batch[i])
        for k in ("binary", "gray", "enhanced", "illumination_normalized"):
            assert np.array_equal(_np(r[k])[i], o[k]), (i, k)
        got = _np(r["info"])[i]
        assert [bool(got[0]), bool(got[1])] == [o["has_glare"], o["has_shadow"]]
        assert scanner.V2_METHODS[int(got[2])] == o["method_used"] and int(got[3]) == o["otsu_level"]
        assert np.array_equal(_np(m)[i], oracle.preprocess_v2(batch[i])[0])
        shadows.append(o["has_shadow"])
    assert shadows == [False, True]


def test_full_size_4k(scanner, oracle):
    """BASELINE configs[3] frame size: k = 385 elliptical close, 480x270 CLAHE tiles."""
    batch = _full_size_frames(2160, 3840, 707)[1:]
    m, info = scanner.preprocess_v2(_t(batch))
    want, glare, shadow = oracle.preprocess_v2(batch[0])
    assert np.array_equal(_np(m)[0], want)
    assert [bool(x) for x in _np(info)[0, :2]] == [glare, shadow] and shadow


def test_unsupported_sizes_raise(scanner):
    import torch

    m, _ = scanner.preprocess_v2(torch.zeros((1, 100, 100, 3), dtype=torch.uint8, device="cuda"))  # sides not / 8: supported
    assert m.shape == (1, 100, 100)
    with pytest.raises(NotImplementedError):  # the 4 x 4100-px illumination kernel is beyond the chord tables
        scanner.preprocess_v2(torch.zeros((1, 64, 4104, 3), dtype=torch.uint8, device="cuda"))
    with pytest.raises(NotImplementedError):
        scanner.v2_stage("dilate_ellipse", torch.zeros((1, 64, 64), dtype=torch.uint8, device="cuda"), 401)


def test_dropin_module_matches_reference_golden(v2pre):
    """sudoku-vision_b200/dropin/cv/preprocess_v2.py: the reference's names and return types."""
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "sudoku-vision_b200", "dropin"))
    from cv import preprocess_v2 as P

    c = "a_shadow"
    img = v2pre[c + "_bgr"]
    assert np.array_equal(P.preprocess_for_grid_detection(img), v2pre[c + "_ref_mask"])
    r = P.preprocess_multi_strategy(img)
    assert isinstance(r, P.PreprocessResult)
    assert np.array_equal(r.binary, v2pre[c + "_ref_binary"]) and np.array_equal(r.enhanced, v2pre[c + "_ref_enhanced"])
    assert np.array_equal(r.illumination_normalized, v2pre[c + "_ref_illum"]) and np.array_equal(r.gray, v2pre[c + "_ref_gray"])
    assert (r.has_glare, r.has_shadow) == tuple(bool(x) for x in v2pre[c + "_ref_flags"]) and r.method_used == str(v2pre[c + "_ref_method"])
    g = v2pre["u_gray"]
    assert np.array_equal(P.normalize_illumination(g), v2pre["u_ref_illum"])
    assert np.array_equal(P.remove_shadow(g), v2pre["u_ref_noshadow"])
    assert np.array_equal(P.apply_clahe(g), v2pre["u_ref_clahe8"])
    assert np.array_equal(P.morphological_cleanup(v2pre["u_mask"]), v2pre["u_ref_cleanup"])
    assert np.array_equal(P.grayscale(img), v2pre[c + "_ref_gray"])
    has, mask = P.detect_glare(v2pre["a_glare_ref_gray"])
    assert has is True and mask.dtype == np.uint8
    with pytest.raises(NotImplementedError):
        P.apply_clahe(g, clip_limit=3.0)
    with pytest.raises(NotImplementedError):
        P.threshold_sauvola(g, window_size=15)


def test_scan_batch_v2_whole_path(scanner, oracle):
    """svb_scan_batch_v2 = run_v2.py's CV + ML sections (--no-quality-check, detection method 1): every stage against
    the oracle's composition on seeded frames; logits within 1e-3, digits / top-3 order identical."""
    import os
    import sys

    import torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from v3_weights import make_v3_state

    from oracle import model_v3_oracle as M
    from oracle import oracle_v2
    from svb200 import frames as F

    sd = make_v3_state()
    scanner.load_weights_v3(sd)
    imgs, _, _ = F.make_frames(3, 544, 960, base_seed=808)
    blank = np.full((1, 544, 960, 3), 127, np.uint8)
    batch = np.concatenate([imgs, blank])
    out = scanner.scan_batch_v2(_t(batch), want_logits=True)
    torch.cuda.synchronize()
    n_found = 0
    for i in range(len(batch)):
        r = oracle.preprocess_multi(batch[i])
        info = _np(out["info"])[i]
        assert scanner.V2_METHODS[int(info[2])] == r["method_used"]
        c = oracle_v2.detect_grid_contour(r["binary"])
        assert bool(_np(out["found"])[i] == 1) == (c is not None)
        if c is None:
            assert not _np(out["digits"])[i].any() and not _np(out["conf"])[i].any()
            continue
        n_found += 1
        assert np.array_equal(_np(out["corners"])[i].astype(np.float32), c)
        board = oracle.warp_perspective(batch[i], c.astype(np.int32))
        x = (oracle.cell_prep(oracle.extract_cells(board)).astype(np.float32) / 255.0 - 0.5) / 0.5
        want = M.forward(sd, x.reshape(81, 1, 28, 28))
        got = _np(out["logits"])[i]
        assert np.abs(got - want).max() < 1e-3
        probs = torch.softmax(torch.from_numpy(want), 1)
        tp, ti = probs.topk(3)
        margin_ok = (tp[:, :-1] - tp[:, 1:]).min(1).values.numpy() > 1e-4  # compare order only where it is not a near-tie
        got_idx = np.concatenate([_np(out["digits"])[i][:, None], _np(out["alt_digits"])[i]], 1)
        got_p = np.concatenate([_np(out["conf"])[i][:, None], _np(out["alt_conf"])[i]], 1)
        assert np.array_equal(got_idx[margin_ok], ti.numpy()[margin_ok].astype(np.uint8))
        assert np.abs(got_p - tp.numpy()).max() < 1e-4
    assert n_found >= 2


QUALITY_TOL = 1e-3  # scores are 0..100 floats; float32 corner arithmetic (acos) may differ in the last bits


def test_grid_quality_golden_and_full_size(scanner, oracle, golden, v2pre):
    """svb_assess_grid_quality vs the reference's scores (golden) and vs the oracle on a 1080p frame."""
    from oracle import oracle_quality as Q
    from oracle import oracle_v2
    from svb200 import frames as F

    g = golden("quality")
    for tag in ("a", "b"):
        names = [c for c in CASES if c.startswith(tag + "_")]
        frames = _t(np.stack([v2pre[c + "_bgr"] for c in names]))
        binary = _t(np.stack([v2pre[c + "_ref_binary"] for c in names]))
        corners = _t(np.stack([g[c + "_corners"].astype(np.int32) for c in names]))
        got = _np(scanner.assess_grid_quality(frames, binary, corners))
        for i, c in enumerate(names):
            assert np.abs(got[i] - g[c + "_ref_scores"]).max() < QUALITY_TOL, (c, got[i], g[c + "_ref_scores"])
        gray = _t(np.stack([v2pre[c + "_ref_gray"] for c in names]))  # gray input: same scores
        assert np.abs(_np(scanner.assess_grid_quality(gray, binary, corners)) - got).max() < 1e-9
    img = F.add_noise_host(F.make_frame(909, 1080, 1920, 10.0).image, 909)
    r = oracle.preprocess_multi(img)
    c = oracle_v2.detect_grid_contour(r["binary"])
    assert c is not None
    want = Q.assess(img, r["binary"], c)
    got = _np(scanner.assess_grid_quality(_t(img[None]), _t(r["binary"][None]), _t(c.astype(np.int32)[None])))[0]
    assert np.abs(got - np.array([want[f] for f in Q.FIELDS])).max() < QUALITY_TOL


def test_scan_batch_v2_quality_gate(scanner, v2pre):
    """min_quality_score: frames below it leave the pipeline as 'quality_failed' (found == 3, zeros), the others are
    scanned exactly as without the gate (run_v2.py:300-308)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from v3_weights import make_v3_state

    scanner.load_weights_v3(make_v3_state())
    names = [c for c in CASES if c.startswith("a_")]
    batch = _t(np.stack([v2pre[c + "_bgr"] for c in names]))
    free = scanner.scan_batch_v2(batch)
    q = _np(free["quality"])[:, 0]
    found0 = _np(free["found"])
    ok = found0 == 1
    assert ok.any()
    thr = float(np.median(q[ok])) + 1e-6 if ok.sum() > 1 else float(q[ok][0]) + 1.0
    gated = scanner.scan_batch_v2(batch, min_quality_score=thr)
    f1 = _np(gated["found"])
    for i in range(len(names)):
        if found0[i] == 1 and q[i] < thr:
            assert f1[i] == 3 and not _np(gated["digits"])[i].any() and not _np(gated["conf"])[i].any()
        else:
            assert f1[i] == found0[i]
            assert np.array_equal(_np(gated["digits"])[i], _np(free["digits"])[i])
    assert (f1 == 3).any()


def test_grid_quality_dropin(golden, v2pre):
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "sudoku-vision_b200", "dropin"))
    from cv import grid_quality as GQ

    g = golden("quality")
    for c in ("a_flat", "b_plain"):
        q = GQ.assess_grid_quality(v2pre[c + "_bgr"], v2pre[c + "_ref_binary"], g[c + "_corners"])
        assert isinstance(q, GQ.QualityScore)
        assert abs(q.overall - g[c + "_ref_scores"][0]) < QUALITY_TOL
        assert GQ.get_user_feedback(q) == str(g[c + "_ref_feedback"])
    with pytest.raises(NotImplementedError):
        GQ.assess_grid_quality(v2pre["a_plain_bgr"], v2pre["a_plain_ref_binary"], g["a_plain_corners"] + 0.5)


def test_batched_solver_vs_reference_golden(scanner, oracle, golden):
    """svb_solve_batch: same status and the same solution (also when several exist) as the reference's solver."""
    g = golden("solver")
    sol, st = scanner.solve_batch(_t(g["grids"]))
    assert np.array_equal(_np(st), g["ref_status"])
    assert np.array_equal(_np(sol), g["ref_solutions"])
    # a large ragged batch against the oracle: 3000 puzzles incl. (9,9)-shaped input
    rng = np.random.default_rng(5)
    base = np.array([[(3 * (r % 3) + r // 3 + c) % 9 + 1 for c in range(9)] for r in range(9)], np.uint8)
    grids = np.stack([(rng.permutation(9) + 1).astype(np.uint8)[base - 1] for _ in range(3000)])
    for k in range(len(grids)):
        flat = grids[k].reshape(-1)
        flat[rng.permutation(81)[: int(rng.integers(35, 60))]] = 0
        if k % 7 == 0:
            idx = np.flatnonzero(flat)
            flat[idx[-1]] = flat[idx[-1]] % 9 + 1
    sol, st = scanner.solve_batch(_t(grids))
    assert sol.shape == (3000, 9, 9)
    for k in range(0, 3000, 13):
        want_st, want = oracle.solve_sudoku(grids[k])
        assert int(_np(st)[k]) == want_st and np.array_equal(_np(sol)[k], want)


def test_v2_preprocess_on_a_reference_photo(scanner, oracle):
    """One of the reference's own photos (2736 x 3648: k = 365 illumination kernel, two 2048-column strips) through
    preprocess_for_grid_detection, against the oracle.  Needs the staged reference copy (baseline/_ref, git-ignored but
    shipped to the GPU box) and cv2 for the JPEG decode only."""
    import os

    cv2 = pytest.importorskip("cv2")
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "sudoku-vision", "data",
                     "test_images", "sample_3.jpg")
    if not os.path.exists(p):
        pytest.skip("staged reference copy not present")
    img = cv2.imread(p)
    assert img.shape[0] % 8 == 0 and img.shape[1] % 8 == 0
    m, info = scanner.preprocess_v2(_t(img[None]))
    want, glare, shadow = oracle.preprocess_v2(img)
    assert np.array_equal(_np(m)[0], want)
    assert [bool(x) for x in _np(info)[0, :2]] == [glare, shadow]


def test_v2_contours_on_full_size_masks_are_complete_and_repeatable(scanner, oracle):
    """The v2 masks of noisy 1080p frames hold thousands of small components, hence thousands of probe crossings per frame:
    the crossing list must hold them (no frame may come back with status 2), and the result must not depend on the
    order in which the crossings were packed (same corners on every repetition, equal to the oracle's)."""
    from oracle import oracle_v2
    from svb200 import frames as F
    import torch

    clean = _t(np.stack([F.make_frame(41000 + i, 1080, 1920).image for i in range(2)]))
    batch = F.noisy_batch_device(clean, 24, seed=11)
    r = scanner.preprocess_multi(batch, want_aux=False)
    ref_c = ref_f = None
    for rep in range(3):
        c, f = scanner.detect_grid_contour_v2(r["binary"])
        torch.cuda.synchronize()
        c, f = _np(c), _np(f)
        assert set(f.tolist()) <= {0, 1}, f"capacity status in {f.tolist()}"
        if ref_c is None:
            ref_c, ref_f = c, f
        assert np.array_equal(c, ref_c) and np.array_equal(f, ref_f)
    for i in (0, 23):
        want = oracle_v2.detect_grid_contour(_np(r["binary"])[i])
        assert (want is not None) == bool(ref_f[i] == 1)
        if want is not None:
            assert np.array_equal(ref_c[i].astype(np.float32), want)
    assert int(ref_f.sum()) >= 20
