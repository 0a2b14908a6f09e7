"""CPU tier: the product's find_contours core (csrc/contours_all_core.cuh, the code contours_all.cu's kernels run,
compiled for the host and orchestrated the same way) against cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) when
cv2 is importable, against the C oracle, and against the reference-minted golden (tests/golden/helpers.npz)."""
import hashlib

import numpy as np
import pytest


def _unpack(g, tag):
    h, w = (int(v) for v in g[f"{tag}_shape"])
    return (np.unpackbits(g[f"{tag}_mask"], axis=1)[:, :w] * 255).astype(np.uint8).reshape(h, w)


@pytest.mark.parametrize("tag", ["syn", "photo"])
def test_find_contours_core_golden(golden, contour_host, tag):
    g = golden("helpers")
    m = _unpack(g, tag)
    got, rounds = contour_host.find_all(m)
    assert len(got) == int(g[f"{tag}_n"]) and sum(len(c) for c in got) == int(g[f"{tag}_npts"])
    assert np.array_equal(np.array([len(c) for c in got], np.int32), g[f"{tag}_lens"])
    allp = np.concatenate(got).astype(np.int32)
    assert hashlib.sha256(allp.tobytes()).digest() == g[f"{tag}_sha"].tobytes()  # every point, cv2's order
    for i in range(min(40, len(got))):
        assert np.array_equal(got[i], g[f"{tag}_c{i}"])
    assert rounds < 64


def test_find_contours_core_vs_oracle_random(oracle, contour_host):
    """nested rings, diagonal links, full / empty masks, 1-px strokes: top-level rule and start pixels"""
    rng = np.random.default_rng(3)
    for t in range(120):
        h, w = int(rng.integers(3, 70)), int(rng.integers(3, 100))
        m = (rng.random((h, w)) < rng.choice([0.15, 0.4, 0.5, 0.6, 0.85])).astype(np.uint8) * 255
        if t % 9 == 0:
            m[:] = 255 if t % 2 else 0
        if t % 5 == 0:  # concentric rings: components inside holes must not be returned
            m[:] = 0
            for k in range(0, min(h, w) // 2, 2):
                m[k, k:w - k] = m[h - 1 - k, k:w - k] = 255
                m[k:h - k, k] = m[k:h - k, w - 1 - k] = 255
        got, _ = contour_host.find_all(m)
        want = oracle.find_contours_external(m)
        assert len(got) == len(want), (t, h, w)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), (t, h, w)


def test_find_contours_core_vs_cv2_random(contour_host):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for t in range(150):
        h, w = int(rng.integers(3, 90)), int(rng.integers(3, 130))
        m = (rng.random((h, w)) < rng.choice([0.2, 0.4, 0.5, 0.6, 0.8])).astype(np.uint8) * 255
        if t % 3 == 0:
            m = cv2.dilate(m, np.ones((2, 2), np.uint8))
        got, _ = contour_host.find_all(m)
        ref, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b.reshape(-1, 2))
