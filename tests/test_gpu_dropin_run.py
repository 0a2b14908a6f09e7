"""GPU tier: the reference's UNCHANGED pipeline/run.py (staged copy under baseline/_ref, git-ignored but shipped to
the GPU box) executed on the B200 drop-in modules through dropin/launch.py's module seeding, on the reference's own
photos (BASELINE config 1).  Expected values were recorded from the same unmodified run.py on cv2 + torch CPU
(tests/golden/make_photo_golden.py).  Skipped when the staged copy is absent."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref" / "sudoku-vision"


@pytest.fixture(scope="module")
def ref_run():
    if not (REF / "pipeline" / "run.py").exists() or not (REF / "ml" / "digit_cnn_v2.pt").exists():
        pytest.skip("staged reference copy (baseline/_ref/sudoku-vision) not present")
    sys.path.insert(0, str(ROOT / "sudoku-vision_b200" / "dropin"))
    for m in ("preprocess", "grid", "extract", "model", "run", "cv", "launch"):
        sys.modules.pop(m, None)
    import launch

    launch.seed_modules()  # our modules under the names run.py imports
    sys.path.insert(0, str(REF / "pipeline"))
    import run

    assert "sudoku-vision_b200" in sys.modules["preprocess"].__file__  # the drop-in, not the reference's cv/
    assert "sudoku-vision_b200" in sys.modules["model"].__file__
    return run


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_unchanged_run_py_on_reference_photos(ref_run, golden, k):
    g = golden("photos_run")
    res = ref_run.run_pipeline(REF / "data" / "test_images" / f"sample_{k}.jpg")
    assert (res.warped_grid is not None) == bool(g[f"s{k}_found"])
    if not g[f"s{k}_found"]:
        assert res.error == "Grid detection failed: no quadrilateral found"  # the reference's own failure text
        return
    assert len(res.cells) == 81 and len(res.predictions) == 81
    assert np.array_equal(res.warped_grid[::90], g[f"s{k}_warp_rows"])          # warp is bit-exact
    assert np.array_equal(np.array(res.recognized_grid, np.uint8), g[f"s{k}_grid"])  # identical digits, all 81 cells
    conf = np.array([p.confidence for p in res.predictions], np.float32)
    assert np.abs(conf - g[f"s{k}_conf"]).max() < 1e-3
