"""Deterministic DigitCNNv3 state_dict (numpy PCG64, platform-stable) shared by tests/golden/make_golden.py and the
tests: no v3 weights ship with the reference, and 698,475 floats are too large for a fixture, so
both sides rebuild them from a seed.  Keys and shapes are those of ml/model_v3.py's state_dict (91 entries)."""
import numpy as np


import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "sudoku-vision_b200"))
from svb200.v3_init import random_v3_state as make_v3_state  # noqa: E402,F401  (one definition, shared with bench.py)


def make_v3_inputs(seed: int = 99, n: int = 40) -> np.ndarray:
    """Half +-1 patterns (what the pipeline feeds), half Gaussian (the drop-in accepts any float tensor)."""
    rng = np.random.default_rng(seed)
    a = np.where(rng.random((n // 2, 1, 28, 28)) < 0.25, 1.0, -1.0)
    b = rng.standard_normal((n - n // 2, 1, 28, 28))
    return np.concatenate([a, b]).astype(np.float32)
