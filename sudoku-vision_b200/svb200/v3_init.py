"""Seeded random initialisation of a DigitCNNv3 state_dict (numpy PCG64, platform-stable).  No v3 weights ship with
the reference (run_v2.load_model keeps the random init when no file exists, run_v2.py:124-125),
so benchmarks and tests build the same 91-entry state_dict (keys and shapes of ml/model_v3.py) from a seed."""
import numpy as np


def random_v3_state(seed: int = 1234) -> dict:
    rng = np.random.default_rng(seed)
    sd = {"temperature": np.ones(1, np.float32)}

    def conv(name, co, ci, k):
        sd[name + ".weight"] = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (co * k * k))).astype(np.float32)

    def bn(name, c):
        sd[name + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[name + ".bias"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        sd[name + ".running_mean"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        sd[name + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[name + ".num_batches_tracked"] = np.array(7, np.int64)

    conv("stem.0", 32, 1, 3)
    bn("stem.1", 32)
    for L, (ci, co) in zip(range(1, 6), ((32, 32), (32, 64), (64, 64), (64, 128), (128, 128))):
        p = f"layer{L}"
        conv(p + ".conv1", co, ci, 3)
        bn(p + ".bn1", co)
        conv(p + ".conv2", co, co, 3)
        bn(p + ".bn2", co)
        sd[p + ".se.excite.0.weight"] = (rng.standard_normal((co // 4, co)) * 0.2).astype(np.float32)
        sd[p + ".se.excite.2.weight"] = (rng.standard_normal((co, co // 4)) * 0.2).astype(np.float32)
        if ci != co:
            conv(p + ".shortcut.0", co, ci, 1)
            bn(p + ".shortcut.1", co)
    sd["fc.weight"] = (rng.standard_normal((10, 128)) * 0.3).astype(np.float32)
    sd["fc.bias"] = (rng.standard_normal(10) * 0.1).astype(np.float32)
    return sd
