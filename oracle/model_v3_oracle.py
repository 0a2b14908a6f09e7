"""CPU oracle of DigitCNNv3.forward in eval mode (ml/model_v3.py:20-37, 40-77, 163-184) written with
torch.nn.functional on CPU tensors — the reference's own dependency, its own operator order.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as np


def forward(sd: dict, x: np.ndarray, return_features: bool = False) -> np.ndarray:
    import torch
    import torch.nn.functional as F

    T = {k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}

    def bn(h, name):
        return F.batch_norm(h, T[name + ".running_mean"], T[name + ".running_var"], T[name + ".weight"], T[name + ".bias"],
                            training=False, eps=1e-5)

    def se(h, p):  # model_v3.py:33-37
        y = h.mean(dim=(2, 3))
        y = torch.sigmoid(F.linear(F.relu(F.linear(y, T[p + ".excite.0.weight"])), T[p + ".excite.2.weight"]))
        return h * y[:, :, None, None]

    def block(h, p, stride):  # model_v3.py:71-77
        out = F.relu(bn(F.conv2d(h, T[p + ".conv1.weight"], stride=stride, padding=1), p + ".bn1"))
        out = bn(F.conv2d(out, T[p + ".conv2.weight"], padding=1), p + ".bn2")
        out = se(out, p + ".se")
        if p + ".shortcut.0.weight" in T:
            h = bn(F.conv2d(h, T[p + ".shortcut.0.weight"], stride=stride), p + ".shortcut.1")
        return F.relu(out + h)

    with torch.no_grad():
        h = torch.as_tensor(np.asarray(x, np.float32))
        h = F.relu(bn(F.conv2d(h, T["stem.0.weight"], padding=1), "stem.1"))
        for L, s in zip(range(1, 6), (1, 2, 1, 2, 1)):
            h = block(h, f"layer{L}", s)
        feat = h.mean(dim=(2, 3))
        if return_features:
            return feat.numpy()
        return F.linear(feat, T["fc.weight"], T["fc.bias"]).numpy()
