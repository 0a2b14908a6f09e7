"""Drop-in for ml/model.py: `DigitCNN` keeps the reference's constructor, parameter names
(conv1/conv2/fc1/fc2 .weight/.bias), `.to()`, `.eval()`, `load_state_dict()`; forward() runs
libsvb200's kernels (svb_digitcnn_forward).  Inference only (eval-mode semantics: dropout is the
identity, ml/model.py:40); tensors must live on a CUDA device — there is no CPU path."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


class DigitCNN(nn.Module):
    """ml/model.py:19-42."""

    def __init__(self, num_classes: int = 10):
        super().__init__()
        if num_classes != 10:
            raise NotImplementedError("DigitCNN: only num_classes=10 is implemented")
        # parameter containers with the reference's names and initialisation
        self.conv1 = nn.Conv2d(1, 32, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
        self.pool = nn.MaxPool2d(2, 2)
        self.fc1 = nn.Linear(64 * 7 * 7, 128)
        self.dropout = nn.Dropout(0.5)
        self.fc2 = nn.Linear(128, num_classes)
        self._packed_versions = None

    def _sync_weights(self):
        ps = [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
              self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias]
        versions = tuple((p.data_ptr(), p._version) for p in ps)
        if versions != self._packed_versions:
            rt.scanner().load_weights({
                "conv1.weight": self.conv1.weight, "conv1.bias": self.conv1.bias,
                "conv2.weight": self.conv2.weight, "conv2.bias": self.conv2.bias,
                "fc1.weight": self.fc1.weight, "fc1.bias": self.fc1.bias,
                "fc2.weight": self.fc2.weight, "fc2.bias": self.fc2.bias})
            self._packed_versions = versions

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("DigitCNN (B200 drop-in) is inference-only: call .eval() first")
        if not x.is_cuda:
            raise RuntimeError("DigitCNN (B200 drop-in): input must be a CUDA tensor; there is no CPU path")
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 28, 28):
            raise ValueError(f"DigitCNN: expected (B,1,28,28), got {tuple(x.shape)}")
        self._sync_weights()
        return rt.scanner().digitcnn_forward(x)


def count_parameters(model: nn.Module) -> int:
    """ml/model.py:45-47."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
