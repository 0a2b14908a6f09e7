"""Drop-in for ml/model_v3.py's DigitCNNv3: same constructor, the same 91 state_dict keys (so the reference's
checkpoints load unchanged), `.to()`, `.eval()`, `forward(x, return_features=False)`, `get_confidence`,
`set_temperature`; forward runs libsvb200's kernels (svb_digitcnn_v3_forward).  Inference only, CUDA tensors only.
DigitCNNv3Light / EmptyClassifier / MC-dropout / calibrate_temperature are not called by any pipeline
 and are not provided."""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


def _se(channels: int, reduction: int = 4) -> nn.Module:
    m = nn.Module()
    m.squeeze = nn.AdaptiveAvgPool2d(1)
    m.excite = nn.Sequential(nn.Linear(channels, channels // reduction, bias=False), nn.ReLU(inplace=True),
                             nn.Linear(channels // reduction, channels, bias=False), nn.Sigmoid())
    return m


def _block(cin: int, cout: int, stride: int, use_se: bool) -> nn.Module:
    """Parameter container with ResidualBlock's names (ml/model_v3.py:40-69)."""
    m = nn.Module()
    m.conv1 = nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False)
    m.bn1 = nn.BatchNorm2d(cout)
    m.conv2 = nn.Conv2d(cout, cout, kernel_size=3, stride=1, padding=1, bias=False)
    m.bn2 = nn.BatchNorm2d(cout)
    m.se = _se(cout) if use_se else nn.Identity()
    if stride != 1 or cin != cout:
        m.shortcut = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False), nn.BatchNorm2d(cout))
    else:
        m.shortcut = nn.Identity()
    return m


class DigitCNNv3(nn.Module):
    """ml/model_v3.py:95-229."""

    def __init__(self, num_classes: int = 10, dropout: float = 0.5, use_se: bool = True):
        super().__init__()
        if num_classes != 10 or not use_se:
            raise NotImplementedError("DigitCNNv3 (B200 drop-in): only num_classes=10, use_se=True are implemented")
        self.stem = nn.Sequential(nn.Conv2d(1, 32, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(32),
                                  nn.ReLU(inplace=True))
        self.layer1 = _block(32, 32, 1, use_se)
        self.layer2 = _block(32, 64, 2, use_se)
        self.layer3 = _block(64, 64, 1, use_se)
        self.layer4 = _block(64, 128, 2, use_se)
        self.layer5 = _block(128, 128, 1, use_se)
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(128, num_classes)
        self.temperature = nn.Parameter(torch.ones(1), requires_grad=False)
        for m in self.modules():  # the reference's initialisation (model_v3.py:151-161)
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self._packed = None

    def _sync_weights(self):
        sd = self.state_dict()
        sig = tuple((v.data_ptr(), v._version) for v in sd.values())
        if sig != self._packed:
            rt.scanner().load_weights_v3(sd)
            self._packed = sig

    def forward(self, x: torch.Tensor, return_features: bool = False) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("DigitCNNv3 (B200 drop-in) is inference-only: call .eval() first")
        if not x.is_cuda:
            raise RuntimeError("DigitCNNv3 (B200 drop-in): input must be a CUDA tensor; there is no CPU path")
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 28, 28):
            raise ValueError(f"DigitCNNv3: expected (B,1,28,28), got {tuple(x.shape)}")
        self._sync_weights()
        return rt.scanner().digitcnn_v3_forward(x, want_features=return_features)

    def get_confidence(self, x: torch.Tensor):
        """model_v3.py:215-225: softmax(logits / temperature) -> (predicted, confidence)."""
        probs = F.softmax(self.forward(x) / self.temperature, dim=1)
        confidence, predicted = probs.max(dim=1)
        return predicted, confidence

    def set_temperature(self, temperature: float):
        self.temperature.data.fill_(temperature)


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
