"""Drop-in for cv/grid_quality.py — same names and return types; the five scores are computed on the GPU
(svb_assess_grid_quality: Laplacian variance, histogram range, line-band coverage of the warped mask, corner geometry
and size) and agree with the reference to float rounding; issues / recommendations / feedback strings are the
reference's thresholds applied to those numbers (cv/grid_quality.py:273-325)."""
import os
import sys
from dataclasses import dataclass, field
from typing import List

import numpy as np
from numpy.typing import NDArray

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _runtime as rt  # noqa: E402


@dataclass
class QualityScore:
    """cv/grid_quality.py:22-45."""
    overall: float
    sharpness: float
    contrast: float
    completeness: float
    geometry: float
    size: float
    issues: List[str] = field(default_factory=list)
    recommendations: List[str] = field(default_factory=list)

    @property
    def is_acceptable(self) -> bool:
        return self.overall >= 50

    @property
    def is_good(self) -> bool:
        return self.overall >= 70


def _int_corners(corners) -> np.ndarray:
    c = np.asarray(corners, dtype=np.float64).reshape(4, 2)
    if not np.array_equal(c, np.round(c)):
        raise NotImplementedError("assess_grid_quality: non-integral corners (only the contour method's corners are supported)")
    return c.astype(np.int32)


def assess_grid_quality(image: NDArray[np.uint8], binary: NDArray[np.uint8], corners: NDArray[np.float32]) -> QualityScore:
    """cv/grid_quality.py:228-306."""
    import torch

    s = rt.scanner()
    scores = s.assess_grid_quality(rt.to_device_u8(image)[None], rt.to_device_u8(binary)[None],
                                   torch.from_numpy(_int_corners(corners))[None].to(s._dev()))
    overall, sharp, contrast, compl, geom, size = (float(x) for x in scores.cpu().numpy()[0])
    issues, recs = [], []
    for bad, issue, rec in ((sharp < 40, "Image is blurry", "Hold camera steady or improve focus"),
                            (contrast < 40, "Low contrast", "Improve lighting conditions"),
                            (compl < 40, "Grid lines not fully visible", "Ensure entire puzzle is in frame"),
                            (geom < 50, "Grid is distorted", "Hold camera more perpendicular to puzzle"),
                            (size < 40, "Puzzle appears too small", "Move camera closer to puzzle")):
        if bad:
            issues.append(issue)
            recs.append(rec)
    return QualityScore(overall, sharp, contrast, compl, geom, size, issues, recs)


def get_user_feedback(quality: QualityScore) -> str:
    """cv/grid_quality.py:309-325."""
    if quality.is_good:
        return "Image quality is good. Processing..."
    if quality.is_acceptable:
        msg = "Image quality is acceptable but could be better."
        if quality.recommendations:
            msg += f" Tip: {quality.recommendations[0]}"
        return msg
    if quality.issues:
        return f"Please retake photo: {quality.issues[0]}. {quality.recommendations[0] if quality.recommendations else ''}"
    return "Image quality is too low. Please retake the photo."
