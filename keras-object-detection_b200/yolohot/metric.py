"""Drop-in mirror of the reference's (stale) yolo_v1/metric.py: the no-argument evaluators
MeanAveragePrecision() and MeanAveragePrecision2() (metric.py:11-56, 59-99; VOC C=20, B=2).
Difference from utils.MeanAveragePrecision that is kept: ground-truth rows are only
thresholded at conf > 0.4, not NMS'd (metric.py:35-37, 81)."""
from __future__ import annotations

from . import utils as _u

__all__ = ["MeanAveragePrecision", "MeanAveragePrecision2"]


class MeanAveragePrecision2(_u.MeanAveragePrecision):
    _nms_true = False

    def __init__(self):
        super().__init__(num_classes=20, num_boxes=2)

    @property
    def all_true_bboxes_variable(self):      # metric.py:61 spelling
        return self.all_true_boxes_variable

    @property
    def all_pred_bboxes_variable(self):      # metric.py:64 spelling
        return self.all_pred_boxes_variable


class MeanAveragePrecision(MeanAveragePrecision2):
    """metric.py:11-56: batches its appends and keeps a `count`; same results."""

    def __init__(self):
        super().__init__()
        self.count = 0

    def reset_states(self):
        super().reset_states()
        self.count = 0

    def update_state(self, y_true, y_pred):
        super().update_state(y_true, y_pred)
        self.count += 1
