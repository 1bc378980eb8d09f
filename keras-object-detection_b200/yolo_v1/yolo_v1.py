"""Module-name shim for the reference's single-file variant (yolo_v1/yolo_v1.py): the hot-path names it defines
(yolo_v1.py:39 intersection_over_union, :75 non_max_suppression, :112 decode_predictions, :177 change_tensor,
:200 mean_average_precision, :355 MeanAveragePrecision, :394 get_tagged_img, :613 YoloV1Loss) served by yolohot on
libyolohot.  The model, the data generator and the training script of that file are outside the path (SURVEY.md
section 8) and are not provided."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from yolohot import utils as _u  # noqa: E402
from yolohot.loss import YoloV1Loss  # noqa: F401,E402
from yolohot.utils import (MeanAveragePrecision, change_tensor, get_tagged_img, intersection_over_union,  # noqa: F401,E402
                           mean_average_precision, non_max_suppression)

__all__ = ["intersection_over_union", "non_max_suppression", "decode_predictions", "change_tensor",
           "mean_average_precision", "MeanAveragePrecision", "get_tagged_img", "YoloV1Loss"]


def decode_predictions(predictions, num_classes=20, num_boxes=2):
    """yolo_v1.py:112: same as utils.decode_predictions, with this file's default of 20 classes."""
    return _u.decode_predictions(predictions, num_classes, num_boxes)
