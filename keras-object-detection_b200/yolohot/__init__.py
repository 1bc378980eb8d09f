"""yolohot - B200-native (sm_100a) YOLOv1 decode / IoU / NMS / loss / mAP hot path behind the
Python call surface of myungsanglee/Keras-Object-Detection (yolo_v1/utils.py, loss.py,
metric.py).  Host code is Python; all arithmetic runs in hand-written CUDA kernels reached
through the C-ABI of libyolohot.so (include/yolohot.h).  No CPU fallback."""
from . import _lib
from ._lib import YoloHotError, launch_count

__version__ = "0.1.0"
__all__ = ["utils", "loss", "metric", "dataset", "dist", "YoloHotError", "launch_count"]
