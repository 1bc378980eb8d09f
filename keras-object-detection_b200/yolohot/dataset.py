"""Label side of the reference's yolo_v1/dataset.py on libyolohot:
YOLO-txt boxes -> (S, S, C+5B) label grids, the tensor the loss and the evaluator consume.

  YoloV1Generator._get_labels(boxes)        dataset.py:88-112   -> get_labels / encode_labels (batched, device)
  YoloV1Generator._get_boxes(label_path)    dataset.py:114-123  -> get_boxes (host text parsing, same float64 rows)

Images, augmentation and the Keras Sequence plumbing stay with the caller (out of scope)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._tensor import on_device, require_cuda, stream_ptr

__all__ = ["get_boxes", "get_labels", "encode_labels", "YoloV1Labels"]


def get_boxes(label_path):
    """dataset.py:114-123: 'class cx cy w h' lines -> (n, 5) float64 rows [cx, cy, w, h, class]."""
    rows = []
    with open(label_path, "r") as f:
        for annot in f.read().splitlines():
            class_id, cx, cy, w, h = map(float, annot.split(" "))
            rows.append([cx, cy, w, h, class_id])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 5)


def encode_labels(boxes, offsets=None, grid=7, num_classes=20, num_boxes=2, device=None, out=None, check=True):
    """Batched dataset.py:88-112.

    boxes: a sequence of per-image box lists (each (k_i, 5) [cx, cy, w, h, class]; what
    albumentations' transformed['bboxes'] holds), or one (total, 5) array/tensor together with
    `offsets` (N+1) so that image i owns rows offsets[i]:offsets[i+1].  Returns the (N, S, S,
    C+5B) float32 CUDA tensor.  With check=True (one host sync) a box the reference would fail
    on (IndexError: cell or class index out of range) raises IndexError here as well."""
    require_cuda()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if offsets is None:
        lists = [np.asarray(b, dtype=np.float64).reshape(-1, 5) for b in boxes]
        counts = np.array([len(b) for b in lists], dtype=np.int64)
        offs_h = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum(counts, out=offs_h[1:])
        flat = np.concatenate(lists, axis=0) if lists else np.zeros((0, 5))
        b_d = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
        o_d = torch.from_numpy(offs_h).to(dev)
    else:
        b_d = (boxes if isinstance(boxes, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(boxes))))
        b_d = b_d.to(device=dev, dtype=torch.float64).reshape(-1, 5).contiguous()     # float32 -> float64 is exact
        o_d = (offsets if isinstance(offsets, torch.Tensor) else torch.from_numpy(np.asarray(offsets)))
        o_d = o_d.to(device=dev, dtype=torch.int64).contiguous()
    n = int(o_d.numel()) - 1
    if n < 0:
        raise ValueError("encode_labels: offsets must hold N+1 entries")
    S, C, B = int(grid), int(num_classes), int(num_boxes)
    if out is None:
        out = torch.empty((n, S, S, C + 5 * B), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n, S, S, C + 5 * B) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("encode_labels: `out` must be a contiguous float32 (N, S, S, C+5B) tensor")
    bad = torch.zeros((1,), dtype=torch.int32, device=dev)
    with on_device(dev):
        _lib.check(_lib.lib().yh_encode_labels(b_d.data_ptr(), o_d.data_ptr(), n, S, B, C, out.data_ptr(),
                                               bad.data_ptr(), stream_ptr(dev)), "encode_labels")
    if check:
        k = int(bad.item())
        if k:
            raise IndexError(f"encode_labels: {k} box(es) fall outside the {S}x{S} grid / the {C + 5 * B} channels "
                             "(dataset.py:107 raises IndexError there)")
    return out


def get_labels(boxes, grid=7, num_classes=20, num_boxes=2, device=None):
    """dataset.py:88-112 for one image -> (S, S, C+5B) float32 CUDA tensor."""
    return encode_labels([boxes], None, grid, num_classes, num_boxes, device)[0]


class YoloV1Labels:
    """The label-producing half of YoloV1Generator (dataset.py:19-32): same attribute names and the
    `_get_labels` / `_get_boxes` methods, so code written against the generator keeps working."""

    def __init__(self, num_classes, num_boxes, grid=7):
        self.grid = grid
        self.num_boxes = num_boxes
        self.num_classes = num_classes
        self.output_shape = (grid, grid, num_classes + (num_boxes * 5))

    def _get_labels(self, boxes):
        return get_labels(boxes, self.grid, self.num_classes, self.num_boxes)

    def _get_boxes(self, label_path):
        return get_boxes(label_path)

    def batch(self, box_lists):
        """dataset.py:72-86 label half of `_get_data`: list of per-image boxes -> (N, S, S, D)."""
        return encode_labels(box_lists, None, self.grid, self.num_classes, self.num_boxes)
