"""Drop-in mirror of the reference's yolo_v1/loss.py: YoloV1Loss(num_classes=20, num_boxes=2),
callable as loss(y_true, y_pred) -> scalar (loss.py:100-215).  Forward and the hand-written
backward run in one CUDA kernel (yh_loss); autograd sees a single node."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._tensor import DL, as_device_f32, dl, give_back, ptr, stream_ptr

__all__ = ["YoloV1Loss", "yolo_v1_loss_terms"]


def _run(y_true, y_pred, C, B, lc, ln, want_grad):
    D = C + 5 * B
    if y_true.shape != y_pred.shape or y_pred.shape[-1] != D:
        raise ValueError(f"YoloV1Loss: expected matching (..., {D}) tensors, got {tuple(y_true.shape)} and {tuple(y_pred.shape)}")
    dev = y_pred.device
    terms = torch.empty((6,), dtype=torch.float32, device=dev)
    grad = torch.empty_like(y_pred) if want_grad else None
    with torch.cuda.device(dev):
        ht, hp, ho, hg = DL(y_true), DL(y_pred), DL(terms), dl(grad)
        _lib.check(_lib.lib().yh_loss_dl(ht.ptr, hp.ptr, B, C, float(lc), float(ln), ho.ptr, ptr(hg), stream_ptr(dev)),
                   "YoloV1Loss")
    return terms, grad


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_true, y_pred, C, B, lc, ln):
        need = y_pred.requires_grad
        terms, grad = _run(y_true, y_pred.detach(), C, B, lc, ln, need)
        if need:
            ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(terms)
        return terms[5].clone(), terms

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        (grad,) = ctx.saved_tensors
        return None, grad * g_total, None, None, None, None


def yolo_v1_loss_terms(y_true, y_pred, num_classes=20, num_boxes=2, lambda_coord=5.0, lambda_noobj=0.5, grad=False):
    """[xy, wh, obj, noobj, cls, total] (loss.py:172-213) and optionally d(total)/d(y_pred)."""
    t, kind = as_device_f32(y_true)
    p, _ = as_device_f32(y_pred, t.device)
    if p.dim() == 2 and t.dim() == 4 and p.numel() == t.numel():
        p = p.view(t.shape)
    terms, g = _run(t, p, int(num_classes), int(num_boxes), lambda_coord, lambda_noobj, grad)
    if grad:
        return give_back(terms, kind), give_back(g, kind)
    return give_back(terms, kind)


class YoloV1Loss:
    """loss.py:100-215.  Same constructor, attributes (lambda_noobj, lambda_coord, num_classes,
    num_boxes, batch_size, name) and call contract as the Keras loss; the result is the batch
    SUM (loss.py:172-213), a 0-d tensor that back-propagates into y_pred."""

    def __init__(self, num_classes=20, num_boxes=2):
        self.name = "YoloV1Loss"
        self.num_classes = num_classes
        self.num_boxes = num_boxes
        self.lambda_noobj = 0.5          # loss.py:115
        self.lambda_coord = 5            # loss.py:116
        self.batch_size = 0              # loss.py:118
        self.last_terms = None           # [xy, wh, obj, noobj, cls, total] of the last call

    def call(self, y_true, y_pred):
        t, kind = as_device_f32(y_true)
        if isinstance(y_pred, torch.Tensor) and y_pred.is_cuda and y_pred.dtype == torch.float32:
            p = y_pred.contiguous()
        else:
            p, _ = as_device_f32(y_pred, t.device)
        if p.dim() == 2 and t.dim() == 4 and p.numel() == t.numel():
            p = p.view(t.shape)              # flat Dense head output (model.py:107; loss.py:122); autograd keeps the view
        self.batch_size = int(t.shape[0])                                   # loss.py:123
        total, terms = _LossFn.apply(t, p, int(self.num_classes), int(self.num_boxes),
                                     float(self.lambda_coord), float(self.lambda_noobj))
        self.last_terms = terms
        if kind == "numpy":
            return np.float32(total.item())
        return give_back(total, kind) if kind == "tf" else total

    __call__ = call
