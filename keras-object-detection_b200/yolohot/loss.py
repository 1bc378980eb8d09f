"""Drop-in mirror of the reference's yolo_v1/loss.py: YoloV1Loss(num_classes=20, num_boxes=2),
callable as loss(y_true, y_pred) -> scalar (loss.py:100-215).  Forward and the hand-written
backward run in one CUDA kernel (yh_loss); the framework's autodiff sees a single node:

  torch        torch.autograd.Function (_LossFn): forward stores d(total)/d(y_pred), backward scales it.
  TF / Keras   when `tensorflow` is importable the class IS a keras.losses.Loss (loss.py:100, compiled into the
               model at yolo_v1.py:810,829) and `call` wraps the same kernel in tf.custom_gradient: the forward
               hands the TF tensors over zero-copy through DLPack inside a tf.py_function (so it also runs
               under Keras' graph mode), returns terms[5] and keeps the gradient tensor; the backward is
               `upstream * saved gradient`.  (SURVEY.md 8b "Differentiability".)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._tensor import DL, as_device_f32, dl, give_back, on_device, ptr, stream_ptr

try:                                          # the reference's host framework, when it is there
    import tensorflow as _tf
    from tensorflow import keras as _keras
    _LossBase = _keras.losses.Loss
except Exception:                             # not installed: torch / NumPy callers only
    _tf = None
    _LossBase = object

__all__ = ["YoloV1Loss", "yolo_v1_loss_terms"]


def _is_tf(x):
    return type(x).__module__.split(".")[0] in ("tensorflow", "keras", "tf_keras")


def _run(y_true, y_pred, C, B, lc, ln, want_grad):
    D = C + 5 * B
    if y_true.shape != y_pred.shape or y_pred.shape[-1] != D:
        raise ValueError(f"YoloV1Loss: expected matching (..., {D}) tensors, got {tuple(y_true.shape)} and {tuple(y_pred.shape)}")
    dev = y_pred.device
    terms = torch.empty((6,), dtype=torch.float32, device=dev)
    grad = torch.empty_like(y_pred) if want_grad else None
    with on_device(dev):
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


def _tf_forward_backward(y_true, y_pred, C, B, lc, ln):
    """Eager TF tensors -> (total, d total / d y_pred) as TF tensors, through DLPack (zero-copy both ways)."""
    t, _ = as_device_f32(y_true)
    p, _ = as_device_f32(y_pred, t.device)
    shape = tuple(p.shape)
    if p.dim() == 2 and t.dim() == 4 and p.numel() == t.numel():
        p = p.view(t.shape)                                                   # loss.py:122, flat head output
    # TF computes on its own CUDA stream: its producers must be done before this kernel reads, and this
    # kernel before TF reads the results
    torch.cuda.synchronize(t.device)
    terms, grad = _run(t, p, C, B, lc, ln, True)
    torch.cuda.synchronize(t.device)
    return give_back(terms[5].clone(), "tf"), give_back(grad.view(shape), "tf")


class YoloV1Loss(_LossBase):
    """loss.py:100-215.  Same constructor, attributes (lambda_noobj, lambda_coord, num_classes,
    num_boxes, batch_size, name) and call contract as the Keras loss; the result is the batch
    SUM (loss.py:172-213), a scalar that back-propagates into y_pred in the caller's framework.
    With TensorFlow installed this is a keras.losses.Loss subclass, usable in model.compile(loss=...)."""

    def __init__(self, num_classes=20, num_boxes=2):
        if _LossBase is not object:
            super().__init__(name="YoloV1Loss")                                # loss.py:111
        else:
            self.name = "YoloV1Loss"
        self.num_classes = num_classes
        self.num_boxes = num_boxes
        self.lambda_noobj = 0.5          # loss.py:115
        self.lambda_coord = 5            # loss.py:116
        self.batch_size = 0              # loss.py:118
        self.last_terms = None           # [xy, wh, obj, noobj, cls, total] of the last call (torch / NumPy callers)

    # ---- TensorFlow / Keras callers
    def _call_tf(self, y_true, y_pred):
        C, B = int(self.num_classes), int(self.num_boxes)
        lc, ln = float(self.lambda_coord), float(self.lambda_noobj)
        y_true = _tf.cast(y_true, _tf.float32)
        y_pred = _tf.cast(y_pred, _tf.float32)
        self.batch_size = y_true.shape[0]                                      # loss.py:123

        @_tf.custom_gradient
        def op(yp):
            total, grad = _tf.py_function(lambda t, p: _tf_forward_backward(t, p, C, B, lc, ln), [y_true, yp],
                                          [_tf.float32, _tf.float32])
            total.set_shape([])
            grad.set_shape(yp.shape)

            def backward(upstream):                                            # loss.py's autodiff backward, precomputed
                return upstream * grad
            return total, backward
        return op(y_pred)

    def call(self, y_true, y_pred):
        if _tf is not None and (_is_tf(y_pred) or _is_tf(y_true)):
            return self._call_tf(y_true, y_pred)
        if _is_tf(y_pred) or _is_tf(y_true):
            raise RuntimeError("YoloV1Loss: got TensorFlow tensors but `import tensorflow` failed in this process")
        t, kind = as_device_f32(y_true)
        if isinstance(y_pred, torch.Tensor) and y_pred.is_cuda and y_pred.dtype == torch.float32:
            p = y_pred.contiguous()
        elif isinstance(y_pred, torch.Tensor) and y_pred.requires_grad:
            # half-precision (or host) head output that trains: widen with an op autograd sees, so the
            # gradient flows back through the cast into the head's own dtype / device
            p = y_pred.to(device=t.device, dtype=torch.float32).contiguous()
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
        return total

    def __call__(self, y_true, y_pred, sample_weight=None):
        """Keras' Loss.__call__ (conversion to TF tensors, reduction - the identity on loss.py:215's scalar) for TF
        callers; torch / NumPy callers go straight to call()."""
        if _LossBase is not object and (_is_tf(y_pred) or _is_tf(y_true)):
            return super().__call__(y_true, y_pred, sample_weight)
        return self.call(y_true, y_pred)
