"""Tensor plumbing between the Python surface and the C-ABI: DLPack capsules in, torch for
device allocation / streams.  torch is plumbing here - no torch op is on the compute path."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

_PyCapsule_GetPointer = C.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = C.c_void_p
_PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("yolohot needs a CUDA device (B200, sm_100a); there is no CPU fallback")


class DL:
    """Holds a DLPack capsule of `obj` and exposes the DLManagedTensor* inside it.

    The capsule keeps its name "dltensor", so when this holder dies the capsule's own
    destructor calls the producer's deleter - the consumer side of the protocol without
    taking ownership (the kernels run on the producer's current stream)."""

    def __init__(self, obj):
        if isinstance(obj, torch.Tensor):
            self.capsule = torch.utils.dlpack.to_dlpack(obj)
        elif hasattr(obj, "__dlpack__"):
            self.capsule = obj.__dlpack__()
        else:  # e.g. a raw capsule from tf.experimental.dlpack.to_dlpack
            self.capsule = obj
        self.ptr = _PyCapsule_GetPointer(self.capsule, b"dltensor")
        self.keep = obj

    def __int__(self):
        return self.ptr


def dl(obj):
    return None if obj is None else DL(obj)


def ptr(h):
    return None if h is None else h.ptr


def as_device_f32(x, device=None):
    """Any array-like -> (contiguous float32 CUDA torch tensor, kind) where kind tells the
    caller what to hand back: 'torch', 'numpy' (host array-likes) or 'tf'."""
    if type(x) is torch.Tensor and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous():
        return x, "torch"                # the common case of a device pipeline: nothing to do
    require_cuda()
    if isinstance(x, torch.Tensor):
        kind = "torch"
        t = x
    elif type(x).__module__.startswith("tensorflow"):
        kind = "tf"
        import tensorflow as tf  # only reached when TF produced the tensor
        t = torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(tf.cast(x, tf.float32)))
    elif hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        kind = "torch"
        t = torch.utils.dlpack.from_dlpack(x)
    else:
        kind = "numpy"
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if not t.is_cuda:
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    if t.dtype in (torch.float16, torch.bfloat16):
        t = head_to_f32(t)               # half-precision head output: exact widening kernel (yh_head_to_f32)
    elif t.dtype != torch.float32:
        t = t.to(torch.float32)          # the reference casts to float32 (utils.py:20-22, 91-92, 169-170)
    return t.contiguous(), kind


def head_to_f32(t):
    """fp16 / bf16 CUDA tensor -> float32 tensor of the same shape."""
    from . import _lib
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    code = _lib.YH_DTYPE_F16 if t.dtype == torch.float16 else _lib.YH_DTYPE_BF16
    with on_device(t.device):
        _lib.check(_lib.lib().yh_head_to_f32(t.data_ptr(), code, t.numel(), out.data_ptr(), stream_ptr(t.device)),
                   "head_to_f32")
    return out


def as_grid(p, num_classes, num_boxes, grid=None):
    """Head adapter (train.py:208 `tf.reshape(predictions, [-1, 7, 7, 30])`): a flat (N, S*S*D)
    head output is viewed - zero-copy - as (N, S, S, D); 4-D input is checked and passed through.
    Returns (tensor, N, S)."""
    D = int(num_classes) + 5 * int(num_boxes)
    if p.dim() == 2:
        per = int(p.shape[1])
        S = int(grid) if grid is not None else int(round((per / D) ** 0.5))
        if S < 1 or S * S * D != per:
            raise ValueError(f"flat head output of {per} values per image is not S*S*{D} for "
                             f"{'grid=' + str(grid) if grid is not None else 'any square grid'}")
        p = p.view(int(p.shape[0]), S, S, D)
    if p.dim() != 4 or p.shape[1] != p.shape[2] or p.shape[3] != D or (grid is not None and int(p.shape[1]) != int(grid)):
        raise ValueError(f"expected (N, S, S, {D}) or flat (N, S*S*{D}) predictions, got {tuple(p.shape)}")
    return p, int(p.shape[0]), int(p.shape[1])


def give_back(t, kind):
    if kind == "numpy":
        return t.detach().cpu().numpy()
    if kind == "tf":
        import tensorflow as tf
        return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))
    return t


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device):
    """cudaStream_t of torch's current stream on `device` (the raw getter when this torch has it: a tenth of the cost of
    building a torch.cuda.Stream object - four of these per evaluator pass were a quarter of its host time)."""
    if _raw_stream is not None:
        idx = device.index if isinstance(device, torch.device) else device
        return C.c_void_p(_raw_stream(torch.cuda.current_device() if idx is None else idx))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class on_device:
    """`with torch.cuda.device(dev)` that does nothing when dev already is the current device (the usual case)."""
    __slots__ = ("dev", "ctx")

    def __init__(self, dev):
        self.dev, self.ctx = dev, None

    def __enter__(self):
        idx = self.dev.index if isinstance(self.dev, torch.device) else self.dev
        if idx is not None and idx != torch.cuda.current_device():
            self.ctx = torch.cuda.device(idx)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False
