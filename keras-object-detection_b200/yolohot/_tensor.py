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
    if t.dtype != torch.float32:
        t = t.to(torch.float32)          # the reference casts to float32 (utils.py:20-22, 91-92, 169-170)
    if not t.is_cuda:
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    return t.contiguous(), kind


def give_back(t, kind):
    if kind == "numpy":
        return t.detach().cpu().numpy()
    if kind == "tf":
        import tensorflow as tf
        return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))
    return t


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
