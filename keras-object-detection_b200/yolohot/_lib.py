"""ctypes binding of libyolohot.so (C-ABI declared in include/yolohot.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
error is raised to the caller.  Nothing here imports the CPU oracle."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YH_LIB_PATH") or os.path.join(_HERE, "libyolohot.so")   # YH_LIB_PATH: A/B builds

YH_OK, YH_ERR_ARG, YH_ERR_CUDA, YH_ERR_NCCL, YH_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
YH_IPC_HANDLE_BYTES = 64
YH_DTYPE_F32, YH_DTYPE_F16, YH_DTYPE_BF16 = 0, 1, 2
YH_MAP_TRUE_ROWS_BY_IMAGE = 1
YH_MAP_ERR_TIMEOUT, YH_MAP_ERR_OVERFLOW = 1, 2
YH_OP_DECODE_NMS, YH_OP_DECODE_NMS_HOST, YH_OP_LOSS, YH_OP_MAP_MATCH, YH_OP_MAP_REDUCE = 1, 2, 3, 4, 5

_lib = None

# every symbol include/yolohot.h declares (tests/test_abi.py checks the two stay in sync)
SYMBOLS = (
    "yh_version", "yh_last_error", "yh_device_info", "yh_launch_count",
    "yh_iou", "yh_decode", "yh_nms", "yh_decode_nms", "yh_decode_nms_ex", "yh_decode_nms_host", "yh_decode_nms_host_typed", "yh_decode_nms_host_rows", "yh_host_alloc", "yh_host_free", "yh_filter_rows", "yh_rows_append",
    "yh_loss", "yh_eval_update", "yh_eval_update_state", "yh_map_match", "yh_map_reduce",
    "yh_encode_labels", "yh_head_to_f32", "yh_decode_nms_typed", "yh_pixel_boxes",
    "yh_comm_init_all", "yh_comm_destroy", "yh_map_allgather", "yh_comm_p2p", "yh_comm_barrier",
    "yh_map_exchange_bytes", "yh_map_exchange", "yh_map_reduce_exchanged",
    "yh_ipc_alloc", "yh_ipc_open", "yh_ipc_close", "yh_ipc_free",
    "yh_workspace_bytes",
    "yh_iou_dl", "yh_decode_dl", "yh_nms_dl", "yh_decode_nms_dl", "yh_loss_dl",
)


class YoloHotError(RuntimeError):
    pass


def lib():
    """Load libyolohot.so once.  Raises if it was not built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise YoloHotError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  yolohot has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, f, i, i64 = C.c_void_p, C.c_float, C.c_int, C.c_int64
    L.yh_version.restype = i
    L.yh_last_error.restype = C.c_char_p
    L.yh_launch_count.restype = i64
    L.yh_device_info.argtypes = [C.POINTER(i)] * 3
    L.yh_iou.argtypes = [vp, vp, i64, vp, vp]
    L.yh_decode.argtypes = [vp, i64, i, i, i, vp, vp]
    L.yh_nms.argtypes = [vp, i64, i, f, f, vp, vp, vp, vp]
    L.yh_decode_nms.argtypes = [vp, i64, i, i, i, f, f, vp, vp, vp, vp]
    L.yh_decode_nms_ex.argtypes = [vp, i64, i, i, i, f, f, i, vp, vp, vp, vp]
    L.yh_decode_nms_host.argtypes = [vp, i64, i, i, i, f, f, vp, vp, vp, i]
    L.yh_decode_nms_host_typed.argtypes = [vp, i, i64, i, i, i, f, f, vp, vp, vp, i]
    L.yh_decode_nms_host_rows.argtypes = [vp, i, i64, i, i, i, f, f, vp, i64, vp, vp, i]
    L.yh_host_alloc.argtypes = [C.c_size_t, i, vp]
    L.yh_host_free.argtypes = [vp]
    L.yh_filter_rows.argtypes = [vp, i64, i, f, vp, vp, vp]
    L.yh_rows_append.argtypes = [vp, vp, i64, i, i64, vp, i64, vp, vp]
    L.yh_loss.argtypes = [vp, vp, i64, i, i, f, f, vp, vp, vp]
    L.yh_eval_update.argtypes = [vp, vp, vp, vp, i64, i, i64, i, f, vp, i64, vp, i64, vp, vp, vp, vp]
    L.yh_eval_update_state.argtypes = [vp, vp, i64, i, i, i, f, f, i64, f, vp, i64, vp, i64, vp, vp, vp, i, vp]
    L.yh_map_match.argtypes = [vp, i64, vp, vp, i64, vp, i, f, i, vp, vp, vp, C.c_size_t, vp]
    L.yh_map_reduce.argtypes = [vp, i64, vp, i64, vp, i, vp, vp, vp, C.c_size_t, vp]
    L.yh_map_exchange_bytes.argtypes = [i, i, i64]
    L.yh_map_exchange_bytes.restype = C.c_size_t
    L.yh_map_exchange.argtypes = [i, i, vp, i, i64, vp, i64, vp, vp, C.c_uint64, vp]
    L.yh_map_reduce_exchanged.argtypes = [i, vp, i, i64, C.c_uint64, i64, vp, vp, vp, vp, C.c_size_t, vp]
    L.yh_encode_labels.argtypes = [vp, vp, i64, i, i, i, vp, vp, vp]
    L.yh_head_to_f32.argtypes = [vp, i, i64, vp, vp]
    L.yh_decode_nms_typed.argtypes = [vp, i, i64, i, i, i, f, f, i, vp, vp, vp, vp]
    L.yh_pixel_boxes.argtypes = [vp, vp, i64, i, i, i, vp, vp]
    L.yh_comm_init_all.argtypes = [i, vp, vp]
    L.yh_comm_destroy.argtypes = [vp]
    L.yh_map_allgather.argtypes = [vp, vp, vp, vp, i, vp, i64, vp]
    L.yh_comm_p2p.argtypes = [vp]
    L.yh_comm_barrier.argtypes = [vp, vp]
    L.yh_ipc_alloc.argtypes = [C.c_size_t, vp, vp]
    L.yh_ipc_open.argtypes = [vp, vp]
    L.yh_ipc_close.argtypes = [vp]
    L.yh_ipc_free.argtypes = [vp]
    L.yh_workspace_bytes.argtypes = [i, i64, i, i, i]
    L.yh_workspace_bytes.restype = C.c_size_t
    L.yh_iou_dl.argtypes = [vp, vp, vp, vp]
    L.yh_decode_dl.argtypes = [vp, i, i, vp, vp]
    L.yh_nms_dl.argtypes = [vp, f, f, vp, vp, vp, vp]
    L.yh_decode_nms_dl.argtypes = [vp, i, i, f, f, vp, vp, vp, vp]
    L.yh_loss_dl.argtypes = [vp, vp, i, i, f, f, vp, vp, vp]
    for s in SYMBOLS:
        fn = getattr(L, s)
        if s not in ("yh_version", "yh_last_error", "yh_launch_count", "yh_workspace_bytes", "yh_map_exchange_bytes"):
            fn.restype = i
    _lib = L
    return L


def check(rc, what=""):
    """Turn a negative return code into the exception the reference surface would raise."""
    if rc == YH_OK:
        return
    msg = lib().yh_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg}" if what else msg
    if rc == YH_ERR_ARG:
        raise ValueError(text)
    if rc == YH_ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    raise YoloHotError(text)


def launch_count():
    return int(lib().yh_launch_count())
