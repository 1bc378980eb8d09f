"""ctypes binding of oracle/libyolo_oracle.so (C restatement; TEST INFRASTRUCTURE ONLY,
see yolo_oracle_c.c header).  Used by tests/ and by bench.py's cpu_baseline / reference arm."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libyolo_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "yolo_oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libyolo_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.yo_iou_many.argtypes = [fp, fp, C.c_int64, fp]
        L.yo_decode.argtypes = [fp, C.c_int64, C.c_int, C.c_int, C.c_int, fp]
        L.yo_decode_nms.argtypes = [fp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                    fp, ip, ip, C.c_int]
        L.yo_loss.argtypes = [fp, fp, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float,
                              C.POINTER(C.c_double), C.c_int]
        L.yo_map.argtypes = [fp, C.c_int64, fp, C.c_int64, C.c_int, C.c_float, fp]
        L.yo_map.restype = C.c_double
        L.yo_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def iou(a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 4)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 4)
    out = np.empty(a.shape[0], np.float32)
    lib().yo_iou_many(_fp(a), _fp(b), a.shape[0], _fp(out))
    return out


def decode(pred, num_classes, num_boxes=2):
    p = np.ascontiguousarray(pred, np.float32)
    n, S = p.shape[0], p.shape[1]
    out = np.empty((n, S * S, 6), np.float32)
    lib().yo_decode(_fp(p), n, S, num_boxes, num_classes, _fp(out))
    return out


def decode_nms(pred, num_classes, num_boxes=2, iou_threshold=0.5, conf_threshold=0.4,
               nthreads=0, want_boxes=True, want_idx=True):
    p = np.ascontiguousarray(pred, np.float32)
    n, S = p.shape[0], p.shape[1]
    M = S * S
    out = np.zeros((n, M, 6), np.float32) if want_boxes else None
    cnt = np.zeros((n,), np.int32)
    kidx = np.empty((n, M), np.int32) if want_idx else None
    L = lib()

    def run(lo, hi):
        if hi <= lo:
            return
        L.yo_decode_nms(_fp(p[lo:hi]), hi - lo, S, num_boxes, num_classes, iou_threshold, conf_threshold,
                        _fp(out[lo:hi]) if want_boxes else None, _ip(cnt[lo:hi]),
                        _ip(kidx[lo:hi]) if want_idx else None, 1)

    nthreads = max(1, int(nthreads) if nthreads else 1)
    if nthreads == 1 or n < 2 * nthreads:
        run(0, n)
    else:   # contiguous image shards, one per host thread (ctypes releases the GIL)
        from concurrent.futures import ThreadPoolExecutor
        edges = [n * i // nthreads for i in range(nthreads + 1)]
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda i: run(edges[i], edges[i + 1]), range(nthreads)))
    return out, cnt, kidx


def loss(y_true, y_pred, num_classes=20, num_boxes=2, lambda_coord=5.0, lambda_noobj=0.5, nthreads=0):
    t = np.ascontiguousarray(y_true, np.float32)
    p = np.ascontiguousarray(y_pred, np.float32)
    D = num_classes + 5 * num_boxes
    terms = (C.c_double * 6)()
    lib().yo_loss(_fp(t), _fp(p), t.size // D, num_boxes, num_classes, lambda_coord, lambda_noobj,
                  terms, nthreads)
    return np.array(list(terms))


def mean_average_precision(true_rows, pred_rows, num_classes, iou_threshold=0.5):
    t = np.ascontiguousarray(true_rows, np.float32).reshape(-1, 7)
    p = np.ascontiguousarray(pred_rows, np.float32).reshape(-1, 7)
    ap = np.zeros(num_classes, np.float32)
    m = lib().yo_map(_fp(t), t.shape[0], _fp(p), p.shape[0], num_classes, iou_threshold, _fp(ap))
    return m, ap


def num_threads():
    """Host threads the baseline may use: the cores this process is allowed on."""
    return len(os.sched_getaffinity(0))
