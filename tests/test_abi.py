"""CPU tests of the boundary: libyolohot.so loads, exports every symbol include/yolohot.h
declares, validates arguments before touching the GPU, and never falls back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "yolohot.h")).read()
    return sorted(set(re.findall(r"YH_API[^;(]*?\b(yh_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from yolohot import _lib
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from yolohot import _lib
    L = _lib.lib()
    for s in header_symbols():
        assert hasattr(L, s), s
    assert L.yh_version() >= 1
    assert isinstance(L.yh_last_error(), bytes)


def test_argument_errors_need_no_gpu():
    from yolohot import _lib
    L = _lib.lib()
    assert L.yh_iou(None, None, -1, None, None) == _lib.YH_ERR_ARG
    assert b"n < 0" in L.yh_last_error()
    assert L.yh_decode_nms(None, 4, 7, 2, 20, 0.5, 0.4, None, None, None, None) == _lib.YH_ERR_ARG
    assert L.yh_decode_nms(None, 4, 17, 2, 20, 0.5, 0.4, None, None, None, None) == _lib.YH_ERR_UNSUPPORTED
    assert L.yh_nms(None, 1, 300, 0.5, 0.4, None, None, None, None) == _lib.YH_ERR_UNSUPPORTED
    assert L.yh_loss(None, None, 10, 2, 20, 5.0, 0.5, None, None, None) == _lib.YH_ERR_ARG
    assert L.yh_comm_init_all(0, None, None) == _lib.YH_ERR_ARG
    assert L.yh_map_allgather(None, None, None, None, 20, None, 0, None) == _lib.YH_ERR_ARG
    assert L.yh_comm_destroy(None) == _lib.YH_OK
    assert L.yh_comm_p2p(None) == 0 and L.yh_comm_barrier(None, None) == _lib.YH_ERR_ARG
    assert L.yh_map_exchange(0, 0, None, 20, 0, None, 0, None, None, 1, None) == _lib.YH_ERR_ARG
    assert L.yh_map_exchange(17, 0, None, 20, 0, None, 0, None, None, 1, None) == _lib.YH_ERR_ARG
    assert b"at most 16 peers" in L.yh_last_error()
    assert L.yh_map_exchange(2, 0, None, 20, 0, None, 0, None, None, 0, None) == _lib.YH_ERR_ARG          # epochs start at 1
    assert L.yh_map_reduce_exchanged(2, None, 20, 1024, 1, 0, None, None, None, None, 0, None) == _lib.YH_ERR_ARG
    assert L.yh_map_exchange_bytes(0, 20, 1024) == 0 and L.yh_map_exchange_bytes(8, 20, 1 << 16) > 8 * 2 * 8 * (1 << 16)
    assert L.yh_map_match(None, -1, None, None, 0, None, 20, 0.5, 0, None, None, None, 0, None) == _lib.YH_ERR_ARG
    assert L.yh_map_match(None, 0, None, None, 0, None, 5000, 0.5, 0, None, None, None, 0, None) == _lib.YH_ERR_ARG   # C limit
    assert L.yh_map_reduce(None, 0, None, 0, None, 20, None, None, None, 0, None) == _lib.YH_ERR_ARG
    assert L.yh_eval_update(None, None, None, None, 4, 300, 0, 20, 0.5, None, 0, None, 0, None, None, None, None) == _lib.YH_ERR_ARG
    assert L.yh_eval_update(None, None, None, None, 4, 49, 0, 20, 0.5, None, 0, None, 0, None, None, None, None) == _lib.YH_ERR_ARG
    # update_state in one launch: sizes, null pointers, the grids it hands to the three-launch path, alignment
    us = L.yh_eval_update_state
    assert us(None, None, -1, 7, 2, 20, 0.5, 0.4, 0, 0.5, None, 0, None, 0, None, None, None, 0, None) == _lib.YH_ERR_ARG
    assert us(None, None, 4, 7, 2, 20, 0.5, 0.4, 0, 0.5, None, 0, None, 0, None, None, None, 0, None) == _lib.YH_ERR_ARG
    assert b"null pointer" in L.yh_last_error()
    assert us(None, None, 4, 9, 2, 20, 0.5, 0.4, 0, 0.5, None, 0, None, 0, None, None, None, 0, None) == _lib.YH_ERR_UNSUPPORTED
    assert b"three-launch path" in L.yh_last_error()
    assert us(None, None, 4, 7, 2, 5000, 0.5, 0.4, 0, 0.5, None, 0, None, 0, None, None, None, 0, None) == _lib.YH_ERR_ARG   # C limit
    assert us(None, None, 0, 7, 2, 20, 0.5, 0.4, 0, 0.5, None, 0, None, 0, None, None, None, 0, None) == _lib.YH_OK       # nothing to do
    assert us(12, 16, 4, 7, 2, 20, 0.5, 0.4, 0, 0.5, None, 0, None, 0, 8, 8, 8, 0, None) == _lib.YH_ERR_ARG             # y_true not 8-byte aligned
    assert b"8-byte aligned" in L.yh_last_error()
    assert L.yh_workspace_bytes(_lib.YH_OP_MAP_REDUCE, 100000, 0, 0, 20) >= 2 * 8 * 100000
    assert L.yh_workspace_bytes(_lib.YH_OP_MAP_MATCH, 100000, 0, 0, 20) >= 12 * 100000
    assert L.yh_ipc_alloc(0, None, None) == _lib.YH_ERR_ARG and L.yh_ipc_open(None, None) == _lib.YH_ERR_ARG
    assert L.yh_ipc_close(None) == _lib.YH_OK and L.yh_ipc_free(None) == _lib.YH_OK
    assert L.yh_decode_nms_host_typed(None, 9, 4, 7, 2, 20, 0.5, 0.4, None, None, None, 0) == _lib.YH_ERR_ARG
    assert b"unknown dtype" in L.yh_last_error()
    assert L.yh_decode_nms_host_rows(None, 0, 4, 7, 2, 20, 0.5, 0.4, None, 0, None, None, 0) == _lib.YH_ERR_ARG
    assert L.yh_host_alloc(0, 0, None) == _lib.YH_ERR_ARG and L.yh_host_free(None) == _lib.YH_OK
    assert L.yh_encode_labels(None, None, -1, 7, 2, 20, None, None, None) == _lib.YH_ERR_ARG
    assert L.yh_head_to_f32(None, 7, 16, None, None) == _lib.YH_ERR_ARG
    assert L.yh_pixel_boxes(None, None, 4, 0, 448, 448, None, None) == _lib.YH_ERR_ARG
    assert L.yh_workspace_bytes(1, 1000, 7, 2, 20) == 0 and L.yh_workspace_bytes(3, 4096, 7, 2, 20) > 0
    with pytest.raises(ValueError):
        _lib.check(_lib.YH_ERR_ARG, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.YH_ERR_UNSUPPORTED, "x")
    with pytest.raises(_lib.YoloHotError):
        _lib.check(_lib.YH_ERR_CUDA, "x")


def test_cpu_tensor_is_rejected_not_computed():
    """kDLCPU input -> YH_ERR_ARG: there is no CPU fallback behind the DLPack front ends."""
    from yolohot import _lib
    from yolohot._tensor import DL
    L = _lib.lib()
    a = torch.zeros(8, 4)
    out = torch.zeros(8, 1)
    rc = L.yh_iou_dl(DL(a).ptr, DL(a).ptr, DL(out).ptr, None)
    assert rc == _lib.YH_ERR_ARG and b"no CPU fallback" in L.yh_last_error()
    p = torch.zeros(2, 7, 7, 30)
    rc = L.yh_decode_nms_dl(DL(p).ptr, 2, 20, 0.5, 0.4, DL(torch.zeros(2, 49, 6)).ptr,
                            DL(torch.zeros(2, dtype=torch.int32)).ptr, None, None)
    assert rc == _lib.YH_ERR_ARG


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_public_surface_fails_loudly_without_gpu():
    from yolohot import loss, utils
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        utils.decode_predictions(np.zeros((1, 7, 7, 30), np.float32), 20, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        utils.decode_nms(np.zeros((1, 7, 7, 30), np.float32), 20, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss.YoloV1Loss()(np.zeros((1, 7, 7, 30), np.float32), np.zeros((1, 7, 7, 30), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "keras-object-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU oracle", "").replace("the CPU oracle", "").lower() or f == "_lib.py", (dirpath, f)


def test_reference_module_names_resolve():
    """`from utils import ...`, `from loss import ...`, `from metric import ...` as the reference's scripts do."""
    import importlib.util
    d = os.path.join(ROOT, "keras-object-detection_b200", "yolo_v1")
    names = {"utils": ["intersection_over_union", "non_max_suppression", "decode_predictions", "mean_average_precision",
                       "MeanAveragePrecision", "MeanAveragePrecisionNumpy", "get_all_bboxes", "non_max_suppression_2",
                       "mean_average_precision_2", "decode_predictions_numpy", "non_max_suppression_numpy",
                       "intersection_over_union_numpy", "mean_average_precision_numpy"],
             "loss": ["YoloV1Loss"], "metric": ["MeanAveragePrecision", "MeanAveragePrecision2"],
             # the single-file variant of the reference (yolo_v1/yolo_v1.py:39,75,112,177,200,355,394,613)
             "yolo_v1": ["intersection_over_union", "non_max_suppression", "decode_predictions", "change_tensor",
                         "mean_average_precision", "MeanAveragePrecision", "get_tagged_img", "YoloV1Loss"]}
    mods = {}
    for mod, syms in names.items():
        spec = importlib.util.spec_from_file_location(f"_shim_{mod}", os.path.join(d, mod + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        for s in syms:
            assert hasattr(m, s), (mod, s)
        mods[mod] = m
    import inspect
    import numpy as np
    assert inspect.signature(mods["yolo_v1"].decode_predictions).parameters["num_classes"].default == 20   # yolo_v1.py:112
    assert inspect.signature(mods["utils"].decode_predictions).parameters["num_classes"].default is inspect.Parameter.empty
    assert np.array_equal(mods["yolo_v1"].change_tensor(np.zeros(4, np.int32), 2), [0, 0, 1, 0])            # yolo_v1.py:177
    from yolohot.loss import YoloV1Loss
    l = YoloV1Loss()
    assert (l.num_classes, l.num_boxes, l.lambda_coord, l.lambda_noobj, l.name) == (20, 2, 5, 0.5, "YoloV1Loss")
