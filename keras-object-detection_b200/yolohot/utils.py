"""Drop-in mirror of the reference's yolo_v1/utils.py hot-path surface on libyolohot.

Same names, positional order and defaults as the reference (file:line = reference repo):
  intersection_over_union(boxes1, boxes2)                               utils.py:9-43
  non_max_suppression(boxes, iou_threshold=0.5, conf_threshold=0.4)     utils.py:79-114
  decode_predictions(predictions, num_classes, num_boxes=2)             utils.py:152-218
  mean_average_precision(true_boxes, pred_boxes, num_classes, iou_threshold=0.5)   utils.py:303-456
  MeanAveragePrecision(num_classes, num_boxes=2)                        utils.py:459-496
  *_numpy / MeanAveragePrecisionNumpy twins                             utils.py:46-76, 117-149, 221-277, 499-620
plus the names the stale metric.py imports (get_all_bboxes, non_max_suppression_2,
mean_average_precision_2; bodies in tmp.py:96-157, 440-595) and one additive call,
decode_nms(), the batched fused path.

Inputs may be torch CUDA tensors (zero-copy via DLPack), anything exposing __dlpack__, TF
tensors, or host array-likes (copied to the current CUDA device).  Results come back in
the producer's kind.  There is no CPU implementation behind any of these names."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._tensor import DL, as_device_f32, as_grid, dl, give_back, on_device, ptr, require_cuda, stream_ptr

__all__ = [
    "intersection_over_union", "intersection_over_union_numpy",
    "non_max_suppression", "non_max_suppression_numpy", "non_max_suppression_2",
    "decode_predictions", "decode_predictions_numpy", "get_all_bboxes", "decode_nms",
    "mean_average_precision", "mean_average_precision_numpy", "mean_average_precision_2",
    "MeanAveragePrecision", "MeanAveragePrecisionNumpy",
    "pixel_boxes", "get_tagged_img", "get_grid_tagged_img",
]


# --------------------------------------------------------------------------- IoU
def intersection_over_union(boxes1, boxes2):
    """utils.py:9-43: element-wise IoU of [cx, cy, w, h] boxes, (...,4),(...,4) -> (...,1)."""
    a, kind = as_device_f32(boxes1)
    b, _ = as_device_f32(boxes2, a.device)
    if a.shape != b.shape:   # the reference broadcasts through TF; expand explicitly
        a, b = torch.broadcast_tensors(a, b)
        a, b = a.contiguous(), b.contiguous()
    if a.shape[-1] != 4:
        raise ValueError(f"intersection_over_union: last dimension must be 4, got {tuple(a.shape)}")
    out = torch.empty(a.shape[:-1] + (1,), dtype=torch.float32, device=a.device)
    with on_device(a.device):
        ha, hb, ho = DL(a), DL(b), DL(out)
        _lib.check(_lib.lib().yh_iou_dl(ha.ptr, hb.ptr, ho.ptr, stream_ptr(a.device)), "intersection_over_union")
    return give_back(out, kind)


def intersection_over_union_numpy(boxes1, boxes2):
    """utils.py:46-76 twin: NumPy in, NumPy out, same device path."""
    return intersection_over_union(np.asarray(boxes1, np.float32), np.asarray(boxes2, np.float32))


# ------------------------------------------------------------------------ decode
def decode_predictions(predictions, num_classes, num_boxes=2, grid=None):
    """utils.py:152-218: (N,S,S,C+5B) -> (N,S*S,6) rows [class_idx, confidence, cx, cy, w, h].
    S is taken from the tensor (the reference hard-codes 7, utils.py:184,200-216).  The flat
    (N, S*S*(C+5B)) head output (model.py:107, reshaped at train.py:208) is accepted as it is."""
    p, kind = as_device_f32(predictions)
    p, n, S = as_grid(p, num_classes, num_boxes, grid)
    out = torch.empty((n, S * S, 6), dtype=torch.float32, device=p.device)
    with on_device(p.device):
        hp, ho = DL(p), DL(out)
        _lib.check(_lib.lib().yh_decode_dl(hp.ptr, int(num_boxes), int(num_classes), ho.ptr, stream_ptr(p.device)),
                   "decode_predictions")
    return give_back(out, kind)


def decode_predictions_numpy(predictions, num_classes, num_boxes=2):
    """utils.py:221-277 twin (the reference's own twin is batch-1 only and float64-leaky)."""
    return decode_predictions(np.asarray(predictions, np.float32), num_classes, num_boxes)


def get_all_bboxes(out, grid=7, num_classes=20):
    """tmp.py:96-157 (imported by the stale metric.py:7): decode with B=2."""
    res = decode_predictions(out, num_classes, 2)
    if int(res.shape[1]) != grid * grid:
        raise ValueError(f"get_all_bboxes: tensor grid {int(round(res.shape[1] ** 0.5))} != grid={grid}")
    return res


# --------------------------------------------------------------------------- NMS
def _nms_device(b, iou_threshold, conf_threshold, want_idx=False):
    """b: (N,M,6) CUDA float32 -> padded (N,M,6), count (N,) int32, optional keep_idx (N,M)."""
    n, M = int(b.shape[0]), int(b.shape[1])
    out = torch.empty((n, M, 6), dtype=torch.float32, device=b.device)      # rows past count[i] are never handed out
    cnt = torch.empty((n,), dtype=torch.int32, device=b.device)
    kidx = torch.full((n, M), -1, dtype=torch.int32, device=b.device) if want_idx else None
    with on_device(b.device):
        hb, ho, hc, hk = DL(b), DL(out), DL(cnt), dl(kidx)
        _lib.check(_lib.lib().yh_nms_dl(hb.ptr, float(iou_threshold), float(conf_threshold), ho.ptr, hc.ptr, ptr(hk),
                                        stream_ptr(b.device)), "non_max_suppression")
    return out, cnt, kidx


def non_max_suppression(boxes, iou_threshold=0.5, conf_threshold=0.4):
    """utils.py:79-114: one image's (M,6) rows -> the (K,6) kept rows in pick order.
    A batched (N,M,6) input returns a list of N such tensors."""
    b, kind = as_device_f32(boxes)
    if b.dim() == 2 and b.shape[-1] == 6:
        out, cnt, _ = _nms_device(b.unsqueeze(0), iou_threshold, conf_threshold)
        return give_back(out[0, :int(cnt.item())], kind)
    if b.dim() == 3 and b.shape[-1] == 6:
        out, cnt, _ = _nms_device(b, iou_threshold, conf_threshold)
        counts = cnt.tolist()
        return [give_back(out[i, :k], kind) for i, k in enumerate(counts)]
    raise ValueError(f"non_max_suppression: expected (M, 6) or (N, M, 6) boxes, got {tuple(b.shape)}")


def non_max_suppression_numpy(boxes, iou_threshold=0.5, conf_threshold=0.4):
    """utils.py:117-149 twin."""
    return non_max_suppression(np.asarray(boxes, np.float32), iou_threshold, conf_threshold)


def non_max_suppression_2(boxes, iou_threshold=0.5, conf_threshold=0.4):
    """Name imported by the stale metric.py:7,77; same contract as non_max_suppression."""
    return non_max_suppression(boxes, iou_threshold, conf_threshold)


def decode_nms(predictions, num_classes, num_boxes=2, iou_threshold=0.5, conf_threshold=0.4,
               return_index=False, out=None, grid=None, score_mode="conf", compact=False):
    """Fused, batched loop body of utils.py:470-480 (decode + per-image NMS).

    Returns (boxes (N,S*S,6), count (N,) int32[, keep_idx (N,S*S) int32]): the first count[i]
    rows of image i are its kept rows in pick order; rows past that are unspecified unless
    `out` buffers were zeroed by the caller.  Host array-likes go through the pipelined
    host entry point (yh_decode_nms_host) and come back as NumPy (rows past count[i] zero).
    compact=True (host inputs only): only the kept rows come back over PCIe - returns
    (rows (K_total, 7) [img, cls, conf, cx, cy, w, h] in image order, count (N,)) (yh_decode_nms_host_rows).

    score_mode="conf" is the reference (a cell's score is its best box confidence).
    score_mode="conf_x_prob" is an extension the reference does not have: confidence x winning
    class probability decides the threshold, the order and the reported confidence."""
    L = _lib.lib()
    if score_mode not in ("conf", "conf_x_prob"):
        raise ValueError(f"decode_nms: score_mode must be 'conf' or 'conf_x_prob', got {score_mode!r}")
    if score_mode != "conf" and (isinstance(predictions, np.ndarray) or not (isinstance(predictions, torch.Tensor) or hasattr(predictions, "__dlpack__"))):
        predictions = torch.from_numpy(np.ascontiguousarray(np.asarray(predictions, dtype=np.float32))).cuda()
    host_half = (score_mode == "conf" and isinstance(predictions, torch.Tensor) and not predictions.is_cuda
                 and predictions.dtype in (torch.float16, torch.bfloat16))
    if host_half or not isinstance(predictions, torch.Tensor) and not hasattr(predictions, "__dlpack__") or isinstance(predictions, np.ndarray):
        require_cuda()
        # element type on the wire: float16 NumPy arrays and float16 / bfloat16 CPU tensors stay half (yh_decode_nms_host_typed)
        if host_half:
            p = predictions.contiguous()
            code = _lib.YH_DTYPE_F16 if p.dtype == torch.float16 else _lib.YH_DTYPE_BF16
        elif isinstance(predictions, np.ndarray) and predictions.dtype == np.float16:
            p, code = np.ascontiguousarray(predictions), _lib.YH_DTYPE_F16
        else:
            p, code = np.ascontiguousarray(np.asarray(predictions, dtype=np.float32)), _lib.YH_DTYPE_F32
        if p.ndim == 2:                                   # flat head output (train.py:208)
            D_ = num_classes + 5 * num_boxes
            S_ = int(grid) if grid is not None else int(round((p.shape[1] / D_) ** 0.5))
            if S_ >= 1 and S_ * S_ * D_ == p.shape[1]:
                p = p.reshape(p.shape[0], S_, S_, D_)
        if p.ndim != 4 or p.shape[1] != p.shape[2] or p.shape[3] != num_classes + 5 * num_boxes:
            raise ValueError(f"expected (N, S, S, {num_classes + 5 * num_boxes}) predictions, got {tuple(p.shape)}")
        n, S = p.shape[0], p.shape[1]
        if compact:
            if return_index:
                raise ValueError("decode_nms: compact=True returns rows, not indices")
            src = p.data_ptr() if host_half else p.ctypes.data
            cnt = np.zeros((n,), np.int32)
            cap = max(1024, 8 * n)                           # a guess; the call says how many rows it needs
            while True:
                rows = np.empty((cap, 7), np.float32)
                total = C.c_int64(0)
                rc = L.yh_decode_nms_host_rows(src, code, n, S, int(num_boxes), int(num_classes), float(iou_threshold),
                                               float(conf_threshold), rows.ctypes.data, cap, cnt.ctypes.data, C.byref(total),
                                               torch.cuda.current_device())
                if rc == _lib.YH_ERR_ARG and total.value > cap:
                    cap = int(total.value)
                    continue
                _lib.check(rc, "decode_nms")
                return rows[:total.value], cnt
        boxes = np.zeros((n, S * S, 6), np.float32)
        cnt = np.zeros((n,), np.int32)
        kidx = np.full((n, S * S), -1, np.int32) if return_index else None
        src = p.data_ptr() if host_half else p.ctypes.data
        _lib.check(L.yh_decode_nms_host_typed(src, code, n, S, int(num_boxes), int(num_classes), float(iou_threshold),
                                              float(conf_threshold), boxes.ctypes.data, cnt.ctypes.data,
                                              kidx.ctypes.data if return_index else None, torch.cuda.current_device()),
                   "decode_nms")
        return (boxes, cnt, kidx) if return_index else (boxes, cnt)
    if compact:
        raise ValueError("decode_nms: compact=True is for host inputs; on the device use MeanAveragePrecision / yh_rows_append")
    half = isinstance(predictions, torch.Tensor) and predictions.is_cuda and predictions.dtype in (torch.float16, torch.bfloat16)
    if half:                        # half-precision head: widened inside the kernel (yh_decode_nms_typed), half the HBM bytes
        p, kind = predictions.contiguous(), "torch"
        if p.data_ptr() % 4:        # the kernels read element pairs: a 2-byte aligned view is re-based
            p = p.clone()
    else:
        p, kind = as_device_f32(predictions)
        if p.data_ptr() % 8:        # the kernels read float32 pairs: a 4-byte aligned view (e.g. an odd slice) is re-based
            p = p.clone()
    p, n, S = as_grid(p, num_classes, num_boxes, grid)
    if out is not None:
        boxes, cnt = out[0], out[1]
        kidx = out[2] if (return_index and len(out) > 2) else None
        if return_index and kidx is None:
            raise ValueError("decode_nms: return_index=True needs out=(boxes, count, keep_idx)")
        for name, t, shape, dt in (("boxes", boxes, (n, S * S, 6), torch.float32), ("count", cnt, (n,), torch.int32),
                                   ("keep_idx", kidx, (n, S * S), torch.int32)):
            if t is None:
                continue                # raw pointers go to the kernels below: every property they rely on is checked here
            if not (isinstance(t, torch.Tensor) and t.device == p.device and t.dtype == dt and tuple(t.shape) == shape
                    and t.is_contiguous()):
                raise ValueError(f"decode_nms: out {name} must be a contiguous {dt} tensor of shape {shape} on {p.device}, got "
                                 f"{getattr(t, 'dtype', type(t))} {tuple(getattr(t, 'shape', ()))} on {getattr(t, 'device', None)}")
    else:
        boxes = torch.empty((n, S * S, 6), dtype=torch.float32, device=p.device)
        cnt = torch.empty((n,), dtype=torch.int32, device=p.device)
        kidx = torch.empty((n, S * S), dtype=torch.int32, device=p.device) if return_index else None
    with on_device(p.device):
        if half:
            code = _lib.YH_DTYPE_F16 if p.dtype == torch.float16 else _lib.YH_DTYPE_BF16
            _lib.check(L.yh_decode_nms_typed(p.data_ptr(), code, n, S, int(num_boxes), int(num_classes), float(iou_threshold),
                                             float(conf_threshold), 0 if score_mode == "conf" else 1, boxes.data_ptr(),
                                             cnt.data_ptr(), kidx.data_ptr() if kidx is not None else None,
                                             stream_ptr(p.device)), "decode_nms")
        elif score_mode == "conf":
            # raw pointers: everything the DLPack front end (yh_decode_nms_dl) would check has been checked above -
            # as_device_f32 / as_grid for the input, the block above (or the allocation) for the outputs - and three
            # capsules per call were a fifth of an evaluator pass's host time
            _lib.check(L.yh_decode_nms(p.data_ptr(), n, S, int(num_boxes), int(num_classes), float(iou_threshold),
                                       float(conf_threshold), boxes.data_ptr(), cnt.data_ptr(),
                                       kidx.data_ptr() if kidx is not None else None, stream_ptr(p.device)), "decode_nms")
        else:
            _lib.check(L.yh_decode_nms_ex(p.data_ptr(), n, S, int(num_boxes), int(num_classes), float(iou_threshold),
                                          float(conf_threshold), 1, boxes.data_ptr(), cnt.data_ptr(),
                                          kidx.data_ptr() if kidx is not None else None, stream_ptr(p.device)), "decode_nms")
    if return_index:
        return give_back(boxes, kind), give_back(cnt, kind), give_back(kidx, kind)
    return give_back(boxes, kind), give_back(cnt, kind)


# --------------------------------------------------------------------------- mAP
def _workspace(dev, nbytes, cache=None, key="ws"):
    """256-byte aligned device scratch (torch's allocator aligns to 512 bytes); cached in `cache` when given."""
    if cache is not None:
        t = cache.get(key)
        if t is not None and t.numel() >= nbytes and t.device == dev:
            return t
    t = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev)
    if cache is not None:
        cache[key] = t
    return t


def map_match(true_rows, pred_rows, num_classes, iou_threshold=0.5, rows_by_image=False):
    """Stage 1 of the mAP on arbitrary rows (yh_map_match): device rows -> (records uint64 held as int64 (np,),
    gt_per_class int32 (C,)).  Records are in ROW order; rows_by_image=True promises ground-truth rows grouped by
    image with nondecreasing image index (skips the radix sort by image)."""
    t, p = true_rows.contiguous(), pred_rows.contiguous()
    dev = p.device
    nt, npred = int(t.shape[0]), int(p.shape[0])
    rec = torch.empty((npred,), dtype=torch.int64, device=dev)
    gtc = torch.empty((num_classes,), dtype=torch.int32, device=dev)
    with on_device(dev):
        _lib.check(_lib.lib().yh_map_match(t.data_ptr(), nt, None, p.data_ptr(), npred, None, int(num_classes),
                                           float(iou_threshold), _lib.YH_MAP_TRUE_ROWS_BY_IMAGE if rows_by_image else 0,
                                           rec.data_ptr(), gtc.data_ptr(), None, 0, stream_ptr(dev)), "map_match")
    return rec, gtc


def map_reduce(rec, gt_per_class, num_classes, nrec_dev=None, workspace=None, n_hint=0):
    """Stage 2 of the mAP (yh_map_reduce, one persistent kernel): records -> (mAP 0-d tensor, AP per class).
    nrec_dev: optional device int64 tensor holding the record count (rec.shape[0] is then only the bound);
    n_hint: expected count (sizes the grid only)."""
    dev = gt_per_class.device
    ap = torch.empty((num_classes,), dtype=torch.float32, device=dev)
    m = torch.empty((1,), dtype=torch.float32, device=dev)
    with on_device(dev):
        _lib.check(_lib.lib().yh_map_reduce(rec.data_ptr(), int(rec.shape[0]), nrec_dev.data_ptr() if nrec_dev is not None else None,
                                            int(n_hint), gt_per_class.data_ptr(), int(num_classes), ap.data_ptr(), m.data_ptr(),
                                            workspace.data_ptr() if workspace is not None else None,
                                            int(workspace.numel()) if workspace is not None else 0, stream_ptr(dev)), "map_reduce")
    return m[0], ap


def _reduce_sharded(rec, nrec_dev, n_hint, gt_per_class, num_classes, group, capacity=None):
    """The exchange step + stage 2 across the ranks of `group` (collective).  Ranks of one box: kernel-level
    exchange over NVLink (yolohot.dist.PeerExchange, no NCCL call, no host sync); otherwise padded all-gathers."""
    from . import dist as _dist
    ex = _dist.peer_exchange(rec.device, num_classes, int(rec.shape[0]), group, capacity)
    if ex is not None:
        return ex.exchange_reduce(rec, nrec_dev, gt_per_class, n_hint)
    if nrec_dev is not None:
        rec = rec[:int(nrec_dev.item())]
    rec_all, gt_all = _dist.gather_records(rec, gt_per_class, group)
    return map_reduce(rec_all, gt_all, num_classes)


def mean_average_precision(true_boxes, pred_boxes, num_classes, iou_threshold=0.5, return_ap=False, sharded=False,
                           group=None):
    """utils.py:303-456: rows [img_idx, class_idx, confidence, cx, cy, w, h] -> scalar mAP.

    Local by default, like the reference.  sharded=True makes the call a COLLECTIVE over `group` (default: the
    world): each rank passes the rows of ITS image shard (image indices only need to be unique within a rank)
    and every rank returns the global value - the per-shard records are exchanged in rank order and the
    per-class ground-truth counts summed (yolohot.dist), then reduced identically everywhere."""
    t, kind = as_device_f32(true_boxes)
    p, _ = as_device_f32(pred_boxes, t.device)
    t = t.reshape(-1, 7)
    p = p.reshape(-1, 7)
    rec, gtc = map_match(t, p, num_classes, iou_threshold)
    from . import dist as _dist
    if sharded and _dist.world_size(group) > 1:
        m, ap = _reduce_sharded(rec, None, int(rec.shape[0]), gtc, num_classes, group)
    else:
        m, ap = map_reduce(rec, gtc, num_classes)
    if kind == "numpy":
        m = np.float32(m.item())
        ap = ap.cpu().numpy()
    elif kind == "tf":
        m, ap = give_back(m, kind), give_back(ap, kind)
    return (m, ap) if return_ap else m


def change_tensor(tensor_1d, idx_col):
    """utils.py:280-299: copy of a 1-D tensor with element `idx_col` set to 1.  The reference's matching loop keeps
    its claimed-ground-truth flags with it (utils.py:411); the matching kernels keep them on the device instead,
    so this is only here for callers that import the name."""
    if isinstance(tensor_1d, torch.Tensor):
        out = tensor_1d.clone()
        out[int(idx_col)] = 1
        return out
    out = np.array(tensor_1d, copy=True)
    out[int(idx_col)] = 1
    return out


def mean_average_precision_numpy(true_boxes, pred_boxes, num_classes, iou_threshold=0.5):
    """utils.py:499-585 twin."""
    return mean_average_precision(np.asarray(true_boxes, np.float32), np.asarray(pred_boxes, np.float32),
                                  num_classes, iou_threshold)


def mean_average_precision_2(true_bboxes, pred_bboxes, iou_threshold=0.5, num_classes=20):
    """tmp.py:440-595 signature (imported by the stale metric.py:7,99)."""
    return mean_average_precision(true_bboxes, pred_bboxes, num_classes, iou_threshold)


class MeanAveragePrecision:
    """utils.py:459-496.  reset_states() / update_state(y_true, y_pred) / result().

    update_state is ONE launch and no host synchronisation (yh_eval_update_state; grids of more than 64 cells: three -
    fused decode + NMS of the predictions, of the ground truth (utils.py:475 / :480), and the accumulation kernel
    yh_eval_update): per image, decode + NMS of both tensors, then both row sets are appended to
    append-only device buffers (instead of the reference's O(total^2) re-concatenation, utils.py:484-489) and
    matches every image on the spot, leaving one packed (class, confidence, TP) record per detection.  result() is
    one more launch (yh_map_reduce: sort + cumulative TP/FP + AP + mean) and returns a 0-d device tensor.
    `all_true_boxes_variable` / `all_pred_boxes_variable` materialise the (rows, 7) views on demand.  The
    reference semantics are kept: ground truth goes through NMS too (utils.py:480), thresholds are the hard-coded
    (0.5, 0.4), reset_states() only rewinds the image counter and the next update overwrites the buffers
    (utils.py:467-468, 484-486).

    sharded=True (not in the reference): every rank of `group` feeds ITS image shard and result() becomes a
    collective that returns the global mAP on every rank; `capacity` = records per rank the exchange buffers are
    sized for (default: four times the largest shard seen at the first result())."""

    _nms_true = True
    _as_numpy = False
    _IOU_THR = 0.5            # utils.py:496 -> mean_average_precision default

    def __init__(self, num_classes, num_boxes=2, sharded=False, group=None, capacity=None):
        self._num_classes = int(num_classes)
        self._num_boxes = int(num_boxes)
        self.img_idx = 0
        self._sharded, self._group, self._capacity = bool(sharded), group, capacity
        self._dev = None
        self._st = None          # device state, see _ensure
        self._cache = {}         # scratch tensors reused across calls

    # -- state helpers
    def _ensure(self, dev, extra):
        """Row / record buffers with room for `extra` more rows of each kind.  Growth copies whole buffers on the
        device - the host only tracks upper bounds (slots, not kept rows), so nothing here synchronises."""
        st = self._st
        if st is None or st["dev"] != dev:
            cap = max(4096, 2 * extra)
            # the two row cursors and the per-class GT counts share one tensor: a restart is ONE fill launch
            state = torch.zeros((2 + (self._num_classes + 1) // 2,), dtype=torch.int64, device=dev)
            st = {"dev": dev, "cap": cap, "bound": 0,
                  "pred": torch.empty((cap, 7), dtype=torch.float32, device=dev),
                  "true": torch.empty((cap, 7), dtype=torch.float32, device=dev),
                  "rec": torch.empty((cap,), dtype=torch.int64, device=dev),
                  "state": state, "cursors": state[:2], "gt": state[2:].view(torch.int32)[:self._num_classes]}
            self._st = st
        if st["bound"] + extra > st["cap"]:
            cap = max(2 * st["cap"], st["bound"] + 2 * extra)
            for name, shape, dt in (("pred", (cap, 7), torch.float32), ("true", (cap, 7), torch.float32), ("rec", (cap,), torch.int64)):
                new = torch.empty(shape, dtype=dt, device=dev)
                new[:st["cap"]] = st[name]
                st[name] = new
            st["cap"] = cap
        return st

    def _view(self, which):
        st = self._st
        if st is None:
            t = torch.full((1, 7), -1.0, dtype=torch.float32)            # utils.py:461-462 initial value
        else:
            t = st[which][:int(st["cursors"][0 if which == "pred" else 1].item())]
        return t.cpu().numpy() if self._as_numpy else t

    @property
    def all_true_boxes_variable(self):
        return self._view("true")

    @property
    def all_pred_boxes_variable(self):
        return self._view("pred")

    # -- reference surface
    def reset_states(self):
        self.img_idx = 0

    def update_state(self, y_true, y_pred):
        yt, _ = as_device_f32(y_true)
        yp, _ = as_device_f32(y_pred, yt.device)
        yp, n, S = as_grid(yp, self._num_classes, self._num_boxes)          # train.py:208: flat head output
        yt, _, _ = as_grid(yt, self._num_classes, self._num_boxes)
        if tuple(yt.shape) != tuple(yp.shape):
            raise ValueError(f"update_state: y_true {tuple(yt.shape)} and y_pred {tuple(yp.shape)} differ")
        dev = yp.device
        self._dev = dev
        M = S * S
        st = self._ensure(dev, n * M)
        restart = self.img_idx == 0                 # utils.py:484-486: first image overwrites
        if restart:
            st["bound"] = 0
        fused_env = os.environ.get("YH_EVAL_FUSED", "")           # "0" / "1" force a path; default: by batch size
        if (self._nms_true and M <= 64 and yt.data_ptr() % 8 == 0 and yp.data_ptr() % 8 == 0
                and (fused_env == "1" or (fused_env != "0" and n * M <= 400_000))):
            # one launch: decode + NMS of both tensors, matching and append per image while its rows are in shared memory.
            # Batches of up to ~8,000 VOC images, where three launches are latency (B200, us per update_state at 2,000 /
            # 5,000 / 10,000 / 20,000 images: 23 / 33 / 65 / 132 fused, 34 / 45 / 59 / 85 as three launches); beyond that the
            # TMA tile kernels of the three-launch path stream the cells faster than one warp per image can fetch them
            with on_device(dev):
                rc = _lib.lib().yh_eval_update_state(yt.data_ptr(), yp.data_ptr(), n, S, self._num_boxes, self._num_classes, 0.5, 0.4,
                                                     int(self.img_idx), self._IOU_THR, st["pred"].data_ptr(), st["cap"],
                                                     st["true"].data_ptr(), st["cap"], st["rec"].data_ptr(), st["cursors"].data_ptr(),
                                                     st["gt"].data_ptr(), 1 if restart else 0, stream_ptr(dev))
            if rc != _lib.YH_ERR_UNSUPPORTED:          # (class tables that do not fit shared memory: the three launches below)
                _lib.check(rc, "update_state")         # (a restart zeroes cursors and counts inside the call: two memsets on
                st["bound"] += n * M                   # the stream instead of a torch fill launch)
                self.img_idx += n                        # utils.py:491
                return
        if restart:
            st["state"].zero_()
        key = (n, M)
        bufs = self._cache.get(key)
        if bufs is None:                            # padded NMS outputs of one batch, reused by the next one
            bufs = [torch.empty((n, M, 6), dtype=torch.float32, device=dev), torch.empty((n,), dtype=torch.int32, device=dev),
                    torch.empty((n, M, 6), dtype=torch.float32, device=dev), torch.empty((n,), dtype=torch.int32, device=dev)]
            self._cache = {k: v for k, v in self._cache.items() if not isinstance(k, tuple)}
            self._cache[key] = bufs
        pb, pc = decode_nms(yp, self._num_classes, self._num_boxes, 0.5, 0.4, out=(bufs[0], bufs[1]))      # utils.py:475
        if self._nms_true:                          # utils.py:480
            tb, tc = decode_nms(yt, self._num_classes, self._num_boxes, 0.5, 0.4, out=(bufs[2], bufs[3]))
        else:                                       # stale metric.py:81: conf > 0.4 only, cell order
            tb, tc = _filter_rows(decode_predictions(yt, self._num_classes, self._num_boxes), 0.4)
        with on_device(dev):
            _lib.check(_lib.lib().yh_eval_update(pb.data_ptr(), pc.data_ptr(), tb.data_ptr(), tc.data_ptr(), n, M, int(self.img_idx),
                                                 self._num_classes, self._IOU_THR, st["pred"].data_ptr(), st["cap"],
                                                 st["true"].data_ptr(), st["cap"], st["rec"].data_ptr(), st["cursors"].data_ptr(),
                                                 st["gt"].data_ptr(), stream_ptr(dev)), "update_state")
        st["bound"] += n * M
        self.img_idx += n                            # utils.py:491

    def result(self):
        """utils.py:496.  A 0-d float32 device tensor; no host synchronisation (read it when you need it)."""
        st = self._st
        from . import dist as _dist
        sharded = self._sharded and _dist.world_size(self._group) > 1
        if st is None:
            if not sharded:                          # mean_average_precision on the initial [-1]*7 rows: no class has GT
                require_cuda()
                return torch.zeros((), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            st = self._ensure(torch.device("cuda", torch.cuda.current_device()), 0)
        dev, C_ = st["dev"], self._num_classes
        rec, nrec = st["rec"], st["cursors"][0:1]        # the whole buffer bounds the count: a captured graph stays valid as rows arrive
        if sharded:
            m, self.last_ap = _reduce_sharded(rec, nrec, st["bound"], st["gt"], C_, self._group, self._capacity)
            return m
        wkey = ("ws_bytes", int(rec.shape[0]))
        nbytes = self._cache.get(wkey)
        if nbytes is None:
            nbytes = self._cache[wkey] = int(_lib.lib().yh_workspace_bytes(_lib.YH_OP_MAP_REDUCE, int(rec.shape[0]), 0, 0, C_))
        m, self.last_ap = map_reduce(rec, st["gt"], C_, nrec_dev=nrec, workspace=_workspace(dev, nbytes, self._cache), n_hint=st["bound"])
        return m


class MeanAveragePrecisionNumpy(MeanAveragePrecision):
    """utils.py:588-620 twin: NumPy in/out on the same device path.  One deviation from the reference twin is kept on
    purpose and pinned by tests: after reset_states() the next update OVERWRITES the buffers, as the TF evaluator does
    (utils.py:484-486); the reference's NumPy twin (utils.py:596-615) keeps appending to the old rows, which double
    counts the previous epoch."""
    _as_numpy = True

    def result(self):
        return np.float32(float(super().result()))


def _filter_rows(boxes, conf_threshold):
    """metric.py:81 (stale evaluator): rows with conf > thr, cell order kept (yh_filter_rows)."""
    n, M = int(boxes.shape[0]), int(boxes.shape[1])
    out = torch.empty((n, M, 6), dtype=torch.float32, device=boxes.device)
    cnt = torch.empty((n,), dtype=torch.int32, device=boxes.device)
    with on_device(boxes.device):
        _lib.check(_lib.lib().yh_filter_rows(boxes.data_ptr(), n, M, float(conf_threshold), out.data_ptr(), cnt.data_ptr(),
                                             stream_ptr(boxes.device)), "filter_rows")
    return out, cnt


# ------------------------------------------------------------- drawing epilogue
def pixel_boxes(boxes, width, height, count=None):
    """utils.py:652-655: rows [cls, conf, cx, cy, w, h] -> int32 [xmin, ymin, xmax, ymax] with
    xmin = int((cx - w/2) * width) etc. (float32, truncation).  boxes (K,6) -> (K,4); padded
    (N,M,6) with `count` (N,) -> (N,M,4), rows past count[i] set to -1."""
    b, kind = as_device_f32(boxes)
    if b.dim() == 2 and b.shape[-1] == 6:
        b3 = b.unsqueeze(0)
    elif b.dim() == 3 and b.shape[-1] == 6:
        b3 = b
    else:
        raise ValueError(f"pixel_boxes: expected (K, 6) or (N, M, 6) rows, got {tuple(b.shape)}")
    n, M = int(b3.shape[0]), int(b3.shape[1])
    out = torch.empty((n, M, 4), dtype=torch.int32, device=b.device)
    cnt = None
    if count is not None:
        cnt = (count if isinstance(count, torch.Tensor) else torch.as_tensor(np.asarray(count)))
        cnt = cnt.to(device=b.device, dtype=torch.int32).contiguous()
        if cnt.numel() != n:
            raise ValueError("pixel_boxes: count must hold one entry per image")
    if n * M:
        with on_device(b.device):
            _lib.check(_lib.lib().yh_pixel_boxes(b3.data_ptr(), cnt.data_ptr() if cnt is not None else None, n, M,
                                                 int(width), int(height), out.data_ptr(), stream_ptr(b.device)),
                       "pixel_boxes")
    out = out[0] if b.dim() == 2 else out
    return out.cpu().numpy() if kind == "numpy" else out


def _draw(img, boxes, names_path, with_grid):
    import cv2                                   # host-side drawing, exactly the reference's calls
    b, _ = as_device_f32(boxes)
    b = b.reshape(-1, 6)
    height, width = img.shape[0], img.shape[1]
    with open(names_path, "r") as f:
        class_name_list = [x.strip() for x in f.readlines()]
    px = pixel_boxes(b, width, height).cpu().numpy()
    rows = b.cpu().numpy()
    for row, (xmin, ymin, xmax, ymax) in zip(rows, px.tolist()):
        class_name = class_name_list[int(row[0])]
        img = cv2.rectangle(img, (xmin, ymin), (xmax, ymax), color=(0, 255, 0))
        if with_grid:                                                                    # utils.py:701
            img = cv2.circle(img, (int(row[2] * np.float32(width)), int(row[3] * np.float32(height))), radius=2,
                             color=(0, 0, 255))
        img = cv2.putText(img, "{:s}, {:.2f}".format(class_name, row[1]), (xmin, ymin + 20),
                          fontFace=cv2.FONT_HERSHEY_PLAIN, fontScale=1, color=(0, 255, 0))
    if with_grid:                                                                        # utils.py:708-711
        for idx in range(6):
            a = int(448 * ((idx + 1) / 7.))
            img = cv2.line(img, (a, 0), (a, height), color=(255, 0, 255))
            img = cv2.line(img, (0, a), (width, a), color=(255, 0, 255))
    return img


def get_tagged_img(img, boxes, names_path):
    """utils.py:623-663: draw the kept boxes; corner arithmetic on the device (yh_pixel_boxes)."""
    return _draw(img, boxes, names_path, False)


def get_grid_tagged_img(img, boxes, names_path):
    """utils.py:666-713: same plus box centres and the 7x7 grid lines."""
    return _draw(img, boxes, names_path, True)
