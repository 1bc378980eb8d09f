"""CPU oracle for the YOLOv1 post-processing / loss / mAP hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``keras-object-detection_b200/`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker or the timed CPU baseline - never as the product path.

PINNING: the reference (myungsanglee/Keras-Object-Detection) holds no tests and no
golden outputs, and TensorFlow is absent from this image and its wheelhouse (even the
``*_numpy`` twins call TF and use the removed ``np.int``).  This file RESTATES the
reference algorithm in float32 NumPy, loop for loop, each function citing the
``yolo_v1/*.py`` lines it follows (paths relative to /root/reference).  It is pinned
against ``tests/golden/ref_golden.npz`` = outputs of the reference's OWN utils.py /
loss.py, executed unmodified in the build container on a NumPy stand-in for the TF
primitives (tests/golden/make_ref_golden.py, tests/golden/tfshim); see
tests/test_ref_golden.py.  Still unpinned: that real TensorFlow kernels agree with
the library semantics below, and the stale metric.py/tmp.py variants (restated only).

Third-party semantics relied on (tensorflow, unpinned; ~2.4-2.6 by API use):
  * ``tf.argsort(direction='DESCENDING')`` is stable (lower index first on ties)
  * ``tf.math.argmax`` returns the first maximum
  * every TF op rounds to float32 separately (no FMA contraction)
  * unwritten ``TensorArray`` slots with a static element_shape read as zeros
  * ``np.trapz`` on float32 arrays stays in float32 (NumPy pairwise summation)
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
_EPS = 1e-6  # python float: weak scalar, becomes float32 next to float32 arrays


# --------------------------------------------------------------------------- #
# IoU                                                    utils.py:9-43 (== 46-76)
# --------------------------------------------------------------------------- #
def intersection_over_union(boxes1, boxes2):
    """Quirky element-wise IoU of [cx, cy, w, h] boxes, (...,4),(...,4)->(...,1).

    Follows utils.py:20-43 line by line: corners are ``(c -/+ w) / 2`` (the
    centre is halved too, Q1), the intersection extents are clipped to [0, 1]
    (Q2), areas go through ``abs`` and the denominator is
    ``((a1 + a2) - inter) + 1e-6``.  All float32 (utils.py:20-22 casts).
    """
    b1 = np.asarray(boxes1, dtype=F32)
    b2 = np.asarray(boxes2, dtype=F32)
    two = F32(2.0)
    b1_xmin = (b1[..., 0:1] - b1[..., 2:3]) / two          # utils.py:24
    b1_ymin = (b1[..., 1:2] - b1[..., 3:4]) / two          # utils.py:25
    b1_xmax = (b1[..., 0:1] + b1[..., 2:3]) / two          # utils.py:26
    b1_ymax = (b1[..., 1:2] + b1[..., 3:4]) / two          # utils.py:27
    b2_xmin = (b2[..., 0:1] - b2[..., 2:3]) / two          # utils.py:29
    b2_ymin = (b2[..., 1:2] - b2[..., 3:4]) / two          # utils.py:30
    b2_xmax = (b2[..., 0:1] + b2[..., 2:3]) / two          # utils.py:31
    b2_ymax = (b2[..., 1:2] + b2[..., 3:4]) / two          # utils.py:32
    ix_min = np.maximum(b1_xmin, b2_xmin)                  # utils.py:34
    iy_min = np.maximum(b1_ymin, b2_ymin)                  # utils.py:35
    ix_max = np.minimum(b1_xmax, b2_xmax)                  # utils.py:36
    iy_max = np.minimum(b1_ymax, b2_ymax)                  # utils.py:37
    inter = (np.clip(ix_max - ix_min, F32(0), F32(1)) *
             np.clip(iy_max - iy_min, F32(0), F32(1)))     # utils.py:39
    a1 = np.abs((b1_xmax - b1_xmin) * (b1_ymax - b1_ymin))  # utils.py:40
    a2 = np.abs((b2_xmax - b2_xmin) * (b2_ymax - b2_ymin))  # utils.py:41
    out = inter / (((a1 + a2) - inter) + F32(_EPS))        # utils.py:43
    return out.astype(F32, copy=False)


def _iou_scalar(a, b):
    """IoU of two 4-vectors -> python-visible float32 scalar (used by the loops)."""
    return intersection_over_union(a, b)[0]


# --------------------------------------------------------------------------- #
# Decode                                               utils.py:152-218 (221-277)
# --------------------------------------------------------------------------- #
def decode_predictions(predictions, num_classes, num_boxes=2, grid=None):
    """(N,S,S,C+5B) -> (N,S*S,6) rows [cls, conf, cx, cy, w, h], float32.

    utils.py:173-216.  ``grid`` generalises the literal 7 of utils.py:184,200-216
    (Q7); default = the tensor's own S.  The one-hot multiply/sum select of
    utils.py:184-197 is restated literally (so 0*x terms are really added).
    """
    p = np.asarray(predictions, dtype=F32)
    assert p.ndim == 4 and p.shape[1] == p.shape[2], p.shape
    S = p.shape[1] if grid is None else int(grid)
    assert p.shape[1] == S
    C, B = int(num_classes), int(num_boxes)
    assert p.shape[3] == C + 5 * B
    n = p.shape[0]

    cls = np.argmax(p[..., :C], axis=-1)[..., None].astype(F32)        # :173-175
    confs = np.stack([p[..., C + 5 * b:C + 5 * b + 1] for b in range(B)])  # :178-182
    best = np.argmax(confs, axis=0)                                    # :183 first max
    onehot = (best == np.arange(B).reshape(1, 1, 1, B)).astype(F32)    # :184
    pred_box = np.zeros((n, S, S, 4), F32)
    pred_conf = np.zeros((n, S, S, 1), F32)
    for b in range(B):                                                 # :191-197
        pred_box = pred_box + onehot[..., b:b + 1] * p[..., C + 1 + 5 * b:C + 5 + 5 * b]
        pred_conf = pred_conf + onehot[..., b:b + 1] * p[..., C + 5 * b:C + 5 * b + 1]
    base = np.tile(np.arange(S, dtype=F32).reshape(1, S), (S, 1))      # :200
    x_idx = base.reshape(S, S, 1)                                      # :201 column
    y_idx = base.T.reshape(S, S, 1)                                    # :203-204 row
    inv = F32(1.0 / S)                                                 # :207 `1 / 7 *`
    x = inv * (pred_box[..., 0:1] + x_idx)                             # :207
    y = inv * (pred_box[..., 1:2] + y_idx)                             # :208
    out = np.concatenate([cls, pred_conf, x, y, pred_box[..., 2:4]], axis=-1)  # :210-213
    return out.reshape(n, S * S, 6).astype(F32, copy=False)            # :216


# --------------------------------------------------------------------------- #
# Greedy per-class NMS                                   utils.py:79-114 (117-149)
# --------------------------------------------------------------------------- #
def non_max_suppression(boxes, iou_threshold=0.5, conf_threshold=0.4, return_index=False):
    """One image: (M,6) -> (K,6) in pick order.  Literal pop/filter loop.

    utils.py:95 strict ``>`` filter; :98 stable descending sort; :103-112 pop the
    head, keep the others iff ``cls != head.cls or iou < thr``.
    Thresholds are compared in float32 (TF converts the python float to the
    tensor's dtype).
    """
    b = np.asarray(boxes, dtype=F32).reshape(-1, 6)
    thr_c = F32(conf_threshold)
    thr_i = F32(iou_threshold)
    idx = np.nonzero(b[:, 1] > thr_c)[0]                               # :95
    order = np.argsort(-b[idx, 1], kind="stable")                      # :98
    idx = idx[order]
    cur = list(idx)
    kept = []
    while len(cur) >= 1:                                               # :103
        head = cur[0]                                                  # :104
        rest = []
        for j in cur[1:]:                                              # :106
            if b[j, 0] != b[head, 0] or _iou_scalar(b[head, 2:], b[j, 2:]) < thr_i:  # :108
                rest.append(j)
        cur = rest                                                     # :110
        kept.append(head)                                              # :112
    kept = np.asarray(kept, dtype=np.int64)
    out = b[kept].reshape(-1, 6)
    if return_index:
        return out, kept
    return out


def decode_nms(predictions, num_classes, num_boxes=2, iou_threshold=0.5, conf_threshold=0.4):
    """Batched driver of utils.py:471-480: returns (padded (N,S*S,6), count (N,), keep_idx (N,S*S))."""
    dec = decode_predictions(predictions, num_classes, num_boxes)
    n, m, _ = dec.shape
    out = np.zeros((n, m, 6), F32)
    cnt = np.zeros((n,), np.int32)
    kidx = np.full((n, m), -1, np.int32)
    for i in range(n):
        rows, keep = non_max_suppression(dec[i], iou_threshold, conf_threshold, return_index=True)
        k = rows.shape[0]
        out[i, :k] = rows
        cnt[i] = k
        kidx[i, :k] = keep
    return out, cnt, kidx


# --------------------------------------------------------------------------- #
# mAP                                          utils.py:303-456 (499-585, 280-299)
# --------------------------------------------------------------------------- #
def mean_average_precision(true_boxes, pred_boxes, num_classes, iou_threshold=0.5,
                           return_details=False):
    """Rows [img, cls, conf, cx, cy, w, h] -> scalar float32 mAP.

    utils.py:325 class loop; :329-330 row selection; :334-336 zero-GT class -> AP 0;
    :367 stable descending sort; :378 GTs of the detection's image (row order);
    :386-393 strict ``>`` best IoU from 0, best index defaults 0; :395-422 TP iff
    ``best > thr`` and that GT unclaimed; :430-446 cumsum / recall / precision /
    prepend (0,1) / ``np.trapz`` (float32); :456 mean over all C classes.
    """
    t = np.asarray(true_boxes, dtype=F32).reshape(-1, 7)
    p = np.asarray(pred_boxes, dtype=F32).reshape(-1, 7)
    thr = F32(iou_threshold)
    eps = F32(_EPS)
    aps = []
    details = []
    for c in np.arange(num_classes, dtype=F32):                        # :325
        det = p[p[:, 1] == c]                                          # :329
        gt = t[t[:, 1] == c]                                           # :330
        total = F32(gt.shape[0])                                       # :333
        if gt.shape[0] == 0:                                           # :334-336
            aps.append(F32(0))
            details.append((0, np.zeros(0, F32)))
            continue
        det = det[np.argsort(-det[:, 2], kind="stable")]               # :367
        tp = np.zeros(det.shape[0], F32)                               # :368 zeros when unwritten
        fp = np.zeros(det.shape[0], F32)                               # :369
        claimed = {}                                                   # :342-364 table+ground_truth_num
        for d_i in range(det.shape[0]):                                # :373
            d = det[d_i]
            g_img = gt[gt[:, 0] == d[0]]                               # :378
            best = F32(0)                                              # :382 unwritten slot -> 0
            best_j = 0                                                 # :383
            for j in range(g_img.shape[0]):                            # :386
                v = _iou_scalar(d[3:], g_img[j, 3:])                   # :387
                if v > best:                                           # :389
                    best = v
                    best_j = j
            if best > thr:                                             # :395
                key = (float(d[0]), best_j)
                if key not in claimed:                                 # :408
                    tp[d_i] = 1                                        # :409
                    claimed[key] = True                                # :412-415
                else:
                    fp[d_i] = 1                                        # :418
            else:
                fp[d_i] = 1                                            # :422
        tpc = np.cumsum(tp, dtype=F32)                                 # :430
        fpc = np.cumsum(fp, dtype=F32)                                 # :431
        rec = tpc / (total + eps)                                      # :434
        prec = tpc / ((tpc + fpc) + eps)                               # :435
        prec = np.concatenate([np.array([1], F32), prec])              # :438
        rec = np.concatenate([np.array([0], F32), rec])                # :439
        ap = np.trapezoid(prec, rec) if hasattr(np, "trapezoid") else np.trapz(prec, rec)  # :444
        aps.append(F32(ap))
        details.append((gt.shape[0], tp))
    aps = np.asarray(aps, F32)
    m = F32(np.sum(aps, dtype=F32) / F32(len(aps)))                    # :456 reduce_mean
    if return_details:
        return m, aps, details
    return m


class MeanAveragePrecision:
    """utils.py:459-496 accumulator (live evaluator: GT rows are NMS'd too, Q9)."""

    def __init__(self, num_classes, num_boxes=2, nms_true=True):
        self._num_classes = num_classes
        self._num_boxes = num_boxes
        self._nms_true = nms_true           # False -> stale metric.py:81 variant (a7)
        self.all_true_boxes_variable = np.full((1, 7), -1, F32)        # :461
        self.all_pred_boxes_variable = np.full((1, 7), -1, F32)        # :462
        self.img_idx = 0

    def reset_states(self):                                            # :467-468 (Q17)
        self.img_idx = 0

    def update_state(self, y_true, y_pred):                            # :470-491
        tb = decode_predictions(y_true, self._num_classes, self._num_boxes)
        pb = decode_predictions(y_pred, self._num_classes, self._num_boxes)
        for i in range(tb.shape[0]):
            pn = non_max_suppression(pb[i], 0.5, 0.4)                  # :475
            if self._nms_true:
                tn = non_max_suppression(tb[i], 0.5, 0.4)              # :480
            else:
                tn = tb[i][tb[i][:, 1] > F32(0.4)]                     # metric.py:81
            pc = np.concatenate([np.full((pn.shape[0], 1), self.img_idx, F32), pn], axis=1)
            tc = np.concatenate([np.full((tn.shape[0], 1), self.img_idx, F32), tn], axis=1)
            if self.img_idx == 0:                                      # :484-486
                self.all_true_boxes_variable = tc
                self.all_pred_boxes_variable = pc
            else:                                                      # :488-489
                self.all_true_boxes_variable = np.concatenate([self.all_true_boxes_variable, tc])
                self.all_pred_boxes_variable = np.concatenate([self.all_pred_boxes_variable, pc])
            self.img_idx += 1                                          # :491

    def result(self):                                                  # :493-496
        return mean_average_precision(self.all_true_boxes_variable,
                                      self.all_pred_boxes_variable, self._num_classes)


# --------------------------------------------------------------------------- #
# Loss                                                          loss.py:100-215
# --------------------------------------------------------------------------- #
def yolo_v1_loss(y_true, y_pred, num_classes=20, num_boxes=2,
                 lambda_coord=5.0, lambda_noobj=0.5):
    """Forward of YoloV1Loss.call.  Returns dict of float32 element-wise terms summed
    in float64 (``*_f64``) and with NumPy's float32 pairwise sum (``*_f32``):
    keys xy, wh, obj, noobj, cls, total.  loss.py:126-213.
    """
    t = np.asarray(y_true, dtype=F32)
    p = np.asarray(y_pred, dtype=F32)
    C, B = int(num_classes), int(num_boxes)
    tb = t[..., C + 1:C + 5]                                           # :129,158
    ious = np.stack([intersection_over_union(tb, p[..., C + 1 + 5 * b:C + 5 + 5 * b])
                     for b in range(B)])                               # :127-133
    best = np.argmax(ious, axis=0)                                     # :136 first max
    onehot = (best == np.arange(B).reshape((1,) * (t.ndim - 1) + (B,))).astype(F32)  # :137
    pred_box = np.zeros(t.shape[:-1] + (4,), F32)
    pred_conf = np.zeros(t.shape[:-1] + (1,), F32)
    pred_iou = np.zeros(t.shape[:-1] + (1,), F32)
    for b in range(B):                                                 # :146-155
        pred_box = pred_box + onehot[..., b:b + 1] * p[..., C + 1 + 5 * b:C + 5 + 5 * b]
        pred_conf = pred_conf + onehot[..., b:b + 1] * p[..., C + 5 * b:C + 5 * b + 1]
        pred_iou = pred_iou + onehot[..., b:b + 1] * ious[b]
    obj = t[..., C:C + 1]                                              # :162
    noobj = F32(1) - obj                                               # :163
    xy = obj * np.square(tb[..., 0:2] - pred_box[..., 0:2])            # :171
    with np.errstate(invalid="ignore"):
        wh = obj * np.square(np.sqrt(tb[..., 2:4]) -
                             (np.sign(pred_box[..., 2:4]) *
                              np.sqrt(np.abs(pred_box[..., 2:4]) + F32(_EPS))))  # :176-178
    ob = obj * np.square(pred_iou - pred_conf)                         # :189
    nb = noobj * np.square(F32(0) - pred_conf)                         # :197
    cl = obj * np.square(t[..., :C] - p[..., :C])                      # :206
    out = {}
    for name, arr in (("xy", xy), ("wh", wh), ("obj", ob), ("noobj", nb), ("cls", cl)):
        out[name + "_f64"] = float(np.sum(arr, dtype=np.float64))
        out[name + "_f32"] = F32(np.sum(arr, dtype=F32))
    lc, ln = F32(lambda_coord), F32(lambda_noobj)
    out["total_f32"] = F32(lc * (out["xy_f32"] + out["wh_f32"]) + out["obj_f32"]
                           + ln * out["noobj_f32"] + out["cls_f32"])   # :210-213
    out["total_f64"] = (float(lambda_coord) * (out["xy_f64"] + out["wh_f64"]) + out["obj_f64"]
                        + float(lambda_noobj) * out["noobj_f64"] + out["cls_f64"])
    out["responsible"] = best[..., 0].astype(np.int32)
    return out


def yolo_v1_loss_grad(y_true, y_pred, num_classes=20, num_boxes=2,
                      lambda_coord=5.0, lambda_noobj=0.5, float32_forward=True):
    """Closed-form d(loss)/d(y_pred), evaluated in float64 from the
    float32 inputs, with TF's sub-gradient conventions: clip passes gradient on
    [0, 1] inclusive; max/min route ties to their FIRST argument (the true box,
    loss.py:128-131 passes it first); abs'(0) = sign(0) = 0.
    The responsible box is chosen exactly as the float32 forward does.
    """
    t32 = np.asarray(y_true, dtype=F32)
    p32 = np.asarray(y_pred, dtype=F32)
    C, B = int(num_classes), int(num_boxes)
    fwd = yolo_v1_loss(t32, p32, C, B, lambda_coord, lambda_noobj)
    k = fwd["responsible"]                                             # (...,)
    t = t32.astype(np.float64)
    p = p32.astype(np.float64)
    g = np.zeros_like(p)
    obj = t[..., C]
    lead = t.shape[:-1]
    idx = np.indices(lead)
    base = C + 5 * k
    def take(off):
        return p[(*idx, base + off)]
    c = take(0); px = take(1); py = take(2); pw = take(3); ph = take(4)
    tx, ty, tw, th = t[..., C + 1], t[..., C + 2], t[..., C + 3], t[..., C + 4]
    # forward IoU pieces: corners and extents are taken from the FLOAT32 forward (utils.py:24-39),
    # because TF's autodiff decides ties of max/min/clip on the float32 values it computed
    # (the loss.py:219-234 fixture sits exactly on such a tie: (0.49-0.09)/2 == (0.5-0.1)/2 in
    # float32 but not in float64); the smooth factors are then evaluated in float64.
    two = F32(2.0)
    def take32(off):
        return p32[(*idx, base + off)]
    px32, py32, pw32, ph32 = take32(1), take32(2), take32(3), take32(4)
    tx32, ty32, tw32, th32 = t32[..., C + 1], t32[..., C + 2], t32[..., C + 3], t32[..., C + 4]
    x1n, x1x = ((tx32 - tw32) / two).astype(np.float64), ((tx32 + tw32) / two).astype(np.float64)
    y1n, y1x = ((ty32 - th32) / two).astype(np.float64), ((ty32 + th32) / two).astype(np.float64)
    x2n, x2x = ((px32 - pw32) / two).astype(np.float64), ((px32 + pw32) / two).astype(np.float64)
    y2n, y2x = ((py32 - ph32) / two).astype(np.float64), ((py32 + ph32) / two).astype(np.float64)
    dx = (np.minimum(x1x, x2x).astype(F32) - np.maximum(x1n, x2n).astype(F32)).astype(np.float64)
    dy = (np.minimum(y1x, y2x).astype(F32) - np.maximum(y1n, y2n).astype(F32)).astype(np.float64)
    if not float32_forward:   # pure float64 evaluation (cross-check against float64 autograd)
        x1n, x1x = (tx - tw) / 2, (tx + tw) / 2
        y1n, y1x = (ty - th) / 2, (ty + th) / 2
        x2n, x2x = (px - pw) / 2, (px + pw) / 2
        y2n, y2x = (py - ph) / 2, (py + ph) / 2
        dx = np.minimum(x1x, x2x) - np.maximum(x1n, x2n)
        dy = np.minimum(y1x, y2x) - np.maximum(y1n, y2n)
    cw, ch = np.clip(dx, 0, 1), np.clip(dy, 0, 1)
    inter = cw * ch
    a1 = np.abs((x1x - x1n) * (y1x - y1n))
    w2, h2 = x2x - x2n, y2x - y2n
    a2 = np.abs(w2 * h2)
    dn = a1 + a2 - inter + 1e-6
    u = inter / dn
    du_dI = 1.0 / dn + inter / dn ** 2
    du_da2 = -inter / dn ** 2
    in_x = ((dx >= 0) & (dx <= 1)).astype(np.float64)
    in_y = ((dy >= 0) & (dy <= 1)).astype(np.float64)
    # min(x1x, x2x): pred gets it only if strictly smaller; max(x1n, x2n): only if strictly larger
    mx = (x2x < x1x).astype(np.float64); nx = (x2n > x1n).astype(np.float64)
    my = (y2x < y1x).astype(np.float64); ny = (y2n > y1n).astype(np.float64)
    ddx_dpx = 0.5 * (mx - nx); ddx_dpw = 0.5 * (mx + nx)
    ddy_dpy = 0.5 * (my - ny); ddy_dph = 0.5 * (my + ny)
    sgn_a2 = np.sign(w2 * h2)
    # x2x - x2n == pw exactly in exact arithmetic: d(w2)/d(pw) = 1, d(w2)/d(px) = 0
    du_dpx = du_dI * in_x * ch * ddx_dpx
    du_dpy = du_dI * in_y * cw * ddy_dpy
    du_dpw = du_dI * in_x * ch * ddx_dpw + du_da2 * sgn_a2 * h2
    du_dph = du_dI * in_y * cw * ddy_dph + du_da2 * sgn_a2 * w2
    e = 2.0 * obj * (u - c)                       # d(obj*(u-c)^2)/du
    lc, ln = float(lambda_coord), float(lambda_noobj)
    g_c = -e + ln * 2.0 * (1.0 - obj) * c
    g_x = -2.0 * lc * obj * (tx - px) + e * du_dpx
    g_y = -2.0 * lc * obj * (ty - py) + e * du_dpy
    def wh_grad(tv, pv):
        s = np.sign(pv)
        r = np.sqrt(np.abs(pv) + 1e-6)
        with np.errstate(invalid="ignore"):
            diff = np.sqrt(tv) - s * r
        # d/dp [s*sqrt(|p|+eps)] = s * sign(p) / (2 r) = s^2/(2r)
        return -2.0 * lc * obj * diff * (s * s) / (2.0 * r)
    g_w = wh_grad(tw, pw) + e * du_dpw
    g_h = wh_grad(th, ph) + e * du_dph
    for off, val in ((0, g_c), (1, g_x), (2, g_y), (3, g_w), (4, g_h)):
        g[(*idx, base + off)] = val
    g[..., :C] = -2.0 * obj[..., None] * (t[..., :C] - p[..., :C])
    return g


# --------------------------------------------------------------------------- #
# Label-grid encoder (defines the y_true layout)          dataset.py:88-112
# --------------------------------------------------------------------------- #
def encode_labels(boxes, grid, num_classes, num_boxes):
    """[[cx, cy, w, h, cls], ...] -> (S,S,C+5B) float64 like dataset.py:88-112
    (first writer wins per cell; x uses the column, y the row)."""
    S, C = int(grid), int(num_classes)
    m = np.zeros((S, S, C + 5 * int(num_boxes)))
    for box in boxes:
        cls = int(box[-1])
        cx, cy, w, h = box[0], box[1], box[2], box[3]
        loc = [S * cy, S * cx]
        li, lj = int(loc[0]), int(loc[1])
        y = loc[0] - li
        x = loc[1] - lj
        if m[li, lj, C] == 0:
            m[li, lj, cls] = 1
            m[li, lj, C + 1:C + 5] = [x, y, w, h]
            m[li, lj, C] = 1
    return m


def encode_labels_batch(boxes, offsets, grid, num_classes, num_boxes):
    """dataset.py:72-86 label half: (total,5) float64 rows + (N+1) offsets -> (N,S,S,C+5B) float32
    (each image's float64 grid is assigned into the float32 batch array, dataset.py:85)."""
    n = len(offsets) - 1
    out = np.zeros((n, grid, grid, num_classes + 5 * num_boxes), F32)
    for i in range(n):
        out[i] = encode_labels(np.asarray(boxes, np.float64)[offsets[i]:offsets[i + 1]], grid, num_classes, num_boxes)
    return out


def pixel_boxes(rows, width, height):
    """utils.py:652-655 on float32 values: xmin = int((x - (w / 2)) * width) ...; int() truncates."""
    r = np.asarray(rows, F32).reshape(-1, 6)
    out = np.zeros((len(r), 4), np.int32)
    for i, box in enumerate(r):
        x, y, w, h = box[2], box[3], box[4], box[5]
        out[i] = [int((x - (w / 2)) * width), int((y - (h / 2)) * height),
                  int((x + (w / 2)) * width), int((y + (h / 2)) * height)]
    return out
