"""CPU tests: the oracle against (a) the values hand-derivable from the reference source for
its own fixtures (asserted as literals), (b) the committed golden file,
(c) its C port, (d) torch autograd for the closed-form loss gradient."""
import os

import numpy as np
import pytest
import torch

from oracle import cport
from oracle import yolo_oracle as O
from tests import fixtures as F

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))
F32 = np.float32


def test_iou_quirk_q1():
    # App. A.0 Q1: reference 0.59999 vs geometric 0.33333 for these two boxes
    v = O.intersection_over_union(np.array([0.30, 0.30, 0.2, 0.2], F32), np.array([0.40, 0.30, 0.2, 0.2], F32))
    assert v.shape == (1,) and abs(float(v[0]) - 0.59999) < 2e-5
    # symmetric bit for bit
    rng = np.random.default_rng(0)
    a, b = rng.random((1000, 4), dtype=F32), rng.random((1000, 4), dtype=F32)
    assert np.array_equal(O.intersection_over_union(a, b), O.intersection_over_union(b, a))


def test_utils_demo_known_answers():
    yt, yp = F.utils_demo()
    nms = O.non_max_suppression(O.decode_predictions(yp, 3, 2)[0])
    want = np.array([[1, 0.9, 0.49857146, 0.49857146, 0.1, 0.1],
                     [0, 0.6, 0.07000001, 0.07000001, 0.1, 0.1],
                     [2, 0.6, 0.92714286, 0.92714286, 0.1, 0.1]], F32)     # App. B-1 (stable-sort witness)
    np.testing.assert_allclose(nms, want, rtol=1e-6)
    assert np.array_equal(nms, G["utils_demo_nms_pred"])
    t = O.non_max_suppression(O.decode_predictions(yt, 3, 2)[0])
    np.testing.assert_allclose(t[:, 2], [0.071428575, 0.5, 0.92857146], rtol=1e-6)
    assert np.array_equal(t, G["utils_demo_nms_true"])
    ev = O.MeanAveragePrecision(3, 2)
    ev.update_state(yt, yp)
    assert abs(float(ev.result()) - 0.99999857) < 1e-7                    # each AP = 1/(1+1e-6) in float32
    assert ev.result() == G["utils_demo_map"]


def test_loss_demo_known_answers():
    yt, yp = F.loss_demo()
    L = O.yolo_v1_loss(yt, yp, 3, 2)
    assert L["responsible"][0, 0, 0] == 0                                  # IoUs (0.8099, 0.3712)
    want = dict(xy=0.00019999962, wh=0.00052657153, obj=0.012082228, noobj=0.0, cls=0.15999998, total=0.17571506)
    for k, v in want.items():
        assert abs(L[k + "_f64"] - v) <= 1e-6 * max(abs(v), 1e-3), (k, L[k + "_f64"], v)
    np.testing.assert_allclose([L[k + "_f64"] for k in ("xy", "wh", "obj", "noobj", "cls", "total")],
                               G["loss_demo_terms"], rtol=1e-12)


def test_test_txt_roundtrip():
    lab = O.encode_labels(F.TEST_TXT_BOXES, 7, 3, 2)
    assert [tuple(x) for x in np.argwhere(lab[..., 3] == 1)] == [(1, 5), (3, 3), (4, 2)]   # App. B-3
    out = O.non_max_suppression(O.decode_predictions(lab[None].astype(F32), 3, 2)[0])
    for row, (cx, cy, w, h, c) in zip(out, F.TEST_TXT_BOXES):
        np.testing.assert_allclose(row, [c, 1, cx, cy, w, h], rtol=2e-6)
    assert np.array_equal(out, G["test_txt_nms"])


def test_metric_demo_golden():
    mt, mp, mp2 = F.metric_demo()
    ev = O.MeanAveragePrecision(20, 2, nms_true=False)
    for i in range(5):
        ev.update_state(mt, mp if i == 0 else mp2)
    assert ev.result() == G["metric_demo_map"]
    # 2 GT classes of 20; class 1: 4 of 5 images detected at conf 0.9 -> AP ~ 0.8, class 5 never detected
    assert 0.0 < float(ev.result()) < 0.1


def test_golden_synthetic():
    b, c, k = O.decode_nms(F.synth_dense(8, seed=1234), 20, 2)
    assert np.array_equal(c, G["dense8_count"]) and np.array_equal(k, G["dense8_idx"]) and np.array_equal(b, G["dense8_boxes"])
    _, c, k = O.decode_nms(F.synth_quantised(4, 14, 3, 80, seed=5), 80, 3, 0.5, 0.05)
    assert np.array_equal(c, G["quant4_count"]) and np.array_equal(k, G["quant4_idx"])


def test_cport_matches_numpy_oracle():
    for (gen, n, S, B, C, it, ct) in ((F.synth_dense, 48, 7, 2, 20, 0.5, 0.4), (F.synth_sparse, 64, 7, 2, 20, 0.5, 0.4),
                                      (F.synth_quantised, 24, 7, 2, 20, 0.5, 0.3), (F.synth_quantised, 6, 14, 3, 80, 0.5, 0.05),
                                      (F.synth_stress, 4, 14, 3, 80, 0.5, 0.05), (F.synth_dense, 16, 4, 1, 3, 0.3, 0.2)):
        p = gen(n, S, B, C)
        a = O.decode_nms(p, C, B, it, ct)
        b = cport.decode_nms(p, C, B, it, ct)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), gen.__name__
        assert np.array_equal(O.decode_predictions(p, C, B), cport.decode(p, C, B))
    rng = np.random.default_rng(1)
    a, b = rng.normal(0.4, 0.3, (5000, 4)).astype(F32), rng.normal(0.4, 0.3, (5000, 4)).astype(F32)
    assert np.array_equal(O.intersection_over_union(a, b)[:, 0], cport.iou(a, b))


def test_cport_threads_equal_single():
    p = F.synth_dense(1000, seed=3)
    a = cport.decode_nms(p, 20, 2, nthreads=1)
    b = cport.decode_nms(p, 20, 2, nthreads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_cport_loss_and_map():
    yt = F.synth_labels(64, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    L = O.yolo_v1_loss(yt, yp)
    np.testing.assert_allclose(cport.loss(yt, yp), [L[k + "_f64"] for k in ("xy", "wh", "obj", "noobj", "cls", "total")], rtol=1e-12)
    yt = F.synth_labels(120, seed=11)
    mp = F.synth_map_pred(yt)
    ev = O.MeanAveragePrecision(20, 2)
    ev.update_state(yt, mp)
    m, _ = cport.mean_average_precision(ev.all_true_boxes_variable, ev.all_pred_boxes_variable, 20)
    assert abs(m - float(ev.result())) < 1e-6
    assert 0.05 < m < 0.95


def test_nms_bitmask_formulation_equals_pop_loop():
    """rank + same-class IoU bits + greedy scan == the reference's pop/filter loop."""
    p = F.synth_quantised(12, 14, 3, 80, seed=9)
    dec = O.decode_predictions(p, 80, 3)
    for i in range(dec.shape[0]):
        rows = dec[i]
        want, widx = O.non_max_suppression(rows, 0.5, 0.05, return_index=True)
        s = rows[:, 1]
        cand = np.nonzero(s > F32(0.05))[0]
        rank = np.array([np.sum(s[cand] > s[j]) + np.sum((s[cand] == s[j]) & (cand < j)) for j in cand])
        order = cand[np.argsort(rank)]
        keep = []
        for j in order:
            if not any(rows[q, 0] == rows[j, 0] and not (O._iou_scalar(rows[q, 2:], rows[j, 2:]) < F32(0.5)) for q in keep):
                keep.append(j)
        assert np.array_equal(np.array(keep), widx)


def test_sharded_map_equals_global():
    """per-shard matching + rank-ordered concatenation == global mAP."""
    yt = F.synth_labels(200, seed=11)
    mp = F.synth_map_pred(yt)
    ev = O.MeanAveragePrecision(20, 2)
    ev.update_state(yt, mp)
    whole = ev.result()
    t_all, p_all = [], []
    for lo, hi in ((0, 50), (50, 100), (100, 150), (150, 200)):
        e = O.MeanAveragePrecision(20, 2)
        e.update_state(yt[lo:hi], mp[lo:hi])
        t, p = e.all_true_boxes_variable.copy(), e.all_pred_boxes_variable.copy()
        t[:, 0] += lo
        p[:, 0] += lo
        t_all.append(t)
        p_all.append(p)
    assert O.mean_average_precision(np.concatenate(t_all), np.concatenate(p_all), 20) == whole


def test_closed_form_grad_matches_autograd():
    """closed form of App. A.6 vs torch autograd of a literal float64 restatement."""
    yt = F.synth_labels(24, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    C, B = 20, 2
    t = torch.from_numpy(yt).double()
    p = torch.from_numpy(yp).double().requires_grad_(True)

    def iou(b1, b2):
        x1n, y1n = (b1[..., 0:1] - b1[..., 2:3]) / 2, (b1[..., 1:2] - b1[..., 3:4]) / 2
        x1x, y1x = (b1[..., 0:1] + b1[..., 2:3]) / 2, (b1[..., 1:2] + b1[..., 3:4]) / 2
        x2n, y2n = (b2[..., 0:1] - b2[..., 2:3]) / 2, (b2[..., 1:2] - b2[..., 3:4]) / 2
        x2x, y2x = (b2[..., 0:1] + b2[..., 2:3]) / 2, (b2[..., 1:2] + b2[..., 3:4]) / 2
        inter = (torch.clamp(torch.minimum(x1x, x2x) - torch.maximum(x1n, x2n), 0, 1) *
                 torch.clamp(torch.minimum(y1x, y2x) - torch.maximum(y1n, y2n), 0, 1))
        a1 = torch.abs((x1x - x1n) * (y1x - y1n))
        a2 = torch.abs((x2x - x2n) * (y2x - y2n))
        return inter / (a1 + a2 - inter + 1e-6)

    ious = torch.stack([iou(t[..., C + 1:C + 5], p[..., C + 1 + 5 * b:C + 5 + 5 * b]) for b in range(B)])
    k = torch.from_numpy(O.yolo_v1_loss(yt, yp)["responsible"]).long()          # float32 choice of the forward
    oh = torch.nn.functional.one_hot(k, B).double()
    pb = sum(oh[..., b:b + 1] * p[..., C + 1 + 5 * b:C + 5 + 5 * b] for b in range(B))
    pc = sum(oh[..., b:b + 1] * p[..., C + 5 * b:C + 5 * b + 1] for b in range(B))
    pi = sum(oh[..., b:b + 1] * ious[b] for b in range(B))
    obj = t[..., C:C + 1]
    tb = t[..., C + 1:C + 5]
    xy = (obj * (tb[..., 0:2] - pb[..., 0:2]) ** 2).sum()
    wh = (obj * (torch.sqrt(tb[..., 2:4]) - torch.sign(pb[..., 2:4]) * torch.sqrt(torch.abs(pb[..., 2:4]) + 1e-6)) ** 2).sum()
    ob = (obj * (pi - pc) ** 2).sum()
    nb = ((1 - obj) * (0 - pc) ** 2).sum()
    cl = (obj * (t[..., :C] - p[..., :C]) ** 2).sum()
    (5 * (xy + wh) + ob + 0.5 * nb + cl).backward()
    g = O.yolo_v1_loss_grad(yt, yp, float32_forward=False)
    np.testing.assert_allclose(g, p.grad.numpy(), rtol=1e-9, atol=1e-9)
    # default mode takes the tie / clip masks from the float32 forward: same away from ties
    np.testing.assert_allclose(O.yolo_v1_loss_grad(yt, yp), p.grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(O.yolo_v1_loss_grad(*F.loss_demo(), 3, 2), G["loss_demo_grad"], rtol=1e-12, atol=1e-15)


def test_division_free_iou_threshold_filter_is_exact():
    """The NMS kernels decide `fl32(inter / den) < thr` (utils.py:108) without dividing whenever
    |fl32(inter - fl32(thr * den))| > 2^-20 * fl32(thr * den) (csrc/yh_decode_nms.cu, suppresses()).
    Emulated here in float32 NumPy on adversarial inputs (quotients within a few ulp of thr, and
    quotients right at the edge of the 2^-20 band): wherever the filter decides, it decides like the division."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    eps = F32(2.0 ** -20)
    for thr in (0.5, 0.45, 0.3, 0.25, 0.2, 0.05, 0.7, 0.999999, 1.0, 1e-3, 3.0):
        t = F32(thr)
        den = (rng.random(n, dtype=F32) * F32(2) + F32(1e-6)).astype(F32)
        base = (den.astype(np.float64) * np.float64(t)).astype(F32)
        inter = base.copy()
        k = rng.integers(-4, 5, n).astype(np.int32)
        inter[: n // 2] = (base[: n // 2].view(np.int32) + k[: n // 2]).view(F32)            # within 4 ulp of thr * den
        edge = (1.0 + rng.choice([-1.0, 1.0], n - n // 2) * (2.0 ** -20) * (1.0 + 1e-3 * rng.standard_normal(n - n // 2)))
        inter[n // 2:] = (base[n // 2:].astype(np.float64) * edge).astype(F32)              # at the edge of the band
        inter[:1000] = rng.random(1000, dtype=F32) * den[:1000]                              # and ordinary values
        ref_keep = (inter / den) < t                                                         # IEEE float32 division
        p = t * den                                                                          # float32 product
        d = inter - p                                                                        # float32 difference
        decided = np.abs(d) > p * eps
        assert decided[:1000].mean() > 0.99 and 0.2 < decided[n // 2:].mean() < 0.8          # both sides of the band hit
        assert np.array_equal((d < 0)[decided], ref_keep[decided]), thr
