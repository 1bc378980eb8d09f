"""Pins against tests/golden/ref_golden.npz = outputs of the reference's OWN source files
(yolo_v1/utils.py, yolo_v1/loss.py) executed in the build container on a NumPy stand-in for
TensorFlow (tests/golden/make_ref_golden.py, tests/golden/tfshim).

CPU part: the oracle (NumPy restatement and its C port) must reproduce those outputs.
GPU part (-m gpu): the CUDA path, through the reference's call surface, must reproduce them.
Bars (BASELINE.json north_star): IoU / decoded rows / kept rows / counts bit-exact; loss within
1e-5 relative; gradient within 1e-4 (float32 autograd vs closed form); mAP within 1e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import cport
from oracle import yolo_oracle as O

R = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))
F32 = np.float32
NMS_CASES = [("dense", 20, 2), ("sparse", 20, 2), ("quant", 20, 2), ("quant_lo", 20, 2), ("b3c5", 5, 3)]


def _kept(rows, cnt):
    m = np.arange(rows.shape[1])[None, :] < np.asarray(cnt)[:, None]
    return rows[m]


# ------------------------------------------------------------------ CPU: oracle vs reference
def test_oracle_iou_matches_reference_source():
    assert np.array_equal(O.intersection_over_union(R["iou_a"], R["iou_b"]), R["iou_out"])
    assert np.array_equal(cport.iou(R["iou_a"], R["iou_b"]).reshape(-1, 1), R["iou_out"])


@pytest.mark.parametrize("name,C,B", NMS_CASES, ids=[c[0] for c in NMS_CASES])
def test_oracle_decode_nms_matches_reference_source(name, C, B):
    p, it, ct = R[f"{name}_in"], float(R[f"{name}_thr"][0]), float(R[f"{name}_thr"][1])
    assert np.array_equal(O.decode_predictions(p, C, B), R[f"{name}_decode"])
    assert np.array_equal(cport.decode(p, C, B), R[f"{name}_decode"])
    for fn in (O.decode_nms, cport.decode_nms):
        rows, cnt, _ = fn(p, C, B, it, ct)
        assert np.array_equal(cnt, R[f"{name}_count"]), name
        assert np.array_equal(_kept(rows, cnt), _kept(R[f"{name}_rows"], R[f"{name}_count"])), name
    if name == "sparse":
        assert R["sparse_count"][5] == 0 and R["sparse_count"][6] == 0      # empty image; conf == thr is dropped


def test_oracle_reference_fixtures():
    yp, yt = R["demo_yp"], R["demo_yt"]
    assert np.array_equal(O.decode_predictions(yp, 3, 2), R["demo_decode_pred"])
    assert np.array_equal(O.non_max_suppression(O.decode_predictions(yp, 3, 2)[0]), R["demo_nms_pred"])
    assert np.array_equal(O.non_max_suppression(O.decode_predictions(yt, 3, 2)[0]), R["demo_nms_true"])
    ev = O.MeanAveragePrecision(3, 2)
    ev.update_state(yt, yp)
    assert np.array_equal(ev.all_true_boxes_variable, R["demo_true_rows"])
    assert np.array_equal(ev.all_pred_boxes_variable, R["demo_pred_rows"])
    assert abs(float(ev.result()) - float(R["demo_map"])) <= 1e-6
    L = O.yolo_v1_loss(R["loss_demo_yt"], R["loss_demo_yp"], 3, 2)
    assert abs(L["total_f64"] - float(R["loss_demo_total"])) <= 1e-5 * float(R["loss_demo_total"])


def test_oracle_loss_and_gradient_match_reference_source():
    for key, C, B in (("loss16", 20, 2), ("loss_b3", 5, 3)):
        L = O.yolo_v1_loss(R[f"{key}_yt"], R[f"{key}_yp"], C, B)
        want = float(R[f"{key}_total"])
        assert abs(L["total_f64"] - want) <= 1e-5 * abs(want), key
        assert abs(float(cport.loss(R[f"{key}_yt"], R[f"{key}_yp"], C, B)[5]) - want) <= 1e-5 * abs(want), key
    g = O.yolo_v1_loss_grad(R["loss16_yt"], R["loss16_yp"])
    np.testing.assert_allclose(g, R["loss16_grad_torch"], rtol=1e-4, atol=1e-4)


def test_oracle_map_matches_reference_source():
    ev = O.MeanAveragePrecision(20, 2)
    ev.update_state(R["map40_yt"][:25], R["map40_yp"][:25])
    ev.update_state(R["map40_yt"][25:], R["map40_yp"][25:])
    assert np.array_equal(ev.all_true_boxes_variable, R["map40_true_rows"])
    assert np.array_equal(ev.all_pred_boxes_variable, R["map40_pred_rows"])
    assert abs(float(ev.result()) - float(R["map40_map"])) <= 1e-6
    assert abs(float(cport.mean_average_precision(R["map40_true_rows"], R["map40_pred_rows"], 20)[0]) - float(R["map40_map"])) <= 1e-6
    for key, thr in (("rows_map", 0.5), ("rows_map_thr03", 0.3)):
        got = O.mean_average_precision(R["rows_true"], R["rows_pred"], 4, thr)
        assert abs(float(got) - float(R[key])) <= 1e-6, (key, got, R[key])
    ev = O.MeanAveragePrecision(20, 2)                                  # reset_states semantics (Q17)
    ev.update_state(R["map40_yt"][:10], R["map40_yp"][:10])
    ev.reset_states()
    ev.update_state(R["map40_yt"][10:16], R["map40_yp"][10:16])
    ev.update_state(R["map40_yt"][16:20], R["map40_yp"][16:20])
    assert np.array_equal(ev.all_true_boxes_variable, R["reset_true_rows"])
    assert np.array_equal(ev.all_pred_boxes_variable, R["reset_pred_rows"])
    assert abs(float(ev.result()) - float(R["reset_map"])) <= 1e-6
    assert np.array_equal(O.intersection_over_union(R["iou4d_a"], R["iou4d_b"]), R["iou4d_out"])


def test_oracle_label_grids_match_reference_source():
    assert np.array_equal(O.encode_labels_batch(R["lab_txt_boxes"], [0, 3], 7, 3, 2), R["lab_txt_out"])
    assert np.array_equal(O.encode_labels_batch(R["lab_boxes"], R["lab_offsets"], 7, 20, 2), R["lab_out"])
    assert np.array_equal(O.encode_labels_batch(R["lab14_boxes"], R["lab14_offsets"], 14, 80, 3), R["lab14_out"])
    assert int(R["lab_out_of_grid_raises"]) == 1
    with pytest.raises(IndexError):
        O.encode_labels(np.array([[1.0, 0.5, 0.1, 0.1, 0]]), 7, 20, 2)


def test_oracle_pixel_boxes_match_reference_source():
    assert np.array_equal(O.pixel_boxes(R["px_rows"], 448, 448), R["px_448"])
    assert np.array_equal(O.pixel_boxes(R["px_rows"], 500, 375), R["px_375x500"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/yolo_v1"), reason="reference sources only exist in the build container")
def test_committed_golden_is_what_the_reference_source_returns(tmp_path):
    """Re-executes the reference's files (make_ref_golden.py) and compares with the committed file."""
    import subprocess
    import sys
    out = str(tmp_path / "regen.npz")
    script = os.path.join(os.path.dirname(__file__), "golden", "make_ref_golden.py")
    subprocess.check_call([sys.executable, script, out], stdout=subprocess.DEVNULL)
    new = np.load(out)
    assert sorted(new.files) == sorted(R.files)
    for k in R.files:
        assert np.array_equal(new[k], R[k], equal_nan=True), k


# ------------------------------------------------------------------ GPU: CUDA path vs reference
@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _cuda(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.gpu
def test_cuda_iou_matches_reference_source(dev):
    from yolohot import utils as yu
    out = yu.intersection_over_union(_cuda(R["iou_a"], dev), _cuda(R["iou_b"], dev)).cpu().numpy()
    assert out.shape == R["iou_out"].shape and np.array_equal(out, R["iou_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,C,B", NMS_CASES, ids=[c[0] for c in NMS_CASES])
def test_cuda_decode_nms_matches_reference_source(dev, name, C, B):
    from yolohot import utils as yu
    p, it, ct = R[f"{name}_in"], float(R[f"{name}_thr"][0]), float(R[f"{name}_thr"][1])
    dec = yu.decode_predictions(_cuda(p, dev), C, B)
    assert np.array_equal(dec.cpu().numpy(), R[f"{name}_decode"])
    rows, cnt = yu.decode_nms(_cuda(p, dev), C, B, it, ct)[:2]
    rows, cnt = rows.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(cnt, R[f"{name}_count"])
    assert np.array_equal(_kept(rows, cnt), _kept(R[f"{name}_rows"], R[f"{name}_count"]))
    # the reference's own per-image call (utils.py:475): decoded rows of one image -> kept rows
    for i in range(min(4, p.shape[0])):
        k = yu.non_max_suppression(dec[i], it, ct).cpu().numpy()
        assert np.array_equal(k, R[f"{name}_rows"][i, :R[f"{name}_count"][i]])


@pytest.mark.gpu
def test_cuda_reference_fixtures(dev):
    from yolohot import loss as yloss, utils as yu
    yp, yt = _cuda(R["demo_yp"], dev), _cuda(R["demo_yt"], dev)
    assert np.array_equal(yu.non_max_suppression(yu.decode_predictions(yp, 3, 2)[0]).cpu().numpy(), R["demo_nms_pred"])
    assert np.array_equal(yu.non_max_suppression(yu.decode_predictions(yt, 3, 2)[0]).cpu().numpy(), R["demo_nms_true"])
    ev = yu.MeanAveragePrecision(3, 2)
    ev.update_state(yt, yp)
    assert np.array_equal(ev.all_true_boxes_variable.cpu().numpy(), R["demo_true_rows"])
    assert np.array_equal(ev.all_pred_boxes_variable.cpu().numpy(), R["demo_pred_rows"])
    assert abs(float(ev.result()) - float(R["demo_map"])) <= 1e-6
    tot = float(yloss.YoloV1Loss(3, 2)(_cuda(R["loss_demo_yt"], dev), _cuda(R["loss_demo_yp"], dev)))
    assert abs(tot - float(R["loss_demo_total"])) <= 1e-5 * float(R["loss_demo_total"])


@pytest.mark.gpu
def test_cuda_loss_and_gradient_match_reference_source(dev):
    from yolohot import loss as yloss
    for key, C, B in (("loss16", 20, 2), ("loss_b3", 5, 3)):
        yp = _cuda(R[f"{key}_yp"], dev).requires_grad_(True)
        tot = yloss.YoloV1Loss(C, B)(_cuda(R[f"{key}_yt"], dev), yp)
        want = float(R[f"{key}_total"])
        assert abs(float(tot.detach()) - want) <= 1e-5 * abs(want), key
        if key == "loss16":
            tot.backward()
            np.testing.assert_allclose(yp.grad.cpu().numpy(), R["loss16_grad_torch"], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_cuda_map_matches_reference_source(dev):
    from yolohot import utils as yu
    ev = yu.MeanAveragePrecision(20, 2)
    ev.update_state(_cuda(R["map40_yt"][:25], dev), _cuda(R["map40_yp"][:25], dev))
    ev.update_state(_cuda(R["map40_yt"][25:], dev), _cuda(R["map40_yp"][25:], dev))
    assert np.array_equal(ev.all_true_boxes_variable.cpu().numpy(), R["map40_true_rows"])
    assert np.array_equal(ev.all_pred_boxes_variable.cpu().numpy(), R["map40_pred_rows"])
    assert abs(float(ev.result()) - float(R["map40_map"])) <= 1e-6
    for key, thr in (("rows_map", 0.5), ("rows_map_thr03", 0.3)):
        got = yu.mean_average_precision(_cuda(R["rows_true"], dev), _cuda(R["rows_pred"], dev), 4, thr)
        assert abs(float(got) - float(R[key])) <= 1e-6, (key, float(got), R[key])
    ev = yu.MeanAveragePrecision(20, 2)                                 # reset_states semantics (Q17)
    ev.update_state(_cuda(R["map40_yt"][:10], dev), _cuda(R["map40_yp"][:10], dev))
    ev.reset_states()
    ev.update_state(_cuda(R["map40_yt"][10:16], dev), _cuda(R["map40_yp"][10:16], dev))
    ev.update_state(_cuda(R["map40_yt"][16:20], dev), _cuda(R["map40_yp"][16:20], dev))
    assert np.array_equal(ev.all_true_boxes_variable.cpu().numpy(), R["reset_true_rows"])
    assert np.array_equal(ev.all_pred_boxes_variable.cpu().numpy(), R["reset_pred_rows"])
    assert abs(float(ev.result()) - float(R["reset_map"])) <= 1e-6
    out = yu.intersection_over_union(_cuda(R["iou4d_a"], dev), _cuda(R["iou4d_b"], dev)).cpu().numpy()
    assert out.shape == R["iou4d_out"].shape and np.array_equal(out, R["iou4d_out"])


@pytest.mark.gpu
def test_cuda_label_grids_match_reference_source(dev):
    from yolohot import dataset as yd
    for key, S, C, B in (("lab", 7, 20, 2), ("lab14", 14, 80, 3)):
        out = yd.encode_labels(R[f"{key}_boxes"], R[f"{key}_offsets"], S, C, B, device=dev)
        assert out.dtype == torch.float32 and np.array_equal(out.cpu().numpy(), R[f"{key}_out"]), key
    # list-of-lists form, the reference's own label file, and the generator-shaped class
    lists = [R["lab_boxes"][a:b] for a, b in zip(R["lab_offsets"][:-1], R["lab_offsets"][1:])]
    assert np.array_equal(yd.encode_labels(lists, None, 7, 20, 2, device=dev).cpu().numpy(), R["lab_out"])
    gen = yd.YoloV1Labels(3, 2)
    assert np.array_equal(gen._get_labels(R["lab_txt_boxes"]).cpu().numpy(), R["lab_txt_out"][0])
    with pytest.raises(IndexError):                       # dataset.py:107 raises for cx = 1.0
        yd.get_labels(np.array([[1.0, 0.5, 0.1, 0.1, 0]]), 7, 20, 2)
    # the labels feed the evaluator: decode + NMS of the encoded grid returns the boxes
    rows, cnt = __import__("yolohot.utils", fromlist=["x"]).decode_nms(gen.batch([R["lab_txt_boxes"]]), 3, 2)
    assert int(cnt[0]) == 3


@pytest.mark.gpu
def test_cuda_pixel_boxes_match_reference_source(dev):
    from yolohot import utils as yu
    rows = _cuda(R["px_rows"], dev)
    assert np.array_equal(yu.pixel_boxes(rows, 448, 448).cpu().numpy(), R["px_448"])
    assert np.array_equal(yu.pixel_boxes(rows, 500, 375).cpu().numpy(), R["px_375x500"])
    padded = torch.zeros((2, 46, 6), device=dev)
    padded[0] = rows
    padded[1, :5] = rows[:5]
    px = yu.pixel_boxes(padded, 448, 448, count=torch.tensor([46, 5])).cpu().numpy()
    assert np.array_equal(px[0], R["px_448"]) and np.array_equal(px[1, :5], R["px_448"][:5]) and (px[1, 5:] == -1).all()


@pytest.mark.gpu
def test_cuda_head_adapter(dev):
    """train.py:208: the flat Dense output is the same memory as (N,7,7,30); half heads widen exactly."""
    from yolohot import loss as yloss, utils as yu
    p = R["dense_in"]
    flat = _cuda(p.reshape(p.shape[0], -1), dev)
    rows, cnt = yu.decode_nms(flat, 20, 2)
    assert np.array_equal(cnt.cpu().numpy(), R["dense_count"])
    assert np.array_equal(_kept(rows.cpu().numpy(), cnt.cpu().numpy()), _kept(R["dense_rows"], R["dense_count"]))
    assert np.array_equal(yu.decode_predictions(flat, 20, 2).cpu().numpy(), R["dense_decode"])
    with pytest.raises(ValueError):
        yu.decode_nms(flat[:, :-1].contiguous(), 20, 2)
    from yolohot._tensor import head_to_f32
    for dt in (torch.float16, torch.bfloat16):
        h = _cuda(p, dev).to(dt)
        assert torch.equal(head_to_f32(h), h.float())                      # the standalone adapter kernel is exact
        want = yu.decode_nms(h.float(), 20, 2, return_index=True)
        got = yu.decode_nms(h, 20, 2, return_index=True)                   # fused: widened inside the kernel
        assert torch.equal(got[1], want[1])
        m = torch.arange(49, device=dev)[None, :] < want[1][:, None]
        assert torch.equal(got[0][m], want[0][m]) and torch.equal(got[2][m], want[2][m])
        assert torch.equal(yu.decode_predictions(h, 20, 2), yu.decode_predictions(h.float(), 20, 2))
    # half-precision heads through every fused kernel: tile ring + direct tail (VOC, odd count), cooperative
    # team kernel (S=14, B=3, C=80), generic shapes, unaligned view, score-mode extension
    from tests import fixtures as F
    for (gen, n, S, B, C, it, ct) in ((F.synth_dense, 1003, 7, 2, 20, 0.5, 0.4), (F.synth_stress, 37, 14, 3, 80, 0.5, 0.05),
                                      (F.synth_quantised, 50, 9, 2, 6, 0.5, 0.3), (F.synth_dense, 21, 13, 5, 7, 0.4, 0.3)):
        x = _cuda(gen(n, S, B, C), dev)
        for dt in (torch.float16, torch.bfloat16):
            h = x.to(dt)
            for mode in ("conf", "conf_x_prob"):
                want = yu.decode_nms(h.float(), C, B, it, ct, return_index=True, score_mode=mode)
                got = yu.decode_nms(h, C, B, it, ct, return_index=True, score_mode=mode)
                assert torch.equal(got[1], want[1]), (S, dt, mode)
                m = torch.arange(S * S, device=dev)[None, :] < want[1][:, None]
                assert torch.equal(got[0][m], want[0][m]) and torch.equal(got[2][m], want[2][m]), (S, dt, mode)
            odd = torch.cat([torch.zeros(1, dtype=dt, device=dev), h.reshape(-1)])[1:].reshape(h.shape)   # 2-byte aligned base
            got = yu.decode_nms(odd, C, B, it, ct)
            want = yu.decode_nms(h.float(), C, B, it, ct)
            assert torch.equal(got[1], want[1])
    yt = _cuda(R["loss16_yt"], dev)
    yp = _cuda(R["loss16_yp"].reshape(16, -1), dev).requires_grad_(True)
    tot = yloss.YoloV1Loss(20, 2)(yt, yp)
    tot.backward()
    assert abs(float(tot.detach()) - float(R["loss16_total"])) <= 1e-5 * abs(float(R["loss16_total"]))
    np.testing.assert_allclose(yp.grad.cpu().numpy().reshape(R["loss16_grad_torch"].shape), R["loss16_grad_torch"],
                               rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_cuda_stale_metric_evaluators_match_executed_source(dev):
    """Row a7 on the device: yolohot.metric / get_all_bboxes against metric.py + tmp.py executed on the stand-in."""
    from yolohot import metric as ym, utils as yu
    assert np.array_equal(yu.get_all_bboxes(_cuda(R["map40_yp"][:8], dev)).cpu().numpy(), R["stale_gab"])
    for cls, key in ((ym.MeanAveragePrecision2, "stale2"), (ym.MeanAveragePrecision, "stale1")):
        ev = cls()
        ev.update_state(_cuda(R["map40_yt"][:25], dev), _cuda(R["map40_yp"][:25], dev))
        ev.update_state(_cuda(R["map40_yt"][25:], dev), _cuda(R["map40_yp"][25:], dev))
        assert np.array_equal(ev.all_true_bboxes_variable.cpu().numpy(), R[f"{key}_true_rows"]), key
        assert np.array_equal(ev.all_pred_bboxes_variable.cpu().numpy(), R[f"{key}_pred_rows"]), key
        assert abs(float(ev.result()) - float(R[f"{key}_map"])) <= 1e-6, key
        got = yu.mean_average_precision_2(ev.all_true_bboxes_variable, ev.all_pred_bboxes_variable)     # tmp.py:440 signature
        assert abs(float(got) - float(R[f"{key}_map"])) <= 1e-6
    assert ev.count == 2


@pytest.mark.gpu
def test_cuda_signed_zero_and_non_finite_inputs(dev):
    """Quirk Q5.  -0.0 / inf in the SELECTED box, the class scores or the confidences: same rows as the reference's
    source.  A non-finite value in a box that was NOT selected: the source's 0 * inf poisons the row with NaN, the
    kernels select and return the selected box - the documented finite-input contract (include/yolohot.h); pinned here
    so that it cannot change silently."""
    from yolohot import utils as yu
    x = _cuda(R["nonfinite_sel_in"], dev)
    assert np.array_equal(yu.decode_predictions(x, 20, 2).cpu().numpy(), R["nonfinite_sel_decode"], equal_nan=True)
    rows, cnt = yu.decode_nms(x, 20, 2)
    assert np.array_equal(cnt.cpu().numpy(), R["nonfinite_sel_count"])
    assert np.array_equal(_kept(rows.cpu().numpy(), cnt.cpu().numpy()), _kept(R["nonfinite_sel_rows"], R["nonfinite_sel_count"]),
                          equal_nan=True)
    for i in range(6):                                                    # the per-image surface takes the same rows
        k = yu.non_max_suppression(_cuda(R["nonfinite_sel_decode"][i], dev)).cpu().numpy()
        assert np.array_equal(k, R["nonfinite_sel_rows"][i, :R["nonfinite_sel_count"][i]], equal_nan=True)
    los = yu.decode_predictions(_cuda(R["nonfinite_los_in"], dev), 20, 2).cpu().numpy()
    ref = R["nonfinite_los_decode"]
    bad = np.isnan(ref)
    assert bad.sum() == 3 and np.array_equal(los[~bad], ref[~bad])        # every other value agrees
    clean = R["nonfinite_los_in"].copy()
    clean[0, 0, 0, 26:30] = 0.0                                           # the same grid without the losing box's junk
    assert np.isfinite(los[0, 0]).all() and np.array_equal(los, yu.decode_predictions(_cuda(clean, dev), 20, 2).cpu().numpy())
