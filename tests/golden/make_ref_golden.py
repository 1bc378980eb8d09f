"""Regenerates tests/golden/ref_golden.npz: what the REFERENCE'S OWN SOURCE returns.

Runs only in the build container (it reads /root/reference, which the GPU box does not have).
`yolo_v1/utils.py` and `yolo_v1/loss.py` are imported unmodified from /root/reference with
`tests/golden/tfshim` first on sys.path, i.e. every line of the reference executes, on top of a
NumPy stand-in for the TensorFlow primitives it calls (TensorFlow itself is absent from this image
and its wheelhouse; see the stand-in's docstring for the exact semantics restated and
DESIGN.md section 5 for what that does and does not pin).  Inputs are stored next to the outputs so
the tests do not depend on NumPy's generator streams.

    python tests/golden/make_ref_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("YH_REFERENCE", "/root/reference/yolo_v1")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "tfshim"))
warnings.simplefilter("ignore")          # np.trapz deprecation (utils.py:444)

import tensorflow as tf  # noqa: E402  (the stand-in)
import torch  # noqa: E402

assert "numpy-standin" in tf.__version__
import types  # noqa: E402

# dataset.py:7-9 imports its augmentation libraries at module level; none is used by the label
# code that is executed here (_get_boxes / _get_labels), so empty modules stand in for them
for _name in ("imgaug", "imgaug.augmenters", "albumentations"):
    sys.modules.setdefault(_name, types.ModuleType(_name))
sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]

import dataset as RD  # noqa: E402  /root/reference/yolo_v1/dataset.py
import loss as RL  # noqa: E402   /root/reference/yolo_v1/loss.py
import utils as RU  # noqa: E402  /root/reference/yolo_v1/utils.py

from tests import fixtures as F  # noqa: E402

assert os.path.realpath(RU.__file__).startswith(os.path.realpath(REF)), RU.__file__
F32 = np.float32


def T(a):
    return tf.cast(np.asarray(a), tf.float32)


def nms_batch(decoded, iou_thr=0.5, conf_thr=0.4):
    """Reference per-image loop (utils.py:474-480) -> padded rows + counts."""
    n, m = decoded.shape[0], decoded.shape[1]
    rows = np.zeros((n, m, 6), F32)
    cnt = np.zeros((n,), np.int32)
    for i in range(n):
        k = np.asarray(RU.non_max_suppression(decoded[i], iou_threshold=iou_thr, conf_threshold=conf_thr))
        k = k.reshape(-1, 6)
        rows[i, :len(k)] = k
        cnt[i] = len(k)
    return rows, cnt


def evaluator(batches, C, B):
    ev = RU.MeanAveragePrecision(C, B)
    for yt, yp in batches:
        ev.update_state(T(yt), T(yp))
    m = np.float32(np.asarray(ev.result()))
    return m, np.asarray(ev.all_true_boxes_variable), np.asarray(ev.all_pred_boxes_variable)


def label_generator(grid, C, B):
    """A YoloV1Generator (dataset.py:19) without its image directory: only what _get_labels reads."""
    gen = RD.YoloV1Generator.__new__(RD.YoloV1Generator)
    gen.grid, gen.num_classes, gen.num_boxes = grid, C, B
    gen.output_shape = (grid, grid, C + B * 5)
    return gen


def labels_of(gen, box_lists):
    """dataset.py:72-86: batch_labels is float32, each image assigned from _get_labels' float64 grid."""
    out = np.zeros((len(box_lists),) + gen.output_shape, dtype=F32)
    for i, b in enumerate(box_lists):
        out[i] = gen._get_labels(b)
    return out


def flat(box_lists):
    offs = np.zeros(len(box_lists) + 1, np.int64)
    offs[1:] = np.cumsum([len(b) for b in box_lists])
    rows = [np.asarray(b, np.float64).reshape(-1, 5) for b in box_lists]
    return (np.concatenate(rows, 0) if rows else np.zeros((0, 5))), offs


def recorded_pixel_boxes(fn, img_shape, boxes, names_path):
    """Runs utils.get_tagged_img / get_grid_tagged_img and records the corners handed to cv2.rectangle."""
    rec = []
    real = RU.cv2.rectangle

    def spy(img, p0, p1, **kw):
        rec.append([p0[0], p0[1], p1[0], p1[1]])
        return real(img, p0, p1, **kw)
    RU.cv2.rectangle = spy
    try:
        fn(np.zeros(img_shape, np.uint8), T(boxes), names_path)
    finally:
        RU.cv2.rectangle = real
    return np.array(rec, np.int32).reshape(-1, 4)


# ---- the stale evaluators ------------------------------------------------------------------
# metric.py:7 imports names that utils.py no longer defines; their bodies survive only as COMMENTED-OUT source in
# tmp.py.  The text of those comment blocks is executed here as it stands (leading "# " stripped), on top of the
# reference's own utils.py, and metric.py is imported unmodified against the result:
#   get_all_bboxes              tmp.py:96-157
#   mean_average_precision_2    tmp.py:440-595           (-> metric.MeanAveragePrecision2.result, metric.py:96-99)
#   mean_average_precision      tmp.py:187-405, the older (true, pred, iou_threshold, num_classes) signature that
#                               metric.MeanAveragePrecision.result (metric.py:55-56) calls with two arguments
#   non_max_suppression_2       NO body survives anywhere; metric.py:77 calls it with the keywords of the live
#                               utils.non_max_suppression, which is bound to the name
#   non_max_suppression(.., threshold=)   metric.py:30 uses the keyword of tmp.py:8-89, whose sort is a
#                               TensorArray.scatter by argsort (the inverse permutation - a bug, presumably why it was
#                               replaced); the live function is bound behind that keyword instead
def _tmp_block(lines, name):
    """The commented-out top-level function `name` of tmp.py, decorator included, as source text."""
    start = next(i for i, l in enumerate(lines) if l.startswith(f"# def {name}("))
    if lines[start - 1].startswith("# @tf.function"):
        start -= 1
    end = start + 2
    while end < len(lines) and not (lines[end].startswith("# @tf.function") or lines[end].startswith("# def ")
                                    or lines[end].startswith("# class ")):
        end += 1
    body = []
    for l in lines[start:end]:
        body.append(l[2:] if l.startswith("# ") else (l[1:] if l.startswith("#") else l))
    return start + 1, end, "\n".join(body) + "\n"


def stale_metric_module(variant):
    """metric.py imported against utils.py + the uncommented tmp.py blocks.  variant 2: MeanAveragePrecision2's
    imports; variant 1: MeanAveragePrecision's (older mean_average_precision, `threshold=` keyword)."""
    import importlib.util
    lines = open(os.path.join(REF, "tmp.py")).read().split("\n")
    u = types.ModuleType("utils")
    u.__file__ = os.path.join(REF, "utils.py")
    exec(compile(open(u.__file__).read(), u.__file__, "exec"), u.__dict__)
    live_nms, live_map = u.non_max_suppression, u.mean_average_precision
    used = {}
    for name in ("get_all_bboxes", "mean_average_precision_2") + (("mean_average_precision",) if variant == 1 else ()):
        lo, hi, src = _tmp_block(lines, name)
        used[name] = (lo, hi)
        exec(compile(src, os.path.join(REF, "tmp.py") + f":{lo}-{hi}", "exec"), u.__dict__)
    u.non_max_suppression_2 = live_nms
    if variant == 1:
        u.non_max_suppression = lambda boxes, iou_threshold=0.5, threshold=0.4: live_nms(boxes, iou_threshold, threshold)
    else:
        u.mean_average_precision = live_map
    saved = sys.modules.get("utils")
    sys.modules["utils"] = u
    try:
        spec = importlib.util.spec_from_file_location(f"metric_v{variant}", os.path.join(REF, "metric.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.modules["utils"] = saved
    return m, u, used



def main():
    g = {}
    # ---- the reference's own fixtures ------------------------------------------------------
    yt, yp = F.utils_demo()                                            # utils.py:717-769
    g["demo_yt"], g["demo_yp"] = yt, yp
    dec = RU.decode_predictions(T(yp), 3, 2)
    g["demo_decode_pred"] = np.asarray(dec)
    g["demo_nms_pred"] = np.asarray(RU.non_max_suppression(dec[0]))
    g["demo_nms_true"] = np.asarray(RU.non_max_suppression(RU.decode_predictions(T(yt), 3, 2)[0]))
    g["demo_map"], g["demo_true_rows"], g["demo_pred_rows"] = evaluator([(yt, yp)], 3, 2)
    lt, lp = F.loss_demo()                                             # loss.py:219-237
    g["loss_demo_yt"], g["loss_demo_yp"] = lt, lp
    g["loss_demo_total"] = np.float32(np.asarray(RL.YoloV1Loss(num_classes=3, num_boxes=2)(T(lt), T(lp))))

    # ---- IoU: random pairs, degenerate and negative extents --------------------------------
    rng = np.random.Generator(np.random.PCG64(42))
    a = rng.random((4096, 4), dtype=F32)
    b = rng.random((4096, 4), dtype=F32)
    a[:64, 2:] *= F32(-1)                      # negative w,h: |area| path (utils.py:40-41)
    b[64:128] = a[64:128]                      # identical boxes
    a[128:192, 2:] = 0                         # zero area
    b[192:256, :2] += F32(3)                   # extents that clip at 1 / disjoint boxes
    g["iou_a"], g["iou_b"] = a, b
    g["iou_out"] = np.asarray(RU.intersection_over_union(T(a), T(b)))

    # ---- decode + NMS (S=7: the reference hard-codes 7, utils.py:184,200-216) --------------
    for name, p, thr in (("dense", F.synth_dense(12, seed=1234), (0.5, 0.4)),
                         ("sparse", F.synth_sparse(48, seed=2025), (0.5, 0.4)),
                         ("quant", F.synth_quantised(12, 7, 2, 20, seed=5), (0.5, 0.4)),
                         ("quant_lo", F.synth_quantised(6, 7, 2, 20, seed=6), (0.25, 0.05)),
                         ("b3c5", F.synth_dense(6, 7, 3, 5, seed=77), (0.3, 0.2))):
        B = 3 if name == "b3c5" else 2
        C = 5 if name == "b3c5" else 20
        if name == "sparse":
            p[5] = 0                            # an image with no candidate at all
            p[6, ..., C] = F32(0.4)             # conf == threshold exactly: strict > (utils.py:95)
            p[6, ..., C + 5] = F32(0.1)
        dec = np.asarray(RU.decode_predictions(T(p), C, B))
        rows, cnt = nms_batch(tf.cast(dec, tf.float32), *thr)
        g[f"{name}_in"], g[f"{name}_decode"], g[f"{name}_rows"], g[f"{name}_count"] = p, dec, rows, cnt
        g[f"{name}_thr"] = np.array(thr, F32)

    # ---- loss (+ gradient of the same source under torch autograd, off ties) ---------------
    yt = F.synth_labels(16, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    g["loss16_yt"], g["loss16_yp"] = yt, yp
    g["loss16_total"] = np.float32(np.asarray(RL.YoloV1Loss(20, 2)(T(yt), T(yp))))
    tp = torch.from_numpy(yp.copy()).requires_grad_(True)
    tot = RL.YoloV1Loss(20, 2)(torch.from_numpy(yt), tp)
    tot.backward()
    g["loss16_total_torch"] = np.float32(tot.item())
    g["loss16_grad_torch"] = tp.grad.numpy().astype(F32)
    yt = F.synth_labels(4, 7, 3, 5, seed=3)
    yp = F.synth_loss_pred(yt.shape, seed=3)
    g["loss_b3_yt"], g["loss_b3_yp"] = yt, yp
    g["loss_b3_total"] = np.float32(np.asarray(RL.YoloV1Loss(5, 3)(T(yt), T(yp))))

    # ---- evaluator + mAP -------------------------------------------------------------------
    yt = F.synth_labels(40, seed=11)
    yp = F.synth_map_pred(yt)
    g["map40_yt"], g["map40_yp"] = yt, yp
    g["map40_map"], g["map40_true_rows"], g["map40_pred_rows"] = evaluator([(yt[:25], yp[:25]), (yt[25:], yp[25:])], 20, 2)
    # reset_states() only rewinds the image counter; the next update overwrites the buffers (utils.py:467-468, 484-486)
    ev = RU.MeanAveragePrecision(20, 2)
    ev.update_state(T(yt[:10]), T(yp[:10]))
    ev.reset_states()
    ev.update_state(T(yt[10:16]), T(yp[10:16]))
    ev.update_state(T(yt[16:20]), T(yp[16:20]))
    g["reset_map"] = np.float32(np.asarray(ev.result()))
    g["reset_true_rows"], g["reset_pred_rows"] = np.asarray(ev.all_true_boxes_variable), np.asarray(ev.all_pred_boxes_variable)
    # IoU on the 4-D shapes the loss feeds it (loss.py:128-131)
    b1 = rng.random((3, 7, 7, 4), dtype=F32)
    b2 = rng.random((3, 7, 7, 4), dtype=F32)
    g["iou4d_a"], g["iou4d_b"] = b1, b2
    g["iou4d_out"] = np.asarray(RU.intersection_over_union(T(b1), T(b2)))
    # direct call on hand-made rows: class without GT, class without detections, two detections
    # fighting for one GT, equal confidences (stable order), detection in an image without GT
    true_rows = np.array([[0, 0, 1, .5, .5, .2, .2], [0, 0, 1, .2, .2, .1, .1], [1, 0, 1, .5, .5, .2, .2],
                          [1, 2, 1, .3, .3, .2, .2], [2, 2, 1, .6, .6, .3, .3]], F32)
    pred_rows = np.array([[0, 0, .9, .5, .5, .2, .2], [0, 0, .9, .51, .5, .2, .2], [0, 0, .8, .2, .2, .1, .1],
                          [1, 0, .7, .9, .9, .1, .1], [3, 0, .95, .5, .5, .2, .2], [1, 2, .6, .3, .3, .2, .2],
                          [2, 2, .6, .6, .6, .3, .3], [2, 2, .6, .6, .61, .3, .3], [0, 1, .99, .5, .5, .2, .2]], F32)
    g["rows_true"], g["rows_pred"] = true_rows, pred_rows
    g["rows_map"] = np.float32(np.asarray(RU.mean_average_precision(T(true_rows), T(pred_rows), 4)))
    g["rows_map_thr03"] = np.float32(np.asarray(RU.mean_average_precision(T(true_rows), T(pred_rows), 4, iou_threshold=0.3)))

    # ---- a7: the stale metric.py evaluators, executed from metric.py + the commented-out text of tmp.py -----------
    yt, yp = g["map40_yt"], g["map40_yp"]
    m2, u2, used = stale_metric_module(2)
    assert used["get_all_bboxes"][0] == 96 and used["mean_average_precision_2"][0] == 440, used      # decorator lines
    g["stale_gab"] = np.asarray(u2.get_all_bboxes(T(yp[:8])))                              # tmp.py:96-157 on 8 images
    ev = m2.MeanAveragePrecision2()
    ev.update_state(T(yt[:25]), T(yp[:25]))
    ev.update_state(T(yt[25:]), T(yp[25:]))
    g["stale2_true_rows"], g["stale2_pred_rows"] = np.asarray(ev.all_true_bboxes_variable), np.asarray(ev.all_pred_bboxes_variable)
    g["stale2_map"] = np.float32(np.asarray(ev.result()))                                  # tmp.py:440-595
    m1, u1, used1 = stale_metric_module(1)
    assert used1["mean_average_precision"][0] == 187, used1
    ev = m1.MeanAveragePrecision()
    ev.update_state(T(yt[:25]), T(yp[:25]))
    ev.update_state(T(yt[25:]), T(yp[25:]))
    g["stale1_true_rows"], g["stale1_pred_rows"] = np.asarray(ev.all_true_bboxes_variable), np.asarray(ev.all_pred_bboxes_variable)
    g["stale1_map"] = np.float32(np.asarray(ev.result()))                                  # tmp.py:187-405
    g["stale1_count"] = np.float32(np.asarray(ev.count))

    # ---- Q5: signed zeros and non-finite values through decode (utils.py:184-197 multiplies the boxes that were NOT
    # selected by 0 and sums) and NMS.  What the source does with them, on the stand-in's IEEE arithmetic:
    #   sel: the SELECTED box / the class scores / the confidences carry -0.0, inf or NaN -> pinned below (the CUDA
    #        path reproduces these: it computes on the same values);
    #   los: a NON-selected box carries inf / NaN -> the source's 0 * inf = NaN poisons the row; the CUDA path selects
    #        instead of multiplying and returns the selected box.  Recorded so that the deviation is pinned, and
    #        documented as "finite inputs only" in include/yolohot.h / INTEGRATION.md.
    q = F.synth_dense(6, seed=77).astype(F32)
    sel = q.copy()
    sel[0, 0, 0, 20:25] = [0.9, -0.0, -0.0, 0.3, 0.3]; sel[0, 0, 0, 25] = 0.1        # selected box: -0.0 coordinates
    sel[1, 1, 1, 20:25] = [0.9, 0.5, 0.5, np.inf, 0.3]; sel[1, 1, 1, 25] = 0.1       # selected box: infinite width
    sel[2, 2, 2, 20] = np.inf; sel[2, 2, 2, 25] = 0.5                                # infinite confidence wins
    sel[3, 3, 3, 0:20] = -0.0                                                        # all class scores -0.0 -> class 0
    sel[4, 4, 4, 5] = np.inf                                                         # infinite class score -> class 5
    sel[5, :, :, 20] = -0.0; sel[5, :, :, 25] = 0.0                                  # +-0 confidences tie -> box 0, never kept
    g["nonfinite_sel_in"] = sel
    with np.errstate(all="ignore"):
        d = np.asarray(RU.decode_predictions(T(sel), 20, 2))
        g["nonfinite_sel_decode"] = d
        g["nonfinite_sel_rows"], g["nonfinite_sel_count"] = nms_batch(d)
        los = q.copy()
        los[0, 0, 0, 20] = 0.9; los[0, 0, 0, 25] = 0.1; los[0, 0, 0, 26:30] = [np.inf, np.nan, -np.inf, 1.0]   # losing box
        g["nonfinite_los_in"] = los
        g["nonfinite_los_decode"] = np.asarray(RU.decode_predictions(T(los), 20, 2))

    # ---- N2: label grids (dataset.py:88-123) ------------------------------------------------
    gen = label_generator(7, 3, 2)
    txt = gen._get_boxes(os.path.join(REF, "data", "test.txt"))          # the reference's own label file
    g["lab_txt_boxes"] = np.asarray(txt, np.float64)
    g["lab_txt_out"] = labels_of(gen, [txt])
    rng = np.random.Generator(np.random.PCG64(21))
    lists = []
    for i in range(24):
        k = int(rng.integers(0, 12)) if i != 3 else 70               # an empty image, and one with > 64 boxes
        b = np.concatenate([rng.random((k, 2)), 0.05 + 0.9 * rng.random((k, 2)), rng.integers(0, 20, (k, 1))], 1)
        if k >= 4:
            b[1, :2] = b[0, :2] + 0.001                                # two boxes in one cell: first one wins (:107)
            b[2, 0], b[2, 1] = 3.0 / 7.0, 5.0 / 7.0                    # centres on cell boundaries
            b[3, 0], b[3, 1] = -0.05, 0.999999                         # int() truncates toward zero; x < 0
        if i == 5 and k:
            b[0, 4] = -1                                               # Python index wrap: channel D-1
            b[0, 0] = -0.2                                             # cell index -1 -> last column
        lists.append(b)
    gen = label_generator(7, 20, 2)
    g["lab_boxes"], g["lab_offsets"] = flat(lists)
    g["lab_out"] = labels_of(gen, lists)
    lists3 = [np.concatenate([rng.random((6, 2)), 0.1 + 0.5 * rng.random((6, 2)), rng.integers(0, 80, (6, 1))], 1)
              for _ in range(5)]
    gen = label_generator(14, 80, 3)
    g["lab14_boxes"], g["lab14_offsets"] = flat(lists3)
    g["lab14_out"] = labels_of(gen, lists3)
    try:                                                                 # cx = 1.0 -> cell index 7
        label_generator(7, 20, 2)._get_labels(np.array([[1.0, 0.5, 0.1, 0.1, 0]]))
        g["lab_out_of_grid_raises"] = np.array(0)
    except IndexError:
        g["lab_out_of_grid_raises"] = np.array(1)

    # ---- N4: pixel corners handed to cv2.rectangle (utils.py:645-657, 688-700) ---------------
    names = os.path.join(REF, "data", "test.names")
    rows = np.concatenate([g["demo_nms_pred"], g["demo_nms_true"]], 0)
    extra = rng.random((40, 6), dtype=F32)
    extra[:, 0] = rng.integers(0, 3, 40)
    extra[:, 4:] *= F32(1.5)                                            # some corners fall outside the image
    rows = np.concatenate([rows, extra], 0).astype(F32)
    g["px_rows"] = rows
    g["px_448"] = recorded_pixel_boxes(RU.get_tagged_img, (448, 448, 3), rows, names)
    g["px_375x500"] = recorded_pixel_boxes(RU.get_grid_tagged_img, (375, 500, 3), rows, names)

    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_golden.npz")
    np.savez_compressed(out, **g)
    for k, v in g.items():
        print(f"{k:22s} {str(np.asarray(v).shape):18s} {np.asarray(v).dtype}")


if __name__ == "__main__":
    main()
