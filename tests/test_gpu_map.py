"""GPU tests of the hand-written mAP stage (csrc/yh_map.cu, yh_map_reduce.cu): the chained-scan append, the fused
per-image matching, the general matching path on arbitrary rows, the persistent radix-sort + AP kernel - each
against an independent restatement (oracle / NumPy / torch.sort), at sizes from empty to several million records."""
import numpy as np
import pytest
import torch

from oracle import yolo_oracle as O
from tests import fixtures as F

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _orderable(conf):
    u = (conf.astype(F32) + F32(0)).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, ~u & 0xffffffff, u | 0x80000000).astype(np.uint64)


def _pack(cls, conf, tp):
    return ((cls.astype(np.uint64) << np.uint64(33)) | ((~_orderable(conf) & np.uint64(0xffffffff)) << np.uint64(1)) | tp.astype(np.uint64))


def _ap_reference(cls, conf, tp, gt, C):
    """utils.py:364-456 on (class, conf, tp) triples in row order: stable sort, cumsum, float32 points, np.trapz."""
    aps = np.zeros(C, F32)
    for c in range(C):
        if gt[c] == 0:
            continue
        sel = np.nonzero(cls == c)[0]
        order = sel[np.argsort(-conf[sel].astype(np.float64), kind="stable")]
        t = tp[order].astype(F32)
        tpc = np.cumsum(t, dtype=F32)
        fpc = np.cumsum(F32(1) - t, dtype=F32)
        rec = tpc / (F32(gt[c]) + F32(1e-6))
        prec = tpc / (tpc + fpc + F32(1e-6))
        prec = np.concatenate([np.ones(1, F32), prec]).astype(F32)
        rec = np.concatenate([np.zeros(1, F32), rec]).astype(F32)
        aps[c] = np.trapz(prec, rec)
    return aps, F32(np.mean(aps.astype(np.float64)))


@pytest.mark.parametrize("n,C,quant", [(0, 5, 0), (1, 1, 0), (777, 3, 8), (4096, 20, 0), (4097, 20, 4), (100_000, 20, 64),
                                       (300_000, 300, 0), (2_500_000, 20, 1 << 16), (50_000, 4096, 0)])
def test_radix_reduce_vs_numpy(dev, n, C, quant):
    """Random records (many equal confidences when `quant`): AP per class and mAP against NumPy's stable sort."""
    from yolohot import utils as yu
    rng = np.random.default_rng(n + C)
    cls = rng.integers(0, C + 1, n)                                   # class C = rows the reference never selects
    if C >= 20:
        cls[rng.random(n) < 0.5] = 3                                  # one dominant class: long same-digit runs
    conf = rng.random(n).astype(F32)
    if quant:
        conf = (np.floor(conf * quant) / quant).astype(F32)
    conf[rng.random(n) < 0.01] = 0.0
    tp = (rng.random(n) < 0.3).astype(np.uint8)
    tp[cls == C] = 0
    gt = rng.integers(0, 50, C).astype(np.int32) + np.bincount(cls[tp == 1], minlength=C + 1)[:C].astype(np.int32)
    gt[rng.random(C) < 0.2] = 0
    rec = torch.from_numpy(_pack(cls, conf, tp).view(np.int64)).to(dev)
    want_ap, want_m = _ap_reference(cls, conf, tp, gt, C)
    m, ap = yu.map_reduce(rec, torch.from_numpy(gt).to(dev), C)
    np.testing.assert_allclose(ap.cpu().numpy(), want_ap, rtol=0, atol=2e-6)
    assert abs(float(m) - float(want_m)) <= 1e-6
    # device-side count (a bound larger than the count), caller workspace, run-to-run identical
    cnt = torch.tensor([n], dtype=torch.int64, device=dev)
    pad = torch.cat([rec, torch.full((1000,), -1, dtype=torch.int64, device=dev)])
    from yolohot import _lib
    ws = torch.empty(int(_lib.lib().yh_workspace_bytes(_lib.YH_OP_MAP_REDUCE, n + 1000, 0, 0, C)), dtype=torch.uint8, device=dev)
    m2, ap2 = yu.map_reduce(pad, torch.from_numpy(gt).to(dev), C, nrec_dev=cnt, workspace=ws)
    assert torch.equal(ap2, ap) and torch.equal(m2, m)


@pytest.mark.parametrize("n,C,quant,skew", [(0, 3, 0, 0.0), (1, 1, 0, 0.0), (5000, 20, 0, 0.0), (73_595, 20, 0, 0.0), (73_595, 20, 16, 0.5),
                                            (200_000, 80, 0, 0.0), (150_000, 512, 3, 0.0), (400_000, 20, 0, 0.9), (30_000, 1, 0, 0.0)])
def test_counting_path_equals_radix_path(dev, n, C, quant, skew, monkeypatch):
    """The sort-free AP (ranks counted inside (class, confidence bucket) groups) against the radix-sort path on the same
    records: AP per class and mAP bit for bit - for continuous and tie-heavy confidences, confidences outside [0, 1],
    infinities and NaNs (which have a place in the sort order), a dominant class, one bucket per class, and the fall-back
    to the radix passes when the pair count known after the histogram is over the limit."""
    from yolohot import utils as yu
    rng = np.random.default_rng(7 * n + C)
    cls = rng.integers(0, C + 1, n)
    if skew:
        cls[rng.random(n) < skew] = C // 2
    conf = rng.random(n).astype(F32)
    if quant:
        conf = (np.floor(conf * quant) / quant).astype(F32)
    odd = rng.random(n)
    conf[odd < 0.01] = 0.0
    conf[(odd >= 0.01) & (odd < 0.02)] = 1.0
    conf[(odd >= 0.02) & (odd < 0.025)] = -0.25
    conf[(odd >= 0.025) & (odd < 0.03)] = 1.75
    conf[(odd >= 0.03) & (odd < 0.032)] = np.inf
    conf[(odd >= 0.032) & (odd < 0.034)] = -np.inf
    conf[(odd >= 0.034) & (odd < 0.036)] = np.nan
    conf[(odd >= 0.036) & (odd < 0.038)] = -np.nan
    tp = (rng.random(n) < 0.35).astype(np.uint8)
    tp[cls == C] = 0
    gt = rng.integers(0, 50, C).astype(np.int32) + np.bincount(cls[tp == 1], minlength=C + 1)[:C].astype(np.int32)
    if C > 2:
        gt[rng.random(C) < 0.1] = 0
    rec = torch.from_numpy(_pack(cls, conf, tp).view(np.int64)).to(dev)
    gtd = torch.from_numpy(gt).to(dev)

    def run(**env):
        for k in ("YH_MAP_COUNT", "YH_MAP_COUNT_PAIRS", "YH_MAP_COUNT_BUCKETS", "YH_MAP_COUNT_NMAX"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, str(v))
        m, ap = yu.map_reduce(rec, gtd, C)
        torch.cuda.synchronize()
        return m.clone(), ap.clone()

    m0, ap0 = run(YH_MAP_COUNT=0)                                     # radix passes only
    for env in ({}, {"YH_MAP_COUNT_BUCKETS": 1}, {"YH_MAP_COUNT_BUCKETS": 4}, {"YH_MAP_COUNT_PAIRS": 1},
                {"YH_MAP_COUNT_PAIRS": 10**15, "YH_MAP_COUNT_NMAX": 10**9}):
        m1, ap1 = run(**env)
        assert torch.equal(ap1.view(torch.int32), ap0.view(torch.int32)), env
        assert torch.equal(m1.view(torch.int32), m0.view(torch.int32)), env


@pytest.mark.parametrize("S,B,C,n1,n2", [(7, 2, 20, 300, 77), (5, 1, 3, 64, 9), (8, 3, 80, 130, 1), (3, 2, 20, 1000, 555)])
def test_fused_update_state_equals_three_launches(dev, monkeypatch, S, B, C, n1, n2):
    """update_state as ONE kernel (yh_eval_update_state: decode + NMS of both tensors, matching, append) against the
    three-launch path (yh_decode_nms x 2 + yh_eval_update) over two batches: row buffers, records, cursors, per-class
    ground-truth counts and the mAP, bit for bit - specialised VOC shape, generic shapes, a 64-cell grid."""
    from yolohot import utils as yu
    out = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("YH_EVAL_FUSED", fused)
        ev = yu.MeanAveragePrecision(C, B)
        for k, (n, seed) in enumerate(((n1, 3), (n2, 4))):
            yt = F.synth_labels(n, S, B, C, seed=seed, lam=2.5)
            yp = F.synth_map_pred(yt, B, C, seed=seed)
            ev.update_state(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev))
        m = ev.result()
        torch.cuda.synchronize()
        st = ev._st
        npred, ntrue = int(st["cursors"][0]), int(st["cursors"][1])
        out[fused] = (npred, ntrue, st["pred"][:npred].clone(), st["true"][:ntrue].clone(), st["rec"][:npred].clone(),
                      st["gt"].clone(), m.clone(), ev.last_ap.clone())
    monkeypatch.delenv("YH_EVAL_FUSED")
    a, b = out["0"], out["1"]
    assert a[0] == b[0] and a[1] == b[1] and a[0] > 0 and a[1] > 0
    for x, y in zip(a[2:], b[2:]):
        assert torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x, y.view(torch.int32) if y.dtype == torch.float32 else y)


def test_rows_append_chained_scan(dev):
    """yh_rows_append / yh_eval_update row buffers against a NumPy compaction, from one tile to a thousand."""
    from yolohot import _lib
    from yolohot._tensor import stream_ptr
    L = _lib.lib()
    for n, M in ((1, 49), (7, 49), (5000, 49), (70_000, 49), (300_000, 16), (1_100_000, 4), (300, 196)):
        g = torch.Generator(device=dev).manual_seed(n)
        boxes = torch.rand((n, M, 6), generator=g, device=dev)
        cnt = torch.randint(0, M + 1, (n,), generator=g, device=dev, dtype=torch.int32)
        if n > 10:
            cnt[3:9] = 0
        total = int(cnt.sum())
        out = torch.full((total + 5 + 3, 7), -7.0, device=dev)
        cur = torch.tensor([5], dtype=torch.int64, device=dev)               # a non-zero cursor: rows land after it
        for rep in range(2):                                                 # the scan state must come back clean
            cur.fill_(5)
            _lib.check(L.yh_rows_append(boxes.data_ptr(), cnt.data_ptr(), n, M, 1000, out.data_ptr(), out.shape[0], cur.data_ptr(),
                                        stream_ptr(dev)), "rows_append")
            assert int(cur) == 5 + total
            mask = torch.arange(M, device=dev)[None, :] < cnt[:, None]
            want = torch.cat([(torch.arange(n, device=dev) + 1000).float()[:, None, None].expand(n, M, 1), boxes], dim=2)[mask]
            assert torch.equal(out[5:5 + total], want), (n, M, rep)
            assert bool((out[:5] == -7).all()) and bool((out[5 + total:] == -7).all())
        # capacity smaller than the batch: rows past it are dropped, the cursor still advances
        small = torch.full((max(total // 2, 1), 7), -7.0, device=dev)
        cur.zero_()
        _lib.check(L.yh_rows_append(boxes.data_ptr(), cnt.data_ptr(), n, M, 1000, small.data_ptr(), small.shape[0], cur.data_ptr(),
                                    stream_ptr(dev)), "rows_append")
        assert int(cur) == total and torch.equal(small[:min(total, small.shape[0])], want[:small.shape[0]])


def _rows_case(n_img, C, seed, ties=False):
    rng = np.random.default_rng(seed)
    t_rows, p_rows = [], []
    for img in range(n_img):
        for _ in range(rng.integers(0, 5)):
            c = rng.integers(0, C)
            box = [rng.random(), rng.random(), 0.2 + 0.3 * rng.random(), 0.2 + 0.3 * rng.random()]
            t_rows.append([img, c, 1.0] + box)
            for _ in range(rng.integers(0, 4)):
                j = 0.03 * rng.standard_normal(4)
                conf = rng.integers(1, 9) / 8.0 if ties else rng.random()
                p_rows.append([img, c if rng.random() < 0.8 else (c + 1) % C, conf] + list(np.array(box) + j))
        for _ in range(rng.integers(0, 3)):                                   # detections on background
            p_rows.append([img, rng.integers(0, C), rng.random() * 0.6, rng.random(), rng.random(), 0.3, 0.3])
    return np.array(t_rows, F32).reshape(-1, 7), np.array(p_rows, F32).reshape(-1, 7)


def test_general_match_unsorted_rows(dev):
    """mean_average_precision on rows in ANY order (radix sort of the ground truths by image, claims by atomicMin)
    against the oracle's literal loop; and against itself with the images' blocks permuted."""
    from yolohot import utils as yu
    for n_img, C, ties in ((1, 3, False), (60, 3, True), (700, 20, False), (4000, 7, True)):
        t_rows, p_rows = _rows_case(n_img, C, n_img, ties)
        want = float(O.mean_average_precision(t_rows, p_rows, C))
        got = float(yu.mean_average_precision(torch.from_numpy(t_rows).to(dev), torch.from_numpy(p_rows).to(dev), C))
        assert abs(got - want) <= 1e-6, (n_img, got, want)
        rng = np.random.default_rng(1)
        # ground-truth rows shuffled as whole rows: the order inside an image only matters for exact IoU ties;
        # prediction rows keep their order (equal confidences resolve by row order, utils.py:367)
        tp_ = t_rows[rng.permutation(t_rows.shape[0])] if not ties else t_rows[np.argsort(-t_rows[:, 0], kind="stable")]
        got2 = float(yu.mean_average_precision(torch.from_numpy(np.ascontiguousarray(tp_)).to(dev), torch.from_numpy(p_rows).to(dev), C))
        assert abs(got2 - want) <= 1e-6, (n_img, got2, want)
        # rows the reference never selects: class out of range / not an integer
        extra = np.array([[0, C, 0.9, .5, .5, .2, .2], [0, -1, 0.9, .5, .5, .2, .2], [0, 0.5, 0.9, .5, .5, .2, .2]], F32)
        got3 = float(yu.mean_average_precision(torch.from_numpy(np.concatenate([t_rows, extra])).to(dev),
                                               torch.from_numpy(np.concatenate([extra, p_rows])).to(dev), C))
        assert abs(got3 - want) <= 1e-6
    rec, gt = yu.map_match(torch.from_numpy(t_rows).to(dev), torch.from_numpy(p_rows).to(dev), C, 0.5, rows_by_image=True)
    rec2, gt2 = yu.map_match(torch.from_numpy(t_rows).to(dev), torch.from_numpy(p_rows).to(dev), C, 0.5, rows_by_image=False)
    assert torch.equal(rec, rec2) and torch.equal(gt, gt2)                    # both ways of finding an image's ground truths


def test_evaluator_records_equal_general_path(dev):
    """The records the evaluator keeps while updating (fused per-image matching) == yh_map_match on its row buffers,
    for VOC-like and dense cfg5-like batches, streamed in ragged batches; result() == oracle."""
    from yolohot import utils as yu
    for (n, S, B, C, seed) in ((1500, 7, 2, 20, 11), (200, 14, 3, 80, 5)):
        if S == 7:
            yt = F.synth_labels(n, seed=seed)
            yp = F.synth_map_pred(yt)
        else:                                                                # many boxes per image, > 32 per chunk
            yt = (F.synth_stress(n, S, B, C, seed=seed) > 0.93).astype(F32) * F.synth_stress(n, S, B, C, seed=seed + 1)
            yp = F.synth_stress(n, S, B, C, seed=seed + 2)
        ev = yu.MeanAveragePrecision(C, B)
        for lo in range(0, n, 333):
            ev.update_state(torch.from_numpy(yt[lo:lo + 333]).to(dev), torch.from_numpy(yp[lo:lo + 333]).to(dev))
        rec, gt = yu.map_match(ev.all_true_boxes_variable, ev.all_pred_boxes_variable, C, 0.5, rows_by_image=True)
        assert rec.shape[0] == int(ev._st["cursors"][0]) and torch.equal(rec, ev._st["rec"][:rec.shape[0]])
        assert torch.equal(gt, ev._st["gt"])
        oe = O.MeanAveragePrecision(C, B)
        oe.update_state(yt[:400], yp[:400])
        e2 = yu.MeanAveragePrecision(C, B)
        e2.update_state(torch.from_numpy(yt[:400]).to(dev), torch.from_numpy(yp[:400]).to(dev))
        assert abs(float(e2.result()) - float(oe.result())) <= 1e-6
        assert float(e2.result()) == float(yu.mean_average_precision(e2.all_true_boxes_variable, e2.all_pred_boxes_variable, C))


def test_result_is_graph_capturable_and_sync_free(dev):
    """update_state + result() issue no host synchronisation and result() can be captured in a CUDA graph."""
    from yolohot import launch_count, utils as yu
    yt = F.synth_labels(256, seed=3)
    yp = F.synth_map_pred(yt)
    a, b = torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)
    ev = yu.MeanAveragePrecision(20, 2)
    ev.update_state(a, b)
    want = float(ev.result())
    torch.cuda.synchronize(dev)
    l0 = launch_count()
    torch.cuda.set_sync_debug_mode("error")                                  # any implicit synchronisation raises
    try:
        ev.reset_states()
        ev.update_state(a, b)
        m = ev.result()
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert launch_count() - l0 <= 5, launch_count() - l0                     # 2 x decode+NMS (+ tail), 1 update, 1 reduce
    assert float(m) == want
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        out = ev.result()
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = ev.result()
    torch.cuda.current_stream(dev).wait_stream(s)
    ev.update_state(a, b)                                                    # more rows, then replay: the graph reads the cursor
    g.replay()
    torch.cuda.synchronize(dev)
    e2 = yu.MeanAveragePrecision(20, 2)
    e2.update_state(torch.cat([a, a]), torch.cat([b, b]))
    assert float(out) == float(e2.result())


def test_numpy_evaluator_overwrites_after_reset(dev):
    """Pinned deviation from the reference's NumPy twin (utils.py:596-615 keeps appending after reset_states()): ours
    overwrites like the TF evaluator (utils.py:484-486)."""
    from yolohot import utils as yu
    yt = F.synth_labels(40, seed=2)
    yp = F.synth_map_pred(yt)
    en = yu.MeanAveragePrecisionNumpy(20, 2)
    en.update_state(yt[:20], yp[:20])
    first = en.all_pred_boxes_variable.shape[0]
    en.reset_states()
    en.update_state(yt[20:], yp[20:])
    e2 = yu.MeanAveragePrecisionNumpy(20, 2)
    e2.update_state(yt[20:], yp[20:])
    assert first > 0 and np.array_equal(en.all_pred_boxes_variable, e2.all_pred_boxes_variable)
    assert float(en.result()) == float(e2.result())


def test_tf_keras_loss_branch_end_to_end(dev):
    """loss.py:100-118 + yolo_v1.py:810,829: the Keras-loss form of YoloV1Loss (tf.custom_gradient over yh_loss_dl,
    tensors through DLPack) gives the torch path's value and gradient.  Runs against the TensorFlow stand-in of
    tests/golden/tfshim in its own process (TensorFlow itself is not in this image)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "tf_branch_check.py"), "gpu"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "tf_branch_check ok (gpu)" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_half_precision_head_trains_through_the_loss(dev):
    """A bf16 / fp16 head output that requires grad keeps its autograd graph through YoloV1Loss."""
    from yolohot import loss as yl
    yt = F.synth_labels(8, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    t = torch.from_numpy(yt).to(dev)
    for dt in (torch.bfloat16, torch.float16):
        w = torch.from_numpy(yp).to(dev).to(dt).requires_grad_(True)         # "the model's parameters"
        head = w * 1.0                                                       # a non-leaf half-precision head output
        total = yl.YoloV1Loss(20, 2)(t, head)
        total.backward()
        ref = w.detach().float().requires_grad_(True)
        yl.YoloV1Loss(20, 2)(t, ref).backward()
        assert w.grad is not None and w.grad.dtype == dt
        assert torch.equal(w.grad, ref.grad.to(dt))
