"""GPU parity tests (-m gpu): the CUDA path, called through the public surface / the C-ABI,
against the oracle on identical seeded inputs.  Bars (BASELINE.json north_star):
kept indices, class labels and kept counts bit-exact; decoded coordinates bit-exact here
(the kernels keep the reference's float32 op order; the stated bar is 1e-5 relative);
loss within 1e-5 relative; mAP within 1e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import cport
from oracle import yolo_oracle as O
from tests import fixtures as F

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))
F32 = np.float32


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _cuda(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _check_nms(got, want, what=""):
    gb, gc, gk = (x.cpu().numpy() if isinstance(x, torch.Tensor) else x for x in got)
    wb, wc, wk = want
    assert np.array_equal(gc, wc), f"{what}: kept counts differ at {np.nonzero(gc != wc)[0][:8]}"
    m = np.arange(wb.shape[1])[None, :] < wc[:, None]
    assert np.array_equal(np.where(m, gk, -1), np.where(m, wk, -1)), f"{what}: kept indices differ"
    assert np.array_equal(gb[m], wb[m]), f"{what}: kept rows differ"


# ------------------------------------------------------------------------------- IoU
def test_iou_bit_exact(dev):
    from yolohot import utils as yu
    rng = np.random.default_rng(0)
    a = rng.normal(0.4, 0.3, (4097, 4)).astype(F32)
    b = rng.normal(0.4, 0.3, (4097, 4)).astype(F32)
    got = yu.intersection_over_union(_cuda(a, dev), _cuda(b, dev))
    assert got.shape == (4097, 1)
    assert np.array_equal(got.cpu().numpy(), O.intersection_over_union(a, b))
    # the reference's other call shapes: (N,S,S,4) (loss.py:128) and two 4-vectors (utils.py:108)
    a4, b4 = a[:4096].reshape(4, 32, 32, 4), b[:4096].reshape(4, 32, 32, 4)
    assert np.array_equal(yu.intersection_over_union(_cuda(a4, dev), _cuda(b4, dev)).cpu().numpy(),
                          O.intersection_over_union(a4, b4))
    v = yu.intersection_over_union_numpy(a[0], b[0])
    assert isinstance(v, np.ndarray) and v.shape == (1,) and v[0] == O.intersection_over_union(a[0], b[0])[0]
    # unaligned view (offset by one float)
    buf = _cuda(np.concatenate([[0], a[:64].ravel()]).astype(F32), dev)[1:].reshape(64, 4)
    assert np.array_equal(yu.intersection_over_union(buf, _cuda(b[:64], dev)).cpu().numpy(),
                          O.intersection_over_union(a[:64], b[:64]))
    assert yu.intersection_over_union(torch.zeros((0, 4), device=dev), torch.zeros((0, 4), device=dev)).shape == (0, 1)


# ---------------------------------------------------------------------------- decode
@pytest.mark.parametrize("S,B,C,n", [(7, 2, 20, 64), (7, 2, 3, 5), (14, 3, 80, 6), (4, 1, 2, 9), (13, 5, 7, 3)])
def test_decode_bit_exact(dev, S, B, C, n):
    from yolohot import utils as yu
    p = F.synth_dense(n, S, B, C, seed=S * 100 + C)
    p[..., :C] = np.round(p[..., :C] * 8) / 8          # ties in the class argmax -> first max
    got = yu.decode_predictions(_cuda(p, dev), C, B)
    assert np.array_equal(got.cpu().numpy(), O.decode_predictions(p, C, B))
    assert np.array_equal(yu.decode_predictions_numpy(p, C, B), O.decode_predictions(p, C, B))


# ------------------------------------------------------------------ fused decode + NMS
CASES = [
    ("dense-voc", F.synth_dense, 64, 7, 2, 20, 0.5, 0.4),
    ("dense-voc-odd-tail", F.synth_dense, 203, 7, 2, 20, 0.5, 0.4),
    ("sparse-voc", F.synth_sparse, 256, 7, 2, 20, 0.5, 0.4),
    ("ties-voc", F.synth_quantised, 128, 7, 2, 20, 0.5, 0.3),
    ("ties-lowthr", F.synth_quantised, 64, 7, 2, 20, 0.2, 0.0),
    ("stress", F.synth_stress, 12, 14, 3, 80, 0.5, 0.05),
    ("stress-ties", F.synth_quantised, 10, 14, 3, 80, 0.5, 0.05),
    ("tiny-grid", F.synth_dense, 33, 4, 1, 3, 0.3, 0.2),
    ("s5-b3", F.synth_dense, 40, 5, 3, 11, 0.4, 0.5),
    ("s9", F.synth_quantised, 20, 9, 2, 6, 0.5, 0.3),
    ("s11-c1", F.synth_dense, 16, 11, 2, 1, 0.3, 0.3),
    ("s16", F.synth_quantised, 5, 16, 2, 4, 0.5, 0.2),
]


@pytest.mark.parametrize("name,gen,n,S,B,C,it,ct", CASES, ids=[c[0] for c in CASES])
def test_decode_nms_bit_exact(dev, name, gen, n, S, B, C, it, ct):
    from yolohot import utils as yu
    p = gen(n, S, B, C)
    want = O.decode_nms(p, C, B, it, ct) if n * S * S <= 4000 else cport.decode_nms(p, C, B, it, ct)
    got = yu.decode_nms(_cuda(p, dev), C, B, it, ct, return_index=True)
    _check_nms(got, want, name)


def test_decode_nms_paths_agree(dev, monkeypatch):
    """TMA ring vs direct kernel, and an 8-byte (not 16) aligned input that must take the direct path."""
    from yolohot import utils as yu
    p = F.synth_dense(1000, seed=21)
    want = cport.decode_nms(p, 20, 2)
    _check_nms(yu.decode_nms(_cuda(p, dev), 20, 2, return_index=True), want, "tma")
    monkeypatch.setenv("YH_TMA", "0")
    _check_nms(yu.decode_nms(_cuda(p, dev), 20, 2, return_index=True), want, "direct")
    monkeypatch.delenv("YH_TMA")
    flat = _cuda(np.concatenate([np.zeros(2, F32), p.ravel()]), dev)[2:].reshape(p.shape)
    assert flat.data_ptr() % 16 == 8
    _check_nms(yu.decode_nms(flat, 20, 2, return_index=True), want, "8B-aligned")
    # big images: cooperative team kernel (default, several team / ring geometries) and the direct kernel
    # behind it (YH_COOP=0); the exact-division variant of the IoU test
    ps = F.synth_stress(40)
    pq = F.synth_quantised(24, 14, 3, 80, seed=9)
    want_s, want_q = cport.decode_nms(ps, 80, 3, 0.5, 0.05), cport.decode_nms(pq, 80, 3, 0.5, 0.05)
    for env in ({}, {"YH_COOP_TEAMS": "1"}, {"YH_COOP_TEAMS": "2", "YH_COOP_STAGES": "3"}, {"YH_COOP_STAGES": "16"},
                {"YH_COOP_TEAMS": "3", "YH_COOP_STAGES": "7"}, {"YH_COOP": "0"}, {"YH_EXACT_DIV": "1"},
                {"YH_COOP": "0", "YH_EXACT_DIV": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        _check_nms(yu.decode_nms(_cuda(ps, dev), 80, 3, 0.5, 0.05, return_index=True), want_s, f"stress {env}")
        _check_nms(yu.decode_nms(_cuda(pq, dev), 80, 3, 0.5, 0.05, return_index=True), want_q, f"stress-ties {env}")
        for k in env:
            monkeypatch.delenv(k)
    for T, W, ST in ((4, 8, 2), (16, 16, 2), (8, 8, 5), (8, 16, 3), (4, 24, 5), (8, 24, 8), (2, 24, 8)):
        monkeypatch.setenv("YH_TMA_T", str(T)); monkeypatch.setenv("YH_TMA_W", str(W)); monkeypatch.setenv("YH_TMA_STAGES", str(ST))
        _check_nms(yu.decode_nms(_cuda(p, dev), 20, 2, return_index=True), want, f"tma T{T} W{W} ST{ST}")


def test_decode_nms_edge_cases(dev):
    from yolohot import utils as yu
    # empty batch, nothing passes, everything passes with one class and identical boxes, single image
    b, c = yu.decode_nms(torch.zeros((0, 7, 7, 30), device=dev), 20, 2)
    assert b.shape == (0, 49, 6) and c.shape == (0,)
    z = np.zeros((3, 7, 7, 30), F32)
    _check_nms(yu.decode_nms(_cuda(z, dev), 20, 2, return_index=True), O.decode_nms(z, 20, 2), "zeros")
    one = np.zeros((2, 7, 7, 30), F32)
    one[..., 0] = 1; one[..., 20] = 0.9; one[..., 21:25] = [0.5, 0.5, 0.2, 0.2]
    _check_nms(yu.decode_nms(_cuda(one, dev), 20, 2, return_index=True), O.decode_nms(one, 20, 2), "all-equal")
    same_center = np.zeros((1, 7, 7, 30), F32)
    same_center[..., 3] = 1; same_center[..., 20] = np.linspace(0.41, 0.99, 49).reshape(7, 7)
    same_center[..., 23:25] = 0.3
    _check_nms(yu.decode_nms(_cuda(same_center, dev), 20, 2, return_index=True), O.decode_nms(same_center, 20, 2), "chain")
    yt, yp = F.utils_demo()
    for y in (yt, yp):
        _check_nms(yu.decode_nms(_cuda(y, dev), 3, 2, return_index=True), O.decode_nms(y, 3, 2), "utils-demo")
    with pytest.raises(ValueError):
        yu.decode_nms(torch.zeros((1, 7, 6, 30), device=dev), 20, 2)
    with pytest.raises(NotImplementedError):
        yu.decode_nms(torch.zeros((1, 17, 17, 30), device=dev), 20, 2)


def test_reference_fixtures_through_reference_surface(dev):
    """utils.py:757-769 and dataset.py:153 as the reference's own smoke blocks drive them."""
    from yolohot import utils as yu
    yt, yp = F.utils_demo()
    nms = yu.non_max_suppression(yu.decode_predictions(_cuda(yp, dev), 3, 2)[0])
    assert np.array_equal(nms.cpu().numpy(), G["utils_demo_nms_pred"])
    nms_np = yu.non_max_suppression_numpy(yu.decode_predictions_numpy(yp, 3, 2)[0])
    assert isinstance(nms_np, np.ndarray) and np.array_equal(nms_np, G["utils_demo_nms_pred"])
    ev = yu.MeanAveragePrecision(3, 2)
    ev.update_state(_cuda(yt, dev), _cuda(yp, dev))
    assert abs(float(ev.result()) - float(G["utils_demo_map"])) <= 1e-6
    assert np.array_equal(ev.all_pred_boxes_variable.cpu().numpy()[:, 1:], G["utils_demo_nms_pred"])
    assert np.array_equal(ev.all_true_boxes_variable.cpu().numpy()[:, 1:], G["utils_demo_nms_true"])
    lab = O.encode_labels(F.TEST_TXT_BOXES, 7, 3, 2)[None].astype(F32)
    out = yu.non_max_suppression(yu.decode_predictions(_cuda(lab, dev), 3, 2)[0])
    assert np.array_equal(out.cpu().numpy(), G["test_txt_nms"])
    got = yu.decode_nms(_cuda(F.synth_dense(8, seed=1234), dev), 20, 2, return_index=True)
    _check_nms(got, (G["dense8_boxes"], G["dense8_count"], G["dense8_idx"]), "golden dense8")


def test_nms_rows_entry(dev):
    """non_max_suppression on already decoded rows, incl. float class values the fused path never sees."""
    from yolohot import utils as yu
    rng = np.random.default_rng(4)
    for M in (1, 31, 49, 100, 196, 256):
        rows = rng.random((6, M, 6), dtype=F32)
        rows[..., 0] = rng.integers(0, 4, (6, M)) * 0.5 - 0.5            # classes -0.5, 0, 0.5, 1.0
        rows[..., 1] = np.round(rows[..., 1] * 16) / 16
        rows[..., 4:6] = 0.2 + 0.4 * rows[..., 4:6]
        outs = yu.non_max_suppression(_cuda(rows, dev), 0.4, 0.3)
        for i in range(6):
            assert np.array_equal(outs[i].cpu().numpy(), O.non_max_suppression(rows[i], 0.4, 0.3)), (M, i)
    # integral classes take the integer-key path; -0.0 == +0.0; a class beyond the table (1000.) falls back
    for M in (49, 196):
        rows = rng.random((5, M, 6), dtype=F32)
        rows[..., 0] = rng.integers(0, 3, (5, M)).astype(F32)
        rows[1, ::3, 0] = -0.0
        rows[2, ::4, 0] = 1000.0
        rows[..., 4:6] = 0.2 + 0.4 * rows[..., 4:6]
        outs = yu.non_max_suppression(_cuda(rows, dev), 0.4, 0.3)
        for i in range(5):
            assert np.array_equal(outs[i].cpu().numpy(), O.non_max_suppression(rows[i], 0.4, 0.3)), (M, i)
    assert yu.non_max_suppression(_cuda(rows[0], dev), 0.5, 2.0).shape == (0, 6)
    assert yu.non_max_suppression_2(_cuda(rows[0], dev)).shape[1] == 6


def test_decode_nms_host_entry(dev):
    """yh_decode_nms_host: NumPy in / NumPy out through the pipelined host path (several chunks)."""
    from yolohot import utils as yu
    p = F.synth_dense(40000, seed=8)
    want = cport.decode_nms(p, 20, 2, nthreads=cport.num_threads())
    got = yu.decode_nms(p, 20, 2, return_index=True)
    assert all(isinstance(x, np.ndarray) for x in got)
    _check_nms(got, want, "host")
    pin = torch.from_numpy(p[:5001]).pin_memory().numpy()
    _check_nms(yu.decode_nms(pin, 20, 2, return_index=True), tuple(w[:5001] for w in want), "host-pinned")
    # half-precision heads stay half on the wire (yh_decode_nms_host_typed): equal to the float32 path on the widened tensor
    for h in (p[:30001].astype(np.float16), torch.from_numpy(p[:30001]).to(torch.bfloat16)):
        wide = h.astype(np.float32) if isinstance(h, np.ndarray) else h.float().numpy()
        wantw = cport.decode_nms(wide, 20, 2, nthreads=cport.num_threads())
        goth = yu.decode_nms(h, 20, 2, return_index=True)
        assert all(isinstance(x, np.ndarray) for x in goth)
        _check_nms(goth, wantw, "host-half")


def test_decode_nms_host_compact_rows(dev):
    """yh_decode_nms_host_rows: only the kept rows cross PCIe; equal to the padded host result, over several chunks,
    for float32 and float16 inputs, with a row buffer that is too small at first."""
    from yolohot import utils as yu
    for gen, n in ((F.synth_sparse, 40_000), (F.synth_dense, 3000), (F.synth_dense, 1)):
        p = gen(n)
        for x in (p, p.astype(np.float16)):
            boxes, cnt = yu.decode_nms(x, 20, 2)
            rows, cnt2 = yu.decode_nms(x, 20, 2, compact=True)
            assert np.array_equal(cnt2, cnt) and rows.shape == (int(cnt.sum()), 7)
            m = np.arange(49)[None, :] < cnt[:, None]
            assert np.array_equal(rows[:, 1:], boxes[m])
            assert np.array_equal(rows[:, 0], np.repeat(np.arange(n, dtype=np.float32), cnt))
    with pytest.raises(ValueError):
        yu.decode_nms(_cuda(F.synth_dense(4), dev), 20, 2, compact=True)


def test_decode_nms_large_vs_cport_and_properties(dev):
    """100k images against the C port; size-independent properties on the same output."""
    from yolohot import utils as yu
    torch.manual_seed(5)
    p = torch.rand((100_000, 7, 7, 30), device=dev)
    boxes, cnt, kidx = yu.decode_nms(p, 20, 2, return_index=True)
    want = cport.decode_nms(p.cpu().numpy(), 20, 2, nthreads=cport.num_threads())
    _check_nms((boxes, cnt, kidx), want, "100k")
    m = torch.arange(49, device=dev)[None, :] < cnt[:, None]
    conf = torch.where(m, boxes[..., 1], torch.full_like(boxes[..., 1], float("inf")))
    assert bool((conf > 0.4).all())                                   # strict filter
    c2 = torch.where(m, boxes[..., 1], torch.zeros_like(boxes[..., 1]))
    assert bool((c2[:, 1:] <= c2[:, :-1]).all())                      # pick order = descending confidence
    # idempotence: NMS of the kept rows keeps them all, in the same order
    b2 = torch.where(m[..., None], boxes, torch.zeros_like(boxes))[:2000]
    again = yu.non_max_suppression(b2)
    for i in range(0, 2000, 97):
        assert torch.equal(again[i], b2[i, :int(cnt[i])])


# ------------------------------------------------------------------------------- loss
@pytest.mark.parametrize("n,S,B,C", [(64, 7, 2, 20), (5, 7, 2, 3), (9, 14, 3, 80), (33, 4, 1, 2)])
def test_loss_forward_backward(dev, n, S, B, C):
    from yolohot import loss as yl
    yt = F.synth_labels(n, S, B, C, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    ref = O.yolo_v1_loss(yt, yp, C, B)
    fn = yl.YoloV1Loss(C, B)
    p = _cuda(yp, dev).requires_grad_(True)
    total = fn(_cuda(yt, dev), p)
    assert total.dim() == 0 and fn.batch_size == n
    terms = fn.last_terms.cpu().numpy()
    for i, k in enumerate(("xy", "wh", "obj", "noobj", "cls", "total")):
        assert abs(terms[i] - ref[k + "_f64"]) <= 1e-5 * max(abs(ref[k + "_f64"]), 1e-6), (k, terms[i], ref[k + "_f64"])
    (2.0 * total).backward()                                   # upstream gradient is honoured
    g = p.grad.cpu().numpy() / 2.0
    g_ref = O.yolo_v1_loss_grad(yt, yp, C, B)
    scale = np.abs(g_ref).max()
    assert np.abs(g - g_ref).max() <= 1e-5 * scale + 1e-6, np.abs(g - g_ref).max()
    assert np.array_equal(g == 0, g_ref == 0) or np.abs(g[g_ref == 0]).max() < 1e-6
    # forward only (no grad requested) gives the same scalar; numpy in -> numpy scalar out
    assert float(fn(_cuda(yt, dev), _cuda(yp, dev))) == float(total)
    assert isinstance(fn(yt, yp), np.float32)


def test_loss_reference_fixture_and_run_to_run(dev):
    from yolohot import loss as yl
    yt, yp = F.loss_demo()
    fn = yl.YoloV1Loss(num_classes=3, num_boxes=2)
    v = float(fn(_cuda(yt, dev), _cuda(yp, dev)))
    assert abs(v - 0.17571506) <= 1e-5 * 0.17571506
    terms, grad = yl.yolo_v1_loss_terms(_cuda(yt, dev), _cuda(yp, dev), 3, 2, grad=True)
    np.testing.assert_allclose(terms.cpu().numpy(), G["loss_demo_terms"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(grad.cpu().numpy(), G["loss_demo_grad"], rtol=1e-4, atol=1e-6)
    yt = F.synth_labels(512, seed=3)
    yp = F.synth_loss_pred(yt.shape, seed=3)
    a = yl.yolo_v1_loss_terms(_cuda(yt, dev), _cuda(yp, dev))
    b = yl.yolo_v1_loss_terms(_cuda(yt, dev), _cuda(yp, dev))
    assert torch.equal(a, b)                                           # deterministic reduction


def test_loss_kernels_agree(dev, monkeypatch):
    """The three loss kernels (TMA ring, gather, stream) against the oracle and each other: terms to 1e-5, gradients bit for bit
    (both evaluate the same per-cell closed form), on aligned, ragged and offset batches."""
    from yolohot import loss as yl
    for n, off in ((4096, 0), (777, 0), (130, 1), (1, 0)):
        yt = F.synth_labels(n + off, seed=5)[off:]
        yp = F.synth_loss_pred((n + off, 7, 7, 30), seed=5)[off:]
        t = _cuda(np.ascontiguousarray(yt), dev) if off == 0 else torch.from_numpy(F.synth_labels(n + off, seed=5)).to(dev)[off:]
        p = _cuda(np.ascontiguousarray(yp), dev) if off == 0 else torch.from_numpy(F.synth_loss_pred((n + off, 7, 7, 30), seed=5)).to(dev)[off:]
        want = cport.loss(np.ascontiguousarray(yt), np.ascontiguousarray(yp), 20, 2)
        res = {}
        for mode in ("0", "1", "2"):                                  # TMA ring, gather, stream (TMA in / TMA out)
            monkeypatch.setenv("YH_LOSS_GATHER", mode)
            res[mode] = yl.yolo_v1_loss_terms(t, p, grad=True)
            np.testing.assert_allclose(res[mode][0].cpu().numpy(), want, rtol=1e-5, err_msg=f"n={n} off={off} gather={mode}")
            fwd = yl.yolo_v1_loss_terms(t, p)
            assert torch.equal(fwd, res[mode][0])
        monkeypatch.delenv("YH_LOSS_GATHER")
        assert torch.equal(res["0"][1], res["1"][1]), (n, off)
        assert torch.equal(res["0"][1], res["2"][1]), (n, off)
        np.testing.assert_allclose(res["0"][0].cpu().numpy(), res["1"][0].cpu().numpy(), rtol=1e-6)
        np.testing.assert_allclose(res["0"][0].cpu().numpy(), res["2"][0].cpu().numpy(), rtol=1e-6)


def test_loss_linearity_property(dev):
    """Batch sum: loss(concat(a, b)) == loss(a) + loss(b) (loss.py:172-213 sums over everything)."""
    from yolohot import loss as yl
    yt = F.synth_labels(300, seed=13)
    yp = F.synth_loss_pred(yt.shape, seed=13)
    whole = yl.yolo_v1_loss_terms(_cuda(yt, dev), _cuda(yp, dev)).double().cpu().numpy()
    parts = sum(yl.yolo_v1_loss_terms(_cuda(yt[lo:hi], dev), _cuda(yp[lo:hi], dev)).double().cpu().numpy()
                for lo, hi in ((0, 77), (77, 300)))
    np.testing.assert_allclose(whole, parts, rtol=2e-6)


# -------------------------------------------------------------------------------- mAP
def test_map_function_vs_oracle(dev):
    from yolohot import utils as yu
    yt = F.synth_labels(300, seed=11)
    mp = F.synth_map_pred(yt)
    oe = O.MeanAveragePrecision(20, 2)
    oe.update_state(yt, mp)
    t_rows, p_rows = oe.all_true_boxes_variable, oe.all_pred_boxes_variable
    want, want_ap, _ = O.mean_average_precision(t_rows, p_rows, 20, return_details=True)
    got, ap = yu.mean_average_precision(_cuda(t_rows, dev), _cuda(p_rows, dev), 20, return_ap=True)
    assert abs(float(got) - float(want)) <= 1e-6
    np.testing.assert_allclose(ap.cpu().numpy(), want_ap, atol=1e-6)
    # row order of the inputs is arbitrary in the reference's contract: shuffle images (stable inside an image)
    rng = np.random.default_rng(0)
    perm = rng.permutation(300)
    t2 = np.concatenate([t_rows[t_rows[:, 0] == i] for i in perm])
    p2 = np.concatenate([p_rows[p_rows[:, 0] == i] for i in perm])
    want2 = O.mean_average_precision(t2, p2, 20)
    assert abs(float(yu.mean_average_precision(_cuda(t2, dev), _cuda(p2, dev), 20)) - float(want2)) <= 1e-6
    # other thresholds, numpy twin, empty detections, class with no GT
    for thr in (0.3, 0.75):
        assert abs(float(yu.mean_average_precision_numpy(t_rows, p_rows, 20, thr)) -
                   float(O.mean_average_precision(t_rows, p_rows, 20, thr))) <= 1e-6
    assert float(yu.mean_average_precision(_cuda(t_rows, dev), torch.zeros((0, 7), device=dev), 20)) == 0.0
    assert abs(float(yu.mean_average_precision_2(_cuda(t_rows, dev), _cuda(p_rows, dev), num_classes=25)) -
               float(O.mean_average_precision(t_rows, p_rows, 25))) <= 1e-6


def test_map_ties_and_duplicates(dev):
    """Equal confidences (stable order) and several detections on one GT (only the first is TP)."""
    from yolohot import utils as yu
    rng = np.random.default_rng(3)
    t_rows, p_rows = [], []
    for img in range(60):
        for _ in range(rng.integers(1, 4)):
            c = rng.integers(0, 3)
            box = [rng.random(), rng.random(), 0.2 + 0.3 * rng.random(), 0.2 + 0.3 * rng.random()]
            t_rows.append([img, c, 1.0] + box)
            for _ in range(rng.integers(0, 4)):
                j = 0.02 * rng.standard_normal(4)
                p_rows.append([img, c if rng.random() < 0.8 else (c + 1) % 3, rng.integers(1, 5) / 5.0] + list(np.array(box) + j))
    t_rows, p_rows = np.array(t_rows, F32), np.array(p_rows, F32)
    want = O.mean_average_precision(t_rows, p_rows, 3)
    got = yu.mean_average_precision(_cuda(t_rows, dev), _cuda(p_rows, dev), 3)
    assert abs(float(got) - float(want)) <= 1e-6 and 0.1 < float(want) < 1.0


def test_evaluator_streaming_and_reset(dev):
    from yolohot import metric as ym
    from yolohot import utils as yu
    yt = F.synth_labels(500, seed=11)
    mp = F.synth_map_pred(yt)
    oe = O.MeanAveragePrecision(20, 2)
    oe.update_state(yt, mp)
    ev = yu.MeanAveragePrecision(20, 2)
    for lo in range(0, 500, 64):                                       # growing buffers, ragged last batch
        ev.update_state(_cuda(yt[lo:lo + 64], dev), _cuda(mp[lo:lo + 64], dev))
    assert ev.img_idx == 500
    assert np.array_equal(ev.all_pred_boxes_variable.cpu().numpy(), oe.all_pred_boxes_variable)
    assert np.array_equal(ev.all_true_boxes_variable.cpu().numpy(), oe.all_true_boxes_variable)
    assert abs(float(ev.result()) - float(oe.result())) <= 1e-6
    ev.reset_states()                                                  # utils.py:467-468 + :484-486
    ev.update_state(_cuda(yt[:32], dev), _cuda(mp[:32], dev))
    o2 = O.MeanAveragePrecision(20, 2)
    o2.update_state(yt[:32], mp[:32])
    assert np.array_equal(ev.all_pred_boxes_variable.cpu().numpy(), o2.all_pred_boxes_variable)
    assert abs(float(ev.result()) - float(o2.result())) <= 1e-6
    en = yu.MeanAveragePrecisionNumpy(20, 2)
    en.update_state(yt[:32], mp[:32])
    assert isinstance(en.all_pred_boxes_variable, np.ndarray) and abs(float(en.result()) - float(o2.result())) <= 1e-6
    # stale metric.py evaluators: GT thresholded only (metric.py:81), driven as metric.py:142-155
    mt, p1, p2 = F.metric_demo()
    for cls in (ym.MeanAveragePrecision2, ym.MeanAveragePrecision):
        e = cls()
        for i in range(5):
            e.update_state(_cuda(mt, dev), _cuda(p1 if i == 0 else p2, dev))
        assert abs(float(e.result()) - float(G["metric_demo_map"])) <= 1e-6
    o3 = O.MeanAveragePrecision(20, 2, nms_true=False)
    o3.update_state(yt[:100], mp[:100])
    e3 = ym.MeanAveragePrecision2()
    e3.update_state(_cuda(yt[:100], dev), _cuda(mp[:100], dev))
    assert np.array_equal(e3.all_true_bboxes_variable.cpu().numpy(), o3.all_true_boxes_variable)
    assert abs(float(e3.result()) - float(o3.result())) <= 1e-6


def test_sharded_map_single_process_emulation(dev):
    """Ranks emulated as contiguous image shards on one GPU: per-shard yh_map_match, records
    concatenated in shard order, yh_map_reduce == the unsharded value (multi-GPU parity rule)."""
    from yolohot import utils as yu
    yt = F.synth_labels(400, seed=11)
    mp = F.synth_map_pred(yt)
    ev = yu.MeanAveragePrecision(20, 2)
    ev.update_state(_cuda(yt, dev), _cuda(mp, dev))
    whole = float(ev.result())
    for w in (3, 8):
        rs, gs = [], 0
        for r in range(w):
            lo, hi = 400 * r // w, 400 * (r + 1) // w
            e = yu.MeanAveragePrecision(20, 2)
            e.update_state(_cuda(yt[lo:hi], dev), _cuda(mp[lo:hi], dev))
            # the general path (arbitrary rows) on the shard's rows == the records the evaluator kept while updating
            rec, g = yu.map_match(e.all_true_boxes_variable, e.all_pred_boxes_variable, 20, 0.5)
            assert torch.equal(rec, e._st["rec"][:rec.shape[0]]) and torch.equal(g, e._st["gt"])
            rs.append(rec); gs = gs + g
        m, _ = yu.map_reduce(torch.cat(rs), gs, 20)
        assert float(m) == whole


# ------------------------------------------------------------ BASELINE.json configs at full size
def _dense_cfg5(n, dev, seed=99):
    """cfg5 data on the device (same recipe as bench.py cfg5_stress)."""
    S, B, C = 14, 3, 80
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    p = torch.rand((n, S, S, C + 5 * B), generator=g, device=dev)
    dom = torch.randint(0, 4, (n, S, S), generator=g, device=dev)
    boost = torch.rand((n, S, S), generator=g, device=dev) < 0.8
    for k in range(4):
        p[..., k] += 1.5 * (boost & (dom == k)).float()
    for b in range(B):
        p[..., C + 5 * b + 3:C + 5 * b + 5] = 0.1 + 0.5 * p[..., C + 5 * b + 3:C + 5 * b + 5]
    return p


def _full_size_checks(yu, p, C, B, it, ct, audit, what):
    """Oracle (C port) on audit slices of the batch + size-independent properties on all of it."""
    n, M = p.shape[0], p.shape[1] * p.shape[2]
    boxes, cnt, kidx = yu.decode_nms(p, C, B, it, ct, return_index=True)
    for lo in (0, n // 2 - audit // 2, n - audit):                       # head, middle and tail of the batch
        want = cport.decode_nms(p[lo:lo + audit].cpu().numpy(), C, B, it, ct, nthreads=cport.num_threads())
        _check_nms((boxes[lo:lo + audit], cnt[lo:lo + audit], kidx[lo:lo + audit]), want, f"{what} slice@{lo}")
    m = torch.arange(M, device=p.device)[None, :] < cnt[:, None]
    conf = boxes[..., 1]
    assert bool((torch.where(m, conf, torch.full_like(conf, float("inf"))) > ct).all())           # strict filter
    c2 = torch.where(m, conf, torch.zeros_like(conf))
    assert bool((c2[:, 1:] <= c2[:, :-1]).all())                                                  # pick order
    # an image's result depends on that image only: any permutation of the batch permutes the result
    perm = torch.randperm(n, device=p.device, generator=torch.Generator(device=p.device).manual_seed(1))
    b2, c2n, k2 = yu.decode_nms(p[perm], C, B, it, ct, return_index=True)
    assert torch.equal(c2n, cnt[perm])
    assert torch.equal(torch.where(m[perm], k2, torch.full_like(k2, -1)), torch.where(m, kidx, torch.full_like(kidx, -1))[perm])
    assert torch.equal(b2[m[perm]], boxes[perm][m[perm]])
    return int(cnt.sum())


def test_cfg2_full_size_one_million_images(dev):
    """BASELINE configs[1]: 1,000,000 dense VOC-shaped outputs, TMA tile ring."""
    from yolohot import utils as yu
    g = torch.Generator(device=dev)
    g.manual_seed(2025)
    p = torch.rand((1_000_000, 7, 7, 30), generator=g, device=dev)
    kept = _full_size_checks(yu, p, 20, 2, 0.5, 0.4, 16384, "cfg2")
    assert 39.0 < kept / 1e6 < 41.0


def test_more_than_2_pow_31_elements(dev):
    """64-bit addressing: 1.6 M VOC images = 2.35e9 input elements (> 2**31); the tail of the batch (elements past
    the 32-bit limit) is audited against the C port, decode / loss / gradient there against a run on the tail alone."""
    from yolohot import utils as yu, loss as yl
    n, audit = 1_600_000, 4096
    g = torch.Generator(device=dev)
    g.manual_seed(77)
    p = torch.rand((n, 7, 7, 30), generator=g, device=dev)
    assert p.numel() > 2 ** 31
    boxes, cnt, kidx = yu.decode_nms(p, 20, 2, 0.5, 0.4, return_index=True)
    for lo in (0, n - audit):
        want = cport.decode_nms(p[lo:lo + audit].cpu().numpy(), 20, 2, 0.5, 0.4, nthreads=cport.num_threads())
        _check_nms((boxes[lo:lo + audit], cnt[lo:lo + audit], kidx[lo:lo + audit]), want, f"2^31 slice@{lo}")
    tb, tc, tk = yu.decode_nms(p[n - 300_000:], 20, 2, 0.5, 0.4, return_index=True)
    assert torch.equal(tc, cnt[n - 300_000:])
    mt = torch.arange(49, device=dev)[None, :] < tc[:, None]
    assert torch.equal(tb[mt], boxes[n - 300_000:][mt]) and torch.equal(tk[mt], kidx[n - 300_000:][mt])
    del boxes, kidx, tb, tk, mt
    d_all = yu.decode_predictions(p, 20, 2)
    assert torch.equal(d_all[n - audit:], yu.decode_predictions(p[n - audit:], 20, 2))
    del d_all
    # loss: label grids = thresholded copies of the predictions' own layout (any y_true works for an addressing check)
    t = torch.zeros_like(p)
    obj = p[..., 20] > 0.9
    t[..., 20] = obj.float()
    t[..., 21:25] = p[..., 25:29] * obj[..., None]
    t[..., 3] = obj.float()
    terms, grad = yl.yolo_v1_loss_terms(t, p, grad=True)
    parts = torch.zeros(6, dtype=torch.float64, device=dev)
    step = 400_000
    for lo in range(0, n, step):
        parts += yl.yolo_v1_loss_terms(t[lo:lo + step], p[lo:lo + step]).double()
    np.testing.assert_allclose(terms.cpu().numpy(), parts.cpu().numpy(), rtol=2e-5)
    _, g_tail = yl.yolo_v1_loss_terms(t[n - audit:], p[n - audit:], grad=True)
    assert torch.equal(grad[n - audit:], g_tail)                         # the sum's gradient is per cell
    want = cport.loss(t[n - audit:].cpu().numpy(), p[n - audit:].cpu().numpy(), 20, 2)
    np.testing.assert_allclose(yl.yolo_v1_loss_terms(t[n - audit:], p[n - audit:]).cpu().numpy(), want, rtol=1e-5)


def test_cfg5_full_size_stress(dev):
    """BASELINE configs[4]: S=14 B=3 C=80, conf threshold 0.05, 131,072 images (9.76 GB), big-image kernel."""
    from yolohot import utils as yu
    p = _dense_cfg5(131_072, dev)
    kept = _full_size_checks(yu, p, 80, 3, 0.5, 0.05, 2048, "cfg5")
    assert kept / 131_072 > 100


def test_cfg3_full_size_loss_batch_4096(dev):
    """BASELINE configs[2]: loss forward + backward at batch 4096 against the C port / NumPy oracle."""
    from yolohot import loss as yl
    yt = F.synth_labels(4096, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    terms, grad = yl.yolo_v1_loss_terms(_cuda(yt, dev), _cuda(yp, dev), grad=True)
    want = cport.loss(yt, yp, 20, 2)
    np.testing.assert_allclose(terms.cpu().numpy(), want, rtol=1e-5)
    g_ref = O.yolo_v1_loss_grad(yt[:256], yp[:256])
    np.testing.assert_allclose(grad[:256].cpu().numpy(), g_ref, rtol=1e-4, atol=1e-4)
    # the gradient is zero wherever the closed form says so (most of the tensor), everywhere in the batch
    nz = (grad != 0).float().mean().item()
    assert 0.01 < nz < 0.15
    again, _ = yl.yolo_v1_loss_terms(_cuda(yt, dev), _cuda(yp, dev), grad=True)
    assert torch.equal(again, terms)                                     # run-to-run identical


def test_cfg4_full_size_map_5000_images(dev):
    """BASELINE configs[3]: mAP@0.5 over 5,000 images, streamed in batches, against the C port."""
    from yolohot import utils as yu
    yt = F.synth_labels(5000, seed=11)
    mp = F.synth_map_pred(yt)
    ev = yu.MeanAveragePrecision(20, 2)
    for lo in range(0, 5000, 512):
        ev.update_state(_cuda(yt[lo:lo + 512], dev), _cuda(mp[lo:lo + 512], dev))
    t_rows = ev.all_true_boxes_variable.cpu().numpy()
    p_rows = ev.all_pred_boxes_variable.cpu().numpy()
    pb, pc, _ = cport.decode_nms(mp, 20, 2, nthreads=cport.num_threads())
    assert p_rows.shape[0] == int(pc.sum())
    want, want_ap = cport.mean_average_precision(t_rows, p_rows, 20)
    got, ap = yu.mean_average_precision(_cuda(t_rows, dev), _cuda(p_rows, dev), 20, return_ap=True)
    assert abs(float(ev.result()) - float(want)) <= 1e-6 and abs(float(got) - float(want)) <= 1e-6
    np.testing.assert_allclose(ap.cpu().numpy(), want_ap, atol=1e-6)


def test_big_image_edge_cases(dev):
    """Shapes that take the cooperative team kernel (> 12 KB per image): empty images, one class with
    identical boxes (longest suppression chain), all cells tied, other team widths (S=13 -> 6 warps,
    S=16 -> 8 warps), single image, image counts around the team / CTA boundaries."""
    from yolohot import utils as yu
    S, B, C = 14, 3, 80
    D = C + 5 * B
    z = np.zeros((5, S, S, D), F32)
    z[2, ..., C] = 0.05                                               # conf == threshold: strict >, nothing passes
    _check_nms(yu.decode_nms(_cuda(z, dev), C, B, 0.5, 0.05, return_index=True), cport.decode_nms(z, C, B, 0.5, 0.05), "big zeros")
    one = np.zeros((3, S, S, D), F32)
    one[..., 7] = 1; one[..., C] = 0.9; one[..., C + 1:C + 5] = [0.5, 0.5, 0.2, 0.2]   # all tied, one class
    one[1, ..., C] = np.linspace(0.06, 0.99, S * S).reshape(S, S)                       # strictly ordered chain
    _check_nms(yu.decode_nms(_cuda(one, dev), C, B, 0.5, 0.05, return_index=True), cport.decode_nms(one, C, B, 0.5, 0.05), "big chain")
    mixed = F.synth_stress(9)
    mixed[4] = 0                                                      # an empty image between full ones
    mixed[7, ..., :C] = 0; mixed[7, ..., 3] = 1                       # one image with a single class everywhere
    _check_nms(yu.decode_nms(_cuda(mixed, dev), C, B, 0.5, 0.05, return_index=True), cport.decode_nms(mixed, C, B, 0.5, 0.05), "big mixed")
    for (s_, b_, c_, thr) in ((13, 5, 7, 0.3), (16, 2, 4, 0.2), (12, 2, 40, 0.4), (10, 3, 100, 0.1)):
        for n in (1, 3, 149, 600):
            p = F.synth_quantised(n, s_, b_, c_, seed=n + s_) if n != 149 else F.synth_dense(n, s_, b_, c_, seed=s_)
            _check_nms(yu.decode_nms(_cuda(p, dev), c_, b_, 0.45, thr, return_index=True),
                       cport.decode_nms(p, c_, b_, 0.45, thr, nthreads=cport.num_threads()), f"S{s_} B{b_} C{c_} n{n}")


def test_score_mode_extension(dev):
    """score_mode='conf_x_prob' (not in the reference): conf x winning class probability drives the
    threshold, the order and the reported confidence.  Checked against the oracle's NMS run on decoded
    rows whose confidence column was multiplied by the float32 class maximum."""
    from yolohot import utils as yu
    for (gen, n, S, B, C, it, ct) in ((F.synth_dense, 300, 7, 2, 20, 0.5, 0.25), (F.synth_stress, 20, 14, 3, 80, 0.5, 0.3)):
        p = gen(n, S, B, C)
        rows = O.decode_predictions(p, C, B)
        rows[..., 1] = (rows[..., 1] * p[..., :C].max(-1).reshape(n, -1)).astype(F32)
        got_b, got_c = yu.decode_nms(_cuda(p, dev), C, B, it, ct, score_mode="conf_x_prob")
        got_b, got_c = got_b.cpu().numpy(), got_c.cpu().numpy()
        for i in range(n):
            want = O.non_max_suppression(rows[i], it, ct)
            assert got_c[i] == len(want) and np.array_equal(got_b[i, :len(want)], want), i
    ref_b, ref_c = yu.decode_nms(_cuda(p, dev), C, B, it, ct)                    # default stays the reference
    _check_nms((ref_b, ref_c, yu.decode_nms(_cuda(p, dev), C, B, it, ct, return_index=True)[2]), cport.decode_nms(p, C, B, it, ct), "default")
    with pytest.raises(ValueError):
        yu.decode_nms(_cuda(p, dev), C, B, score_mode="prob")


def test_concurrent_threads_and_streams(dev):
    """Re-entrancy: four host threads, each on its own CUDA stream, run decode+NMS, the loss and
    the mAP stages at the same time; every thread must get the single-threaded results bit for bit."""
    import threading
    from yolohot import loss as yl, utils as yu
    p = _cuda(F.synth_dense(20000, seed=21), dev)
    yt = _cuda(F.synth_labels(600, seed=21), dev)
    yp = _cuda(F.synth_loss_pred((600, 7, 7, 30), seed=21), dev)
    mt = F.synth_labels(300, seed=22)
    mp_ = _cuda(F.synth_map_pred(mt), dev)
    mt = _cuda(mt, dev)

    def work():
        b, c, k = yu.decode_nms(p, 20, 2, return_index=True)
        terms, grad = yl.yolo_v1_loss_terms(yt, yp, grad=True)
        ev = yu.MeanAveragePrecision(20, 2)
        ev.update_state(mt, mp_)
        m = ev.result()
        msk = torch.arange(49, device=dev)[None, :] < c[:, None]
        return (c.clone(), b[msk].clone(), k[msk].clone(), terms.clone(), grad.clone(), float(m))

    want = work()
    torch.cuda.synchronize(dev)
    out, errs = {}, []

    def run(i):
        try:
            s = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(s):
                for _ in range(5):
                    out[i] = work()
            s.synchronize()
        except Exception as e:                                   # surfaced below, in the main thread
            errs.append(repr(e))

    ths = [threading.Thread(target=run, args=(i,)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for i in range(4):
        got = out[i]
        assert all(torch.equal(g, w) for g, w in zip(got[:5], want[:5])), i
        assert got[5] == want[5], i


def test_images_not_a_multiple_of_16_bytes(dev, monkeypatch):
    """Team kernel on images whose byte size is 4, 8 or 12 mod 16 (odd cell / channel counts, 16-bit heads): chunks are
    fetched as the aligned range around them.  Odd and even batch sizes (the last image goes to the direct kernel when the
    tensor ends off a 16-byte boundary), against the C port and against the direct kernel."""
    from yolohot import utils as yu
    cases = [(7, 3, 80, torch.float32), (13, 3, 80, torch.float32), (14, 3, 80, torch.float16), (14, 3, 80, torch.bfloat16),
             (11, 1, 20, torch.float16), (9, 2, 33, torch.float32)]
    for S, B, C, dt in cases:
        for n in (1, 2, 5, 300, 301):
            p = F.synth_stress(n, S, B, C, seed=S + n, dominant=4)
            t = torch.from_numpy(p).to(dev).to(dt)
            img_bytes = S * S * (C + 5 * B) * t.element_size()
            assert img_bytes % 16 != 0
            want = cport.decode_nms(t.float().cpu().numpy(), C, B, 0.5, 0.05, nthreads=cport.num_threads())
            got = yu.decode_nms(t, C, B, 0.5, 0.05, return_index=True)
            _check_nms(got, want, f"shifted S={S} B={B} C={C} {dt} n={n}")
            monkeypatch.setenv("YH_COOP_SHIFTED", "0")
            ref = yu.decode_nms(t, C, B, 0.5, 0.05, return_index=True)
            monkeypatch.delenv("YH_COOP_SHIFTED")
            assert torch.equal(ref[1], got[1])
