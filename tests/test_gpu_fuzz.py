"""Randomised differential test (-m gpu): the fused decode+NMS entry points against the C port over random
shapes, batch sizes, thresholds, value distributions, element types and alignments - every kernel behind
yh_decode_nms (tile ring, cooperative team kernel, direct kernel) gets hit.  Bit-exact bar as everywhere."""
import os

import numpy as np
import pytest
import torch

from oracle import cport
from tests import fixtures as F

pytestmark = pytest.mark.gpu
F32 = np.float32


def _check(got, want, what):
    gb, gc, gk = (x.cpu().numpy() for x in got)
    wb, wc, wk = want
    assert np.array_equal(gc, wc), f"{what}: counts differ at {np.nonzero(gc != wc)[0][:6]}"
    m = np.arange(wb.shape[1])[None, :] < wc[:, None]
    assert np.array_equal(gk[m], wk[m]), f"{what}: kept indices differ"
    assert np.array_equal(gb[m], wb[m]), f"{what}: kept rows differ"


def test_decode_nms_random_shapes():
    from yolohot import utils as yu
    dev = torch.device("cuda:0")
    # YH_FUZZ_SEED / YH_FUZZ_TRIALS: longer one-off campaigns (the defaults keep the suite short)
    rng = np.random.default_rng(int(os.environ.get("YH_FUZZ_SEED", 20261018)))
    gens = (F.synth_dense, F.synth_quantised, F.synth_sparse, F.synth_stress)
    for trial in range(int(os.environ.get("YH_FUZZ_TRIALS", 120))):
        S = int(rng.integers(1, 17))
        B = int(rng.integers(1, 5))
        C = int(rng.choice([1, 2, 3, 5, 7, 20, 21, 33, 80, 90]))
        n = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 31, 64, 100, 257]))
        if S * S * (C + 5 * B) * n > 6_000_000:
            n = max(1, 6_000_000 // (S * S * (C + 5 * B)))
        it = float(rng.choice([0.2, 0.3, 0.5, 0.5, 0.7, 1.0]))
        ct = float(rng.choice([0.0, 0.05, 0.3, 0.4, 0.6]))
        gen = gens[int(rng.integers(0, len(gens)))]
        p = gen(n, S, B, C, seed=1000 + trial) if gen is not F.synth_stress else gen(n, S, B, C, seed=1000 + trial, dominant=min(4, C))
        if trial % 5 == 0:
            p[rng.integers(0, n)] = 0                                     # an empty image somewhere
        what = f"trial {trial}: S={S} B={B} C={C} n={n} iou={it} conf={ct} {gen.__name__}"
        want = cport.decode_nms(p, C, B, it, ct, nthreads=cport.num_threads())
        t = torch.from_numpy(p).to(dev)
        _check(yu.decode_nms(t, C, B, it, ct, return_index=True), want, what)
        if trial % 3 == 0:                                                # 8-byte (not 16) aligned base -> direct kernel
            flat = torch.cat([torch.zeros(2, device=dev), t.reshape(-1)])[2:].reshape(t.shape)
            _check(yu.decode_nms(flat, C, B, it, ct, return_index=True), want, what + " (8B-aligned)")
        if trial % 4 == 0:                                                # half-precision heads: exact widening inside the kernel
            for dt in (torch.float16, torch.bfloat16):
                h = t.to(dt)
                wh = cport.decode_nms(h.float().cpu().numpy(), C, B, it, ct, nthreads=cport.num_threads())
                _check(yu.decode_nms(h, C, B, it, ct, return_index=True), wh, what + f" ({dt})")


def test_loss_random_shapes(monkeypatch):
    """Random grids / boxes / classes / batch sizes through every loss kernel (TMA ring, gather, stream - the stream kernel
    hands shapes whose tiles do not fit its shared memory to the gather kernel): terms against the C port, gradients
    bit-identical between the kernels."""
    from yolohot import loss as yl
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(int(os.environ.get("YH_FUZZ_SEED", 7)))
    for trial in range(int(os.environ.get("YH_FUZZ_TRIALS_LOSS", 40))):
        S = int(rng.integers(1, 15))
        B = int(rng.integers(1, 4))
        C = int(rng.choice([1, 3, 20, 80]))
        n = int(rng.choice([1, 2, 5, 33, 130, 1000]))
        if S * S * (C + 5 * B) * n > 3_000_000:
            n = max(1, 3_000_000 // (S * S * (C + 5 * B)))
        yt = F.synth_labels(n, S, B, C, seed=trial, lam=float(rng.choice([0.3, 2.5, 20.0])))
        yp = F.synth_loss_pred(yt.shape, seed=trial)
        want = cport.loss(yt, yp, C, B)
        td, pd = torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)
        grads = []
        for mode in (None, "0", "1", "2"):
            if mode is None:
                monkeypatch.delenv("YH_LOSS_GATHER", raising=False)
            else:
                monkeypatch.setenv("YH_LOSS_GATHER", mode)
            terms, g = yl.yolo_v1_loss_terms(td, pd, C, B, grad=True)
            np.testing.assert_allclose(terms.cpu().numpy(), want, rtol=1e-5, atol=1e-6,
                                       err_msg=f"trial {trial}: S={S} B={B} C={C} n={n} kernel={mode}")
            grads.append(g)
        monkeypatch.delenv("YH_LOSS_GATHER", raising=False)
        for g in grads[1:]:
            assert torch.equal(g, grads[0]), f"trial {trial}: S={S} B={B} C={C} n={n}: gradients differ between the kernels"


def test_map_random_rows():
    """mean_average_precision on random row sets (ties in confidence, duplicate detections on one ground truth,
    classes without ground truth or without detections, images without ground truth) against the C port."""
    from yolohot import utils as yu
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(int(os.environ.get("YH_FUZZ_SEED", 99)))
    for trial in range(int(os.environ.get("YH_FUZZ_TRIALS_MAP", 40))):
        C = int(rng.choice([1, 3, 20]))
        n_img = int(rng.choice([1, 5, 60, 400]))
        t_rows, p_rows = [], []
        for img in range(n_img):
            for _ in range(int(rng.integers(0, 4))):
                c = int(rng.integers(0, C))
                box = [rng.random(), rng.random(), 0.1 + 0.5 * rng.random(), 0.1 + 0.5 * rng.random()]
                t_rows.append([img, c, 1.0] + box)
                for _ in range(int(rng.integers(0, 4))):
                    jit = 0.03 * rng.standard_normal(4)
                    cc = c if rng.random() < 0.8 else int(rng.integers(0, C))
                    p_rows.append([img, cc, int(rng.integers(1, 9)) / 8.0] + list(np.array(box) + jit))
            if rng.random() < 0.3:                                          # a detection in an image / class without ground truth
                p_rows.append([img, int(rng.integers(0, C)), rng.random(), rng.random(), rng.random(), 0.2, 0.2])
        t = np.array(t_rows, F32).reshape(-1, 7)
        p = np.array(p_rows, F32).reshape(-1, 7)
        for thr in (0.5, 0.3):
            want, want_ap = cport.mean_average_precision(t, p, C, thr)
            got, ap = yu.mean_average_precision(torch.from_numpy(t).to(dev), torch.from_numpy(p).to(dev), C, thr, return_ap=True)
            assert abs(float(got) - float(want)) <= 1e-6, (trial, thr, float(got), float(want))
            np.testing.assert_allclose(ap.cpu().numpy(), want_ap, atol=1e-6)


def test_evaluator_random_shapes(monkeypatch):
    """MeanAveragePrecision over random grids / boxes / classes / batch splits: update_state as ONE kernel and as three
    launches must leave identical row buffers, records and ground-truth counts (bit for bit), result() on the counting path
    and on the radix passes must agree bit for bit, and rows and mAP must equal the oracle evaluator's (NumPy restatement
    of utils.py:459-496): rows bit-exact, mAP to 1e-6."""
    from oracle import yolo_oracle as O
    from yolohot import utils as yu
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(int(os.environ.get("YH_FUZZ_SEED", 4242)))
    for trial in range(int(os.environ.get("YH_FUZZ_TRIALS_EVAL", 24))):
        S = int(rng.integers(1, 9))
        B = int(rng.integers(1, 4))
        C = int(rng.choice([1, 3, 20, 80]))
        batches = [int(rng.choice([1, 2, 7, 19, 40])) for _ in range(int(rng.integers(1, 4)))]
        data = []
        for k, n in enumerate(batches):
            yt = F.synth_labels(n, S, B, C, seed=1000 * trial + k, lam=float(rng.choice([0.5, 2.5, 8.0])))
            yp = F.synth_map_pred(yt, B, C, seed=1000 * trial + k)
            if rng.random() < 0.3:                                       # tie-heavy confidences
                yp[..., C] = np.round(yp[..., C] * 8) / 8
            data.append((yt, yp))
        oe = O.MeanAveragePrecision(C, B)
        for yt, yp in data:
            oe.update_state(yt, yp)
        want_m = float(oe.result())
        got = {}
        for fused in ("1", "0"):
            monkeypatch.setenv("YH_EVAL_FUSED", fused)
            ev = yu.MeanAveragePrecision(C, B)
            for yt, yp in data:
                ev.update_state(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev))
            monkeypatch.delenv("YH_MAP_COUNT", raising=False)
            m_count = ev.result().clone()
            monkeypatch.setenv("YH_MAP_COUNT", "0")
            m_radix = ev.result().clone()
            monkeypatch.delenv("YH_MAP_COUNT")
            torch.cuda.synchronize()
            st = ev._st
            npred, ntrue = int(st["cursors"][0]), int(st["cursors"][1])
            got[fused] = (st["pred"][:npred].clone(), st["true"][:ntrue].clone(), st["rec"][:npred].clone(), st["gt"].clone(), m_count)
            what = f"trial {trial}: S={S} B={B} C={C} batches={batches} fused={fused}"
            assert torch.equal(m_count.view(torch.int32), m_radix.view(torch.int32)), what
            assert abs(float(m_count) - want_m) <= 1e-6, (what, float(m_count), want_m)
            assert np.array_equal(st["pred"][:npred].cpu().numpy(), oe.all_pred_boxes_variable.reshape(-1, 7)[:npred] if npred else np.zeros((0, 7), F32)), what
            assert np.array_equal(st["true"][:ntrue].cpu().numpy(), oe.all_true_boxes_variable.reshape(-1, 7)[:ntrue] if ntrue else np.zeros((0, 7), F32)), what
        monkeypatch.delenv("YH_EVAL_FUSED")
        for x, y in zip(got["1"], got["0"]):
            assert torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x, y.view(torch.int32) if y.dtype == torch.float32 else y), (trial, S, B, C)
