"""Inputs of the reference's own ``__main__`` smoke blocks (the only fixtures it holds)
and the seeded synthetic generators of SURVEY.md section 8(d).  NumPy only."""
from __future__ import annotations

import numpy as np

F32 = np.float32


def utils_demo():
    """yolo_v1/utils.py:717-754 -> (y_true, y_pred) (1,7,7,13), C=3, B=2."""
    C = 3
    y_true = np.zeros((1, 7, 7, C + 10), F32)
    for (r, c, k) in ((0, 0, 0), (3, 3, 1), (6, 6, 2)):
        y_true[:, r, c, k] = 1
        y_true[:, r, c, C] = 1
        y_true[:, r, c, C + 1:C + 5] = [0.5, 0.5, 0.1, 0.1]
    y_pred = np.zeros((1, 7, 7, C + 10), F32)
    y_pred[:, 0, 0, :C] = [0.8, 0.5, 0.1]
    y_pred[:, 0, 0, C] = 0.6
    y_pred[:, 0, 0, C + 1:C + 5] = [0.49, 0.49, 0.1, 0.1]
    y_pred[:, 0, 0, C + 5] = 0.2
    y_pred[:, 0, 0, C + 6:C + 10] = [0.45, 0.45, 0.1, 0.1]
    y_pred[:, 3, 3, :C] = [0.2, 0.8, 0.1]
    y_pred[:, 3, 3, C] = 0.1
    y_pred[:, 3, 3, C + 1:C + 5] = [0.45, 0.45, 0.1, 0.1]
    y_pred[:, 3, 3, C + 5] = 0.9
    y_pred[:, 3, 3, C + 6:C + 10] = [0.49, 0.49, 0.1, 0.1]
    y_pred[:, 6, 6, :C] = [0.1, 0.5, 0.8]
    y_pred[:, 6, 6, C] = 0.6
    y_pred[:, 6, 6, C + 1:C + 5] = [0.49, 0.49, 0.1, 0.1]
    y_pred[:, 6, 6, C + 5] = 0.2
    y_pred[:, 6, 6, C + 6:C + 10] = [0.45, 0.45, 0.1, 0.1]
    return y_true, y_pred


def loss_demo():
    """yolo_v1/loss.py:219-234 -> (y_true, y_pred) (1,7,7,13); note :229 is overwritten by :230."""
    y_true = np.zeros((1, 7, 7, 13))
    y_true[:, 0, 0, 2] = 1
    y_true[:, 0, 0, 3] = 1
    y_true[:, 0, 0, 4:8] = (0.5, 0.5, 0.1, 0.1)
    y_pred = np.zeros((1, 7, 7, 13))
    y_pred[:, 0, 0, 2] = 0.6
    y_pred[:, 0, 0, 3] = 0.7
    y_pred[:, 0, 0, 4:8] = (0.49, 0.49, 0.09, 0.09)
    y_pred[:, 0, 0, 9] = 0.4
    y_pred[:, 0, 0, 9:13] = (0.45, 0.45, 0.09, 0.09)
    return y_true.astype(F32), y_pred.astype(F32)


def metric_demo():
    """yolo_v1/metric.py:103-140 -> (y_true, y_pred, y_pred_2) (1,7,7,30), C=20."""
    t = np.zeros((1, 7, 7, 30))
    for (r, c, k) in ((1, 1, 1), (4, 4, 5)):
        t[:, r, c, k] = 1
        t[:, r, c, 20] = 1
        t[:, r, c, 21:23] = 0.5
        t[:, r, c, 23:25] = 0.1
    p = np.zeros((1, 7, 7, 30))
    for (r, c, k, v) in ((1, 1, 1, 0.8), (4, 4, 5, 0.9)):
        p[:, r, c, k] = v
        p[:, r, c, 20] = 0.3
        p[:, r, c, 21:23] = 0.49
        p[:, r, c, 23:25] = 0.1
        p[:, r, c, 25] = 0.2
        p[:, r, c, 26:28] = 0.45
        p[:, r, c, 28:30] = 0.1
    p2 = np.zeros((1, 7, 7, 30))
    p2[:, 1, 1, 1] = 0.8
    p2[:, 1, 1, 20] = 0.9
    p2[:, 1, 1, 21:23] = 0.49
    p2[:, 1, 1, 23:25] = 0.1
    p2[:, 1, 1, 25] = 0.2
    p2[:, 1, 1, 26:28] = 0.45
    p2[:, 1, 1, 28:30] = 0.1
    return t.astype(F32), p.astype(F32), p2.astype(F32)


# yolo_v1/data/test.txt ("class cx cy w h"), restated as [cx, cy, w, h, cls] rows
TEST_TXT_BOXES = [
    [0.756250, 0.210417, 0.293750, 0.179167, 0],
    [0.450000, 0.480556, 0.582812, 0.505556, 1],
    [0.287891, 0.661806, 0.233594, 0.556944, 2],
]


# ----------------------------- synthetic configs ---------------------------- #
def synth_dense(n, S=7, B=2, C=20, seed=1234):
    """cfg1 / cfg2-dense: U[0,1) on every channel."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.random((n, S, S, C + 5 * B), dtype=F32)


def synth_sparse(n, S=7, B=2, C=20, seed=2025):
    """cfg2-sparse: confidences u**32 (about 5.6 % of cells pass 0.4), w,h = 0.05+0.45u."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.random((n, S, S, C + 5 * B), dtype=F32)
    for b in range(B):
        p[..., C + 5 * b] = p[..., C + 5 * b] ** 32
        p[..., C + 5 * b + 3:C + 5 * b + 5] = F32(0.05) + F32(0.45) * p[..., C + 5 * b + 3:C + 5 * b + 5]
    return p


def synth_quantised(n, S=7, B=2, C=20, seed=5, levels=8, dominant=4):
    """Tie-heavy stress: values quantised to 1/levels, few dominant classes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = (rng.integers(0, levels + 1, (n, S, S, C + 5 * B)) / levels).astype(F32)
    dom = rng.integers(0, dominant, (n, S, S))
    boost = rng.random((n, S, S)) < 0.8
    for k in range(dominant):
        p[..., k] = np.where(boost & (dom == k), F32(2.0), p[..., k])
    for b in range(B):
        p[..., C + 5 * b + 3:C + 5 * b + 5] = F32(0.1) + F32(0.5) * p[..., C + 5 * b + 3:C + 5 * b + 5]
    return p


def synth_stress(n, S=14, B=3, C=80, seed=99, dominant=4):
    """cfg5: dense survivors at conf_thr=0.05, ~80 % of argmaxes in `dominant` classes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.random((n, S, S, C + 5 * B), dtype=F32)
    dom = rng.integers(0, dominant, (n, S, S))
    boost = rng.random((n, S, S)) < 0.8
    for k in range(dominant):
        p[..., k] = np.where(boost & (dom == k), F32(1.5) + p[..., k], p[..., k])
    for b in range(B):
        p[..., C + 5 * b + 3:C + 5 * b + 5] = F32(0.1) + F32(0.5) * p[..., C + 5 * b + 3:C + 5 * b + 5]
    return p


def synth_labels(n, S=7, B=2, C=20, seed=7, lam=2.5):
    """cfg3/cfg4 y_true in the dataset.py:107-110 layout: Poisson(lam) objects per image
    (clipped to [1, S*S]) in distinct cells, conf=1, one-hot class, x,y~U[0,1), w,h~U[0.05,0.9)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.zeros((n, S, S, C + 5 * B), F32)
    cnt = np.clip(rng.poisson(lam, n), 1, S * S)
    for i in range(n):
        cells = rng.choice(S * S, size=int(cnt[i]), replace=False)
        for cell in cells:
            r, c = divmod(int(cell), S)
            t[i, r, c, rng.integers(0, C)] = 1
            t[i, r, c, C] = 1
            t[i, r, c, C + 1:C + 3] = rng.random(2, dtype=F32)
            t[i, r, c, C + 3:C + 5] = F32(0.05) + F32(0.85) * rng.random(2, dtype=F32)
    return t


def synth_loss_pred(shape, seed=7):
    """cfg3 y_pred ~ N(0.3, 0.3) on every channel (sign-varying w,h)."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    return (F32(0.3) + F32(0.3) * rng.standard_normal(shape, dtype=F32)).astype(F32)


def synth_map_pred(y_true, B=2, C=20, seed=11):
    """cfg4 y_pred = jittered y_true + background confidences."""
    rng = np.random.Generator(np.random.PCG64(seed + 2000))
    t = y_true
    n, S = t.shape[0], t.shape[1]
    obj = t[..., C] > 0
    p = np.zeros_like(t)
    p[..., :C] = t[..., :C] * F32(0.6) + F32(0.5) * rng.random(t.shape[:-1] + (C,), dtype=F32)
    flip = obj & (rng.random((n, S, S)) < 0.1)
    fl_cls = rng.integers(0, C, (n, S, S))
    ii, rr, cc = np.nonzero(flip)
    p[ii, rr, cc, fl_cls[ii, rr, cc]] += F32(1.0)
    for b in range(B):
        jit = (F32(0.08) * rng.standard_normal(t.shape[:-1] + (4,), dtype=F32)).astype(F32)
        p[..., C + 5 * b + 1:C + 5 * b + 5] = np.where(obj[..., None], t[..., C + 1:C + 5] + jit,
                                                      rng.random(t.shape[:-1] + (4,), dtype=F32))
    u0 = rng.random((n, S, S), dtype=F32)
    u1 = rng.random((n, S, S), dtype=F32)
    bg = rng.random((n, S, S), dtype=F32)
    p[..., C] = np.where(obj, F32(0.2) + F32(0.8) * u0, F32(0.55) * bg)
    if B > 1:
        p[..., C + 5] = np.where(obj, F32(0.6) * u1, F32(0.3) * bg)
    for b in range(2, B):
        p[..., C + 5 * b] = F32(0.1) * bg
    return p.astype(F32)
