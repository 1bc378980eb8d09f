"""Times the adapters either side of the hot path on one GPU, CUDA events."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from yolohot import _lib, dataset as yd, utils as yu  # noqa: E402
from yolohot._tensor import head_to_f32, stream_ptr  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.mean(ts), min(ts)


n = int(os.environ.get("YH_PROF_IMAGES", 1_000_000))
rng = np.random.Generator(np.random.PCG64(3))
cnt = np.clip(rng.poisson(2.5, n), 1, 49)
offs = np.zeros(n + 1, np.int64); offs[1:] = np.cumsum(cnt)
tot = int(offs[-1])
boxes = np.concatenate([rng.random((tot, 2)), 0.05 + 0.85 * rng.random((tot, 2)), rng.integers(0, 20, (tot, 1))], 1)
b_d = torch.from_numpy(boxes).to(dev); o_d = torch.from_numpy(offs).to(dev)
out = torch.empty((n, 7, 7, 30), device=dev)
bad = torch.zeros(1, dtype=torch.int32, device=dev)
sp = stream_ptr(dev)
f = lambda: _lib.check(L.yh_encode_labels(b_d.data_ptr(), o_d.data_ptr(), n, 7, 2, 20, out.data_ptr(), bad.data_ptr(), sp))
ms, mn = timed(f)
by = n * 5880 + tot * 40 + (n + 1) * 8
print(f"encode_labels  {n} images, {tot} boxes: {ms:.3f} ms (min {mn:.3f}) = {by / ms / 1e6:.0f} GB/s, {n / ms / 1e3:.1f} M images/s")
h = torch.rand((n, 7, 7, 30), device=dev).to(torch.bfloat16)
ms, mn = timed(lambda: head_to_f32(h))
print(f"head_to_f32    bf16 {h.numel()} values: {ms:.3f} ms = {h.numel() * 6 / ms / 1e6:.0f} GB/s (includes torch.empty)")
rows = torch.rand((n, 49, 6), device=dev); c = torch.randint(0, 49, (n,), device=dev, dtype=torch.int32)
px = torch.empty((n, 49, 4), dtype=torch.int32, device=dev)
f = lambda: _lib.check(L.yh_pixel_boxes(rows.data_ptr(), c.data_ptr(), n, 49, 448, 448, px.data_ptr(), sp))
ms, mn = timed(f)
print(f"pixel_boxes    {n * 49} rows: {ms:.3f} ms (min {mn:.3f})")
