"""Throughput of the reference's SEPARATE calls on the device path (yh_decode, yh_nms on decoded rows, yh_iou),
next to the fused kernel: python profiles/prof_separate_calls.py"""
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch  # noqa: E402

from yolohot import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
n = int(os.environ.get("YH_PROF_IMAGES", 1_000_000))
g = torch.Generator(device=dev); g.manual_seed(2025)
p = torch.rand((n, 7, 7, 30), generator=g, device=dev)
rows = torch.empty((n, 49, 6), device=dev)
kept = torch.empty((n, 49, 6), device=dev)
cnt = torch.empty((n,), device=dev, dtype=torch.int32)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.mean(ts)


ms = timed(lambda: _lib.check(L.yh_decode(p.data_ptr(), n, 7, 2, 20, rows.data_ptr(), sp)))
print(f"yh_decode      {ms:.3f} ms  {(n * (5880 + 1176)) / ms / 1e6:.0f} GB/s")
ms = timed(lambda: _lib.check(L.yh_nms(rows.data_ptr(), n, 49, 0.5, 0.4, kept.data_ptr(), cnt.data_ptr(), None, sp)))
k = int(cnt.sum())
print(f"yh_nms (rows)  {ms:.3f} ms  {(n * (1176 + 4) + 24 * k) / ms / 1e6:.0f} GB/s  kept/img {k / n:.2f}")
ms = timed(lambda: _lib.check(L.yh_decode_nms(p.data_ptr(), n, 7, 2, 20, 0.5, 0.4, kept.data_ptr(), cnt.data_ptr(), None, sp)))
print(f"yh_decode_nms  {ms:.3f} ms  {(n * (5880 + 4) + 24 * k) / ms / 1e6:.0f} GB/s")
m = 50_000_000
a4 = torch.rand((m, 4), generator=g, device=dev); b4 = torch.rand((m, 4), generator=g, device=dev); o = torch.empty((m,), device=dev)
ms = timed(lambda: _lib.check(L.yh_iou(a4.data_ptr(), b4.data_ptr(), m, o.data_ptr(), sp)))
print(f"yh_iou         {ms:.3f} ms  {m * 36 / ms / 1e6:.0f} GB/s  ({m} pairs)")
