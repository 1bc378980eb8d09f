"""Mid-size grids (64 < S*S <= 256, images <= 12 KB): one-warp-per-image tile ring (YH_COOP_MIN_BYTES=12288) against
the cooperative team kernel (default): python profiles/prof_midsize.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch  # noqa: E402

from yolohot import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (S, B, C, thr) in ((9, 2, 20, 0.4), (13, 2, 6, 0.4), (10, 1, 4, 0.3), (16, 1, 3, 0.4), (12, 2, 10, 0.2)):
    D = C + 5 * B
    n = int(2.0e9 // (4 * S * S * D))
    g = torch.Generator(device=dev); g.manual_seed(1)
    p = torch.rand((n, S, S, D), generator=g, device=dev)
    boxes = torch.empty((n, S * S, 6), device=dev); cnt = torch.empty((n,), device=dev, dtype=torch.int32)
    res = []
    for env in ("12288", None):
        if env is None:
            os.environ.pop("YH_COOP_MIN_BYTES", None)
        else:
            os.environ["YH_COOP_MIN_BYTES"] = env
        ts = []
        for i in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(L.yh_decode_nms(p.data_ptr(), n, S, B, C, 0.5, thr, boxes.data_ptr(), cnt.data_ptr(), None, st))
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        kept = int(cnt.sum())
        res.append((min(ts[1:]), kept))
    by = n * (4 * S * S * D + 4) + 24 * res[0][1]
    print(f"S={S} B={B} C={C} ({4 * S * S * D} B/image, {n} images, kept/img {res[0][1] / n:.1f}): tile ring {res[0][0]:.3f} ms = {by / res[0][0] / 1e6:.0f} GB/s, "
          f"cooperative {res[1][0]:.3f} ms = {by / res[1][0] / 1e6:.0f} GB/s, same counts {res[0][1] == res[1][1]}")
