"""Short driver for ncu: a few launches of the fused decode+NMS kernel on the cfg2 workload
(1M dense images unless YH_PROF_IMAGES is set).  `python profiles/prof_decode_nms.py [dense|sparse|stress|stress5]` (stress = uniform classes, stress5 = cfg5 data)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch  # noqa: E402

from yolohot import _lib  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "dense"
n = int(os.environ.get("YH_PROF_IMAGES", 1_000_000))
S, B, C, thr = 7, 2, 20, 0.4
if mode in ("stress", "stress5"):
    S, B, C, thr = 14, 3, 80, 0.05
    n = int(os.environ.get("YH_PROF_IMAGES", 65_536))
D = C + 5 * B
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(2025)
p = torch.rand((n, S, S, D), generator=g, device=dev)
if mode == "stress5":        # cfg5 as specified: ~80 % of the argmaxes in 4 dominant classes, w,h in [0.1, 0.6)
    dom = torch.randint(0, 4, (n, S, S), generator=g, device=dev)
    boost = torch.rand((n, S, S), generator=g, device=dev) < 0.8
    for k in range(4):
        p[..., k] += 1.5 * (boost & (dom == k)).float()
    for b in range(B):
        p[..., C + 5 * b + 3:C + 5 * b + 5] = 0.1 + 0.5 * p[..., C + 5 * b + 3:C + 5 * b + 5]
    del dom, boost
if mode == "sparse":
    for b in range(B):
        p[..., C + 5 * b] = p[..., C + 5 * b] ** 32
boxes = torch.empty((n, S * S, 6), device=dev)
cnt = torch.empty((n,), device=dev, dtype=torch.int32)
L = _lib.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = int(os.environ.get("YH_PROF_REPS", 5))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    _lib.check(L.yh_decode_nms(p.data_ptr(), n, S, B, C, 0.5, thr, boxes.data_ptr(), cnt.data_ptr(), None, st))
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
kept = int(cnt.sum())
print(f"{mode}: n={n} S={S} B={B} C={C} ms={['%.3f' % m for m in ms]} kept/img={kept / n:.2f} "
      f"GB/s(best)={(n * (4 * S * S * D + 4) + 24 * kept) / min(ms) / 1e6:.0f}")

# half-precision heads (YH_PROF_HALF=1): the same workload as bfloat16 / float16 through yh_decode_nms_typed
if os.environ.get("YH_PROF_HALF"):
    for dt, code in ((torch.bfloat16, _lib.YH_DTYPE_BF16), (torch.float16, _lib.YH_DTYPE_F16)):
        ph = p.to(dt)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            _lib.check(L.yh_decode_nms_typed(ph.data_ptr(), code, n, S, B, C, 0.5, thr, 0, boxes.data_ptr(), cnt.data_ptr(), None, st))
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        kept = int(cnt.sum())
        print(f"{mode} {dt}: ms={['%.3f' % m for m in ms]} kept/img={kept / n:.2f} "
              f"GB/s(best)={(n * (2 * S * S * D + 4) + 24 * kept) / min(ms) / 1e6:.0f} images/s(best)={n / min(ms) / 1e3:.1f} M")
