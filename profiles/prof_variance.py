"""Per-launch times of the headline kernel over 60 back-to-back launches (same setup as bench.py)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch
from yolohot import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
n = 1_000_000
g = torch.Generator(device=dev); g.manual_seed(2025)
p = torch.rand((n, 7, 7, 30), generator=g, device=dev)
boxes = torch.empty((n, 49, 6), device=dev); cnt = torch.empty((n,), device=dev, dtype=torch.int32)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
f = lambda: _lib.check(L.yh_decode_nms(p.data_ptr(), n, 7, 2, 20, 0.5, 0.4, boxes.data_ptr(), cnt.data_ptr(), None, st))
for _ in range(5): f()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(60)]
for a, b in ev:
    a.record(); f(); b.record()
torch.cuda.synchronize()
ts = [a.elapsed_time(b) for a, b in ev]
print(" ".join("%.3f" % t for t in ts))
print("mean %.4f min %.4f max %.4f" % (sum(ts) / len(ts), min(ts), max(ts)))
