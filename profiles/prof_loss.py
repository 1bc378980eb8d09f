"""Times yh_loss on BASELINE cfg3 (batch 4096, 72.25 MB) with 8 rotating buffer sets (578 MB > L2)."""
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import fixtures as F  # noqa: E402
from yolohot import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
n = int(os.environ.get("YH_PROF_BATCH", 4096))
yt0 = torch.from_numpy(F.synth_labels(n, seed=7)).to(dev)
yp0 = torch.from_numpy(F.synth_loss_pred(tuple(yt0.shape), seed=7)).to(dev)
sets = [(yt0.clone(), yp0.clone(), torch.empty_like(yp0)) for _ in range(8)]
terms = torch.empty(6, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rnd(g=True):
    for a, b, c in sets:
        _lib.check(L.yh_loss(a.data_ptr(), b.data_ptr(), n * 49, 2, 20, 5.0, 0.5, terms.data_ptr(), c.data_ptr() if g else None, sp))


for g in (True, False):
    for _ in range(3):
        rnd(g)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rnd(g); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 8)
    print("fwd+bwd" if g else "fwd    ", "us/launch mean %.1f min %.1f" % (1e3 * statistics.mean(ts), 1e3 * min(ts)),
          "GB/s %.0f" % ((3 if g else 2) * n * 5880 / min(ts) / 1e6), "loss", float(terms[5]))

# size-matched copy: a plain device-to-device copy moving the same 72.25 MB of traffic (36.1 MB read + 36.1 MB
# written), same rotation over 8 buffer sets - what the memory system delivers for a launch of this size
nbytes = 3 * n * 5880 // 2
srcs = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(8)]
dsts = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(8)]
for _ in range(3):
    for a, b in zip(srcs, dsts):
        b.copy_(a)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, b in zip(srcs, dsts):
        b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 8)
print("copy    us/launch mean %.1f min %.1f" % (1e3 * statistics.mean(ts), 1e3 * min(ts)), "GB/s %.0f" % (2 * nbytes / min(ts) / 1e6),
      "(same traffic as fwd+bwd)")
