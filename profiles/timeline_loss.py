"""Per-CTA timeline of loss_gather_kernel at cfg3 (build csrc/yh_loss.cu with -DYH_LOSS_TIMELINE, run with YH_LOSS_DBG=1)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200")); sys.path.insert(0, ROOT)
import torch
from tests import fixtures as F
from yolohot import _lib
L = _lib.lib(); dev = torch.device("cuda:0"); n = 4096
yt0 = torch.from_numpy(F.synth_labels(n, seed=7)).to(dev)
yp0 = torch.from_numpy(F.synth_loss_pred(tuple(yt0.shape), seed=7)).to(dev)
sets = [(yt0.clone(), yp0.clone(), torch.empty_like(yp0)) for _ in range(8)]
terms = torch.empty(6, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(3):
    for a, b, c in sets:
        _lib.check(L.yh_loss(a.data_ptr(), b.data_ptr(), n * 49, 2, 20, 5.0, 0.5, terms.data_ptr(), c.data_ptr(), sp))
torch.cuda.synchronize()
