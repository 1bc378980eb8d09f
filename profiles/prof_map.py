"""mAP pipeline at scale: evaluator update (decode+NMS of y_true and y_pred, row append) and result
(match + reduce) for YH_PROF_IMAGES images: python profiles/prof_map.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch  # noqa: E402

from tests import fixtures as F  # noqa: E402
from yolohot import utils as yu  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("YH_PROF_IMAGES", 100_000))
base = 5000
yt0 = F.synth_labels(base, seed=11)
yp0 = F.synth_map_pred(yt0)
reps = (n + base - 1) // base
yt = torch.from_numpy(yt0).to(dev).repeat(reps, 1, 1, 1)[:n].contiguous()
yp = torch.from_numpy(yp0).to(dev).repeat(reps, 1, 1, 1)[:n].contiguous()


def run():
    ev = yu.MeanAveragePrecision(20, 2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for lo in range(0, n, 65536):
        ev.update_state(yt[lo:lo + 65536], yp[lo:lo + 65536])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    m = float(ev.result())
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return t1 - t0, t2 - t1, m, ev.all_pred_boxes_variable.shape[0], ev.all_true_boxes_variable.shape[0]


run()
u, r, m, npred, ngt = run()
print(f"{n} images: update_state {u * 1e3:.2f} ms, result {r * 1e3:.2f} ms, mAP {m:.6f}, {npred} detections, {ngt} ground truths; "
      f"{n / (u + r) / 1e6:.2f} M images/s end to end, result stage {28 * (npred + ngt) / r / 1e9:.1f} GB/s of rows")
