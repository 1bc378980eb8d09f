"""Host-side cost of one evaluator pass (cfg4): cProfile over 300 passes of reset_states / update_state / result."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch  # noqa: E402

from tests import fixtures as F  # noqa: E402
from yolohot import utils as yu  # noqa: E402

dev = torch.device("cuda:0")
yt0 = F.synth_labels(5000, seed=11)
a, b = torch.from_numpy(yt0).to(dev), torch.from_numpy(F.synth_map_pred(yt0)).to(dev)
ev = yu.MeanAveragePrecision(20, 2)


def one():
    ev.reset_states()
    ev.update_state(a, b)
    return ev.result()


for _ in range(20):
    one()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    one()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per pass {1e6 * (t1 - t0) / 300:.1f} us (launch side only), {1e6 * (t2 - t0) / 300:.1f} us until the device is done")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    one()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
