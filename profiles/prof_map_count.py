"""result() alone (the reduce stage, one launch) at several accumulated sizes, counting path against the radix passes:
python profiles/prof_map_count.py [images ...]   (YH_MAP_COUNT / YH_MAP_COUNT_BUCKETS / YH_MAP_COUNT_GRID are read per call)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch  # noqa: E402

from tests import fixtures as F  # noqa: E402
from yolohot import utils as yu  # noqa: E402

dev = torch.device("cuda:0")
sizes = [int(v) for v in sys.argv[1:]] or [5000, 20000, 50000, 100000, 200000]
yt0 = F.synth_labels(5000, seed=11)
a, b = torch.from_numpy(yt0).to(dev), torch.from_numpy(F.synth_map_pred(yt0)).to(dev)


def timed(ev, reps=30):
    for _ in range(3):
        m = ev.result()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m = ev.result()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, float(m)


for n in sizes:
    ev = yu.MeanAveragePrecision(20, 2)
    for lo in range(0, n, 5000):
        k = min(5000, n - lo)
        ev.update_state(a[:k], b[:k])
    nrec = int(ev._st["cursors"][0])
    out = []
    for label, env in (("radix", {"YH_MAP_COUNT": "0"}), ("count16", {}), ("count4", {"YH_MAP_COUNT_BUCKETS": "4"}),
                       ("count1", {"YH_MAP_COUNT_BUCKETS": "1"})):
        for k in ("YH_MAP_COUNT", "YH_MAP_COUNT_BUCKETS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        os.environ["YH_MAP_COUNT_PAIRS"] = str(10**15)
        os.environ["YH_MAP_COUNT_NMAX"] = str(10**9)
        us, m = timed(ev)
        out.append(f"{label} {us:.1f} us (mAP {m:.6f})")
    print(f"{n} images, {nrec} records: " + ", ".join(out), flush=True)
