"""mAP pipeline: python profiles/prof_map.py [cfg4|big]
cfg4: BASELINE configs[3] - 5,000 images through MeanAveragePrecision.update_state + result(), 10 passes
      (4 launches per pass: decode+NMS x2, yh_eval_update, yh_map_reduce; no host synchronisation);
big:  YH_PROF_IMAGES images (default 1,000,000) accumulated in 50k-image batches, then result() alone."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import torch  # noqa: E402

from tests import fixtures as F  # noqa: E402
from yolohot import launch_count, utils as yu  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
if mode == "cfg4":
    yt0 = F.synth_labels(5000, seed=11)
    a, b = torch.from_numpy(yt0).to(dev), torch.from_numpy(F.synth_map_pred(yt0)).to(dev)
    ev = yu.MeanAveragePrecision(20, 2)

    def one():
        ev.reset_states()
        ev.update_state(a, b)
        return ev.result()
    for _ in range(3):
        m = one()
    torch.cuda.synchronize()
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(10):
        m = one()
    e1.record()
    torch.cuda.synchronize()
    print(f"cfg4: update_state + result() {1e3 * (time.perf_counter() - t0) / 10:.4f} ms wall, {e0.elapsed_time(e1) / 10:.4f} ms on the device, "
          f"{(launch_count() - l0) / 10:.1f} launches per pass, mAP {float(m):.6f}, {int(ev._st['cursors'][0])} records")
else:
    n = int(os.environ.get("YH_PROF_IMAGES", 1_000_000))
    yt0 = F.synth_labels(50_000, seed=11)
    a, b = torch.from_numpy(yt0).to(dev), torch.from_numpy(F.synth_map_pred(yt0)).to(dev)
    ev = yu.MeanAveragePrecision(20, 2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for lo in range(0, n, 50_000):
        k = min(50_000, n - lo)
        ev.update_state(a[:k], b[:k])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ev.result()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m = ev.result(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    nrec = int(ev._st["cursors"][0])
    print(f"big: {n} images, update_state total {1e3 * (t1 - t0):.2f} ms ({n / (t1 - t0) / 1e6:.1f} M images/s), result() {min(ts):.3f} ms "
          f"(min of 5) over {nrec} records = {8 * nrec * 3 * 5 / (min(ts) * 1e-3) / 1e9:.0f} GB/s of sort traffic (3 x 8 B x 5 passes), mAP {float(m):.6f}")
