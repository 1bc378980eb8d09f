"""update_state throughput by batch size (YH_PROF_BATCH, default 50,000 images; 1 M images per epoch), second epoch (buffers already sized), fused
one-launch path against the three-launch path (YH_EVAL_FUSED is read per call)."""
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
NB = int(os.environ.get("YH_PROF_BATCH", 50_000))
REPS = max(1, 1_000_000 // NB)
yt0 = F.synth_labels(NB, seed=11)
a, b = torch.from_numpy(yt0).to(dev), torch.from_numpy(F.synth_map_pred(yt0)).to(dev)
for fused in ("1", "0"):
    os.environ["YH_EVAL_FUSED"] = fused
    ev = yu.MeanAveragePrecision(20, 2)
    for epoch in range(3):
        ev.reset_states()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(REPS):
            ev.update_state(a, b)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        m = float(ev.result())
        print(f"fused={fused} batch {NB} epoch {epoch}: {REPS * NB} images in {1e3 * (t1 - t0):.2f} ms = {REPS * NB / (t1 - t0) / 1e6:.1f} M images/s, {1e6 * (t1 - t0) / REPS:.1f} us per update_state, mAP {m:.6f}", flush=True)
