"""Context numbers for BASELINE.json configs 4 and 5 (not the bench headline):
  cfg4  VOC-style mAP@0.5 over 5k synthetic images, sharded over the ranks with the NCCL all-gather
  cfg5  stress S=14 B=3 C=80, conf_thr 0.05 (dense survivors), per-GPU image shards
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
          --master-port 29520 profiles/bench_configs.py          (or plain python for N=1)
Rank 0 prints one JSON line per config; times are CUDA events, max over ranks."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
from tests import fixtures as F  # noqa: E402
from yolohot import dist as yd  # noqa: E402
from yolohot import utils as yu  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rank = yd.rank()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---------------- cfg4: mAP over 5k images ----------------
    n = 5000
    yt = F.synth_labels(n, seed=11)
    yp = F.synth_map_pred(yt)
    lo, hi = yd.shard_range(n)
    a, b = torch.from_numpy(yt[lo:hi]).to(dev), torch.from_numpy(yp[lo:hi]).to(dev)

    def run():
        ev = yu.MeanAveragePrecision(20, 2)
        ev.update_state(a, b)
        return ev.result()

    for _ in range(3):
        m = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        m = run()
    mval = float(m)
    dt = maxr((time.perf_counter() - t0) / reps)
    if rank == 0:
        print(json.dumps({"config": "cfg4 mAP@0.5, 5000 images, evaluator update+result (host-timed, includes the "
                                    "all-gather when n_gpus>1)", "n_gpus": world, "ms": dt * 1e3, "mAP": mval,
                          "images_per_s": n / dt}), flush=True)

    # ---------------- cfg5: stress decode+NMS ----------------
    S, B, C = 14, 3, 80
    n5 = int(os.environ.get("YH_CFG5_IMAGES", 131072))
    g = torch.Generator(device=dev)
    g.manual_seed(99 + rank)
    p = torch.rand((n5, S, S, C + 5 * B), generator=g, device=dev)
    dom = torch.randint(0, 4, (n5, S, S), generator=g, device=dev)
    boost = torch.rand((n5, S, S), generator=g, device=dev) < 0.8
    for k in range(4):
        p[..., k] += 1.5 * (boost & (dom == k)).float()
    for bb in range(B):
        p[..., C + 5 * bb + 3:C + 5 * bb + 5] = 0.1 + 0.5 * p[..., C + 5 * bb + 3:C + 5 * bb + 5]
    del dom, boost
    boxes = torch.empty((n5, S * S, 6), device=dev)
    cnt = torch.empty((n5,), device=dev, dtype=torch.int32)
    for _ in range(2):
        yu.decode_nms(p, C, B, 0.5, 0.05, out=(boxes, cnt))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yu.decode_nms(p, C, B, 0.5, 0.05, out=(boxes, cnt))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = maxr(sum(ts) / len(ts))
    kept = float(cnt.sum())
    bytes_ = n5 * (4 * S * S * (C + 5 * B) + 4) + 24 * kept
    if rank == 0:
        print(json.dumps({"config": f"cfg5 stress S=14 B=3 C=80 conf_thr 0.05, {n5} images per GPU (9.76 GB), 4 dominant classes",
                          "n_gpus": world, "ms": ms, "images_per_s": world * n5 / (ms * 1e-3),
                          "GBps_per_gpu": bytes_ / (ms * 1e-3) / 1e9, "frac_hbm": bytes_ / (ms * 1e-3) / 1e9 / peak(),
                          "kept_per_image": kept / n5}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
