"""Multi-GPU check, run under torchrun on an N-GPU box (NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
Each rank evaluates ITS contiguous image shard (decode+NMS, loss, mAP matching on its own GPU);
only the mAP records cross NVLink (yolohot.dist).  Asserts, on every rank, that the sharded results
equal the single-GPU results computed on rank-local copies of the whole input."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
from tests import fixtures as F  # noqa: E402
from yolohot import dist as yd  # noqa: E402
from yolohot import loss as yl  # noqa: E402
from yolohot import utils as yu  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 5000                                                   # BASELINE cfg4
    yt = F.synth_labels(n, seed=11)
    yp = F.synth_map_pred(yt)
    lo, hi = yd.shard_range(n)
    # sharded evaluator: result() all-gathers the records and returns the global mAP on every rank
    ev = yu.MeanAveragePrecision(20, 2)
    ev.update_state(torch.from_numpy(yt[lo:hi]).to(dev), torch.from_numpy(yp[lo:hi]).to(dev))
    m_sharded = float(ev.result())
    # single-GPU value of the whole set (temporarily without the process group's view: local match+reduce)
    e1 = yu.MeanAveragePrecision(20, 2)
    e1.update_state(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev))
    k, t, g = yu.map_match(e1.all_true_boxes_variable, e1.all_pred_boxes_variable, 20, 0.5)
    m_single = float(yu.map_reduce(k, t, g, 20)[0])
    assert m_sharded == m_single, (rank, m_sharded, m_single)
    # decode+NMS and loss need no communication: shard results are slices of the whole
    b_all, c_all = yu.decode_nms(torch.from_numpy(yp).to(dev), 20, 2)
    b_sh, c_sh = yu.decode_nms(torch.from_numpy(yp[lo:hi]).to(dev), 20, 2)
    assert torch.equal(c_sh, c_all[lo:hi])
    terms = yl.yolo_v1_loss_terms(torch.from_numpy(yt[lo:hi]).to(dev), torch.from_numpy(yp[lo:hi]).to(dev)).double()
    dist.all_reduce(terms)                                      # optional scalar all-reduce
    whole = yl.yolo_v1_loss_terms(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)).double()
    assert torch.allclose(terms, whole, rtol=2e-6), (terms, whole)
    vals = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(vals, torch.tensor([m_sharded], device=dev, dtype=torch.float64))
    assert all(float(v) == m_sharded for v in vals)
    if rank == 0:
        print(f"multigpu_check ok: world={world} mAP sharded={m_sharded:.9f} single={m_single:.9f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
