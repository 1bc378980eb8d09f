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
    # the two exchange implementations (yolohot.dist): records stored into the peers' buffers by the matching kernel
    # (CUDA IPC + NVLink, the default on one box) against the padded NCCL all-gathers - record for record
    tr, pr = ev.all_true_boxes_variable, ev.all_pred_boxes_variable
    k, t, g = yu.map_match(tr, pr, 20, 0.5)
    k_n, t_n, g_n = yd.gather_records(k, t, g)
    ex = yd.peer_exchange(dev)
    p2p = ex is not None
    if os.environ.get("YH_DIST_P2P", "1") != "0":
        assert p2p, "PeerExchange unavailable on this box"
    if p2p:
        for rnd in range(3):                                    # buffer reuse across calls
            k_p, t_p, g_p = ex.match_gather(tr, pr, 20, 0.5)
            assert torch.equal(k_p, k_n) and torch.equal(t_p, t_n) and torch.equal(g_p, g_n), (rank, rnd)
        # latency of the two paths (stage 1 + exchange), wall clock around a device sync, max over ranks
        import time

        def lat(fn, reps=20):
            fn()
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize(dev)
            v = torch.tensor([(time.perf_counter() - t0) / reps], device=dev, dtype=torch.float64)
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            return float(v) * 1e3
        ms_peer = lat(lambda: ex.match_gather(tr, pr, 20, 0.5))
        ms_nccl = lat(lambda: yd.gather_records(*yu.map_match(tr, pr, 20, 0.5)))
        if rank == 0:
            print(f"  stage 1 + exchange, {k_n.shape[0]} records, world {world}: peer stores {ms_peer:.3f} ms, "
                  f"padded NCCL all-gathers {ms_nccl:.3f} ms")
        small = yd.PeerExchange(dev, capacity=1024)             # forced collective growth of the buffers
        k_p, t_p, g_p = small.match_gather(tr, pr, 20, 0.5)
        assert small.capacity >= k_n.shape[0] and torch.equal(k_p, k_n) and torch.equal(t_p, t_n) and torch.equal(g_p, g_n)
        m_small = float(yu.map_reduce(k_p, t_p, g_p, 20)[0])
        assert m_small == m_single
        small.close()
        empty = ex.match_gather(tr[:0], pr[:0], 20, 0.5)        # no ground truth and no detections anywhere
        assert empty[0].shape[0] == 0 and int(empty[2].sum()) == 0
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
        print(f"multigpu_check ok: world={world} mAP sharded={m_sharded:.9f} single={m_single:.9f} "
              f"exchange={'peer stores (CUDA IPC) == NCCL all-gather' if p2p else 'NCCL all-gather'} records={k_n.shape[0]}")
    yd.shutdown()                                               # unmap / free the IPC exchange buffers (collective)
    assert yd.peer_exchange(dev) is not None or os.environ.get("YH_DIST_P2P", "1") == "0"   # and they can be set up again
    yd.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
