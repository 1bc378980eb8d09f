"""Multi-GPU check, run under torchrun on an N-GPU box (NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
Each rank evaluates ITS contiguous image shard (decode+NMS, loss, mAP matching on its own GPU);
only the mAP records cross NVLink (yolohot.dist).  Asserts, on every rank, that the sharded results
equal the single-GPU results computed on rank-local copies of the whole input."""
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
    yt_d, yp_d = torch.from_numpy(yt[lo:hi]).to(dev), torch.from_numpy(yp[lo:hi]).to(dev)
    # single-GPU value of the whole set: a default (local, like the reference) evaluator on this rank
    e1 = yu.MeanAveragePrecision(20, 2)
    e1.update_state(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev))
    m_single = float(e1.result())
    # sharded evaluator: result() exchanges the records and returns the global mAP on every rank
    ev = yu.MeanAveragePrecision(20, 2, sharded=True)
    ev.update_state(yt_d, yp_d)
    m_sharded = float(ev.result())
    assert m_sharded == m_single, (rank, m_sharded, m_single)
    ex = yd.peer_exchange(dev, 20, 0)
    p2p = ex is not None
    if os.environ.get("YH_DIST_P2P", "1") != "0":
        assert p2p, "PeerExchange unavailable on this box"
    for rnd in range(4):                                       # buffer reuse across epochs (both parities), streaming updates
        ev.reset_states()
        for a in range(0, hi - lo, 997):
            ev.update_state(yt_d[a:a + 997], yp_d[a:a + 997])
        assert float(ev.result()) == m_single, (rank, rnd)
    if p2p:
        assert ex.error() == 0
    # the padded all-gather fallback gives the same records, hence the same value
    nrec = int(ev._st["cursors"][0].item())
    rec_all, gt_all = yd.gather_records(ev._st["rec"][:nrec], ev._st["gt"])
    assert float(yu.map_reduce(rec_all, gt_all, 20)[0]) == m_single
    # the function form on this rank's rows, sharded
    m_fn = float(yu.mean_average_precision(ev.all_true_boxes_variable, ev.all_pred_boxes_variable, 20, sharded=True))
    assert m_fn == m_single, (rank, m_fn, m_single)
    # and the default stays local (reference behaviour): a rank-0-only call must not hang
    if rank == 0:
        float(yu.mean_average_precision(ev.all_true_boxes_variable, ev.all_pred_boxes_variable, 20))

    def lat(fn, reps=20):                                      # wall clock around a device sync, max over ranks
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

    def upd_res():
        ev.reset_states()
        ev.update_state(yt_d, yp_d)
        return ev.result()
    ms_peer = lat(upd_res)
    ms_nccl = lat(lambda: yu.map_reduce(*yd.gather_records(ev._st["rec"][:nrec], ev._st["gt"]), 20))
    if rank == 0:
        print(f"  world {world}, {rec_all.shape[0]} records: update_state + sharded result() {ms_peer:.3f} ms; "
              f"exchange + reduce through padded NCCL all-gathers alone {ms_nccl:.3f} ms")
    # an evaluator whose shard outgrows the exchange capacity: every rank gets NaN, none hangs
    if p2p:
        small = yd.PeerExchange(dev, 20, 1024)
        m_bad, _ = small.exchange_reduce(ev._st["rec"][:ev._st["bound"]], ev._st["cursors"][0:1], ev._st["gt"])
        assert torch.isnan(m_bad).item() and small.error() != 0
        small.close()
    # decode+NMS and loss need no communication: shard results are slices of the whole
    b_all, c_all = yu.decode_nms(torch.from_numpy(yp).to(dev), 20, 2)
    b_sh, c_sh = yu.decode_nms(yp_d, 20, 2)
    assert torch.equal(c_sh, c_all[lo:hi])
    terms = yl.yolo_v1_loss_terms(yt_d, yp_d).double()
    dist.all_reduce(terms)                                      # optional scalar all-reduce
    whole = yl.yolo_v1_loss_terms(torch.from_numpy(yt).to(dev), torch.from_numpy(yp).to(dev)).double()
    assert torch.allclose(terms, whole, rtol=2e-6), (terms, whole)
    vals = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(vals, torch.tensor([m_sharded], device=dev, dtype=torch.float64))
    assert all(float(v) == m_sharded for v in vals)
    if rank == 0:
        print(f"multigpu_check ok: world={world} mAP sharded={m_sharded:.9f} single={m_single:.9f} "
              f"exchange={'kernel-level (CUDA IPC + NVLink)' if p2p else 'NCCL all-gather'} records={rec_all.shape[0]}")
    yd.shutdown()                                               # unmap / free the IPC exchange buffers (collective)
    assert yd.peer_exchange(dev, 20, 1000) is not None or os.environ.get("YH_DIST_P2P", "1") == "0"   # and they can be set up again
    yd.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
