"""Single-process multi-GPU check of the C-ABI exchange step (yh_comm_init_all / yh_map_allgather),
run on an N-GPU box:   python tests/multigpu_capi_check.py [N]
One process drives all devices: every device evaluates its contiguous image shard (decode+NMS,
yh_map_match), the records are all-gathered in device order over NCCL, every device reduces
(yh_map_reduce) - and every device must return the single-GPU mAP of the whole set, bit for bit."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
from tests import fixtures as F  # noqa: E402
from yolohot import _lib  # noqa: E402
from yolohot import utils as yu  # noqa: E402


def main():
    ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    assert torch.cuda.device_count() >= ndev >= 2, "needs at least 2 GPUs"
    L = _lib.lib()
    n, CLS = 5000, 20
    yt = F.synth_labels(n, seed=11)
    yp = F.synth_map_pred(yt)
    # single-GPU value
    ev = yu.MeanAveragePrecision(CLS, 2)
    ev.update_state(torch.from_numpy(yt).cuda(0), torch.from_numpy(yp).cuda(0))
    m_single = float(ev.result())

    comm = C.c_void_p()
    _lib.check(L.yh_comm_init_all(ndev, None, C.byref(comm)), "yh_comm_init_all")
    keys, tps, gts, nrec = [], [], [], []
    for d in range(ndev):
        lo, hi = n * d // ndev, n * (d + 1) // ndev
        with torch.cuda.device(d):
            e = yu.MeanAveragePrecision(CLS, 2)
            e.img_idx = 0
            e.update_state(torch.from_numpy(yt[lo:hi]).cuda(d), torch.from_numpy(yp[lo:hi]).cuda(d))
            k, t, g = yu.map_match(e.all_true_boxes_variable, e.all_pred_boxes_variable, CLS, 0.5)
        keys.append(k); tps.append(t); gts.append(g); nrec.append(int(k.shape[0]))
    total = sum(nrec)
    out_k = [torch.empty(total, dtype=torch.int64, device=f"cuda:{d}") for d in range(ndev)]
    out_t = [torch.empty(total, dtype=torch.uint8, device=f"cuda:{d}") for d in range(ndev)]
    for d in range(ndev):
        torch.cuda.synchronize(d)
    arr = lambda ts: (C.c_void_p * ndev)(*[t.data_ptr() for t in ts])
    streams = (C.c_void_p * ndev)(*[torch.cuda.current_stream(d).cuda_stream for d in range(ndev)])
    nrec_c = (C.c_int64 * ndev)(*nrec)
    _lib.check(L.yh_map_allgather(comm, arr(keys), arr(tps), nrec_c, arr(gts), CLS, arr(out_k), arr(out_t), total, streams),
               "yh_map_allgather")
    vals = []
    for d in range(ndev):
        with torch.cuda.device(d):
            m, _ = yu.map_reduce(out_k[d], out_t[d], gts[d], CLS)
            vals.append(float(m))
    assert all(v == m_single for v in vals), (vals, m_single)
    assert all(torch.equal(out_k[d].cpu(), out_k[0].cpu()) for d in range(ndev))

    # the same exchange fused into stage 1 over peer-mapped memory (yh_map_match_p2p): no collective call
    p2p = int(L.yh_comm_p2p(comm))
    if p2p:
        f32 = lambda t: t.contiguous().float()
        pk = [torch.zeros(total, dtype=torch.int64, device=f"cuda:{d}") for d in range(ndev)]
        pt = [torch.full((total,), 7, dtype=torch.uint8, device=f"cuda:{d}") for d in range(ndev)]
        pg = [torch.zeros(CLS, dtype=torch.int32, device=f"cuda:{d}") for d in range(ndev)]
        rows = []
        for d in range(ndev):
            lo, hi = n * d // ndev, n * (d + 1) // ndev
            with torch.cuda.device(d):
                e = yu.MeanAveragePrecision(CLS, 2)
                e.img_idx = 0
                e.update_state(torch.from_numpy(yt[lo:hi]).cuda(d), torch.from_numpy(yp[lo:hi]).cuda(d))
                rows.append((f32(e.all_true_boxes_variable), f32(e.all_pred_boxes_variable)))
        for d in range(ndev):
            torch.cuda.synchronize(d)
        _lib.check(L.yh_comm_barrier(comm, streams), "yh_comm_barrier")
        off = 0
        for d in range(ndev):
            tr, pr = rows[d]
            assert pr.shape[0] == nrec[d]
            with torch.cuda.device(d):
                _lib.check(L.yh_map_match_p2p(comm, d, tr.data_ptr(), tr.shape[0], pr.data_ptr(), pr.shape[0], CLS, 0.5,
                                              arr(pk), arr(pt), off, arr(pg), streams[d]), "yh_map_match_p2p")
            off += nrec[d]
        _lib.check(L.yh_comm_barrier(comm, streams), "yh_comm_barrier")
        pvals = []
        for d in range(ndev):
            with torch.cuda.device(d):
                m, _ = yu.map_reduce(pk[d], pt[d], pg[d], CLS)
                pvals.append(float(m))
        assert all(v == m_single for v in pvals), (pvals, m_single)
        for d in range(ndev):
            assert torch.equal(pk[d].cpu(), out_k[0].cpu()) and torch.equal(pt[d].cpu(), out_t[0].cpu()), d
            assert torch.equal(pg[d].cpu(), gts[0].cpu()), d
        print(f"  peer-scatter path (yh_map_match_p2p, no collective): records and mAP identical on all {ndev} devices")
    else:
        print("  devices cannot peer-access each other: yh_map_match_p2p not checked")
    _lib.check(L.yh_comm_destroy(comm), "yh_comm_destroy")
    print(f"multigpu_capi_check ok: {ndev} devices in one process, {total} records, mAP {vals[0]:.9f} == single-GPU {m_single:.9f}")


if __name__ == "__main__":
    main()
