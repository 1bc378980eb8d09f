"""Single-process multi-GPU check of the C-ABI exchange step, run on an N-GPU box:
    python tests/multigpu_capi_check.py [N]
One process drives all devices: every device evaluates its contiguous image shard (decode + NMS +
yh_eval_update), then the records are exchanged (a) as a collective (yh_comm_init_all /
yh_map_allgather over NCCL, then yh_map_reduce) and (b) by kernels over peer-mapped memory
(yh_map_exchange / yh_map_reduce_exchanged, no collective and no host synchronisation) - and every
device must return the single-GPU mAP of the whole set, bit for bit, on both paths."""
import ctypes as C
import os
import sys

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
    ev = yu.MeanAveragePrecision(CLS, 2)                      # single-GPU value
    ev.update_state(torch.from_numpy(yt).cuda(0), torch.from_numpy(yp).cuda(0))
    m_single = float(ev.result())

    comm = C.c_void_p()
    _lib.check(L.yh_comm_init_all(ndev, None, C.byref(comm)), "yh_comm_init_all")
    evs, recs, gts, nrec = [], [], [], []
    for d in range(ndev):
        lo, hi = n * d // ndev, n * (d + 1) // ndev
        with torch.cuda.device(d):
            e = yu.MeanAveragePrecision(CLS, 2)
            e.update_state(torch.from_numpy(yt[lo:hi]).cuda(d), torch.from_numpy(yp[lo:hi]).cuda(d))
            k = int(e._st["cursors"][0].item())
        evs.append(e); recs.append(e._st["rec"][:k].clone()); gts.append(e._st["gt"].clone()); nrec.append(k)
    total = sum(nrec)
    out = [torch.empty(total, dtype=torch.int64, device=f"cuda:{d}") for d in range(ndev)]
    for d in range(ndev):
        torch.cuda.synchronize(d)
    arr = lambda ts: (C.c_void_p * ndev)(*[t.data_ptr() for t in ts])
    streams = (C.c_void_p * ndev)(*[torch.cuda.current_stream(d).cuda_stream for d in range(ndev)])
    gsum = [g.clone() for g in gts]
    _lib.check(L.yh_map_allgather(comm, arr(recs), (C.c_int64 * ndev)(*nrec), arr(gsum), CLS, arr(out), total, streams),
               "yh_map_allgather")
    vals = []
    for d in range(ndev):
        with torch.cuda.device(d):
            vals.append(float(yu.map_reduce(out[d], gsum[d], CLS)[0]))
    assert all(v == m_single for v in vals), (vals, m_single)
    assert all(torch.equal(out[d].cpu(), out[0].cpu()) for d in range(ndev))

    if int(L.yh_comm_p2p(comm)):                               # kernel-level exchange over peer-mapped memory
        cap = max(nrec) + 100
        nbytes = int(L.yh_map_exchange_bytes(ndev, CLS, cap))
        bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in range(ndev)]
        wsb = int(L.yh_workspace_bytes(_lib.YH_OP_MAP_REDUCE, ndev * cap, 0, 0, CLS))
        ws = [torch.empty(wsb, dtype=torch.uint8, device=f"cuda:{d}") for d in range(ndev)]
        err = [torch.zeros(1, dtype=torch.int32, device=f"cuda:{d}") for d in range(ndev)]
        m = [torch.empty(1, device=f"cuda:{d}") for d in range(ndev)]
        for d in range(ndev):
            torch.cuda.synchronize(d)
        for epoch in (1, 2, 3):                                # buffer reuse, both parities
            for d in range(ndev):
                st = evs[d]._st
                with torch.cuda.device(d):
                    _lib.check(L.yh_map_exchange(ndev, d, arr(bufs), CLS, cap, st["rec"].data_ptr(), st["bound"],
                                                 st["cursors"].data_ptr(), st["gt"].data_ptr(), epoch, streams[d]), "yh_map_exchange")
            for d in range(ndev):
                with torch.cuda.device(d):
                    _lib.check(L.yh_map_reduce_exchanged(ndev, bufs[d].data_ptr(), CLS, cap, epoch, total, None, m[d].data_ptr(),
                                                         err[d].data_ptr(), ws[d].data_ptr(), wsb, streams[d]), "yh_map_reduce_exchanged")
            pv = [float(m[d].cpu()) for d in range(ndev)]
            assert all(v == m_single for v in pv) and all(int(e.cpu()) == 0 for e in err), (epoch, pv, m_single)
        print(f"  kernel-level exchange (yh_map_exchange, no collective): mAP identical on all {ndev} devices, 3 epochs")
    else:
        print("  devices cannot peer-access each other: yh_map_exchange not checked")
    _lib.check(L.yh_comm_destroy(comm), "yh_comm_destroy")
    print(f"multigpu_capi_check ok: {ndev} devices in one process, {total} records, mAP {vals[0]:.9f} == single-GPU {m_single:.9f}")


if __name__ == "__main__":
    main()
