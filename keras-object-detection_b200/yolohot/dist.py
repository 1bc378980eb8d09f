"""Multi-GPU plumbing of the mAP reduction: one process per GPU (torch.distributed, NCCL over
NVLink on a B200 box; gloo in the CPU tests).  The decode / NMS / loss / matching kernels
need no communication - images are sharded by contiguous range - so the only exchange is
  (1) the per-class ground-truth counts (C int32 per rank, summed), and
  (2) the per-detection records (one packed uint64: class, confidence, TP flag), concatenated IN RANK
      ORDER (rank order == image order, which the stable sort of stage 2 relies on for equal confidences).
Volume is 8 B x detections (a few MB at most): latency-bound.

Two implementations:
  PeerExchange     ranks of ONE box (NCCL group, NVLink / NVSwitch): every rank owns an exchange buffer that its
                   peers map through CUDA IPC; an exchange is two kernels and nothing else - yh_map_exchange stores
                   the rank's records, their count and its GT counts into every peer's buffer and releases a flag;
                   the reduce kernel (yh_map_reduce_exchanged) waits on the flags inside the kernel.  No NCCL call,
                   no host synchronisation, no size exchange on the host: torch.distributed is used once, to swap
                   the IPC handles.
  gather_records   anything else (gloo, several hosts, YH_DIST_P2P=0): padded all-gathers through
                   torch.distributed, sizes through the host."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._tensor import on_device


def world_size(group=None):
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def rank(group=None):
    return dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0


def shard_range(n, r=None, w=None):
    """Contiguous image range [lo, hi) of rank r out of w."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return (n * r) // w, (n * (r + 1)) // w


def _comm_device(t, group=None):
    """gloo moves CPU tensors, NCCL moves CUDA tensors."""
    if dist.get_backend(group) == "gloo":
        return torch.device("cpu")
    return t.device


def gather_records(rec, gt_per_class, group=None):
    """(rec int64 (n_r,), gt_per_class int32 (C,)) of this rank -> the rank-ordered concatenation over all ranks
    and the summed GT counts, on every rank.  Collective; synchronises the host (sizes travel through it)."""
    w = dist.get_world_size(group)
    home = rec.device
    cd = _comm_device(rec, group)
    n_local = torch.tensor([rec.shape[0]], dtype=torch.int64, device=cd)
    sizes = [torch.zeros_like(n_local) for _ in range(w)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(max(sizes), 1)
    r_pad = torch.zeros((n_max,), dtype=torch.int64, device=cd)
    r_pad[:rec.shape[0]] = rec.to(cd)
    r_all = [torch.empty_like(r_pad) for _ in range(w)]
    dist.all_gather(r_all, r_pad, group=group)
    g = gt_per_class.to(cd).clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    rec_cat = torch.cat([r_all[r][:sizes[r]] for r in range(w)]).to(home)
    return rec_cat, g.to(home)


class PeerExchange:
    """This rank's exchange buffer, mapped by every peer of the box (CUDA IPC), plus the peers' buffers mapped here.
    Construction and close() are collective over `group`; exchange_reduce() must be called by every rank the same
    number of times (it is what MeanAveragePrecision(sharded=True).result() does) but involves no host-side
    collective.  `capacity` = records per rank the buffers hold; a shard that outgrows it makes every rank's result
    NaN (and exchange_reduce raises on the rank that can tell)."""

    def __init__(self, device, num_classes, capacity, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib, self.L = C, _lib, _lib.lib()
        self.group, self.device = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.num_classes, self.capacity = int(num_classes), int(capacity)
        self.epoch = 0
        self.bufs = [None] * self.world          # device addresses of every rank's exchange buffer, index = rank
        self._alloc()
        total = self.world * self.capacity
        nbytes = int(self.L.yh_workspace_bytes(_lib.YH_OP_MAP_REDUCE, total, 0, 0, self.num_classes))
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        self.err = torch.zeros((1,), dtype=torch.int32, device=self.device)

    def _alloc(self):
        """Collective; raises the same RuntimeError on EVERY rank if any rank could not allocate or map."""
        C, L = self._C, self.L
        nbytes = int(L.yh_map_exchange_bytes(self.world, self.num_classes, self.capacity))
        h = (C.c_ubyte * self._lib.YH_IPC_HANDLE_BYTES)()
        p = C.c_void_p()
        err = None
        with on_device(self.device):
            try:
                self._lib.check(L.yh_ipc_alloc(nbytes, C.byref(p), h), "ipc_alloc")       # zero-filled
            except (RuntimeError, ValueError) as e:
                err = str(e)
            self.bufs[self.rank] = p.value
            handles = [None] * self.world
            dist.all_gather_object(handles, (err, bytes(h)), group=self.group)
            if all(x[0] is None for x in handles):
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            continue
                        q = C.c_void_p()
                        hb = handles[r][1]
                        self._lib.check(L.yh_ipc_open((C.c_ubyte * len(hb)).from_buffer_copy(hb), C.byref(q)), "ipc_open")
                        self.bufs[r] = q.value
                except (RuntimeError, ValueError) as e:
                    err = str(e)
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None and all(x[0] is None for x in handles), group=self.group)
        if not all(oks):
            self.close()
            raise RuntimeError(f"PeerExchange: a rank could not allocate or map the exchange buffers ({err or 'peer failure'})")

    def close(self):
        """Collective: unmap the peers' buffers, then free the own one."""
        C, L = self._C, self.L
        torch.cuda.synchronize(self.device)
        for r in range(self.world):
            if r != self.rank and self.bufs[r]:
                self._lib.check(L.yh_ipc_close(C.c_void_p(self.bufs[r])), "ipc_close")
                self.bufs[r] = None
        if self.bufs[self.rank]:
            dist.barrier(group=self.group)               # every peer unmapped before the owner frees
            self._lib.check(L.yh_ipc_free(C.c_void_p(self.bufs[self.rank])), "ipc_free")
            self.bufs[self.rank] = None

    def exchange_reduce(self, rec, nrec_dev, gt_per_class, n_hint=0):
        """This rank's records (int64 tensor, the first *nrec_dev of it, or all of it) and GT counts ->
        (global mAP 0-d tensor, AP per class) on every rank.  Two kernel launches, asynchronous."""
        C, L = self._C, self.L
        from ._tensor import stream_ptr
        dev = self.device
        n_max = int(rec.shape[0])
        if nrec_dev is None and n_max > self.capacity:
            raise RuntimeError(f"PeerExchange: {n_max} records exceed the exchange capacity of {self.capacity} per rank; "
                               "create the evaluator / exchange with a larger capacity")
        self.epoch += 1
        ap = torch.empty((self.num_classes,), dtype=torch.float32, device=dev)
        m = torch.empty((1,), dtype=torch.float32, device=dev)
        ba = (C.c_void_p * self.world)(*self.bufs)
        with on_device(dev):
            sp = stream_ptr(dev)
            self._lib.check(L.yh_map_exchange(self.world, self.rank, ba, self.num_classes, self.capacity, rec.data_ptr(), n_max,
                                              nrec_dev.data_ptr() if nrec_dev is not None else None, gt_per_class.data_ptr(),
                                              self.epoch, sp), "map_exchange")
            self._lib.check(L.yh_map_reduce_exchanged(self.world, C.c_void_p(self.bufs[self.rank]), self.num_classes,
                                                      self.capacity, self.epoch, int(n_hint) * self.world, ap.data_ptr(),
                                                      m.data_ptr(), self.err.data_ptr(), self.ws.data_ptr(), int(self.ws.numel()),
                                                      sp), "map_reduce_exchanged")
        return m[0], ap

    def error(self):
        """Host check (synchronises): 0, or YH_MAP_ERR_TIMEOUT / YH_MAP_ERR_OVERFLOW of a failed exchange."""
        return int(self.err.item())


_EXCHANGES = {}


def peer_exchange(device, num_classes, n_bound, group=None, capacity=None):
    """The cached PeerExchange of (group, device, classes), created collectively on first use, or None when the
    exchange has to stay on the padded all-gathers: CPU / gloo, YH_DIST_P2P=0, ranks on several hosts, more peers
    than the kernels take, or IPC mapping refused.  Capacity (records per rank): `capacity` if given, else four times
    the largest n_bound over the ranks at creation, at least 65,536."""
    import os
    import socket
    device = torch.device(device)
    key = (id(group), device.index, int(num_classes))
    if key in _EXCHANGES:
        return _EXCHANGES[key]
    ex = None
    ok = (device.type == "cuda" and dist.get_backend(group) == "nccl" and os.environ.get("YH_DIST_P2P", "1") != "0"
          and dist.get_world_size(group) <= 16)
    if ok:
        info = [None] * dist.get_world_size(group)
        with on_device(device):
            dist.all_gather_object(info, (socket.gethostname(), int(capacity) if capacity else 4 * int(n_bound)), group=group)
        if len({h for h, _ in info}) == 1:
            cap = max(max(c for _, c in info), 1 << 16)
            try:
                ex = PeerExchange(device, num_classes, cap, group=group)      # fails on every rank or on none
            except RuntimeError:
                ex = None
    _EXCHANGES[key] = ex
    return ex


def shutdown():
    """Collective: unmaps and frees the exchange buffers of every cached PeerExchange.  Call before
    torch.distributed.destroy_process_group() when the process goes on to create another group."""
    for ex in list(_EXCHANGES.values()):
        if ex is not None:
            ex.close()
    _EXCHANGES.clear()
