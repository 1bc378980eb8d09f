"""Multi-GPU plumbing of the mAP reduction: one process per GPU (torch.distributed, NCCL over
NVLink on a B200 box; gloo in the CPU tests).  The decode / NMS / loss / matching kernels
need no communication - images are sharded by contiguous range - so the only exchange is
  (1) all-reduce(sum) of the per-class ground-truth counts (C int32), and
  (2) all-gather of the per-detection records (uint64 sort key, uint8 TP flag), padded to the
      largest shard, concatenated IN RANK ORDER (rank order == image order, which the stable
      sort of stage 2 relies on for equal confidences).
Volume is <= 9 B x detections (a few MB at most): latency-bound.

Two implementations of (2):
  gather_records   padded all-gathers through torch.distributed (NCCL, or gloo on CPU);
  PeerExchange     NCCL ranks of ONE box: every rank owns record buffers that its peers map through CUDA IPC, and the
                   last kernel of the matching stage stores each record straight into all of them over NVLink
                   (yh_map_match_peers) - the exchange is fused into the compute kernel.  What is left on NCCL is the
                   control plane: an all-gather of the shard sizes (gives the offsets and orders the reuse of the
                   buffers) and the all-reduce of (1), which also orders the peers' stores before stage 2."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def shard_range(n, r=None, w=None):
    """Contiguous image range [lo, hi) of rank r out of w."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return (n * r) // w, (n * (r + 1)) // w


def _comm_device(t):
    """gloo moves CPU tensors, NCCL moves CUDA tensors."""
    if dist.get_backend() == "gloo":
        return torch.device("cpu")
    return t.device


def gather_records(keys, tp, gt_per_class, group=None):
    """(keys int64 (n_r,), tp uint8 (n_r,), gt_per_class int32 (C,)) of this rank ->
    the rank-ordered concatenation over all ranks and the summed GT counts, on every rank."""
    w = dist.get_world_size(group)
    home = keys.device
    cd = _comm_device(keys)
    n_local = torch.tensor([keys.shape[0]], dtype=torch.int64, device=cd)
    sizes = [torch.zeros_like(n_local) for _ in range(w)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(max(sizes), 1)
    # one padded buffer per record field; key and flag travel as int64 / uint8
    k_pad = torch.zeros((n_max,), dtype=torch.int64, device=cd)
    t_pad = torch.zeros((n_max,), dtype=torch.uint8, device=cd)
    k_pad[:keys.shape[0]] = keys.to(cd)
    t_pad[:tp.shape[0]] = tp.to(cd)
    k_all = [torch.empty_like(k_pad) for _ in range(w)]
    t_all = [torch.empty_like(t_pad) for _ in range(w)]
    dist.all_gather(k_all, k_pad, group=group)
    dist.all_gather(t_all, t_pad, group=group)
    g = gt_per_class.to(cd).clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    keys_cat = torch.cat([k_all[r][:sizes[r]] for r in range(w)]).to(home)
    tp_cat = torch.cat([t_all[r][:sizes[r]] for r in range(w)]).to(home)
    return keys_cat, tp_cat, g.to(home)


class _DeviceArray:
    """Zero-copy torch view of a raw device range (buffers come from yh_ipc_alloc, not from torch's allocator)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """Record buffers of this rank, mapped by every peer of the box (CUDA IPC), plus the peers' buffers mapped here.
    All methods are collective over `group`.  See the module docstring for the protocol."""

    def __init__(self, device, capacity=1 << 18, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib, self.L = C, _lib, _lib.lib()
        self.group, self.device = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.capacity = 0
        self.keys = [None] * self.world        # device addresses, index = rank
        self.tp = [None] * self.world
        self._alloc(int(capacity))

    # -- buffers ------------------------------------------------------------------------------
    def _release(self):
        C, L = self._C, self.L
        torch.cuda.synchronize(self.device)
        for r in range(self.world):
            if r != self.rank:
                for arr in (self.keys, self.tp):
                    if arr[r]:
                        self._lib.check(L.yh_ipc_close(C.c_void_p(arr[r])), "ipc_close")
                        arr[r] = None
        if self.capacity:
            dist.barrier(group=self.group)               # every peer unmapped before the owner frees
            for arr in (self.keys, self.tp):
                if arr[self.rank]:
                    self._lib.check(L.yh_ipc_free(C.c_void_p(arr[self.rank])), "ipc_free")
                    arr[self.rank] = None
        self.capacity = 0

    def _alloc(self, capacity):
        """Collective; raises the same RuntimeError on EVERY rank if any rank could not allocate or map."""
        C, L = self._C, self.L
        self._release()
        hk = (C.c_ubyte * self._lib.YH_IPC_HANDLE_BYTES)()
        ht = (C.c_ubyte * self._lib.YH_IPC_HANDLE_BYTES)()
        pk, pt = C.c_void_p(), C.c_void_p()
        err = None
        with torch.cuda.device(self.device):
            try:
                self._lib.check(L.yh_ipc_alloc(capacity * 8, C.byref(pk), hk), "ipc_alloc")
                self._lib.check(L.yh_ipc_alloc(capacity, C.byref(pt), ht), "ipc_alloc")
            except (RuntimeError, ValueError) as e:
                err = str(e)
            self.keys[self.rank], self.tp[self.rank] = pk.value, pt.value
            self.capacity = capacity
            handles = [None] * self.world
            dist.all_gather_object(handles, (err, bytes(hk), bytes(ht)), group=self.group)
            if all(h[0] is None for h in handles):
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            continue
                        for arr, h in ((self.keys, handles[r][1]), (self.tp, handles[r][2])):
                            q = C.c_void_p()
                            self._lib.check(L.yh_ipc_open((C.c_ubyte * len(h)).from_buffer_copy(h), C.byref(q)), "ipc_open")
                            arr[r] = q.value
                except (RuntimeError, ValueError) as e:
                    err = str(e)
            flag = torch.tensor([0 if (err or any(h[0] for h in handles)) else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self._release()
            raise RuntimeError(f"PeerExchange: a rank could not allocate or map the exchange buffers ({err or 'peer failure'})")

    def close(self):
        self._release()

    # -- stage 1 + exchange -------------------------------------------------------------------
    def match_gather(self, true_rows, pred_rows, num_classes, iou_threshold=0.5):
        """This rank's rows -> (keys int64, tp uint8, gt_per_class int32) of ALL ranks in rank order, on every rank.
        The returned key / flag tensors are views of this rank's exchange buffers, valid until the next call."""
        C, L = self._C, self.L
        from ._tensor import stream_ptr
        dev = self.device
        t, p = true_rows.contiguous(), pred_rows.contiguous()
        n_local = torch.tensor([p.shape[0]], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n_local) for _ in range(self.world)]
        # also the fence between the previous call's readers and this call's writers (stream-ordered on every rank)
        dist.all_gather(sizes, n_local, group=self.group)
        sizes = [int(x.item()) for x in sizes]
        total, offset = sum(sizes), sum(sizes[:self.rank])
        if total > self.capacity:                       # the same decision on every rank
            cap = self.capacity
            while cap < total:
                cap *= 2
            self._alloc(cap)
        gt = torch.empty((num_classes,), dtype=torch.int32, device=dev)
        ka = (C.c_void_p * self.world)(*self.keys)
        ta = (C.c_void_p * self.world)(*self.tp)
        with torch.cuda.device(dev):
            self._lib.check(L.yh_map_match_peers(self.world, self.rank, t.data_ptr(), int(t.shape[0]), p.data_ptr(),
                                                 int(p.shape[0]), int(num_classes), float(iou_threshold), ka, ta, offset,
                                                 None, gt.data_ptr(), stream_ptr(dev)), "map_match_peers")
        # sums the GT counts; completes on a rank only after every peer's contribution, which follows that peer's
        # match kernel in stream order - so all records are in this rank's buffers when stage 2 starts
        dist.all_reduce(gt, op=dist.ReduceOp.SUM, group=self.group)
        if total == 0:
            return (torch.empty((0,), dtype=torch.int64, device=dev), torch.empty((0,), dtype=torch.uint8, device=dev), gt)
        keys = torch.as_tensor(_DeviceArray(self.keys[self.rank], total, "<i8"), device=dev)
        tp = torch.as_tensor(_DeviceArray(self.tp[self.rank], total, "|u1"), device=dev)
        return keys, tp, gt


_EXCHANGES = {}


def peer_exchange(device, group=None):
    """The cached PeerExchange of (group, device), or None when the exchange has to stay on the padded all-gathers:
    CPU / gloo, YH_DIST_P2P=0, ranks on several hosts, more peers than the kernel takes, or IPC mapping refused."""
    import os
    import socket
    device = torch.device(device)
    key = (id(group), device.index)
    if key in _EXCHANGES:
        return _EXCHANGES[key]
    ex = None
    ok = (device.type == "cuda" and dist.get_backend(group) == "nccl" and os.environ.get("YH_DIST_P2P", "1") != "0"
          and dist.get_world_size(group) <= 16)
    if ok:
        hosts = [None] * dist.get_world_size(group)
        with torch.cuda.device(device):
            dist.all_gather_object(hosts, socket.gethostname(), group=group)
        if len(set(hosts)) == 1:
            try:
                ex = PeerExchange(device, group=group)      # fails on every rank or on none
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
