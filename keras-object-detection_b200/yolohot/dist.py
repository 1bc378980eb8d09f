"""Multi-GPU plumbing of the mAP reduction: one process per GPU (torch.distributed, NCCL over
NVLink on a B200 box; gloo in the CPU tests).  The decode / NMS / loss / matching kernels
need no communication - images are sharded by contiguous range - so the only exchange is
  (1) all-reduce(sum) of the per-class ground-truth counts (C int32), and
  (2) all-gather of the per-detection records (uint64 sort key, uint8 TP flag), padded to the
      largest shard, concatenated IN RANK ORDER (rank order == image order, which the stable
      sort of stage 2 relies on for equal confidences).
Volume is <= 9 B x detections (a few MB at most): latency-bound, so it stays on NCCL."""
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
