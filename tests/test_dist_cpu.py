"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard ranges and the rank-ordered
all-gather of mAP records + all-reduce of GT counts (yolohot.dist)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "keras-object-detection_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolohot import dist as yd
    assert yd.world_size() == world and yd.rank() == rank
    n = 11
    lo, hi = yd.shard_range(n)
    # ragged shards: rank r contributes r*3 + 2 records; one rank may be empty in the second round
    for rnd, cnt in enumerate((rank * 3 + 2, 0 if rank == 0 else 4)):
        rec = ((torch.arange(cnt, dtype=torch.int64) + 1000 * (rank + 1) + 100 * rnd) << 1) | (torch.arange(cnt) % 2)
        gt = torch.tensor([rank + 1, 10 * (rank + 1), 0], dtype=torch.int32)
        k, g = yd.gather_records(rec, gt)
        np.savez(os.path.join(out_dir, f"r{rank}_{rnd}.npz"), k=k.numpy(), g=g.numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


def test_gather_records_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rnd, counts in enumerate(((2, 5), (0, 4))):
        want_k = np.concatenate([((np.arange(c) + 1000 * (r + 1) + 100 * rnd) << 1) | (np.arange(c) % 2) for r, c in enumerate(counts)])
        for r in range(world):
            z = np.load(tmp_path / f"r{r}_{rnd}.npz")
            assert np.array_equal(z["k"], want_k)
            assert np.array_equal(z["g"], [3, 30, 0])
    z0, z1 = np.load(tmp_path / "r0_0.npz"), np.load(tmp_path / "r1_0.npz")
    assert (int(z0["lo"]), int(z0["hi"]), int(z1["lo"]), int(z1["hi"])) == (0, 5, 5, 11)
