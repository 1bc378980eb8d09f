#!/usr/bin/env python
"""bench.py - post-processing images/sec (decode + IoU + NMS) on B200, per BASELINE.json.

  python bench.py --gpus N --steps K --warmup W            # own arm (libyolohot, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference's algorithm
                                                            #   (oracle C port) on the host cores

A step = one pass of the fused decode+IoU+NMS hot path (yh_decode_nms, utils.py:470-480 loop
body) over BASELINE.json configs[1]: 1,000,000 synthetic YOLOv1 outputs (S=7, B=2, C=20, dense
U[0,1), seed 2025) PER GPU (weak scaling, images sharded by contiguous range, no collective on
this path).  `value` times the kernel with inputs resident in HBM (CUDA events on the launch
stream, 5.88 GB per step >> L2); `e2e` times the same workload through the C-ABI host entry
point yh_decode_nms_host with pinned HOST buffers (H2D + kernel + D2H inside the timed region).
One JSON line on rank 0.  Nothing here reads /root/reference."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))

S, B, C = 7, 2, 20
M, D = S * S, C + 5 * B
IMG_IN = 4 * M * D            # 5,880 B read per image
IMG_OUT_ROW = 24              # bytes written per kept row
CONF_THR, IOU_THR = 0.4, 0.5  # utils.py:475
METRIC = "post-processing images/sec (decode+IoU+NMS)"
UNIT = "images/s"


def base_config(world, n_img):
    return {"workload": "cfg2: YOLOv1 S=7 B=2 C=20 decode+IoU+NMS, 1M synthetic images per GPU, dense U[0,1) seed 2025",
            "images_per_gpu": n_img, "S": S, "B": B, "C": C, "conf_threshold": CONF_THR, "iou_threshold": IOU_THR,
            "sharding": f"images by contiguous range x{world}, no data-path collective",
            "l2": "inputs larger than L2 (5.88 GB read per step per GPU; no flush needed)"}


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("decode_nms_tma_kernel_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """NVML sampler of SM clock / throttle reasons during a timed region."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.mem, self.power = [], []
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.mem.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sm_mhz_min": (min(self.samples) if self.samples else None),
                "mem_mhz": (statistics.median(self.mem) if self.mem else None),
                "power_w_max": (max(self.power) if self.power else None)}


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU implementation of the path = the oracle C port (the real one is TF
    Python, not importable here), all host threads, bounded sample per step."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import numpy as np
    from oracle import cport
    cport.build()
    cores = cport.num_threads()
    n_step = env_int("YH_BENCH_REF_IMAGES", 1_000_000)          # the stated config: 1M images per step
    try:
        import psutil
        while n_step > 65_536 and n_step * (IMG_IN + M * 24 + 8) * 2 > psutil.virtual_memory().available:
            n_step //= 2                                         # host RAM decides, and the line says so
    except Exception:
        pass
    rng = np.random.Generator(np.random.PCG64(2025))
    p = np.empty((n_step, S, S, D), dtype=np.float32)
    for lo in range(0, n_step, 65_536):                          # generated in slices: no 2x float64 temporary
        p[lo:lo + 65_536] = rng.random((min(65_536, n_step - lo), S, S, D), dtype=np.float32)
    for _ in range(max(1, min(args.warmup, 2))):
        cport.decode_nms(p[: n_step // 4], C, B, IOU_THR, CONF_THR, nthreads=cores, want_idx=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cport.decode_nms(p, C, B, IOU_THR, CONF_THR, nthreads=cores, want_idx=False)
    dt = time.perf_counter() - t0
    value = n_step * args.steps / dt
    sample = f"{n_step} images/step x {args.steps} steps of the cfg2 workload (dense U[0,1), seed 2025), {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args.gpus, 1_000_000),           # the same config as the own arm ...
        "reference_sample": {"images_per_step": n_step, "full_config": n_step == 1_000_000},   # ... and what was timed of it
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is TF-eager Python (not importable: no TensorFlow); this arm times oracle/yolo_oracle_c.c, "
                "a C restatement of the same algorithm, which is far faster than the real reference would be",
    }
    print(json.dumps(line), flush=True)
    return 0


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _gpu_numa_node(local):
    """(node, how) of GPU `local`, trying sysfs, NVML's memory affinity and `nvidia-smi topo -m` in turn."""
    tried = []
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                       # nvml pads the PCI domain to 8 hex digits
            bus = bus[4:]
        try:
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
            if node >= 0:
                return node, "sysfs numa_node"
            tried.append("sysfs numa_node = -1")
        except Exception as e:
            tried.append(f"sysfs: {type(e).__name__}")
        try:
            mask = pynvml.nvmlDeviceGetMemoryAffinity(h, 4, pynvml.NVML_AFFINITY_SCOPE_NODE)
            for w, word in enumerate(mask):
                for b in range(64):
                    if (int(word) >> b) & 1:
                        return 64 * w + b, "nvmlDeviceGetMemoryAffinity"
            tried.append("nvml memory affinity empty")
        except Exception as e:
            tried.append(f"nvml affinity: {type(e).__name__}")
    except Exception as e:
        tried.append(f"nvml: {type(e).__name__}")
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        hdr = None
        for line in out.splitlines():
            cols = [c for c in line.replace("\x1b[4m", "").replace("\x1b[0m", "").split("\t") if c != ""]
            if hdr is None and any("NUMA Affinity" in c for c in cols):
                hdr = [c.strip() for c in cols]
                continue
            if hdr is not None and cols and cols[0].strip() == f"GPU{local}":
                k = [i for i, c in enumerate(hdr) if "NUMA Affinity" in c][0] + 1     # the header row has no row label
                v = cols[k].strip() if k < len(cols) else ""
                if v and v.split(",")[0].split("-")[0].isdigit():
                    return int(v.split(",")[0].split("-")[0]), "nvidia-smi topo -m"
                tried.append(f"topo NUMA Affinity = {v!r}")
        if hdr is None:
            tried.append("topo: no NUMA Affinity column")
    except Exception as e:
        tried.append(f"topo: {type(e).__name__}")
    return None, "; ".join(tried)


def bind_to_gpu_numa_node(local):
    """Place this rank - its threads and, through the memory policy, the pages of the pinned host buffers it allocates
    afterwards - on the NUMA node its GPU hangs off, so that the e2e H2D / D2H copies of 8 ranks do not cross the socket
    interconnect.  Placement only.  Always returns a description of what was (or could not be) done for the JSON line."""
    info = {"nodes_online": None, "gpu_node": None, "how": None, "cpus_bound": None, "mempolicy": None}
    try:
        with open("/sys/devices/system/node/online") as f:
            info["nodes_online"] = f.read().strip()
    except Exception:
        info["nodes_online"] = "unreadable"
    node, how = _gpu_numa_node(local)
    info["gpu_node"], info["how"] = node, how
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            allowed = os.sched_getaffinity(0) & _cpulist(f.read())
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus_bound"] = len(allowed)
        else:
            info["cpus_bound"] = "none of the node's CPUs are in this process's cpuset"
    except Exception as e:
        info["cpus_bound"] = f"{type(e).__name__}"
    try:                                                          # set_mempolicy(MPOL_PREFERRED, {node}): also works when the
        libc = ctypes.CDLL(None, use_errno=True)                  # cpuset is a single socket
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, 1, mask, 16 * 64 + 1)              # __NR_set_mempolicy (x86-64), MPOL_PREFERRED
        info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy errno %d" % ctypes.get_errno()
    except Exception as e:
        info["mempolicy"] = f"{type(e).__name__}"
    return info


# ------------------------------------------------------------------------------------ own arm
def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from yolohot import _lib
    from yolohot import utils as yu

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    numa = bind_to_gpu_numa_node(local) if env_int("YH_BENCH_NUMA_BIND", 1) else {"how": "disabled (YH_BENCH_NUMA_BIND=0)"}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    n_img = env_int("YH_BENCH_IMAGES", 1_000_000)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2025 + rank)
    pred = torch.rand((n_img, S, S, D), generator=gen, device=dev, dtype=torch.float32)
    boxes = torch.empty((n_img, M, 6), device=dev, dtype=torch.float32)
    cnt = torch.empty((n_img,), device=dev, dtype=torch.int32)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def step():
        _lib.check(L.yh_decode_nms(pred.data_ptr(), n_img, S, B, C, IOU_THR, CONF_THR, boxes.data_ptr(), cnt.data_ptr(),
                                   None, sp), "yh_decode_nms")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    kept = int(cnt.sum().item())

    # ---- timed region: K steps, one event pair per step (= per launch of the dominant kernel)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        barrier()
        t_all0 = torch.cuda.Event(enable_timing=True)
        t_all1 = torch.cuda.Event(enable_timing=True)
        t_all0.record(stream)
        for a, b in ev:
            a.record(stream)
            step()
            b.record(stream)
        t_all1.record(stream)
        barrier()
    launches = _lib.launch_count() - l0
    total_ms = t_all0.elapsed_time(t_all1)
    per_launch_ms = [a.elapsed_time(b) for a, b in ev]
    ms_step = total_ms / args.steps
    if world > 1:
        t = torch.tensor([ms_step, float(launches), float(kept)], device=dev, dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_step = float(mx[0])
        launches_job = int(sm[1])
    else:
        launches_job = launches
    value = world * n_img / (ms_step / 1000.0)

    # ---- roofline of the dominant kernel (this rank's launches)
    peak, peak_src = measured_peak()
    algo_bytes = n_img * (IMG_IN + 4) + IMG_OUT_ROW * kept          # 4*S^2*D read + 24K + 4 written
    k_ms = statistics.mean(per_launch_ms)
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "kernel": "decode_nms_tma_kernel<2,20,2>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms_mean": k_ms, "kernel_ms_min": min(per_launch_ms),
                "kept_rows_per_image": kept / n_img}

    # ---- end to end through the C-ABI with pinned host buffers
    e2e_n = env_int("YH_BENCH_E2E_IMAGES", n_img)
    try:
        import psutil
        avail = psutil.virtual_memory().available
        while e2e_n > 65_536 and world * e2e_n * (IMG_IN + M * 24 + 4) * 2 > avail:
            e2e_n //= 2
    except Exception:
        pass
    wc_ptr = None
    if env_int("YH_BENCH_E2E_WC", 0):       # experiment: write-combined pinned input buffer (yh_host_alloc)
        import numpy as _np
        wc_ptr = ctypes.c_void_p()
        _lib.check(L.yh_host_alloc(e2e_n * IMG_IN, 1, ctypes.byref(wc_ptr)), "yh_host_alloc")
        buf = (ctypes.c_float * (e2e_n * M * D)).from_address(wc_ptr.value)
        h_in = torch.from_numpy(_np.frombuffer(buf, dtype=_np.float32).reshape(e2e_n, S, S, D))
    else:
        h_in = torch.empty((e2e_n, S, S, D), dtype=torch.float32, pin_memory=True)
    h_boxes = torch.empty((e2e_n, M, 6), dtype=torch.float32, pin_memory=True)
    h_cnt = torch.empty((e2e_n,), dtype=torch.int32, pin_memory=True)
    if wc_ptr is None:
        h_in.copy_(pred[:e2e_n])
    else:                                    # the CPU only ever writes this buffer
        for lo in range(0, e2e_n, 65_536):
            hi = min(e2e_n, lo + 65_536)
            h_in[lo:hi] = pred[lo:hi].cpu()
    torch.cuda.synchronize(dev)

    def e2e_step():
        _lib.check(L.yh_decode_nms_host(h_in.data_ptr(), e2e_n, S, B, C, IOU_THR, CONF_THR, h_boxes.data_ptr(),
                                        h_cnt.data_ptr(), None, local), "yh_decode_nms_host")

    e2e_step()
    # the e2e path's own roofline: a bare pinned H2D copy of the same input (PCIe), device-timed - once with every rank
    # copying at the same time (what the host memory system / PCIe fabric delivers to N GPUs at once: the ceiling of the
    # N-GPU e2e figure) and once rank by rank (what one GPU gets when it has the host to itself)
    d_tmp = torch.empty_like(pred[:e2e_n])
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def bare_h2d():
        c0.record(stream)
        d_tmp.copy_(h_in, non_blocking=True)
        c1.record(stream)
        torch.cuda.synchronize(dev)
        return e2e_n * IMG_IN / (c0.elapsed_time(c1) * 1e-3) / 1e9
    bare_h2d()
    barrier()
    h2d_gbs = bare_h2d()                                             # all ranks at once
    h2d_all = h2d_gbs
    h2d_solo = h2d_gbs
    if world > 1:
        t = torch.tensor([h2d_gbs], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        h2d_all = float(t[0])                                        # aggregate over the ranks
        solo = torch.zeros(world, device=dev, dtype=torch.float64)
        for r in range(world):                                       # one rank at a time
            dist.barrier()
            if r == rank:
                solo[r] = bare_h2d()
        dist.all_reduce(solo)
        h2d_solo = float(solo[rank])
        solo_all = [float(v) for v in solo]
    del d_tmp
    e2e_steps = max(2, min(args.steps, env_int("YH_BENCH_E2E_STEPS", 5)))
    with ClockSampler(local) as clk2:
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
    ok_e2e = bool(torch.equal(h_cnt, cnt[:e2e_n].cpu()))
    zero_tail = bool((h_boxes[0, int(h_cnt[0]):] == 0).all())        # rows past count[i] come back as zeros
    numa_all = [numa]
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)
    e2e = {"value": world * e2e_n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": e2e_n * IMG_IN,
           "d2h_bytes_per_step": e2e_n * (M * 24 + 4), "images_per_step_per_gpu": e2e_n, "steps": e2e_steps,
           "ms_per_step": 1000.0 * e2e_s, "api": "yh_decode_nms_host (C-ABI, pinned host buffers in/out)",
           "counts_match_device_path": ok_e2e, "rows_past_count_are_zero": zero_tail,
           "numa_binding_rank0": numa, "numa_gpu_node_per_rank": [(x or {}).get("gpu_node") for x in numa_all],
           "bare_pinned_h2d_copy_GBps": h2d_gbs, "bare_pinned_h2d_all_ranks_at_once_GBps_total": h2d_all,
           "bare_pinned_h2d_one_rank_at_a_time_GBps": h2d_solo,
           "frac_of_bare_h2d_copy": (e2e_n * IMG_IN / e2e_s / 1e9) / h2d_gbs,
           "frac_of_concurrent_h2d_ceiling": (world * e2e_n * IMG_IN / e2e_s / 1e9) / h2d_all,
           "limit": "PCIe / host memory: e2e runs at the rate a bare pinned H2D copy of the same bytes gets when all ranks copy at once"}
    if world > 1:
        e2e["bare_pinned_h2d_one_rank_at_a_time_GBps_per_rank"] = solo_all
    # the same call with a float16 head (half the bytes over PCIe; widened exactly inside the kernel), at every N; the
    # headline e2e above stays on the reference's float32 tensors
    e2e_half = None
    if env_int("YH_BENCH_E2E_HALF", 1):
        h_half = torch.empty((e2e_n, S, S, D), dtype=torch.float16, pin_memory=True)
        h_half.copy_(h_in)

        def half_step():
            _lib.check(L.yh_decode_nms_host_typed(h_half.data_ptr(), _lib.YH_DTYPE_F16, e2e_n, S, B, C, IOU_THR, CONF_THR,
                                                  h_boxes.data_ptr(), h_cnt.data_ptr(), None, local), "yh_decode_nms_host_typed")
        half_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            half_step()
        torch.cuda.synchronize(dev)
        hs = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([hs], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            hs = float(t[0])
        e2e_half = {"value": world * e2e_n / hs, "unit": UNIT, "ms_per_step": 1000.0 * hs, "h2d_bytes_per_step": e2e_n * IMG_IN // 2,
                    "n_gpus": world, "api": "yh_decode_nms_host_typed (float16 head, pinned host buffers)",
                    "note": "float16 rounding of the synthetic inputs changes the results; parity is against the widened tensor"}
        del h_half
    # VOC-like sparse outputs (~2.7 kept rows per image): the padded block is 94 % slot space; the compact form
    # (yh_decode_nms_host_rows) sends only the kept rows back
    e2e_sparse = None
    if env_int("YH_BENCH_E2E_SPARSE", 1) and wc_ptr is None:
        for lo in range(0, e2e_n, 65_536):
            hi = min(e2e_n, lo + 65_536)
            q = pred[lo:hi].clone()
            for b in range(B):
                q[..., C + 5 * b] = q[..., C + 5 * b] ** 32
            h_in[lo:hi].copy_(q)
            del q
        torch.cuda.synchronize(dev)
        cap_rows = 8 * e2e_n
        h_rows = torch.empty((cap_rows, 7), dtype=torch.float32, pin_memory=True)
        tot = ctypes.c_int64(0)

        def compact_step():
            _lib.check(L.yh_decode_nms_host_rows(h_in.data_ptr(), _lib.YH_DTYPE_F32, e2e_n, S, B, C, IOU_THR, CONF_THR, h_rows.data_ptr(),
                                                 cap_rows, h_cnt.data_ptr(), ctypes.byref(tot), local), "yh_decode_nms_host_rows")
        res = {}
        for name, fn in (("padded", e2e_step), ("compact_rows", compact_step)):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fn()
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) / e2e_steps
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t[0])
            res[name] = dt
        kept_s = int(tot.value)
        e2e_sparse = {"workload": "cfg2 sparse variant (confidences u^32), float32, pinned host buffers", "n_gpus": world,
                      "kept_rows_per_image": kept_s / e2e_n,
                      "padded": {"value": world * e2e_n / res["padded"], "unit": UNIT, "d2h_bytes_per_step": e2e_n * (M * 24 + 4),
                                 "api": "yh_decode_nms_host"},
                      "compact_rows": {"value": world * e2e_n / res["compact_rows"], "unit": UNIT, "d2h_bytes_per_step": 28 * kept_s + 4 * e2e_n,
                                       "api": "yh_decode_nms_host_rows", "rows_equal_counts": bool(int(h_cnt.sum()) == kept_s)}}
        del h_rows
    del h_in, h_boxes

    # ---- sustained: >= 1,000 back-to-back launches (the K timed steps above are a burst of a few tens of ms)
    sustained = None
    n_sus = env_int("YH_BENCH_SUSTAINED", 1000)
    if n_sus > 0:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_sus + 1)]
        with ClockSampler(local) as clk3:
            barrier()
            evs[0].record(stream)
            for k in range(n_sus):
                step()
                evs[k + 1].record(stream)
            barrier()
        ts = sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(n_sus))
        tot = evs[0].elapsed_time(evs[n_sus])
        gbs = algo_bytes / (tot / n_sus * 1e-3) / 1e9
        sustained = {"launches": n_sus, "seconds": tot / 1e3, "ms_mean": tot / n_sus, "ms_p50": ts[n_sus // 2],
                     "ms_p99": ts[min(n_sus - 1, int(0.99 * n_sus))], "ms_min": ts[0], "images_per_s_per_gpu": n_img / (tot / n_sus * 1e-3),
                     "GBps": gbs, "frac_hbm": gbs / peak, "ms_first_20_mean": statistics.mean(evs[k].elapsed_time(evs[k + 1]) for k in range(min(20, n_sus))),
                     "ms_last_100_mean": statistics.mean(evs[k].elapsed_time(evs[k + 1]) for k in range(max(0, n_sus - 100), n_sus)),
                     "clocks": clk3.summary()}
        if world > 1:
            t = torch.tensor([sustained["ms_mean"]], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sustained["ms_mean_max_over_ranks"] = float(t[0])
            sustained["images_per_s"] = world * n_img / (float(t[0]) * 1e-3)

    # ---- CPU baseline (rank 0, N=1 only): oracle C port on a bounded slice of the same inputs
    cpu_baseline = None
    extras = {}
    if rank == 0 and world == 1:
        from oracle import cport
        cport.build()
        cores = cport.num_threads()
        n_cpu = env_int("YH_BENCH_CPU_IMAGES", 262_144)
        sl = pred[:n_cpu].cpu().numpy()
        cport.decode_nms(sl[:4096], C, B, IOU_THR, CONF_THR, nthreads=cores, want_idx=False)
        reps, t0 = 0, time.perf_counter()
        while True:
            _, c_cpu, _ = cport.decode_nms(sl, C, B, IOU_THR, CONF_THR, nthreads=cores, want_idx=False)
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 1000:
                break
        dt = time.perf_counter() - t0
        audit = bool(np.array_equal(c_cpu, cnt[:n_cpu].cpu().numpy()))
        n1 = min(n_cpu, 32_768)                                     # the reference itself is single-threaded Python: 1-core figure too
        t1 = time.perf_counter()
        cport.decode_nms(sl[:n1], C, B, IOU_THR, CONF_THR, nthreads=1, want_idx=False)
        one_core = n1 / (time.perf_counter() - t1)
        cpu_baseline = {"value": n_cpu * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {n_cpu} images of the same batch x {reps} passes ({dt:.1f} s), oracle C port "
                                  f"(reference itself is TF-eager Python, not importable: no TensorFlow)",
                        "kept_counts_equal_gpu": audit, "value_1_core": one_core}
    del pred, boxes
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs (kernel time, resident inputs): cfg4 (sharded when N > 1) and cfg5 at EVERY N,
    # cfg1 / cfg2-sparse / cfg3 at N = 1
    ctx = dict(torch=torch, dist=dist, dev=dev, L=L, _lib=_lib, yu=yu, sp=sp, rank=rank, world=world)
    if env_int("YH_BENCH_EXTRAS", 1):
        extras.update(cfg5_stress(ctx))
        extras.update(cfg4_map(ctx))
        if world == 1:
            extras.update(other_configs(ctx))

    clocks = clk.summary()
    clocks["e2e"] = clk2.summary()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": base_config(world, n_img),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches_job, "clocks": clocks,
    }
    if sustained is not None:
        line["sustained"] = sustained
    line.update(extras)
    if e2e_half is not None:
        line["e2e_float16_head"] = e2e_half
    if e2e_sparse is not None:
        line["e2e_sparse"] = e2e_sparse
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _timed(torch, dev, fn, reps):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return statistics.mean(ts), min(ts)


def _max_over_ranks(ctx, v):
    if ctx["world"] == 1:
        return float(v)
    t = ctx["torch"].tensor([float(v)], device=ctx["dev"], dtype=ctx["torch"].float64)
    ctx["dist"].all_reduce(t, op=ctx["dist"].ReduceOp.MAX)
    return float(t[0])


def cfg5_stress(ctx):
    """BASELINE configs[4]: S=14 B=3 C=80, conf threshold 0.05, ~80 % of the argmaxes in 4 dominant classes,
    131,072 images PER GPU (9.76 GB each), images sharded by range - no collective on this path."""
    torch, dev, L, _lib, sp, world = ctx["torch"], ctx["dev"], ctx["L"], ctx["_lib"], ctx["sp"], ctx["world"]
    peak, _ = measured_peak()
    S5, B5, C5 = 14, 3, 80
    n5 = env_int("YH_BENCH_CFG5_IMAGES", 131_072)
    gen = torch.Generator(device=dev); gen.manual_seed(99 + ctx["rank"])
    p5 = torch.rand((n5, S5, S5, C5 + 5 * B5), generator=gen, device=dev)
    for lo in range(0, n5, 16_384):                                   # in slices: the masks are as large as a channel plane
        v = p5[lo:lo + 16_384]
        dom = torch.randint(0, 4, v.shape[:3], generator=gen, device=dev)
        boost = torch.rand(v.shape[:3], generator=gen, device=dev) < 0.8
        for k in range(4):
            v[..., k] += 1.5 * (boost & (dom == k)).float()
        for b in range(B5):
            v[..., C5 + 5 * b + 3:C5 + 5 * b + 5] = 0.1 + 0.5 * v[..., C5 + 5 * b + 3:C5 + 5 * b + 5]
        del dom, boost
    boxes5 = torch.empty((n5, S5 * S5, 6), device=dev); cnt5 = torch.empty((n5,), device=dev, dtype=torch.int32)
    f5 = lambda: _lib.check(L.yh_decode_nms(p5.data_ptr(), n5, S5, B5, C5, IOU_THR, 0.05, boxes5.data_ptr(), cnt5.data_ptr(), None, sp))
    ms, mn = _timed(torch, dev, f5, 10)
    kept5 = int(cnt5.sum().item())
    gbs = (n5 * (4 * S5 * S5 * (C5 + 5 * B5) + 4) + 24 * kept5) / (ms * 1e-3) / 1e9
    ms_job = _max_over_ranks(ctx, ms)
    out = {"images_per_gpu": n5, "images_per_s": world * n5 / (ms_job * 1e-3), "ms": ms_job, "ms_rank0": ms, "ms_min_rank0": mn,
           "GBps_per_gpu": gbs, "frac_hbm": gbs / peak, "kept_per_image": kept5 / n5, "n_gpus": world,
           "kernel": "decode_nms_coop_kernel<80,3>"}
    del p5, boxes5, cnt5
    torch.cuda.empty_cache()
    return {"cfg5_stress": out}


def cfg4_map(ctx):
    """BASELINE configs[3]: VOC-style mAP@0.5 over 5k synthetic images through the evaluator (update_state + result()).
    N = 1: one GPU.  N > 1: the images are sharded by contiguous range, every rank feeds its shard to a sharded
    evaluator; result() exchanges the records with kernels over NVLink (no NCCL call) and every rank must get the
    single-GPU value bit for bit (checked here, on every rank, against an unsharded local evaluator)."""
    torch, dist, dev, _lib, yu, world, rank = ctx["torch"], ctx["dist"], ctx["dev"], ctx["_lib"], ctx["yu"], ctx["world"], ctx["rank"]
    from tests import fixtures as F
    from yolohot import dist as yd
    n = 5000
    yt5 = F.synth_labels(n, seed=11); mp5 = F.synth_map_pred(yt5)
    a_all, b_all = torch.from_numpy(yt5).to(dev), torch.from_numpy(mp5).to(dev)
    e1 = yu.MeanAveragePrecision(C, B)
    e1.update_state(a_all, b_all)
    m_single = float(e1.result())
    lo, hi = yd.shard_range(n, rank, world)
    a, b_ = a_all[lo:hi].contiguous(), b_all[lo:hi].contiguous()
    ev = yu.MeanAveragePrecision(C, B, sharded=world > 1)

    def run_map():
        ev.reset_states()
        ev.update_state(a, b_)
        return ev.result()
    for _ in range(3):
        m = run_map()
    mval = float(m)
    nrec = int(ev._st["cursors"][0].item())
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    reps = 20
    l0 = _lib.launch_count()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        m = run_map()
    e1_.record()
    torch.cuda.synchronize(dev)
    wall_ms = (time.perf_counter() - t0) / reps * 1e3
    gpu_ms = e0.elapsed_time(e1_) / reps
    launches = (_lib.launch_count() - l0) / reps
    sync_ms = []                                                     # one pass at a time, result read on the host
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        float(run_map())
        sync_ms.append((time.perf_counter() - t0) * 1e3)
    out = {"images": n, "n_gpus": world, "mAP": mval, "mAP_single_gpu": m_single, "records_this_rank": nrec,
           "ms_update_plus_result": _max_over_ranks(ctx, wall_ms), "gpu_ms_update_plus_result": _max_over_ranks(ctx, gpu_ms),
           "ms_update_plus_result_host_read": _max_over_ranks(ctx, statistics.median(sync_ms)),
           "kernel_launches_per_pass": launches, "host_syncs_per_pass": 0,
           "path": "yh_eval_update_state (decode+NMS of y_pred and y_true, match, append: one kernel) -> " +
                   ("yh_map_exchange (peer stores over NVLink) -> yh_map_reduce_exchanged" if world > 1 else "yh_map_reduce")}
    if world == 1:
        # the two stages on their own (back to back, CUDA events): update_state = one kernel, result() = one kernel on its
        # counting path - and, for comparison, on its radix-sort passes (YH_MAP_COUNT=0 is read per call)
        def timed(fn, reps=30):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                fn()
            a1.record()
            torch.cuda.synchronize(dev)
            return a0.elapsed_time(a1) / reps * 1e3

        def upd():
            ev.reset_states()
            ev.update_state(a, b_)
        out["us_update_state_alone"] = timed(upd)
        out["us_result_alone"] = timed(ev.result)
        os.environ["YH_MAP_COUNT"] = "0"
        try:
            out["us_result_alone_radix_passes_only"] = timed(ev.result)
            out["mAP_radix_passes_only"] = float(ev.result())
        finally:
            del os.environ["YH_MAP_COUNT"]
    if world > 1:
        ok = torch.tensor([1 if mval == m_single else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        tot = torch.tensor([nrec], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        ex = yd.peer_exchange(dev, C, 0)
        rec_l, gt_l = ev._st["rec"][:nrec], ev._st["gt"]

        def nccl_path():
            r_all, g_all = yd.gather_records(rec_l, gt_l)
            return yu.map_reduce(r_all, g_all, C)[0]
        m_nccl = float(nccl_path())
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            nccl_path()
        torch.cuda.synchronize(dev)
        nccl_ms = (time.perf_counter() - t0) / reps * 1e3

        def peer_path():
            return ex.exchange_reduce(ev._st["rec"], ev._st["cursors"][0:1], gt_l, nrec)[0]
        peer_ms = None
        if ex is not None:
            float(peer_path())
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                peer_path()
            torch.cuda.synchronize(dev)
            peer_ms = _max_over_ranks(ctx, (time.perf_counter() - t0) / reps * 1e3)
        out.update({"equal_single_gpu": bool(int(ok[0])), "records_exchanged": int(tot[0]),
                    "bytes_over_nvlink_per_rank": 8 * nrec * (world - 1) + 4 * C * (world - 1),
                    "exchange": "kernel-level peer stores (CUDA IPC)" if ex is not None else "padded NCCL all-gathers",
                    "ms_exchange_plus_reduce_peer_stores": peer_ms,
                    "ms_exchange_plus_reduce_padded_nccl": _max_over_ranks(ctx, nccl_ms), "mAP_padded_nccl_path": m_nccl,
                    "exchange_error_flag": ex.error() if ex is not None else None})
        key = "cfg4_map_sharded"
    else:
        # the same evaluator over 1 M images: result() alone (records stay on the device)
        key = "cfg4_map_5k_images"
    return {key: out}


def map_one_million(ctx):
    """result() over the records of 1 M images (context for the reduce kernel at scale)."""
    torch, dev, yu = ctx["torch"], ctx["dev"], ctx["yu"]
    from tests import fixtures as F
    yt = F.synth_labels(50_000, seed=11); mp = F.synth_map_pred(yt)
    a, b_ = torch.from_numpy(yt).to(dev), torch.from_numpy(mp).to(dev)
    ev = yu.MeanAveragePrecision(C, B)
    for _ in range(20):                                               # 20 x 50k images
        ev.update_state(a, b_)
    ev.result()
    ms, mn = _timed(torch, dev, ev.result, 5)
    return {"map_1M_images": {"images": ev.img_idx, "records": int(ev._st["cursors"][0].item()), "ms_result": ms, "ms_result_min": mn}}


def other_configs(ctx):
    """cfg1 (batch 64), cfg2-sparse and cfg3 (loss fwd+bwd, batch 4096) - context numbers, N = 1."""
    import numpy as np
    from tests import fixtures as F
    torch, dev, L, _lib, sp = ctx["torch"], ctx["dev"], ctx["L"], ctx["_lib"], ctx["sp"]
    out = {}
    peak, _ = measured_peak()
    timed = lambda fn, reps: _timed(torch, dev, fn, reps)

    # cfg1: batch 64 (BASELINE configs[0], the reference's own CPU-runnable case): launch-bound, so also as a CUDA graph
    p64 = torch.from_numpy(F.synth_dense(64, seed=1234)).to(dev)
    b64 = torch.empty((64, M, 6), device=dev); c64 = torch.empty((64,), device=dev, dtype=torch.int32)
    f64 = lambda spx=sp: _lib.check(L.yh_decode_nms(p64.data_ptr(), 64, S, B, C, IOU_THR, CONF_THR, b64.data_ptr(), c64.data_ptr(), None, spx))

    def per_call_us(fn, reps=300):
        for _ in range(20):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return 1e3 * a.elapsed_time(b) / reps
    plain_us = per_call_us(f64)
    gs = torch.cuda.Stream(device=dev)
    gs.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(gs):
        f64(ctypes.c_void_p(gs.cuda_stream))
        gs.synchronize()
        with torch.cuda.graph(graph, stream=gs):
            f64(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.current_stream(dev).wait_stream(gs)
    graph_us = per_call_us(graph.replay)
    from oracle import cport as _cp
    x64 = p64.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(50):
        _, c_cpu, _ = _cp.decode_nms(x64, C, B, IOU_THR, CONF_THR, nthreads=1, want_idx=False)
    cpu_us = (time.perf_counter() - t0) / 50 * 1e6
    out["cfg1_batch64"] = {"us_per_call_back_to_back": plain_us, "us_per_cuda_graph_replay": graph_us,
                           "cpu_port_1_thread_us": cpu_us, "counts_equal_cpu": bool(np.array_equal(c_cpu, c64.cpu().numpy()))}
    # cfg2 sparse: confidences u^32, ~2.7 candidates per image
    n = env_int("YH_BENCH_IMAGES", 1_000_000)
    gen = torch.Generator(device=dev); gen.manual_seed(2025)
    p = torch.rand((n, S, S, D), generator=gen, device=dev)
    for b in range(B):
        p[..., C + 5 * b] = p[..., C + 5 * b] ** 32
        p[..., C + 5 * b + 3:C + 5 * b + 5] = 0.05 + 0.45 * p[..., C + 5 * b + 3:C + 5 * b + 5]
    boxes = torch.empty((n, M, 6), device=dev); cnt = torch.empty((n,), device=dev, dtype=torch.int32)
    f = lambda: _lib.check(L.yh_decode_nms(p.data_ptr(), n, S, B, C, IOU_THR, CONF_THR, boxes.data_ptr(), cnt.data_ptr(), None, sp))
    ms, mn = timed(f, 10)
    kept = int(cnt.sum().item())
    gbs = (n * (IMG_IN + 4) + 24 * kept) / (ms * 1e-3) / 1e9
    out["cfg2_sparse"] = {"images_per_s": n / (ms * 1e-3), "ms": ms, "GBps": gbs, "frac_hbm": gbs / peak, "kept_per_image": kept / n}
    del p, boxes, cnt
    # cfg3: loss fwd+bwd batch 4096.  The 72 MB working set fits the 126 MB L2, so the timed loop
    # rotates over 8 independent buffer sets (578 MB in total): every launch reads cold data.
    yt0 = torch.from_numpy(F.synth_labels(4096, seed=7)).to(dev)
    yp0 = torch.from_numpy(F.synth_loss_pred(tuple(yt0.shape), seed=7)).to(dev)
    sets = [(yt0.clone(), yp0.clone(), torch.empty_like(yp0)) for _ in range(8)]
    terms = torch.empty(6, device=dev)

    def loss_round():
        for (a_, b_, g_) in sets:
            _lib.check(L.yh_loss(a_.data_ptr(), b_.data_ptr(), 4096 * M, B, C, 5.0, 0.5, terms.data_ptr(), g_.data_ptr(), sp))
    ms, mn = timed(loss_round, 10)
    ms, mn = ms / len(sets), mn / len(sets)
    gbs = 3 * 4096 * IMG_IN / (ms * 1e-3) / 1e9
    del sets
    # what the memory system delivers for a launch of THIS size: a plain device copy with the same traffic
    half = 3 * 4096 * IMG_IN // 2
    cs = [(torch.empty(half, dtype=torch.uint8, device=dev), torch.empty(half, dtype=torch.uint8, device=dev)) for _ in range(8)]

    def copy_round():
        for (a_, b_) in cs:
            b_.copy_(a_)
    cms, _ = timed(copy_round, 10)
    cms /= len(cs)
    out["cfg3_loss_fwd_bwd_b4096"] = {"ms": ms, "ms_min": mn, "GBps": gbs, "frac_hbm": gbs / peak, "loss": float(terms[5]),
                                      "l2": "8 rotating buffer sets (578 MB > L2), back-to-back launches",
                                      "device_copy_same_traffic_ms": cms, "frac_of_same_size_copy": cms / ms}
    del cs
    out.update(map_one_million(ctx))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
