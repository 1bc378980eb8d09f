"""Multi-GPU checks as pytest cases (-m gpu): skipped on a one-GPU box, run as subprocesses on a box with two or more
(tests/multigpu_check.py under torchrun: one process per GPU, NCCL + the CUDA-IPC peer-store exchange;
tests/multigpu_capi_check.py: one process driving all devices through the C-ABI exchange entry points)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_one_process_per_gpu_world2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "multigpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_single_process_all_devices():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multigpu_capi_check.py"), "2"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "multigpu_capi_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
