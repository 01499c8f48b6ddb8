"""Multi-GPU checks on real GPUs (timestep / sample / instance sharding, the fused peer-memory exchange
and its failure behaviour): tests/multi_gpu_worker.py under torchrun, one rank per GPU.  Skipped on a
single-GPU box; the host-side plumbing is covered on CPU by tests/test_distributed_cpu.py (gloo)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_paths_on_all_local_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (found %d)" % n)
    world = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout)
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert "ALL PASS" in r.stdout and "FAIL" not in r.stdout
