"""Multi-GPU end to end on hardware: sharding.generate_sharded under torchrun with NCCL - independent
seeds per rank, no collective in the loop, one all_gather of the uint8 images - equals the union of single-GPU runs
byte for byte. Skipped on boxes with fewer than two GPUs (the CPU suite covers the host logic over gloo)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("seeds", ["42,43,44,45", "7,8,9"])
def test_generate_sharded_over_nccl_equals_single_gpu_runs(seeds):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs two GPUs, found {n}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_worker.py"), seeds, "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout
