"""Data-parallel training on >= 2 real GPUs (NCCL): bucketed, overlapped gradient all-reduce == plain all-reduce of the
local gradients, replicas stay bit-identical (replaces DistributedDataParallel, multi_gpu_train2.py:89). Skipped on a
single-GPU box; the host logic is covered on CPU by tests/test_multiproc_cpu.py (gloo, world size 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_bucketed_allreduce_matches_plain_allreduce_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ddp_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
