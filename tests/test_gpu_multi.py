"""Multi-GPU test (-m gpu; skipped with fewer than 2 visible GPUs): one process per GPU via torchrun, NCCL.
Row-sharded fit (one all-reduce of the totals), row-sharded fold batch (two all-reduces) and fold-sharded batch,
all against the numpy oracle."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_paths():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_fit_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_OK 2" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
