"""-m gpu, needs >= 2 GPUs (skipped otherwise): the partitioned path under torchrun + NCCL against the
oracle (scripts/multi_gpu_check.py): distributed vmult <= 1e-12, merged CG iteration parity."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_parity_against_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAIL" not in r.stdout and r.stdout.count("OK  ") >= 10     # 5 cases x (peer, nccl) transports
