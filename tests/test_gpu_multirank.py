"""Multi-GPU parity (elements partitioned by Nek5000's rule, NCCL halo exchange + all-reduce): runs tests/mr_worker.py under
torchrun on 2 GPUs when the box has them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(here, "mr_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MULTIRANK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
