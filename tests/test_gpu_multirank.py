"""Multi-GPU parity (elements partitioned by Nek5000's rule, NCCL halo exchange + all-reduce): runs tests/mr_worker.py under
torchrun on 2 GPUs when the box has them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_parity(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + world), os.path.join(here, "mr_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MULTIRANK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
