"""N>1 host logic on CPU: partition + the C++ gather-scatter plan of csrc/gs.cu driven over gloo (world_size 2 and 3)."""
import os
import subprocess
import sys

import pytest


@pytest.mark.parametrize("world", [2, 3])
def test_gs_plan_over_gloo(world):
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(here, "gloo_worker.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
