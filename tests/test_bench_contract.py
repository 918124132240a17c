"""The bench.py output contract on the CPU-runnable arm: `--impl reference` prints exactly one JSON line carrying the keys the
driver reads (metric, value, unit, impl, cpu_baseline, e2e, config.workload), under plain python and as rank 1 of a
2-process launch (which must print nothing and exit 0)."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "linearized-NS DOF*timesteps/s" and d["unit"] == "DOF*steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "hexahedra" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "DOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the CPU arm runs to tolerance and reports its own measured iteration counts (no forced counts)
    assert d["config"]["pres_iters_per_step"] > 0 and d["config"]["helm_iters_per_comp_per_step"] > 0
    assert "forced" not in cb["sample"] and "solvers run to tol" in cb["sample"]
    assert d["config"]["workload"] == "cyl3d_1996x10_lx8" and d["config"]["elements"] == 19960      # = the GPU arm's workload name


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""


@pytest.mark.gpu
def test_gpu_and_cpu_arms_do_the_same_work_on_the_small_mesh():
    """on the same (small) mesh, seed, dt, tolerances and steps the CPU arm's measured iteration counts must
    agree with the GPU arm's (same algorithm, same preconditioner): within 10 %."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--small", "--steps", "3", "--warmup", "2", "--arnoldi", "0",
                        "--cpu-seconds", "150"], capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    cb = d["cpu_baseline"]
    assert cb["value"] and cb["steps_done"] == 3, cb
    gp, cp = d["config"]["pres_iters_per_step"], cb["pres_iters_per_step"]
    gh, ch = d["config"]["helm_iters_per_comp_per_step"], cb["helm_iters_per_comp_per_step"]
    assert abs(gp - cp) <= 0.10 * cp, (gp, cp)
    assert abs(gh - ch) <= 0.10 * ch, (gh, ch)
    assert d["gpu_launches"] > 0 and d["e2e"]["value"] > 0 and d["roofline"]["frac"] > 0
