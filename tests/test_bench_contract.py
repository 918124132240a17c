"""The bench.py output contract on the CPU-runnable arm: `--impl reference` prints exactly one JSON line carrying the keys the
driver reads (metric, value, unit, impl, cpu_baseline, e2e, config.workload), under plain python and as rank 1 of a
2-process launch (which must print nothing and exit 0)."""
import json
import os
import subprocess
import sys

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


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""
