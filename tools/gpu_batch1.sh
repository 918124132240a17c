#!/bin/bash
# GPU batch 1 (round 2): full GPU test suite, default bench line, cfg-1 tolerance study
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/b1_smi.txt 2>&1
nproc > gpurun_out/b1_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 ) > gpurun_out/b1_pytest.log 2>&1
tail -5 gpurun_out/b1_pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/b1_bench.json 2> gpurun_out/b1_bench.err
tail -c 1500 gpurun_out/b1_bench.json
( timeout 300 python - <<'PY'
import sys, json
sys.path.insert(0, 'tools')
import run_arnoldi_cfg1
a = run_arnoldi_cfg1.run(200, 1e-7, 1e-9, "pmg", 0, "direct", tag_suffix="_loose")
b = run_arnoldi_cfg1.run(200, 1e-11, 1e-11, "pmg", 0, "direct", tag_suffix="_tight")
import numpy as np
va = np.array([x[0] + 1j * x[1] for x in a["ritz_values_first_24"]]); vb = np.array([x[0] + 1j * x[1] for x in b["ritz_values_first_24"]])
d = [float(abs(v - vb[np.argmin(abs(vb - v))]) / abs(v)) for v in va]
json.dump({"loose": a, "tight": b, "rel_diff_between_our_two_runs_first_24": d}, open('gpurun_out/b1_cfg1_tolerance_study.json', 'w'), indent=1)
print("tolerance study rel diffs:", ["%.1e" % x for x in d])
PY
) > gpurun_out/b1_tolstudy.log 2>&1
tail -3 gpurun_out/b1_tolstudy.log
