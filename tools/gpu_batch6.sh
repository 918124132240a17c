#!/bin/bash
# GPU batch 6 (1 GPU): full GPU test suite on the current build, A/B of the surface-first Helmholtz layout, default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q -s 2>&1 ) > gpurun_out/b6_pytest.log 2>&1
tail -6 gpurun_out/b6_pytest.log
grep -E "floquet multipliers|cavity:|cfg5 nz" gpurun_out/b6_pytest.log | cut -c1-600
for pm in 1 0; do
  echo "== NSB_PERM=$pm"
  NSB_PERM=$pm timeout 600 python bench.py --steps 20 --warmup 5 --arnoldi 0 --no-cpu-baseline > gpurun_out/b6_bench_perm$pm.json 2> gpurun_out/b6_bench_perm$pm.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/b6_bench_perm$pm.json') if l.startswith('{')][-1])
    k = d['roofline']['kernels']
    print('ms/step', round(d['ms_per_step'], 3), 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'its', d['config']['pres_iters_per_step'], d['config']['helm_iters_per_comp_per_step'], 'step frac', round(d['roofline']['step']['frac'], 4), round(d['roofline']['step']['survey_contract_frac'], 4), 'setup', round(d['config']['setup_s'], 1))
    print({a: round(b['avg_ms'], 4) for a, b in k.items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/b6_bench_perm$pm.err').read()[-1500:])
PY
done
