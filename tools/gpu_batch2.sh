#!/bin/bash
# GPU batch 2 (round 2): fused pressure-CG tail + persistent axhelm: parity, then A/B bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pmg.py tests/test_gpu_matvec.py tests/test_gpu_cfg5_oracle.py tests/test_gpu_fullsize.py tests/test_gpu_cavity.py -x -q -s 2>&1 ) > gpurun_out/b2_pytest.log 2>&1
tail -8 gpurun_out/b2_pytest.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== NSB_PCG_FUSED=$1 NSB_AX_PERSISTENT=$2"
  NSB_PCG_FUSED=$1 NSB_AX_PERSISTENT=$2 timeout 600 python bench.py --steps 20 --warmup 5 --arnoldi 0 --no-cpu-baseline > gpurun_out/b2_bench_f$1_a$2.json 2> gpurun_out/b2_bench_f$1_a$2.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/b2_bench_f$1_a$2.json') if l.startswith('{')][-1])
    k = d['roofline']['kernels']
    print('ms/step', round(d['ms_per_step'], 3), 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'its', d['config']['pres_iters_per_step'], d['config']['helm_iters_per_comp_per_step'], 'step frac', round(d['roofline']['step']['frac'], 4))
    print({a: round(b['avg_ms'], 4) for a, b in k.items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/b2_bench_f$1_a$2.err').read()[-1500:])
PY
done
