#!/bin/bash
# GPU batch 5 (1 GPU): surface-first layout parity + A/B, ncu launch list / full captures inside the timed call, cfg-5 Arnoldi run
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pmg.py tests/test_gpu_matvec.py tests/test_gpu_cfg5_oracle.py tests/test_gpu_fullsize.py tests/test_gpu_newton.py -x -q -s 2>&1 ) > gpurun_out/b5_pytest.log 2>&1
tail -6 gpurun_out/b5_pytest.log
for pm in 1 0; do
  echo "== NSB_PERM=$pm"
  NSB_PERM=$pm timeout 600 python bench.py --steps 20 --warmup 5 --arnoldi 0 --no-cpu-baseline > gpurun_out/b5_bench_perm$pm.json 2> gpurun_out/b5_bench_perm$pm.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/b5_bench_perm$pm.json') if l.startswith('{')][-1])
    k = d['roofline']['kernels']
    print('ms/step', round(d['ms_per_step'], 3), 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'its', d['config']['pres_iters_per_step'], d['config']['helm_iters_per_comp_per_step'], 'step frac', round(d['roofline']['step']['frac'], 4), round(d['roofline']['step']['survey_contract_frac'], 4), 'setup', round(d['config']['setup_s'], 1))
    print({a: round(b['avg_ms'], 4) for a, b in k.items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/b5_bench_perm$pm.err').read()[-1500:])
PY
done
echo "== ncu launch list (timed call only)"
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/b5_ncu_launch.log 2>&1
wc -l gpurun_out/r2_launches.csv
echo "== ncu full A (advection / residual / Helmholtz loop)"
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_advab2|k_axhelm3|k_hcg_update|k_gs_sum|k_make_rhs" -s 0 -c 9 -f -o gpurun_out/r2_prof_helm python bench.py --steps 1 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/b5_ncu_a.log 2>&1
echo "== ncu full B (pressure loop)"
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_div3q|k_gradt3|k_pcg_fused_p|k_gs_sum|k_pm_" -s 80 -c 14 -f -o gpurun_out/r2_prof_pres python bench.py --steps 1 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/b5_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== cfg-5 Arnoldi (k_dim 100, schur_tgt 2)"
( time timeout 1500 python tools/run_arnoldi_cfg5.py 100 2 1300 ) > gpurun_out/b5_arnoldi_cfg5.log 2>&1
tail -25 gpurun_out/b5_arnoldi_cfg5.log
