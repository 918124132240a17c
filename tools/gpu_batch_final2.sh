#!/bin/bash
# last GPU batch of round 2: evidence for the FINAL sources (ncu launch list + full captures stamped with the source hash),
# fallback-path test, cfg-5 matvec vs the C oracle at the full 19 960-element size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import bench; print(bench.kernel_source_hash())" > gpurun_out/bf2_source_hash.txt; cat gpurun_out/bf2_source_hash.txt
( time timeout 900 python -m pytest tests/test_gpu_matvec.py tests/test_gpu_ops.py tests/test_gpu_pmg.py -x -q -s 2>&1 ) > gpurun_out/bf2_pytest.log 2>&1
tail -4 gpurun_out/bf2_pytest.log
echo "== cfg-5 matvec vs the C oracle, full size (nz = 10)"
( time NSB_FULLSIZE_NZ=10 timeout 1200 python -m pytest tests/test_gpu_cfg5_oracle.py -x -q -s 2>&1 ) > gpurun_out/bf2_cfg5_fullsize.log 2>&1
grep -E "cfg5 nz|passed|failed" gpurun_out/bf2_cfg5_fullsize.log | cut -c1-700
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf2_bench.json 2> gpurun_out/bf2_bench.err
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/bf2_bench.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 3), {a: round(b['avg_ms'], 4) for a, b in d['roofline']['kernels'].items()})
PY
echo "== ncu launch list (timed call only)"
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf2_ncu_launch.log 2>&1
wc -l gpurun_out/r2_launches.csv
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_advab2|k_axhelm3|k_hcg_update|k_gs_sum|k_make_rhs" -s 0 -c 9 -f -o gpurun_out/r2_prof_helm python bench.py --steps 1 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf2_ncu_a.log 2>&1
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_div3q|k_gradt3|k_pcg_fused_p|k_gs_sum|k_pm_" -s 80 -c 14 -f -o gpurun_out/r2_prof_pres python bench.py --steps 1 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf2_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
