#!/bin/bash
# final 1-GPU batch of round 2: full GPU test suite, the driver's bench commands (both arms), ncu launch list + full captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import bench; print('kernel source hash', bench.kernel_source_hash())" > gpurun_out/bf_source_hash.txt
( time timeout 1800 python -m pytest tests -m gpu -x -q -s 2>&1 ) > gpurun_out/bf_pytest.log 2>&1
tail -5 gpurun_out/bf_pytest.log
echo "== bench (ours)"
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bf_bench.json 2> gpurun_out/bf_bench.err
tail -c 1200 gpurun_out/bf_bench.json; tail -3 gpurun_out/bf_bench.err
echo "== A/B: direction kernel at 4 CTAs per SM (40 registers)"
NSB_GRADT_LB=4 timeout 600 python bench.py --steps 20 --warmup 5 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf_bench_lb4.json 2> gpurun_out/bf_bench_lb4.err
python - <<PY
import json
for f in ('gpurun_out/bf_bench.json', 'gpurun_out/bf_bench_lb4.json'):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), {a: round(b['avg_ms'], 4) for a, b in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'failed', e)
PY
echo "== bench (reference arm)"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bf_bench_ref.json 2> gpurun_out/bf_bench_ref.err
tail -c 1500 gpurun_out/bf_bench_ref.json; tail -3 gpurun_out/bf_bench_ref.err
echo "== ncu launch list (timed call only)"
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf_ncu_launch.log 2>&1
wc -l gpurun_out/r2_launches.csv
echo "== ncu full A (advection / residual / Helmholtz loop)"
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_advab2|k_axhelm3|k_hcg_update|k_gs_sum|k_make_rhs" -s 0 -c 9 -f -o gpurun_out/r2_prof_helm python bench.py --steps 1 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf_ncu_a.log 2>&1
echo "== ncu full B (pressure loop)"
NSB_GRAPHS=0 NSB_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_div3q|k_gradt3|k_pcg_fused_p|k_gs_sum|k_pm_" -s 80 -c 14 -f -o gpurun_out/r2_prof_pres python bench.py --steps 1 --warmup 1 --arnoldi 0 --no-cpu-baseline > gpurun_out/bf_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
