#!/bin/bash
# GPU batch 3 (round 2), 2 GPUs: parity of the new kernels, A/B of k_div3q, multi-rank tests, first N=2 bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pmg.py tests/test_gpu_matvec.py tests/test_gpu_cfg5_oracle.py tests/test_gpu_fullsize.py tests/test_gpu_multirank.py -x -q -s 2>&1 ) > gpurun_out/b3_pytest.log 2>&1
tail -6 gpurun_out/b3_pytest.log
for dq in 1 0; do
  echo "== NSB_DIVQ=$dq"
  NSB_DIVQ=$dq CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --arnoldi 0 --no-cpu-baseline > gpurun_out/b3_bench_q$dq.json 2> gpurun_out/b3_bench_q$dq.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/b3_bench_q$dq.json') if l.startswith('{')][-1])
    k = d['roofline']['kernels']
    print('ms/step', round(d['ms_per_step'], 3), 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'its', d['config']['pres_iters_per_step'], d['config']['helm_iters_per_comp_per_step'], 'step frac', round(d['roofline']['step']['frac'], 4), round(d['roofline']['step']['survey_contract_frac'], 4))
    print({a: round(b['avg_ms'], 4) for a, b in k.items()})
except Exception as e:
    print('failed', e); print(open('gpurun_out/b3_bench_q$dq.err').read()[-1500:])
PY
done
echo "== 2-GPU bench"
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/b3_bench_2gpu.json 2> gpurun_out/b3_bench_2gpu.err
tail -c 3000 gpurun_out/b3_bench_2gpu.json
tail -5 gpurun_out/b3_bench_2gpu.err
