#!/bin/bash
# GPU batch 4b (2 GPUs): multi-rank parity (world 2), N=2 bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -s 2>&1 ) > gpurun_out/b4b_pytest.log 2>&1
tail -6 gpurun_out/b4b_pytest.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/b4b_bench_2gpu.json 2> gpurun_out/b4b_bench_2gpu.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/b4b_bench_2gpu.json') if l.startswith('{')][-1])
    k = d['roofline']['kernels']
    print({a: d[a] for a in ('value', 'ms_per_step', 'data_plane', 'parity_n')}, 'its', d['config']['pres_iters_per_step'], 'e2e %.4g' % d['e2e']['value'])
    print({a: round(b['avg_ms'], 4) for a, b in k.items()})
    print('strong', d['strong'])
    print('arnoldi', {a: d['arnoldi'][a] for a in ('wall_s_per_iteration', 'pres_iters_per_step')})
except Exception as e:
    print('failed', e)
PY
tail -5 gpurun_out/b4b_bench_2gpu.err
