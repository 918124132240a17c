#!/bin/bash
# multi-GPU batch: usage  bash tools/gpu_batch_multi.sh N   (N = 2, 4, 8): multi-rank parity tests (world <= N) + bench line at N
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -q -s 2>&1 ) > gpurun_out/bm${N}_pytest.log 2>&1
tail -4 gpurun_out/bm${N}_pytest.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bm${N}_bench.json 2> gpurun_out/bm${N}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bm${N}_bench.json') if l.startswith('{')][-1])
    k = d['roofline']['kernels']
    print({a: d[a] for a in ('n_gpus', 'value', 'ms_per_step', 'data_plane', 'parity_n')}, 'its', d['config']['pres_iters_per_step'], 'e2e %.4g' % d['e2e']['value'], 'setup', round(d['config']['setup_s'], 1))
    print({a: round(b['avg_ms'], 4) for a, b in k.items()})
    print('strong', d['strong'])
    print('arnoldi', {a: d['arnoldi'][a] for a in ('wall_s_per_iteration', 'pres_iters_per_step')})
except Exception as e:
    print('failed', e)
PY
tail -4 gpurun_out/bm${N}_bench.err
