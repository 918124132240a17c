#!/bin/bash
# CUDA-graph replay on the peer-memory data plane (N > 1): multi-rank parity tests, then the bench line with and without graph replay
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -s 2>&1 ) > gpurun_out/bg${N}_pytest.log 2>&1
tail -4 gpurun_out/bg${N}_pytest.log
for G in ${GRAPH_MODES:-1 0}; do
  echo "== NSB_GRAPHS=$G"
  ( time NSB_GRAPHS=$G timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$G bench.py --gpus $N --steps 20 --warmup 5 --arnoldi 0 --no-cpu-baseline ) > gpurun_out/bg${N}_bench_g$G.json 2> gpurun_out/bg${N}_bench_g$G.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bg${N}_bench_g$G.json') if l.startswith('{')][-1])
    print({a: d[a] for a in ('n_gpus', 'value', 'ms_per_step', 'data_plane', 'parity_n')}, 'its', d['config']['pres_iters_per_step'], 'e2e %.4g' % d['e2e']['value'], d['config'].get('timing'))
    print('strong', d.get('strong'))
except Exception as e:
    print('failed', e)
PY
  tail -3 gpurun_out/bg${N}_bench_g$G.err
done
