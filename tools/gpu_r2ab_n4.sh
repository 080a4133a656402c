#!/bin/bash
# round-2 GPU call AB (4 GPUs): sample generation + [A] ahead of the render while a dense all-reduce is in flight -- A/B at N = 4 and N = 2
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "two_stream" 2>&1 | tail -2
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $2 --steps 20 --warmup 3 $3 > gpurun_out/r2ab_$1.json 2> gpurun_out/r2ab_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ab_$1.json').read().strip().splitlines()[-1])
    print('$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('splat_exchange_steps'))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/r2ab_$1.err').read()[-800:])
PY
}
run n4_cover 4 ""
run n4_nocover 4 "--no-cover"
run n2_dense_cover 2 "--dense-allreduce"
run n2_sparse 2 ""
