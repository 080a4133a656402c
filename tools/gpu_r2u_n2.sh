#!/bin/bash
# round-2 GPU call U (2 GPUs): sparse visible-row exchange -> unit tests (rows ops, NCCL sparse vs dense), bench sparse vs dense
timeout 900 python -m pytest tests/test_gpu_parallel_nccl.py tests/test_gpu_round2.py -q -m gpu -x -k "rows_pack or sparse_row" > gpurun_out/r2u_tests.log 2>&1
tail -4 gpurun_out/r2u_tests.log
grep -n "Error\|assert" gpurun_out/r2u_tests.log | head -10
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 $2 > gpurun_out/r2u_$1.json 2> gpurun_out/r2u_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2u_$1.json').read().strip().splitlines()[-1])
    print('$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('splat_exchange_steps'), (d.get('stage_ms') or {}).get('wait_splat_allreduce'))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/r2u_$1.err').read()[-1200:])
PY
}
run sparse ""
run dense "--dense-allreduce"
run sparse_inline "--overlap 0"
run dense_inline "--overlap 0 --dense-allreduce"
