#!/bin/bash
# round-2 GPU call Y (4 GPUs): final data-parallel lines (N = 4, 2, 1) + the 2-GPU NCCL test
timeout 600 python -m pytest tests/test_gpu_parallel_nccl.py -q -m gpu -x 2>&1 | tail -2
run() {  # name, nproc, extra flags
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $2 --steps 20 --warmup 3 $3 > gpurun_out/r2y_$1.json 2> gpurun_out/r2y_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2y_$1.json').read().strip().splitlines()[-1])
    print('$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('splat_exchange_steps'), (d.get('stage_ms') or {}).get('wait_splat_allreduce'))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/r2y_$1.err').read()[-800:])
PY
}
run n4 4 ""
run n4_dense 4 "--dense-allreduce"
run n2 2 ""
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2y_n1.json 2> gpurun_out/r2y_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2y_n1.json').read().strip().splitlines()[-1])
print('n1', d['ms_per_step'], d['value'])
PY
