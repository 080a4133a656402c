#!/bin/bash
# round-2 GPU call T (4 GPUs): data-parallel bench at N = 4 and N = 2 with the shared camera pool, + identical-pose control
run() {  # name, nproc, extra flags
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $2 --steps 20 --warmup 3 $3 > gpurun_out/r2t_$1.json 2> gpurun_out/r2t_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2t_$1.json').read().strip().splitlines()[-1])
    print('$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], (d.get('stage_ms') or {}).get('wait_splat_allreduce'))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/r2t_$1.err').read()[-800:])
PY
}
run n4 4 ""
run n4_samecam 4 "--same-cameras"
run n2 2 ""
run n2_samecam 2 "--same-cameras"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2t_n1.json 2> gpurun_out/r2t_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_n1.json').read().strip().splitlines()[-1])
print('n1', d['ms_per_step'], d['value'])
PY
