#!/bin/bash
# round-2 GPU call S (2 GPUs): where the data-parallel step loses time -- per-stage table under DP, same cameras vs own cameras, NCCL priority
run() {  # name, extra flags
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 $2 > gpurun_out/r2s_$1.json 2> gpurun_out/r2s_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s_$1.json').read().strip().splitlines()[-1])
    print('$1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('stage_ms'))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/r2s_$1.err').read()[-800:])
PY
}
run inline "--overlap 0"
run inline_samecam "--overlap 0 --same-cameras"
run inline_hiprio "--overlap 0 --nccl-high-priority"
run twostream_hiprio "--overlap 1 --nccl-high-priority"
