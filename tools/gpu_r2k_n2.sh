#!/bin/bash
# round-2 GPU call K (2 GPUs): data-parallel bench, two-stream schedule (default) and in-line, each rank its own cameras
for o in 1 0; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --overlap $o > gpurun_out/r2k_bench_n2_o$o.json 2> gpurun_out/r2k_bench_n2_o$o.err
tail -c 600 gpurun_out/r2k_bench_n2_o$o.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2k_bench_n2_o$o.json').read().strip().splitlines()[-1])
    print('n2 overlap $o', d['ms_per_step'], d['value'], d['e2e'])
except Exception as e: print('failed', e)
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2k_bench_n1.json').read().strip().splitlines()[-1])
print('n1', d['ms_per_step'], d['value'])
PY
