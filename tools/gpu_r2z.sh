#!/bin/bash
# round-2 GPU call Z: SDF Adam on the SDF stream -> schedule test, bench
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "two_stream" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2z_bench_b.json 2> gpurun_out/r2z_bench_b.err
for f in r2z_bench r2z_bench_b; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['loss_end'])
PY
done
tail -c 300 gpurun_out/r2z_bench.err
