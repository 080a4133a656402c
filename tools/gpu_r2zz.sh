#!/bin/bash
# round-2 GPU call ZZ: the other workloads with the final build (c2, c3, c5, c1) + the radii mismatch counts of the live reference comparison
timeout 600 python -m pytest tests/test_gpu_splat_parity.py -q -m gpu -s -k "live_at_baseline" 2>&1 | grep -i "radii differ\|passed\|failed" | head
for w in c2 c3 c5 c1; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2zz_bench_$w.json 2> gpurun_out/r2zz_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2zz_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['counts'])
except Exception as e:
    print('$w failed', e); print(open('gpurun_out/r2zz_bench_$w.err').read()[-600:])
PY
done
