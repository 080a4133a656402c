#!/bin/bash
# round-2 GPU call Q: state after the revert -> parity tests, bench, launch list
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2q_tests.log 2>&1
tail -3 gpurun_out/r2q_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2q_bench_inline.json 2> gpurun_out/r2q_bench_inline.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2q_ncu_bench.log 2>&1
for f in r2q_bench r2q_bench_inline; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['stage_ms'])
PY
done
grep -i "tile_s" gpurun_out/r2q_launches.csv | tail -2 | cut -c1-200
