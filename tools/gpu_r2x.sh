#!/bin/bash
# round-2 GPU call X: final single-GPU verification -- full GPU suite, smoke, bench (default + in-line), reference arm, launch list
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2x_tests.log 2>&1
tail -3 gpurun_out/r2x_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1
tail -1 gpurun_out/r2x_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2x_bench_inline.json 2> gpurun_out/r2x_bench_inline.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2x_ncu_bench.log 2>&1
for f in r2x_bench r2x_bench_inline; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('cpu_baseline',{}).get('value'), (d.get('stock_cuda') or {}).get('total_ms'))
PY
done
tail -c 400 gpurun_out/r2x_bench_reference.json
