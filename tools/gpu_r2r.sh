#!/bin/bash
# round-2 GPU call R: two-level grid walk of large rects -> parity tests, bench, launch list
timeout 900 python -m pytest tests/test_gpu_splat_parity.py tests/test_gpu_shim.py -q -m gpu -x > gpurun_out/r2r_tests.log 2>&1
tail -3 gpurun_out/r2r_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2r_bench_inline.json 2> gpurun_out/r2r_bench_inline.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2r_ncu_bench.log 2>&1
for f in r2r_bench r2r_bench_inline; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['stage_ms'])
PY
done
grep -i "tile_s" gpurun_out/r2r_launches.csv | tail -2 | cut -c1-200
