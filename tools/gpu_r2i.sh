#!/bin/bash
# round-2 GPU call I: full GPU test suite, smoke, bench (default = two-stream schedule) + in-line A/B, launch list
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2i_tests.log 2>&1
rc=$?
tail -4 gpurun_out/r2i_tests.log
if [ $rc -ne 0 ]; then echo "tests failed"; grep -n "Error\|assert\|FAILED" gpurun_out/r2i_tests.log | head -30; fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
tail -1 gpurun_out/r2i_smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2i_bench_inline.json 2> gpurun_out/r2i_bench_inline.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap 0 > gpurun_out/r2i_ncu_bench.log 2>&1
for f in r2i_bench r2i_bench_inline; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['stage_ms'])
PY
done
