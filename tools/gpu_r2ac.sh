#!/bin/bash
# round-2 GPU call AC: projection with the reference build's fp32 structure -> radii mismatch counts, full GPU suite, smoke, bench
timeout 900 python -m pytest tests/test_gpu_splat_parity.py -q -m gpu -s -k "live_at_baseline" 2>&1 | grep -i "radii differ\|passed\|failed" | head
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2ac_tests.log 2>&1
tail -3 gpurun_out/r2ac_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ac_smoke.log 2>&1
tail -1 gpurun_out/r2ac_smoke.log | cut -c1-160
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['counts'])
PY
