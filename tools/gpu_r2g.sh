#!/bin/bash
# round-2 GPU call G: two-stream schedule A/B (same kernels; SDF-only work beside the render)
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "two_stream or gate_compaction" > gpurun_out/r2g_tests.log 2>&1
for o in 0 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap $o > gpurun_out/r2g_bench_o$o.json 2> gpurun_out/r2g_bench_o$o.err
done
tail -3 gpurun_out/r2g_tests.log
for o in 0 1 2 3; do python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_bench_o$o.json').read().strip().splitlines()[-1])
print($o, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['raster_fwd']['kernel_ms'])
PY
done
