#!/bin/bash
# round-2 GPU call D: full GPU suite after the persistent forward / sampling changes, bench, other BASELINE configs
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2d_tests.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
for w in c2 c3 c5; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline --no-stock-cuda > gpurun_out/r2d_bench_$w.json 2>> gpurun_out/r2d_bench.err
done
tail -3 gpurun_out/r2d_tests.log
tail -c 400 gpurun_out/r2d_bench.err
