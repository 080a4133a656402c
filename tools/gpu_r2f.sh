#!/bin/bash
# round-2 GPU call F: staged traversal + multi-block scan -> octree/gate tests, smoke, bench, launch list
timeout 900 python -m pytest tests/test_gpu_octree.py tests/test_gpu_round2.py tests/test_gpu_densify.py -q -m gpu -x > gpurun_out/r2f_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2f_ncu_bench.log 2>&1
tail -3 gpurun_out/r2f_tests.log; tail -1 gpurun_out/r2f_smoke.log | cut -c1-200
tail -c 300 gpurun_out/r2f_bench.err
