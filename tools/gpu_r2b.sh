#!/bin/bash
# round-2 GPU call B: fixed tests + octree tests + bench with sample generation in the step
cp gpurun_out/golden/tcnn_grid_ref*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_octree.py -q -m gpu -s > gpurun_out/r2b_tests.log 2>&1
timeout 900 python -m pytest tests/test_gpu_splat_parity.py -q -m gpu -s -k "live" >> gpurun_out/r2b_tests.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
grep -E "passed|failed" gpurun_out/r2b_tests.log
tail -c 400 gpurun_out/r2b_bench.err
