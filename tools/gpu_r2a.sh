#!/bin/bash
# round-2 GPU call A: goldens, all GPU tests, first bench line with the new step
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
python oracle/gen_golden_tcnn.py gpurun_out/golden/tcnn_grid_ref.npz > gpurun_out/r2a_golden.log 2>&1
cp gpurun_out/golden/tcnn_grid_ref*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -s > gpurun_out/r2a_tests_round2.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu -s --deselect tests/test_gpu_round2.py > gpurun_out/r2a_tests_all.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_tests_round2.log gpurun_out/r2a_tests_all.log
tail -c 600 gpurun_out/r2a_bench.err
