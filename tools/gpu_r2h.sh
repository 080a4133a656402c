#!/bin/bash
# round-2 GPU call H: ncu --set full of the streaming DSSIM kernels + launch list of the step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dssim_fwd|dssim_bwd" --launch-skip 8 -c 2 -o gpurun_out/r2h_dssim -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2h_ncu.log 2>&1
tail -2 gpurun_out/r2h_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2h_ncu_bench.log 2>&1
grep -c dssim gpurun_out/r2h_launches.csv
