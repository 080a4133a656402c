#!/bin/bash
# round-2 GPU call C: new tests + bench + A/B variants + shim bench + ncu launch list + full captures
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_densify.py tests/test_gpu_sdf_parity.py tests/test_gpu_octree.py -q -m gpu -s > gpurun_out/r2c_tests.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
GSSDF_RASTER_BWD_VARIANT=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2c_bench_bwdvar1.json 2>> gpurun_out/r2c_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda --no-l2-persist > gpurun_out/r2c_bench_nol2.json 2>> gpurun_out/r2c_bench.err
timeout 600 python tools/shim_bench.py > gpurun_out/r2c_shim.json 2>> gpurun_out/r2c_bench.err
GSSDF_SHIM_PRESORT_CULL=1 timeout 600 python tools/shim_bench.py > gpurun_out/r2c_shim_cull.json 2>> gpurun_out/r2c_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2c_ncu_bench.log 2>&1
for k in raster2dgs_bwd_kernel raster2dgs_fwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/r2c_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdf_bwd_tc_kernel -s 13 -c 1 -o gpurun_out/r2c_sdf_train_splats -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda > /dev/null 2>&1
grep -E "passed|failed" gpurun_out/r2c_tests.log
tail -c 300 gpurun_out/r2c_bench.err
tail -2 gpurun_out/r2c_smoke.log
ls -la gpurun_out/*.ncu-rep | tail -4
