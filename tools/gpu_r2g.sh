#!/bin/bash
# round-2 GPU call G: 8-lane ray traversal + two-stream schedule A/B + ncu --set full of the mid-size kernels
timeout 900 python -m pytest tests/test_gpu_octree.py tests/test_gpu_round2.py tests/test_gpu_splat_parity.py -q -m gpu -x > gpurun_out/r2g_tests.log 2>&1
rc=$?
tail -5 gpurun_out/r2g_tests.log
if [ $rc -ne 0 ]; then echo "tests failed: stopping"; grep -n "Error\|assert" gpurun_out/r2g_tests.log | head -20; exit 1; fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
tail -1 gpurun_out/r2g_smoke.log | cut -c1-200
for o in 0 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stock-cuda --overlap $o > gpurun_out/r2g_bench_o$o.json 2> gpurun_out/r2g_bench_o$o.err
done
for o in 0 1 2 3; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2g_bench_o$o.json').read().strip().splitlines()[-1])
    print($o, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['raster_fwd']['kernel_ms'], (d.get('stage_ms') or {}).get('rng+sample_generation[A0]'))
except Exception as e:
    print($o, 'failed', e)
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dssim_fwd|dssim_bwd|tile_count|tile_scatter|tile_cull|ray_count|normal_consistency|tile_sort" --launch-skip 40 -c 10 -o gpurun_out/r2g_mid -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-cuda > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
