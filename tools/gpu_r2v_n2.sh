#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parallel_nccl.py -q -m gpu -x -s > gpurun_out/r2v_tests.log 2>&1
grep -n "sparse_replicas\|passed\|failed" gpurun_out/r2v_tests.log | head
