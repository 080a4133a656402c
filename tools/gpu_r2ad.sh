#!/bin/bash
# round-2 GPU call AD: full GPU suite on the final build
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2ad_tests.log 2>&1
tail -4 gpurun_out/r2ad_tests.log
grep -n "^FAILED\|^ERROR" gpurun_out/r2ad_tests.log | head
