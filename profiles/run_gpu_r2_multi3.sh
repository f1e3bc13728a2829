#!/bin/bash
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/tests_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_multi.log
grep -E "^E  |FAILED|ERROR|passed|failed|skipped|pytest exit" gpurun_out/tests_multi.log | tail -12
