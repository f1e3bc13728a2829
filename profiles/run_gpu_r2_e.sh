#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_store.py tests/test_gpu_builder.py -q > gpurun_out/tests_store.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_store.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_store.log | tail -25
