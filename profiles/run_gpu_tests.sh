#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "FAILED|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -15
