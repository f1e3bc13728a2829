#!/bin/bash
# quick GPU visit: CGConv parity tests + phase profile / switch A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_models.py -x -q > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
tail -4 gpurun_out/tests.log
timeout 600 python profiles/phase_profile.py ${1:-16384} > gpurun_out/phase_profile.txt 2>&1
cat gpurun_out/phase_profile.txt
