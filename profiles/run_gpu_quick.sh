#!/bin/bash
# quick GPU visit: CGConv parity tests first (new kernels trap instead of hanging: umma::mbar_wait is bounded),
# then the whole GPU suite, then the phase profile / switch A/B at roofline size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cgconv.py -x -q > gpurun_out/tests_cgconv.log 2>&1
echo "pytest cgconv exit $?" >> gpurun_out/tests_cgconv.log
tail -15 gpurun_out/tests_cgconv.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_cgconv.py > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
tail -25 gpurun_out/tests.log
timeout 600 python profiles/phase_profile.py ${1:-16384} > gpurun_out/phase_profile.txt 2>&1
cat gpurun_out/phase_profile.txt
