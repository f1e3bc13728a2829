#!/bin/bash
# One GPU-box visit: parity tests of the CGConv path, phase profile + switch A/B, bench line.
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 1200 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_models.py tests/test_gpu_engine.py tests/test_gpu_store.py -x -q > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 600 python profiles/phase_profile.py ${1:-16384} > gpurun_out/phase_profile.txt 2>&1
cat gpurun_out/phase_profile.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
