#!/bin/bash
# memcheck + racecheck of the pipelined forward kernel's parity tests, then the whole GPU suite
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_cgconv.py -q -k "pipe or tc_det or crystal" > gpurun_out/memcheck_cgconv.txt 2>&1
tail -4 gpurun_out/memcheck_cgconv.txt
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_cgconv.py -q -k "pipe and (tiny or crystal or add_aggr)" > gpurun_out/racecheck_cgconv_fwd.txt 2>&1
tail -4 gpurun_out/racecheck_cgconv_fwd.txt
rm -f gpurun_out/parity_errors.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "FAILED|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -8
