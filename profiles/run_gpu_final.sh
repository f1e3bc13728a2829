#!/bin/bash
# Round-end GPU visit: full GPU test suite, bench line, phase profile, memcheck of the CGConv parity tests.
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "FAILED|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -8
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
head -c 600 gpurun_out/bench_n1.json; echo
timeout 200 python profiles/phase_profile.py 16384 > gpurun_out/phase_profile.txt 2>&1
tail -8 gpurun_out/phase_profile.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_cgconv.py -q -k "pipe or tc_det or crystal" > gpurun_out/memcheck_cgconv.txt 2>&1
tail -4 gpurun_out/memcheck_cgconv.txt
