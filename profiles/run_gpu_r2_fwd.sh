#!/bin/bash
# forward-kernel iteration: CGConv parity tests (both forms), phase profile + A/B at roofline size, racecheck of the ws kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_cgconv_smear.py tests/test_gpu_engine.py tests/test_gpu_store.py -x -q > gpurun_out/tests_cgconv.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_cgconv.log
tail -6 gpurun_out/tests_cgconv.log
timeout 300 python profiles/phase_profile.py 16384 > gpurun_out/phase_profile.txt 2>&1
grep -A9 "== fwd" gpurun_out/phase_profile.txt; grep "A/B" gpurun_out/phase_profile.txt
timeout 300 python bench.py --roofline-only > gpurun_out/roofline_only.json 2> gpurun_out/roofline_only.err; cat gpurun_out/roofline_only.json | head -c 1500; echo
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_cgconv.py -q -k "ws and (tiny or crystal or hub)" > gpurun_out/racecheck_ws.txt 2>&1
tail -4 gpurun_out/racecheck_ws.txt
