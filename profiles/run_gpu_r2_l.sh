#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_store.py -q -x > gpurun_out/tests_engine.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_engine.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_engine.log | tail -6
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline --no-other-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"; head -c 200 gpurun_out/bench_quick.json; echo
