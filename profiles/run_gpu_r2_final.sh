#!/bin/bash
# final build of round 2: whole GPU suite, smoke, the default bench line and the full lines of configs 2..4
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt gpurun_out/parity_errors_baseline_shapes.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -15
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "bench c1 exit $?"; head -c 260 gpurun_out/bench_c1.json; echo; tail -2 gpurun_out/bench_c1.err
for c in 2 3 4; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 --cpu-steps 2 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  echo "config $c exit $?"; head -c 260 gpurun_out/bench_c$c.json; echo
done
