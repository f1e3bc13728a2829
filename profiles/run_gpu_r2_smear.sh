#!/bin/bash
# smearing-fused CGConv: its parity tests first, then the whole GPU suite, the config-1 bench line and the serialised
# launch lists of one eager step of configs 2..4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cgconv_smear.py -x -q > gpurun_out/tests_smear.log 2>&1
echo "pytest smear exit $?" >> gpurun_out/tests_smear.log
tail -25 gpurun_out/tests_smear.log
rm -f gpurun_out/parity_errors.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -15
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 --cpu-steps 2 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "bench exit $?"; head -c 600 gpurun_out/bench_c1.json; echo; tail -3 gpurun_out/bench_c1.err
for c in 2 3 4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c$c.csv \
     python profiles/model_step_launches.py $c > gpurun_out/launches_c$c.log 2>&1
  echo "ncu config $c exit $?"
done
