#!/bin/bash
# bench lines of every BASELINE config (short runs) + serialised launch lists of one step of configs 2..4
mkdir -p gpurun_out
for c in 1 2 3 4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --cpu-steps 2 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  echo "config $c exit $?"; tail -c 1500 gpurun_out/bench_c$c.json; tail -3 gpurun_out/bench_c$c.err
done
for c in 2 3 4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c$c.csv \
     python profiles/model_step_launches.py $c > gpurun_out/launches_c$c.log 2>&1
  echo "ncu config $c exit $?"
done
python profiles/summarize_step_launches.py gpurun_out/launches_c2.csv | head -40
