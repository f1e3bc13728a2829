#!/bin/bash
# Round-2 state of the build on one B200: whole GPU suite, bench lines of configs 1..4, phase profile,
# serialised launch list of the config-1 bench and one `ncu --set full` capture of the CGConv forward / backward kernels.
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -15
for c in 1 2 3 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --cpu-steps 2 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  echo "config $c exit $?"; head -c 400 gpurun_out/bench_c$c.json; echo; tail -2 gpurun_out/bench_c$c.err
done
timeout 300 python profiles/phase_profile.py 16384 > gpurun_out/phase_profile.txt 2>&1
tail -30 gpurun_out/phase_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c1.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/launches_c1.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_cgconv_fwd|k_cgconv_bwd" -s 12 -c 2 \
   -o gpurun_out/r2_ncu_full_cgconv -f python bench.py --roofline-only > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
