#!/bin/bash
# Round-end GPU visit: full GPU test suite, bench line, step launch list, ncu --set full of both edge kernels.
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
tail -3 gpurun_out/tests_gpu_full.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1500 gpurun_out/bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cgconv_fwd_pipe -s 3 -c 1 -f -o gpurun_out/ncu_fwd_pipe \
  python bench.py --roofline-only > gpurun_out/ncu_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cgconv_tc -s 3 -c 1 -f -o gpurun_out/ncu_bwd_dst \
  python bench.py --roofline-only > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out | tail -12
