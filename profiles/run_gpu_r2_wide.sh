#!/bin/bash
# wide-layer (C > 64) tensor-core CGConv + everything else that changed: CGConv suites first, then the whole GPU suite,
# then timings of C = 128 / 100 layers (tensor-core chunks vs SIMT) and the 4 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_cgconv_smear.py -x -q > gpurun_out/tests_cgconv.log 2>&1
echo "pytest cgconv exit $?" >> gpurun_out/tests_cgconv.log
tail -8 gpurun_out/tests_cgconv.log
rm -f gpurun_out/parity_errors.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_cgconv.py --deselect tests/test_gpu_cgconv_smear.py > gpurun_out/tests_gpu_rest.log 2>&1
echo "pytest rest exit $?" >> gpurun_out/tests_gpu_rest.log
grep -E "FAILED|ERROR|passed|failed|pytest rest exit" gpurun_out/tests_gpu_rest.log | tail -15
timeout 300 python profiles/wide_layer_timing.py > gpurun_out/wide_layer_timing.txt 2>&1; cat gpurun_out/wide_layer_timing.txt
for c in 1 2 3 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --cpu-steps 2 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  echo "config $c exit $?"; head -c 300 gpurun_out/bench_c$c.json; echo; tail -2 gpurun_out/bench_c$c.err
done
