#!/bin/bash
# NNConv kernels A/B + parity, whole GPU suite, the four bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_baseline_shapes.py tests/test_gpu_graphops.py -q > gpurun_out/tests_models.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_models.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_models.log | tail -20
for w in cta default; do
  if [ $w = cta ]; then export MDL_NNCONV=cta; else unset MDL_NNCONV; fi
  timeout 300 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_nnconv_$w.json 2> gpurun_out/ab_nnconv_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_nnconv_$w.json"))
print("MDL_NNCONV=$w", "ms/step", round(d["ms_per_step"],4), "roofline kernel ms", d["roofline"]["ms"], "store", d.get("store_step"))
PY
done
unset MDL_NNCONV
rm -f gpurun_out/parity_errors.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -15
for c in 1 2 3; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --cpu-steps 2 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  echo "config $c exit $?"; head -c 250 gpurun_out/bench_c$c.json; echo; tail -2 gpurun_out/bench_c$c.err
done
