#!/bin/bash
# CGConv suites (wide layers), padded replay of the other model families, config-1 step A/B (weight-gradient kernel choice)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_cgconv_smear.py tests/test_gpu_store.py tests/test_gpu_engine.py -q > gpurun_out/tests_cgconv.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_cgconv.log
grep -E "FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_cgconv.log | tail -25
for w in default tc; do
  if [ $w = tc ]; then export MDL_WGRAD=tc; else unset MDL_WGRAD; fi
  timeout 300 python bench.py --config 1 --steps 50 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/ab_wgrad_$w.json 2> gpurun_out/ab_wgrad_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_wgrad_$w.json"))
print("MDL_WGRAD=$w", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "store", round(d["store_step"]["value"]))
PY
done
unset MDL_WGRAD
