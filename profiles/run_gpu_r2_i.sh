#!/bin/bash
# weight-gradient kernels of the CGConv backward on a side stream: parity / engine tests, step A/B, smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_store.py tests/test_gpu_cgconv_smear.py tests/test_gpu_models.py -q -x > gpurun_out/tests_engine.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_engine.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_engine.log | tail -12
for o in 1 0; do
  MDL_BWD_OVERLAP=$o timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-roofline --no-other-configs > gpurun_out/ab_overlap_$o.json 2> gpurun_out/ab_overlap_$o.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_overlap_$o.json"))
print("MDL_BWD_OVERLAP=$o", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "shipped", round(d["e2e"]["distances_shipped"]["value"]), "store", round(d["store_step"]["value"]))
PY
done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
