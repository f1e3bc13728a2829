#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_cgconv_smear.py -q -x > gpurun_out/tests_cgconv.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_cgconv.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_cgconv.log | tail -8
timeout 200 python profiles/phase_profile.py 16384 > gpurun_out/phase_profile.txt 2>&1
grep -A9 "== fwd" gpurun_out/phase_profile.txt; grep "A/B" gpurun_out/phase_profile.txt | head -1
timeout 200 python bench.py --roofline-only > gpurun_out/roofline_only.json 2> gpurun_out/roofline_only.err
python - <<PY
import json
d=json.load(open("gpurun_out/roofline_only.json"))
print("fwd", round(d["fwd"]["ms"],4), round(d["fwd"]["frac"],4), "fwd fused", round(d["fwd_smear_fused"]["ms"],4), "bwd", round(d["bwd_both_passes"]["ms"],4), round(d["bwd_both_passes"]["frac"],4))
PY
