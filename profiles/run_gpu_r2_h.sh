#!/bin/bash
# node rows by cp.async (forward loaders, backward value tile): parity, A/B, phase profile; whole suite; final config-1 line;
# ncu --set full of the two CGConv kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_cgconv_smear.py tests/test_gpu_engine.py tests/test_gpu_graphops.py -q -x > gpurun_out/tests_cgconv.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_cgconv.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_cgconv.log | tail -12
for q in cpasync bulk; do
  if [ $q = bulk ]; then export MDL_CGCONV_ROWS=bulk; else unset MDL_CGCONV_ROWS; fi
  timeout 200 python bench.py --roofline-only > gpurun_out/roofline_rows_$q.json 2> gpurun_out/roofline_rows_$q.err
  python - <<PY
import json
d=json.load(open("gpurun_out/roofline_rows_$q.json"))
print("rows=$q", "fwd", round(d["fwd"]["ms"],4), "fwd fused", round(d["fwd_smear_fused"]["ms"],4), "bwd", round(d["bwd_both_passes"]["ms"],4), "bwd fused", round(d["bwd_smear_fused"]["ms"],4))
PY
done
unset MDL_CGCONV_ROWS
timeout 200 python profiles/phase_profile.py 16384 > gpurun_out/phase_profile.txt 2>&1
grep -A9 "== fwd" gpurun_out/phase_profile.txt; grep -A17 "== bwd" gpurun_out/phase_profile.txt | head -18; grep "A/B" gpurun_out/phase_profile.txt | head -2
rm -f gpurun_out/parity_errors.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu_full.log
grep -E "FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_gpu_full.log | tail -10
timeout 600 python bench.py > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
echo "bench exit $?"; head -c 300 gpurun_out/bench_c1.json; echo; tail -2 gpurun_out/bench_c1.err
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_cgconv_fwd|k_cgconv_bwd" -s 12 -c 2 \
   -o gpurun_out/r2_ncu_full_cgconv_final -f python bench.py --roofline-only > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
