#!/bin/bash
# backward with bulk-reduction dQ: parity, A/B, phase profile; edge-level kernels: timings + one ncu --set full capture each
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cgconv.py tests/test_gpu_cgconv_smear.py tests/test_gpu_engine.py -q -x > gpurun_out/tests_cgconv.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_cgconv.log
grep -E "^E  |FAILED|ERROR|passed|failed|pytest exit" gpurun_out/tests_cgconv.log | tail -12
for q in bulk red; do
  if [ $q = red ]; then export MDL_CGCONV_DQ=red; else unset MDL_CGCONV_DQ; fi
  timeout 300 python bench.py --roofline-only > gpurun_out/roofline_dq_$q.json 2> gpurun_out/roofline_dq_$q.err
  python - <<PY
import json
d=json.load(open("gpurun_out/roofline_dq_$q.json"))
print("dQ=$q", "fwd", round(d["fwd"]["ms"],4), "bwd", round(d["bwd_both_passes"]["ms"],4), "bwd fused", round(d["bwd_smear_fused"]["ms"],4))
PY
done
unset MDL_CGCONV_DQ
timeout 300 python profiles/phase_profile.py 16384 > gpurun_out/phase_profile.txt 2>&1
grep -A18 "== bwd" gpurun_out/phase_profile.txt | head -20
timeout 300 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c4.json"))
print("config 4 ms/step", round(d["ms_per_step"],4), "roofline kernel ms", d["roofline"]["ms"])
PY
timeout 300 python profiles/edge_kernel_timing.py > gpurun_out/edge_kernel_timing.txt 2>&1; cat gpurun_out/edge_kernel_timing.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_wgrad_tc|k_linear_tc|k_edge_mlp2_fwd|k_nnconv_msg_w" -s 40 -c 6 \
   -o gpurun_out/r2_ncu_full_edge -f python profiles/edge_kernel_timing.py > gpurun_out/ncu_edge.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_edge.log
