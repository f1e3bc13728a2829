#!/bin/bash
# 2 GPUs, bounded: engine test across ranks, then the config-1 bench line at N = 2 (in-graph all-reduce, strong-scaling leg)
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/tests_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_multi.log
grep -E "^E  |FAILED|ERROR|passed|failed|skipped|pytest exit" gpurun_out/tests_multi.log | tail -12
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_c1.json 2> gpurun_out/bench_n2_c1.err
echo "n2 bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n2_c1.json"))
    print("  ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "store", round(d["store_step"]["value"]), "strong", d.get("strong_scaling"))
except Exception as e: print("parse failed", e)
PY
tail -3 gpurun_out/bench_n2_c1.err
