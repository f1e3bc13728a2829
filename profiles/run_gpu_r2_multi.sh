#!/bin/bash
# 2 GPUs: engine test across ranks (NCCL), bench lines of configs 1 and 2 (weak + strong-scaling leg), reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/tests_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_multi.log
grep -E "^E  |FAILED|ERROR|passed|failed|skipped|pytest exit" gpurun_out/tests_multi.log | tail -12
for c in 1 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 20 --warmup 3 > gpurun_out/bench_n2_c$c.json 2> gpurun_out/bench_n2_c$c.err
  echo "n2 config $c exit $?"; head -c 400 gpurun_out/bench_n2_c$c.json; echo; tail -3 gpurun_out/bench_n2_c$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n2_c$c.json"))
    print("  ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "strong", d.get("strong_scaling"))
except Exception as e: print("parse failed", e)
PY
done
MDL_GRAPH_ALLREDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 1 --steps 20 --warmup 3 --no-roofline > gpurun_out/bench_n2_c1_eager_allreduce.json 2> gpurun_out/bench_n2_c1_eager_allreduce.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n2_c1_eager_allreduce.json"))
    print("eager all-reduce: ms/step", round(d["ms_per_step"],4), "value", round(d["value"]))
except Exception as e: print("parse failed", e)
PY
