#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-roofline > gpurun_out/bench_n2_c1.json 2> gpurun_out/bench_n2_c1.err
echo "n2 bench exit $?"
python - <<PY
import json
txt=open("gpurun_out/bench_n2_c1.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][0])
print("  ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "store", round(d["store_step"]["value"]), "strong", round(d["strong_scaling"]["value"]), d["config"]["step"][:110])
PY
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
echo "n2 reference arm exit $?"; head -c 300 gpurun_out/bench_n2_ref.json; echo
