"""A few eager training steps of a BASELINE config's model (run under ncu for the launch list of one step):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/model_step_launches.py <config>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                  # noqa: E402
from matdeeplearn_b200 import models as M                      # noqa: E402
from matdeeplearn_b200.engine import TrainStep                 # noqa: E402

cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = bench.CONFIGS[cfgno]
dev = torch.device("cuda", 0)
ds, hb = bench.make_workload(0, c["graphs"], c["kind"], c.get("sweep", (50,))[0])
b = hb.to(dev)
b.num_graphs = c["graphs"]
torch.manual_seed(0)
model = getattr(M, c["model"])(ds, **c["cfg"]).to(dev).train()
step = TrainStep(model, lr=1e-3)
for _ in range(4):
    step.eager(b)
torch.cuda.synchronize()
