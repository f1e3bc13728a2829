"""Launch list of the GraphStore training step (run under ncu):  a few TrainStep.from_store replays."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matdeeplearn_b200 import models as M, process as pr   # noqa: E402
from matdeeplearn_b200.engine import TrainStep               # noqa: E402
from matdeeplearn_b200.store import GraphStore               # noqa: E402

dev = torch.device("cuda", 0)
ds = pr.synthetic_dataset("bulk", 512, seed=pr.BENCH_SEED)
cfg = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=4, post_fc_count=1, pool="global_mean_pool",
           pool_order="early", batch_norm="True", batch_track_stats="True", act="relu", dropout_rate=0.0)
torch.manual_seed(0)
model = M.CGCNN(ds, **cfg).to(dev).train()
step = TrainStep(model, lr=1e-3)
store = GraphStore.from_dataset(ds, dev)
rng = np.random.default_rng(0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    step.from_store(store, rng.permutation(512)[:256])
torch.cuda.synchronize()
print("kernels per store step:", step.store_kernels_per_step)
