"""Secondary measurement: one training step (zero_grad, forward, l1_loss, backward, AdamW; CUDA-graph
replayed, device-resident batch) for the four model families on the BASELINE.json config shapes, 1 GPU.
The bench line (bench.py) is configs[1]; these are the other configs' shapes for the record.
  python profiles/model_throughput.py  > gpurun_out/model_throughput.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matdeeplearn_b200 import models as M, process as pr  # noqa: E402
from matdeeplearn_b200.engine import TrainStep  # noqa: E402

CASES = [
    ("CGCNN", "bulk", 256, 50, dict(dim1=64, dim2=64, gc_count=4, post_fc_count=1)),
    ("SchNet", "bulk", 256, 50, dict(dim1=128, dim2=128, dim3=128, gc_count=4, post_fc_count=1, cutoff=8)),
    ("MEGNet", "mof", 64, 50, dict(dim1=128, dim2=128, dim3=128, gc_count=3, gc_fc_count=2, post_fc_count=1)),
    ("MPNN", "bulk", 256, 50, dict(dim1=64, dim2=64, dim3=64, gc_count=4, post_fc_count=1)),
    ("MPNN", "bulk", 256, 100, dict(dim1=64, dim2=64, dim3=64, gc_count=4, post_fc_count=1)),
    ("MPNN", "bulk", 256, 200, dict(dim1=64, dim2=64, dim3=64, gc_count=4, post_fc_count=1)),
]
dev = torch.device("cuda:0")
flush = torch.zeros(128 * 1024 * 1024, device=dev)
out = []
for name, kind, graphs, G, cfg in CASES:
    ds = pr.synthetic_dataset(kind, graphs, seed=pr.BENCH_SEED, edge_length=G)
    b = ds.batch().to(dev)
    b.num_graphs = graphs
    torch.manual_seed(0)
    model = getattr(M, name)(ds, **cfg).to(dev).train()
    step = TrainStep(model, lr=1e-3)
    replay = step.resident(b, warmup=3)
    for _ in range(5):
        replay()
    ts = []
    for _ in range(20):
        flush.add_(1.0)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); replay(); c.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    ms = sum(ts) / len(ts)
    rec = {"model": name, "graphs": graphs, "shape": kind, "G": G, "nodes": int(b.x.shape[0]),
           "edges": int(b.edge_index.shape[1]), "params": sum(p.numel() for p in model.parameters()),
           "ms_per_step": ms, "graphs_per_s": graphs / ms * 1e3, "edges_per_s": b.edge_index.shape[1] / ms * 1e3}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del model, step, replay
    torch.cuda.empty_cache()
