"""engine MPNN vs fp64 oracle: do ReLU masks differ (a pre-activation within rounding of the kink)?"""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from matdeeplearn_b200 import models as M, process as pr, nn as mnn
from matdeeplearn_b200.models import _prepare
from oracle import models as OM, pyg_ops as P
dev = "cuda:0"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ds = pr.synthetic_dataset("bulk", 32, seed=7, edge_length=G)
b = ds.batch(); b.num_graphs = 32
cfg = dict(dim1=64, dim2=64, dim3=64, pre_fc_count=1, gc_count=3, post_fc_count=1)
torch.manual_seed(0)
ref = OM.MPNN(ds, **cfg)
m = M.MPNN(ds, **cfg); m.load_state_dict(ref.state_dict()); m = m.to(dev).train()
r = copy.deepcopy(ref).double().train()
gb, b64 = b.to(dev), b.double()

def trace(model, data, eng):
    rec = {}
    with torch.no_grad():
        if eng:
            csr = _prepare(data)
            out = model._embed(data)
        else:
            out = model._pre(data)
        rec["pre_fc"] = out
        hidden = out.unsqueeze(0)
        for i, conv in enumerate(model.conv_list):
            # edge-net hidden pre-activation
            rec[f"edge_pre{i}"] = conv.nn[0](data.edge_attr)
            mm = conv(out, data.edge_index, data.edge_attr, csr=csr) if eng else conv(out, data.edge_index, data.edge_attr)
            mm = model.bn_list[i](mm)
            rec[f"bn{i}"] = mm
            mm = F.relu(mm)
            out, hidden = model.gru_list[i](mm.unsqueeze(0), hidden)
            out = out.squeeze(0)
        pooled = (mnn if eng else P).global_mean_pool(out, data.batch)
        rec["pooled"] = pooled
        rec["post_pre"] = F.linear(pooled, model.post_lin_list[0].weight, model.post_lin_list[0].bias)
    return rec
ta, tb = trace(m, gb, True), trace(r, b64, False)
for k in ta:
    a, c = ta[k].double().cpu(), tb[k]
    flips = int(((a > 0) != (c > 0)).sum())
    print(f"{k:10s} max|err| {(a - c).abs().max().item():.2e}  min|ref| {c.abs().min().item():.2e}  exact zeros in ref {int((c == 0).sum())}  sign flips {flips}")
