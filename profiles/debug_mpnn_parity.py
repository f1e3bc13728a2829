"""Bisect the MPNN gradient mismatch: engine pieces vs the oracle's (both fp32 on the GPU) -- debug aid."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from matdeeplearn_b200 import models as M, process as pr, nn as mnn
from oracle import models as OM, pyg_ops as P
dev = "cuda:0"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ds = pr.synthetic_dataset("bulk", 32, seed=7, edge_length=G)
b = ds.batch(); b.num_graphs = 32
cfg = dict(dim1=64, dim2=64, dim3=64, pre_fc_count=1, gc_count=3, post_fc_count=1)
torch.manual_seed(0)
ref = OM.MPNN(ds, **cfg)
gb = b.to(dev)
mo = copy.deepcopy(ref).to(dev).train()
m = M.MPNN(ds, **cfg); m.load_state_dict(ref.state_dict()); m = m.to(dev).train()

def rel(a, r):
    return (a.double() - r.double()).abs().max().item() / max(r.double().abs().max().item(), 1e-30)

# (1) readout only
torch.manual_seed(1)
h0 = torch.randn(gb.x.shape[0], 64, device=dev)
ha, hb_ = h0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
M.MPNN.forward  # noqa
from matdeeplearn_b200.models import _prepare
_prepare(gb)
oa = m._readout(ha, gb); ob = mo._post(hb_, gb)
F.l1_loss(oa, gb.y).backward(); F.l1_loss(ob, gb.y).backward()
print("(1) readout: out", rel(oa, ob), "dh", rel(ha.grad, hb_.grad), "post_lin.w", rel(m.post_lin_list[0].weight.grad, mo.post_lin_list[0].weight.grad),
      "post_lin.b", rel(m.post_lin_list[0].bias.grad, mo.post_lin_list[0].bias.grad), "lin_out.w", rel(m.lin_out.weight.grad, mo.lin_out.weight.grad))
m.zero_grad(); mo.zero_grad()
# (2) one NNConv layer
xa, xb = h0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
ya = m.conv_list[0](xa, gb.edge_index, gb.edge_attr); yb = mo.conv_list[0](xb, gb.edge_index, gb.edge_attr)
w = torch.randn_like(yb)
(ya * w).sum().backward(); (yb * w).sum().backward()
print("(2) NNConv: out", rel(ya, yb), "dx", rel(xa.grad, xb.grad), {k: f"{rel(p.grad, dict(mo.conv_list[0].named_parameters())[k].grad):.2e}" for k, p in m.conv_list[0].named_parameters()})
m.zero_grad(); mo.zero_grad()
# (3) conv + bn + relu + gru, one layer
xa, xb = h0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
def layer(model, x, conv, eng):
    mm = conv(x, gb.edge_index, gb.edge_attr)
    mm = model.bn_list[0](mm)
    mm = F.relu(mm)
    out, hh = model.gru_list[0](mm.unsqueeze(0), x.unsqueeze(0))
    return out.squeeze(0)
ya = layer(m, xa, m.conv_list[0], True); yb = layer(mo, xb, mo.conv_list[0], False)
(ya * w).sum().backward(); (yb * w).sum().backward()
print("(3) layer: out", rel(ya, yb), "dx", rel(xa.grad, xb.grad), "gru.w_ih", rel(m.gru_list[0].weight_ih_l0.grad, mo.gru_list[0].weight_ih_l0.grad),
      "bn.w", rel(m.bn_list[0].weight.grad, mo.bn_list[0].weight.grad))
m.zero_grad(); mo.zero_grad()
# (4) whole model, engine vs oracle on GPU, then engine with oracle pool / oracle conv swapped in
def grads(model):
    model.zero_grad()
    out = model(gb); F.l1_loss(out, gb.y).backward()
    return out.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}
o_ref, g_ref = grads(mo)
def report(tag, model):
    o, g = grads(model)
    worst = sorted(((rel(g[k], g_ref[k]), k) for k in g if g_ref[k].abs().max() > 1e-12), reverse=True)[:3]
    print(tag, "out", rel(o, o_ref), "worst", [(f"{a:.2e}", k) for a, k in worst], "post_lin.b", f"{rel(g['post_lin_list.0.bias'], g_ref['post_lin_list.0.bias']):.2e}")
report("(4a) engine", m)
report("(4a') engine again", m)
import matdeeplearn_b200.nn as mnn_mod
saved = mnn_mod.global_mean_pool
mnn_mod.global_mean_pool = P.global_mean_pool
report("(4b) engine + oracle pool", m)
mnn_mod.global_mean_pool = saved
m2 = copy.deepcopy(m)
for i in range(3):
    m2.conv_list[i] = copy.deepcopy(mo.conv_list[i])
m2.forward = lambda data, mm=m2: M.MPNN.forward(mm, data)
class _Wrap(torch.nn.Module):
    def __init__(s, c): super().__init__(); s.c = c
    def forward(s, x, ei, ea, csr=None): return s.c(x, ei, ea)
for i in range(3):
    m2.conv_list[i] = _Wrap(m2.conv_list[i])
o, _ = None, None
m2.zero_grad(); out = m2(gb); F.l1_loss(out, gb.y).backward()
g2 = {k.replace(".c.", "."): p.grad for k, p in m2.named_parameters()}
worst = sorted(((rel(g2[k], g_ref[k]), k) for k in g2 if g_ref[k].abs().max() > 1e-12), reverse=True)[:3]
print("(4c) engine + oracle NNConv: out", rel(out, o_ref), "worst", [(f"{a:.2e}", k) for a, k in worst])
