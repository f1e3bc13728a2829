"""MEGNet (MOF-shaped batch) on a capacity-padded static batch vs the exactly assembled batch: activations and their
gradients after every op of every _MegnetStack, real rows only (development aid)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from matdeeplearn_b200 import models as M, process as pr, functional as MF
from matdeeplearn_b200.store import GraphStore
dev = "cuda:0"
cfg = dict(dim1=64, dim2=32, dim3=64, pre_fc_count=1, gc_count=2, gc_fc_count=2, post_fc_count=1)
ds = pr.synthetic_dataset("mof", 14, seed=5)
store = GraphStore.from_dataset(ds, dev)
torch.manual_seed(0)
model = M.MEGNet(ds, **cfg)
idx = np.random.default_rng(4).permutation(len(ds))[:12]
REC = []

def _run(self, h, first_done=False, n_valid=None):
    layers = getattr(self, self._name)
    for i, lin in enumerate(layers):
        if not (i == 0 and first_done):
            pre = MF.linear(h, lin.weight, lin.bias)
            pre.retain_grad(); REC.append((f"{self._name}.{i}.pre", pre))
            h = getattr(F, self.act)(pre)
        h.retain_grad(); REC.append((f"{self._name}.{i}.act", h))
        if self.batch_norm == "True":
            h = MF.masked_batch_norm(self.bn_list[i], h, n_valid) if (n_valid is not None or os.environ.get("FORCE_OWN_BN")) else self.bn_list[i](h)
        h.retain_grad(); REC.append((f"{self._name}.{i}.bn", h))
    return h
M._MegnetStack._run = _run

def go(batch):
    REC.clear()
    m = copy.deepcopy(model).to(dev).train()
    out = m(batch)
    loss = F.l1_loss(out, batch.y)
    loss.backward()
    return [(n, t.detach(), t.grad.detach()) for n, t in REC], loss.item()

static = store.static_batch(12, lazy=True)
assert store.load(static, idx)
store.assemble(static)
exact = store.batch(idx)
# isolated check: the same rows through torch mm at the two row counts
torch.manual_seed(1)
gA = torch.randn(2625, 64, device=dev) * 1e-3; W = torch.randn(64, 64, device=dev) * 0.1
d1 = gA.mm(W)[:2380]; d2 = gA[:2380].contiguous().mm(W)
print("mm rows 2625 vs 2380: max diff", (d1 - d2).abs().max().item(), "scale", d2.abs().max().item(),
      "vs fp64", (d2.double() - gA[:2380].double().mm(W.double())).abs().max().item())
for force in ("1",):
    if force: os.environ["FORCE_OWN_BN"] = "1"
    ra, la = go(static)
    rb, lb = go(exact)
    print(f"== exact path BN = {'own kernels' if force else 'torch'}: loss padded {la:.7f} exact {lb:.7f}")
    for (n, ta, ga), (_, tb, gb) in zip(ra, rb):
        r = tb.shape[0]
        dt = (ta[:r] - tb).abs().max().item(); dg = (ga[:r] - gb).abs().max().item()
        pad_t = ta[r:].abs().max().item() if ta.shape[0] > r else 0.0
        pad_g = ga[r:].abs().max().item() if ga.shape[0] > r else 0.0
        print(f"   {n:22s} rows {r:6d}/{ta.shape[0]:6d}  act diff {dt:.2e} (scale {tb.abs().max().item():.2e})  grad diff {dg:.2e} (scale {gb.abs().max().item():.2e})  padding: act {pad_t:.2e} grad {pad_g:.2e}")
