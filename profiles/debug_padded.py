"""Per-parameter gradient difference between a model run on a capacity-padded static batch and on the exactly
assembled batch (development aid for the padded replay of SchNet / MPNN / MEGNet)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from matdeeplearn_b200 import models as M, process as pr
from matdeeplearn_b200.store import GraphStore
from matdeeplearn_b200.engine import TrainStep
dev = "cuda:0"
CASES = [("SchNet", "bulk", dict(dim1=32, dim2=32, dim3=48, cutoff=8, pre_fc_count=1, gc_count=2, post_fc_count=1)),
         ("MPNN", "bulk", dict(dim1=32, dim2=32, dim3=32, pre_fc_count=1, gc_count=2, post_fc_count=1)),
         ("MEGNet", "bulk", dict(dim1=32, dim2=32, dim3=32, pre_fc_count=1, gc_count=2, gc_fc_count=1, post_fc_count=1)),
         ("MEGNet", "mof", dict(dim1=64, dim2=32, dim3=64, pre_fc_count=1, gc_count=2, gc_fc_count=2, post_fc_count=1)),
         ("CGCNN", "bulk", dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1))]
for name, kind, cfg in CASES:
    ds = pr.synthetic_dataset(kind, 48 if kind == "bulk" else 14, seed=5)
    store = GraphStore.from_dataset(ds, dev)
    torch.manual_seed(0)
    model = getattr(M, name)(ds, **cfg)
    idx = np.random.default_rng(4).permutation(len(ds))[:12]
    for direct in (False, True):
        m1, m2 = copy.deepcopy(model).to(dev).train(), copy.deepcopy(model).to(dev).train()
        if direct:
            s1, s2 = TrainStep(m1, lr=1e-3), TrainStep(m2, lr=1e-3)
        static = store.static_batch(12, lazy=True)
        assert store.load(static, idx)
        store.assemble(static)
        exact = store.batch(idx)
        try:
            if direct:
                l1 = s1._fwd_bwd(static); l2 = s2._fwd_bwd(exact)
                g1 = {n: p._mdl_grad_dest.clone() for n, p in m1.named_parameters()}
                g2 = {n: p._mdl_grad_dest.clone() for n, p in m2.named_parameters()}
            else:
                l1 = torch.nn.functional.l1_loss(m1(static), static.y); l1.backward()
                l2 = torch.nn.functional.l1_loss(m2(exact), exact.y); l2.backward()
                g1 = {n: p.grad for n, p in m1.named_parameters()}
                g2 = {n: p.grad for n, p in m2.named_parameters()}
        except Exception as e:
            print(name, "direct" if direct else "autograd", "FAILED", repr(e)[:300])
            continue
        print(f"== {name} {'direct' if direct else 'autograd'}: loss padded {l1.item():.7f} exact {l2.item():.7f}")
        for n in g1:
            d = (g1[n] - g2[n]).abs().max().item()
            sc = g2[n].abs().max().item()
            flag = "  <<<" if d > 1e-4 * max(sc, 1e-6) else ""
            print(f"   {n:40s} max|diff| {d:.3e}  scale {sc:.3e}{flag}")
