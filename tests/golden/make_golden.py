"""Generate the golden fixtures in tests/golden/ by EXECUTING THE REFERENCE'S OWN
PYTHON SOURCES from /root/reference (read-only; nothing is copied).

The reference cannot be imported as a package here: torch_geometric,
torch_scatter and ase are not installed.  This script
registers stub modules for those names and loads the reference files one by
one with importlib:

  * matdeeplearn/process/process.py  -> threshold_sort, GaussianSmearing,
    OneHotDegree, NormalizeEdge run AS-IS.  The three PyG utilities they call
    (dense_to_sparse, degree, add_self_loops) are stubbed with their documented
    semantics (nonzero scan / bincount / append loops).
  * matdeeplearn/models/{cgcnn,schnet,mpnn,megnet,gcn}.py -> the model glue (pre/post
    FC, BN placement, residuals, GRU threading, MEGNet wiring) runs AS-IS, with
    torch_geometric.nn.{CGConv,NNConv,MetaLayer,global_*_pool},
    InteractionBlock and torch_scatter.scatter* bound to oracle/pyg_ops.py.

So the fixtures pin (a) the builder functions outright and (b) the glue, given
the oracle's operator restatement.  The PyG operator arithmetic itself stays
"parity unpinned" (oracle/__init__.py).

Run from the repo root in the dev container:  python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/matdeeplearn"
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import pyg_ops as O  # noqa: E402
from matdeeplearn_b200 import process as pr  # noqa: E402  (only to synthesise INPUT graphs)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    class _Dummy:
        def __init__(self, *a, **k):
            pass

    def dense_to_sparse(adj):
        idx = adj.nonzero(as_tuple=False).t().contiguous()
        return idx, adj[idx[0], idx[1]]

    def degree(index, num_nodes=None, dtype=None):
        out = torch.bincount(index, minlength=int(num_nodes))
        return out.to(dtype) if dtype is not None else out

    def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
        loops = torch.arange(num_nodes, dtype=torch.long)
        ei = torch.cat([edge_index, torch.stack([loops, loops])], dim=1)
        ew = torch.cat([edge_weight, torch.full((num_nodes,), float(fill_value), dtype=edge_weight.dtype)])
        return ei, ew

    ase = _module("ase")
    ase.io = _module("ase.io")
    tg = _module("torch_geometric")
    tg.data = _module("torch_geometric.data", DataLoader=_Dummy, Dataset=_Dummy, Data=_Dummy,
                      InMemoryDataset=_Dummy)
    tg.utils = _module("torch_geometric.utils", dense_to_sparse=dense_to_sparse, degree=degree,
                       add_self_loops=add_self_loops)
    tg.transforms = _module("torch_geometric.transforms")
    tg.nn = _module("torch_geometric.nn", Set2Set=O.Set2Set, global_mean_pool=O.global_mean_pool,
                    global_add_pool=O.global_add_pool, global_max_pool=O.global_max_pool,
                    CGConv=O.CGConv, NNConv=O.NNConv, MetaLayer=O.MetaLayer, GCNConv=O.GCNConv)
    tg.nn.models = _module("torch_geometric.nn.models")
    tg.nn.models.schnet = _module("torch_geometric.nn.models.schnet", InteractionBlock=O.InteractionBlock)
    _module("torch_scatter", scatter=O.scatter, scatter_mean=O.scatter_mean, scatter_add=O.scatter_add,
            scatter_max=O.scatter_max)


def load_ref(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_process(P):
    rng = np.random.default_rng(11)
    out = {}
    for i, (n, radius, k) in enumerate([(10, 8.0, 12), (37, 8.0, 12), (60, 5.0, 6), (25, 3.0, 4)]):
        numbers, pos, cell = pr._random_structure(rng, n, 0.06)
        D = pr.pairwise_distances(pos, cell)
        if i == 3:  # exact ties and a coincident pair: exercises ordinal ranking / zero dropping
            D = np.round(D, 0)
        out[f"dist_{i}"] = D
        out[f"args_{i}"] = np.array([radius, k])
        out[f"trimmed_{i}"] = P.threshold_sort(D, radius, k, adj=False)
    d = torch.linspace(0, 1, 23) ** 2
    for G in (50, 100, 200):
        gs = P.GaussianSmearing(0, 1, G, 0.2)
        out[f"smear_{G}"] = gs(d).numpy()
        out[f"smear_offset_{G}"] = gs.offset.numpy()
        out[f"smear_coeff_{G}"] = np.array(gs.coeff)
    out["smear_in"] = d.numpy()

    # OneHotDegree + NormalizeEdge on reference-style data objects
    class Obj:
        pass
    objs = []
    for i in range(3):
        D = out[f"dist_{i}"]
        t = torch.Tensor(P.threshold_sort(D, 8.0, 12, adj=False))
        ei, ew = sys.modules["torch_geometric.utils"].dense_to_sparse(t)
        ei, ew = sys.modules["torch_geometric.utils"].add_self_loops(ei, ew, num_nodes=D.shape[0], fill_value=0)
        o = Obj()
        o.edge_index, o.x, o.num_nodes = ei, torch.zeros(D.shape[0], 3), D.shape[0]
        o.edge_descriptor = {"distance": ew}
        objs.append(o)
        out[f"ei_{i}"] = ei.numpy()
        out[f"ew_{i}"] = ew.numpy()
    for i, o in enumerate(objs):
        o2 = P.OneHotDegree(o, 13)
        out[f"onehotdeg_{i}"] = o2.x.numpy()
    P.NormalizeEdge(objs, "distance")
    for i, o in enumerate(objs):
        out[f"norm_{i}"] = o.edge_descriptor["distance"].numpy()
    np.savez_compressed(os.path.join(OUT, "process_golden.npz"), **out)
    print("process_golden.npz:", len(out), "arrays")


MODEL_CFGS = {
    "CGCNN": dict(dim1=32, dim2=24, pre_fc_count=1, gc_count=3, post_fc_count=2),
    "SchNet": dict(dim1=32, dim2=24, dim3=40, cutoff=8, pre_fc_count=1, gc_count=3, post_fc_count=2),
    "MPNN": dict(dim1=16, dim2=24, dim3=20, pre_fc_count=1, gc_count=2, post_fc_count=1),
    "MEGNet": dict(dim1=32, dim2=24, dim3=28, pre_fc_count=1, gc_count=3, gc_fc_count=2, post_fc_count=2),
    "CGCNN_late_add": dict(dim1=32, dim2=24, pre_fc_count=1, gc_count=2, post_fc_count=1,
                           pool="global_add_pool", pool_order="late"),
    "SchNet_nobn_max": dict(dim1=32, dim2=24, dim3=32, cutoff=8, pre_fc_count=1, gc_count=2,
                            post_fc_count=1, batch_norm="False", pool="global_max_pool"),
    "MEGNet_fc1": dict(dim1=32, dim2=24, dim3=28, pre_fc_count=1, gc_count=2, gc_fc_count=1,
                       post_fc_count=1),
    "GCN": dict(dim1=32, dim2=24, pre_fc_count=1, gc_count=3, post_fc_count=2),
    "CGCNN_set2set": dict(dim1=32, dim2=24, pre_fc_count=1, gc_count=2, post_fc_count=1, pool="set2set"),
    "GCN_set2set_late": dict(dim1=32, dim2=24, pre_fc_count=1, gc_count=2, post_fc_count=1, pool="set2set",
                             pool_order="late"),
    "GCN_nobn_late": dict(dim1=32, dim2=24, pre_fc_count=1, gc_count=2, post_fc_count=1, batch_norm="False",
                          pool="global_add_pool", pool_order="late"),
}


def golden_models(only=None):
    ds = pr.synthetic_dataset("bulk", 6, seed=5, edge_length=50)
    batch = ds.batch()
    if only is None:
        np.savez_compressed(os.path.join(OUT, "batch_inputs.npz"),
                            x=batch.x.numpy(), edge_index=batch.edge_index.numpy(),
                            edge_attr=batch.edge_attr.numpy(), edge_weight=batch.edge_weight.numpy(),
                            batch=batch.batch.numpy(), u=batch.u.numpy(), y=batch.y.numpy())
    mods = {n: load_ref(f"models/{n}.py", f"ref_{n}") for n in ("cgcnn", "schnet", "mpnn", "megnet", "gcn")}
    classes = {"CGCNN": mods["cgcnn"].CGCNN, "SchNet": mods["schnet"].SchNet,
               "MPNN": mods["mpnn"].MPNN, "MEGNet": mods["megnet"].MEGNet, "GCN": mods["gcn"].GCN}
    for tag, cfg in MODEL_CFGS.items():
        if only is not None and tag.split("_")[0] not in only and tag not in only:
            continue
        cls = classes[tag.split("_")[0]]
        torch.manual_seed(1234)
        model = cls(data=ds, **cfg).double()
        state0 = {k: v.clone() for k, v in model.state_dict().items()}  # before BN stats move
        model.train()
        out = model(batch.double())
        loss = torch.nn.functional.l1_loss(out, batch.y.double())
        loss.backward()
        rec = {"out_train": out.detach().numpy(), "loss": np.array(loss.item())}
        for k, v in state0.items():
            rec["param/" + k] = v.numpy()
        for k, p in model.named_parameters():
            rec["grad/" + k] = p.grad.numpy() if p.grad is not None else np.zeros(0)
        model.eval()
        with torch.no_grad():
            rec["out_eval"] = model(batch.double()).numpy()
        np.savez_compressed(os.path.join(OUT, f"model_{tag}.npz"), **rec)
        print(f"model_{tag}.npz: out {rec['out_train'][:3]} loss {loss.item():.6f}")


def golden_test_data():
    """BASELINE configs[0]: the reference's bundled data/test_data fixture (Pt10 clusters), first 32
    structures, default CGCNN_demo hyper-parameters (config.yml:121-136).  Inputs built by the product's
    host builder from the tarball; outputs by the reference's cgcnn.py glue on the oracle ops, fp64."""
    tar = "/root/reference/data/test_data/test_data.tar.gz"
    ds = pr.load_ase_json_tar(tar, limit=32)
    batch = ds.batch()
    cfg = dict(dim1=100, dim2=150, pre_fc_count=1, gc_count=4, post_fc_count=3)
    mod = load_ref("models/cgcnn.py", "ref_cgcnn_td")
    torch.manual_seed(4321)
    model = mod.CGCNN(data=ds, **cfg).double()
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    model.train()
    out = model(batch.double())
    loss = torch.nn.functional.l1_loss(out, batch.y.double())
    loss.backward()
    rec = {"out_train": out.detach().numpy(), "loss": np.array(loss.item())}
    for k in ("x", "edge_index", "edge_attr", "edge_weight", "batch", "u", "y"):
        rec["in/" + k] = getattr(batch, k).numpy()
    for k, v in state0.items():
        rec["param/" + k] = v.numpy().astype(np.float32) if v.is_floating_point() else v.numpy()
    for k, p in model.named_parameters():
        rec["grad/" + k] = p.grad.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "testdata_cgcnn_demo_b32.npz"), **rec)
    print("testdata_cgcnn_demo_b32.npz: loss", loss.item(), "E", batch.edge_index.shape[1])


if __name__ == "__main__":
    install_stubs()
    if len(sys.argv) > 2 and sys.argv[1] == "--only":  # e.g. --only GCN : add fixtures of one model family
        golden_models(only=set(sys.argv[2:]))
    else:
        P = load_ref("process/process.py", "ref_process")
        golden_process(P)
        golden_models()
        golden_test_data()
