"""Read datasets the reference has already processed, without torch_geometric.

The reference's `process_data` ends by writing either ONE file `processed/data.pt`
holding `(data, slices) = InMemoryDataset.collate(data_list)` (dataset_type
"inmemory", matdeeplearn/process/process.py:521-523; read back by
`StructureDataset`, process.py:66-98) or one `data_{i}.pt` per structure
(dataset_type "large", process.py:525-532).  Both are pickles of
`torch_geometric.data.Data` objects, so `torch.load` needs that package just to
rebuild the container class -- the tensors inside are plain torch tensors.

`load_processed(path)` unpickles them with every `torch_geometric.*` class
replaced by an attribute bag, finds the attribute mapping whatever the PyG
version's internal layout (1.x: attributes in `__dict__`; 2.x: `_store._mapping`),
cuts the collated tensors back into graphs along `slices`, and returns the
engine's `GraphDataset` -- ready for `GraphStore.from_dataset` / model
constructors.  Non-tensor attributes (`structure_id`, `length`) are kept as
Python objects where they slice cleanly and dropped otherwise.
"""
from __future__ import annotations

import glob
import os
import pickle
import re

import torch

from .data import Data, GraphDataset

_GRAPH_KEYS = ("x", "edge_index", "edge_attr", "edge_weight", "u", "y", "z", "pos")


class _Bag:
    """Stand-in for any torch_geometric class met while unpickling: keeps whatever state it is given."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict, slots)
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch_geometric" or module.startswith("torch_geometric."):
            return type(name, (_Bag,), {"__module__": "matdeeplearn_b200.compat"})
        return super().find_class(module, name)


class _PickleModule:
    """What torch.load(pickle_module=...) expects: a module-like object with Unpickler / load."""
    __name__ = "matdeeplearn_b200.compat._PickleModule"
    Unpickler = _Unpickler

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()


def _mapping(obj):
    """attribute name -> value of a (stubbed) PyG Data object, for PyG 1.x and 2.x layouts."""
    if isinstance(obj, dict):
        return obj
    d = getattr(obj, "__dict__", {})
    store = d.get("_store")
    if store is not None:                       # PyG >= 2.0: Data._store (GlobalStorage) ._mapping
        m = getattr(store, "__dict__", {}).get("_mapping")
        if isinstance(m, dict):
            return m
        if isinstance(store, dict):
            return store
    return {k: v for k, v in d.items() if not k.startswith("_")}


def _cut(key, value, lo, hi):
    if torch.is_tensor(value):
        if key == "edge_index" or (value.dim() == 2 and key.endswith("index")):
            return value[:, lo:hi].clone()       # concatenated along the edge axis, node ids stay graph-local
        return value[lo:hi].clone()
    if isinstance(value, (list, tuple)):
        seg = value[lo:hi]
        return seg[0] if len(seg) == 1 else list(seg)
    return value


def _to_graph(m, target_index=0):
    g = {}
    for k, v in m.items():
        if v is None or k in ("ase", "edge_descriptor"):
            continue
        g[k] = v
    d = Data(**g)
    if hasattr(d, "u") and torch.is_tensor(d.u) and d.u.dim() == 1:
        d.u = d.u.reshape(1, -1)
    if hasattr(d, "y") and torch.is_tensor(d.y) and target_index != -1 and d.y.dim() == 2:
        # the file holds y as [1, T] (process.py:321-322); StructureDataset applies GetY(index) on access
        # (process.py:93, 695-703): a 0-dim target.  target_index = -1 keeps all T columns, as there.
        d.y = d.y[0][target_index]
    return d


def graphs_from_collated(data, slices, target_index=0):
    """(data, slices) of InMemoryDataset.collate -> list of per-graph `Data`."""
    m = _mapping(data)
    s = _mapping(slices)
    if "x" not in s and "edge_index" not in s:
        raise ValueError("data.pt: no 'x' / 'edge_index' slices -- not a MatDeepLearn processed dataset?")
    n_graphs = int(s["x" if "x" in s else "edge_index"].shape[0]) - 1
    graphs = []
    for i in range(n_graphs):
        g = {}
        for k, sl in s.items():
            if k not in m or m[k] is None:
                continue
            lo, hi = int(sl[i]), int(sl[i + 1])
            g[k] = _cut(k, m[k], lo, hi)
        graphs.append(_to_graph(g, target_index))
    return graphs


def _torch_load(path):
    return torch.load(path, map_location="cpu", pickle_module=_PickleModule, weights_only=False)


def load_processed(path, target_index=0):
    """`path`: a `data.pt` file, or a `processed/` directory holding `data.pt` (in-memory datasets) or
    `data_{i}.pt` files (large datasets).  Returns a GraphDataset of float32 / int64 host tensors in the
    reference's layout (x, edge_index [2,E] graph-local, edge_attr, edge_weight, u [1,3], y); `target_index`
    is the reference's `target_index` job setting (which column of y a model trains on)."""
    if os.path.isdir(path):
        single = os.path.join(path, "data.pt")
        if os.path.exists(single):
            path = single
        else:
            files = glob.glob(os.path.join(path, "data_*.pt"))
            files.sort(key=lambda f: int(re.search(r"data_(\d+)\.pt$", f).group(1)))
            if not files:
                raise FileNotFoundError(f"{path}: neither data.pt nor data_<i>.pt found")
            return GraphDataset([_to_graph(_mapping(_torch_load(f)), target_index) for f in files])
    obj = _torch_load(path)
    if isinstance(obj, (tuple, list)) and len(obj) == 2:
        return GraphDataset(graphs_from_collated(obj[0], obj[1], target_index))
    return GraphDataset([_to_graph(_mapping(obj), target_index)])
